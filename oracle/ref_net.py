"""fp32 eager PyTorch restatement of the reference evaluator (CPU ORACLE -- test infrastructure only).

Follows network.py:21-104: five pre-activation residual blocks (BatchNorm -> LeakyReLU -> 3x3 conv,
twice; 1x1 projection on the skip when the channel count changes), 50 filters, one Linear head over
the NCHW flatten, softmax over the first A outputs and tanh on the last.  state_dict key names match the
shipped checkpoints so `load_state_dict` of models/*.pth works.  It is the numerics
reference for the product's bf16 evaluator and the network used by the CPU baseline.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

FILTERS = 50


class _PreActBlock(nn.Module):  # network.py:83-104
    def __init__(self, cin, cout):
        super().__init__()
        self.bn1, self.bn2 = nn.BatchNorm2d(cin), nn.BatchNorm2d(cout)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv3 = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        y = self.conv2(F.leaky_relu(self.bn2(self.conv1(F.leaky_relu(self.bn1(x))))))
        skip = self.conv3(x) if hasattr(self, "conv3") else x
        return skip + y


class RefNet(nn.Module):  # network.py:21-80
    def __init__(self, state_shape, num_distinct_actions, device=None):
        super().__init__()
        self.state_shape = list(state_shape)
        self.num_distinct_actions = num_distinct_actions
        self.device = device or torch.device("cpu")
        c, h, w = state_shape
        chans = [c + 1] + [FILTERS] * 5
        for k in range(5):
            setattr(self, "resblock%d" % (k + 1), _PreActBlock(chans[k], chans[k + 1]))
        self.fc1 = nn.Linear(FILTERS * h * w, num_distinct_actions + 1)

    def forward(self, x):
        for k in range(1, 6):
            x = getattr(self, "resblock%d" % k)(x)
        out = self.fc1(x.reshape(x.shape[0], -1))
        logits, v = out[:, :self.num_distinct_actions], out[:, self.num_distinct_actions:]
        return F.softmax(logits, dim=1), torch.tanh(v)

    def predict(self, state):  # network.py:66-80 -- a valid policy_fn
        from .ref_port import board_planes
        with torch.no_grad():
            x = torch.from_numpy(board_planes(state, self.state_shape)).float().to(self.device).unsqueeze(0)
            p, v = self.forward(x)
        return p.tolist()[0], float(v)
