"""Policy-value network of the self-play path: drop-in for the reference's network.py.

  state_to_board(state, state_shape)            network.py:9-18
  Net(state_shape, num_distinct_actions, ...)   network.py:21-80   (.forward, .predict; checkpoint-compatible keys)
  BatchedEvaluator(net, ...)                    replaces Evaluator.evaluate_nn + handle_gpu
                                                (examplegenerator.py:39-77): eval_batch(obs) -> (priors, values)

The ResNet is the only dense contraction on the path and, per the north star, runs bf16 on the tensor cores
through PyTorch (cuDNN/cuBLAS).  BatchedEvaluator prepares the weights for that: eval-mode BatchNorm folded
into the preceding conv where that is exact (bn2 into conv1), the other BatchNorms applied as per-channel
affines, 50 filters zero-padded to 64, channels-last bf16, FC weights permuted to the channels-last
flatten (network.py:60 flattens NCHW), and the whole forward captured in one CUDA graph at a fixed batch.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

N_FILTERS = 50


def state_to_board(state, state_shape):
    """(C+1, H, W) float64: OpenSpiel planes + a current-player plane (network.py:9-18)."""
    c, h, w = state_shape
    player = state.current_player()
    board = np.empty((c + 1, h, w), dtype=np.float64)
    board[:c] = np.asarray(state.information_state_as_normalized_vector()).reshape(c, h, w)
    board[c] = player
    return board


class ResidualBlock(nn.Module):
    """Pre-activation residual block (network.py:83-104); attribute names fixed by the shipped checkpoints."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(in_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.use_1x1conv = in_channels != out_channels
        if self.use_1x1conv:
            self.conv3 = nn.Conv2d(in_channels, out_channels, 1, padding=0)

    def forward(self, x):
        branch = self.conv1(F.leaky_relu(self.bn1(x)))
        branch = self.conv2(F.leaky_relu(self.bn2(branch)))
        return (self.conv3(x) if self.use_1x1conv else x) + branch


class Net(nn.Module):
    """Same constructor, state_dict keys (resblock{1..5}.*, fc1.*) and outputs as the reference Net."""

    def __init__(self, state_shape, num_distinct_actions, **kwargs):
        super().__init__()
        self.state_shape = list(state_shape)
        self.num_filters_input = state_shape[0] + 1
        self.height, self.width = state_shape[1], state_shape[2]
        self.num_distinct_actions = num_distinct_actions
        self.device = kwargs.get("device", torch.device("cpu"))
        self.n_filts = N_FILTERS
        widths = [self.num_filters_input] + [N_FILTERS] * 5
        for k in range(5):
            setattr(self, "resblock%d" % (k + 1), ResidualBlock(widths[k], widths[k + 1]))
        self.fc1 = nn.Linear(N_FILTERS * self.width * self.height, num_distinct_actions + 1)

    def blocks(self):
        return [getattr(self, "resblock%d" % k) for k in range(1, 6)]

    def forward(self, x):
        for blk in self.blocks():
            x = blk(x)
        out = self.fc1(x.reshape(-1, self.height * self.width * self.n_filts))
        logits, v = out.split(self.num_distinct_actions, 1)
        return F.softmax(logits, dim=1), torch.tanh(v)

    def predict(self, state):
        """policy_fn for a single state (network.py:66-80): (list[A] of floats, float)."""
        with torch.no_grad():
            x = torch.from_numpy(state_to_board(state, self.state_shape)).float().to(self.device).unsqueeze(0)
            p, v = self.forward(x)
        return p.tolist()[0], float(v)


def _fold_bn(bn):
    """eval-mode BatchNorm as y = a*x + b (fp64 for the folding arithmetic)."""
    a = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
    return a, bn.bias.double() - bn.running_mean.double() * a


class BatchedEvaluator:
    """eval_batch(obs bf16 [B,H,W,4]) -> (priors fp32 [B,A], values fp32 [B]) at a fixed batch, CUDA-graph captured.

    Replaces the reference's per-state pipe round trip (examplegenerator.py:44-77).  Outputs feed az_step directly.
    """

    CPAD = 64

    def __init__(self, net, batch, device, use_graph=True, dtype=torch.bfloat16):
        self.batch, self.device, self.dtype = batch, torch.device(device), dtype
        self.h, self.w, self.A = net.height, net.width, net.num_distinct_actions
        self.use_graph = use_graph
        self._graph = None
        self.load(net)
        B, H, W = batch, self.h, self.w
        self.obs = torch.zeros((B, H, W, 4), dtype=dtype, device=self.device)
        self.priors = torch.zeros((B, self.A), dtype=torch.float32, device=self.device)
        self.values = torch.zeros((B,), dtype=torch.float32, device=self.device)

    @torch.no_grad()
    def load(self, net):
        """(Re)load weights from a (CPU or GPU) Net: fold, pad, permute, cast.  In-place when shapes match so a
        captured graph stays valid (weight broadcast between generations)."""
        dev, dt, CP = self.device, self.dtype, self.CPAD
        new = {}
        for k, blk in enumerate([getattr(net, "resblock%d" % i) for i in range(1, 6)]):
            cin = blk.conv1.in_channels
            cin_p = 4 if cin == 4 else CP
            a1, b1 = _fold_bn(blk.bn1)
            a2, b2 = _fold_bn(blk.bn2)
            s1 = torch.zeros(cin_p, dtype=torch.float64)
            t1 = torch.zeros(cin_p, dtype=torch.float64)
            s1[:cin], t1[:cin] = a1.cpu(), b1.cpu()
            w1 = blk.conv1.weight.double().cpu() * a2.cpu().view(-1, 1, 1, 1)      # bn2 folded into conv1
            c1b = blk.conv1.bias.double().cpu() * a2.cpu() + b2.cpu()
            w1p = torch.zeros((CP, cin_p, 3, 3), dtype=torch.float64)
            w1p[:N_FILTERS, :cin] = w1
            b1p = torch.zeros(CP, dtype=torch.float64)
            b1p[:N_FILTERS] = c1b
            w2p = torch.zeros((CP, CP, 3, 3), dtype=torch.float64)
            w2p[:N_FILTERS, :N_FILTERS] = blk.conv2.weight.double().cpu()
            b2p = torch.zeros(CP, dtype=torch.float64)
            b2p[:N_FILTERS] = blk.conv2.bias.double().cpu()
            new["s1_%d" % k] = s1.view(1, -1, 1, 1)
            new["t1_%d" % k] = t1.view(1, -1, 1, 1)
            new["w1_%d" % k], new["b1_%d" % k] = w1p, b1p
            new["w2_%d" % k], new["b2_%d" % k] = w2p, b2p
            if blk.use_1x1conv:
                w3p = torch.zeros((CP, cin_p, 1, 1), dtype=torch.float64)
                w3p[:N_FILTERS, :cin] = blk.conv3.weight.double().cpu()
                b3p = torch.zeros(CP, dtype=torch.float64)
                b3p[:N_FILTERS] = blk.conv3.bias.double().cpu()
                new["w3_%d" % k], new["b3_%d" % k] = w3p, b3p
        # FC: reference flattens NCHW (c*H*W + h*W + w); activations here are NHWC with CP channels
        fw = net.fc1.weight.double().cpu().view(self.A + 1, N_FILTERS, self.h, self.w)
        fwp = torch.zeros((self.A + 1, self.h, self.w, CP), dtype=torch.float64)
        fwp[..., :N_FILTERS] = fw.permute(0, 2, 3, 1)
        new["fw"] = fwp.reshape(self.A + 1, -1)
        new["fb"] = net.fc1.bias.double().cpu()
        if not hasattr(self, "par"):
            self.par = {}
            for name, t in new.items():
                t = t.to(dev, dt)
                if t.dim() == 4 and t.shape[2] in (1, 3) and name[0] == "w":
                    t = t.contiguous(memory_format=torch.channels_last)
                self.par[name] = t
        else:
            for name, t in new.items():
                self.par[name].copy_(t.to(dev, dt))

    def _forward(self, obs):
        p = self.par
        x = obs.permute(0, 3, 1, 2)  # logical NCHW view of channels-last memory
        for k in range(5):
            y = F.leaky_relu(x * p["s1_%d" % k] + p["t1_%d" % k])
            y = F.leaky_relu(F.conv2d(y, p["w1_%d" % k], p["b1_%d" % k], padding=1))
            y = F.conv2d(y, p["w2_%d" % k], p["b2_%d" % k], padding=1)
            if ("w3_%d" % k) in p:
                x = F.conv2d(x, p["w3_%d" % k], p["b3_%d" % k])
            x = x + y
        flat = x.permute(0, 2, 3, 1).reshape(self.batch, -1)
        out = F.linear(flat, p["fw"], p["fb"]).float()
        return F.softmax(out[:, :self.A], dim=1), torch.tanh(out[:, self.A])

    @torch.no_grad()
    def _run_into(self):
        pr, v = self._forward(self.obs)
        self.priors.copy_(pr)
        self.values.copy_(v)

    @torch.no_grad()
    def __call__(self):
        """Evaluate self.obs into self.priors / self.values on the current stream."""
        if not self.use_graph:
            self._run_into()
            return self.priors, self.values
        if self._graph is None:
            s = torch.cuda.Stream(self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(3):
                    self._run_into()
            torch.cuda.current_stream(self.device).wait_stream(s)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._run_into()
        self._graph.replay()
        return self.priors, self.values

    @torch.no_grad()
    def eval_batch(self, obs):
        """Generic entry: obs [B,H,W,4] (any float dtype, B == batch) -> (priors, values) clones."""
        self.obs.copy_(obs.to(self.dtype))
        pr, v = self()
        return pr.clone(), v.clone()
