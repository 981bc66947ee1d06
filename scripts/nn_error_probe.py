"""Observed numerics of the bf16 tcgen05 evaluator against fp32 references (run on the GPU box; prints one JSON line).
The bounds asserted in tests/test_gpu_nn.py are <= 2x the values this prints."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alphazero_openspiel_b200 import engine as E, _lib as L  # noqa: E402
from alphazero_openspiel_b200.network import Net  # noqa: E402
from alphazero_openspiel_b200.nn_fused import FusedEvaluator  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
out = {}


def pins(game, shape, A, ckpt, hists, p_ref, v_ref):
    net = Net(shape, A).eval()
    net.load_state_dict(torch.load(os.path.join(G, ckpt), map_location="cpu", weights_only=True))
    obs = E.game_replay(game, hists, L.OBS_BF16_NHWC)["obs"]
    p, v = FusedEvaluator(net, len(hists), "cuda:0").eval_batch(obs)
    p, v = p.cpu().numpy(), v.cpu().numpy()
    return {"max_dp": float(np.abs(p - p_ref).max()), "max_dv": float(np.abs(v - v_ref).max()),
            "argmax_agree": float((p.argmax(1) == p_ref.argmax(1)).mean())}


g = json.load(open(os.path.join(G, "reference_golden.json")))["encoding_pins"]["c4"]
out["c4_pins"] = pins("connect_four", [3, 6, 7], 7, "example_model_connect_four.pth", g["histories"],
                      np.array(g["p"]), np.array(g["v"]))
z = np.load(os.path.join(G, "bt6_pins.npz"))
out["bt6_pins"] = pins("breakthrough(rows=6,columns=6)", [3, 6, 6], 432, "example_model_breakthrough_6x6.pth",
                       [[int(a) for a in h if a >= 0] for h in z["histories"]], z["p"], z["v"])


def random_init(game):
    """The construction of tests/test_gpu_nn.py::test_fused_evaluator_matches_fp32_reference."""
    from oracle import ref_net
    shape, A = E.game_shape(game)
    torch.manual_seed(11)
    ref = ref_net.RefNet(shape, A).eval()
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.3, 0.3)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    net = Net(shape, A).eval()
    net.load_state_dict(ref.state_dict())
    B = 1000
    hist, lens = E.game_random_playouts(game, B, seed=5, max_plies=24)
    x32 = E.game_replay_dev(game, hist, lens, L.OBS_F32_NCHW)["obs"]
    xbf = E.game_replay_dev(game, hist, lens, L.OBS_BF16_NHWC)["obs"]
    with torch.no_grad():
        p_ref, v_ref = ref(x32.cpu())
    p, v = FusedEvaluator(net, B, "cuda:0").eval_batch(xbf)
    p, v = p.cpu(), v.cpu()
    return {"max_dp": float((p - p_ref).abs().max()), "max_dv": float((v - v_ref[:, 0]).abs().max()),
            "argmax_agree": float((p.argmax(1) == p_ref.argmax(1)).float().mean())}


for game in ["connect_four", "breakthrough(rows=6,columns=6)", "breakthrough"]:
    out["random_init/" + game] = random_init(game)
print(json.dumps(out))
