"""Small exercise of every kernel added in round 2, for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
k_head_mma (both modes), k_block, k_conv8 (packed N = 160), k_step_vl, the rollout evaluator, k_observations, MCTS.playout."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import engine as E, _lib as L
from alphazero_openspiel_b200.network import Net
from alphazero_openspiel_b200.nn_fused import FusedEvaluator
from alphazero_openspiel_b200.examplegenerator import SelfPlayRunner
dev = "cuda:0"
torch.manual_seed(0)
for game, B in [("connect_four", 37), ("breakthrough(rows=6,columns=6)", 130), ("breakthrough", 33)]:
    shape, A = E.game_shape(game)
    net = Net(shape, A).eval()
    obs = (torch.rand((B, shape[1], shape[2], 4), device=dev) > 0.5).to(torch.bfloat16)
    for blk in ("0", "1"):
        os.environ["AZ_NN_BLOCK"] = blk
        p, v = FusedEvaluator(net, B, dev).eval_batch(obs)
        assert torch.isfinite(p).all() and torch.isfinite(v).all()
    os.environ["AZ_NN_BLOCK"] = "0"
    if A > 7:
        os.environ["AZ_NN_HEAD_MODE"] = "0"
        p0, _ = FusedEvaluator(net, B, dev).eval_batch(obs)
        os.environ["AZ_NN_HEAD_MODE"] = "1"
        assert (p0 - p).abs().max().item() < 1e-5
for vl in (0, 4):
    r = SelfPlayRunner(Net([3, 6, 6], 432).eval(), "breakthrough(rows=6,columns=6)", dev, 24, n_playouts=12, seed=1,
                       max_sims_per_step=8, use_graph=False, virtual_loss=vl)
    r.round(40)
    c = r.counters(); r.close()
    assert c["sims"] > 0 and c["overflow"] == 0
eng = E.Engine("connect_four", 16, n_playouts=20, noise_mode=L.NOISE_COUNTER, eval_mode=L.EVAL_ROLLOUT,
               flags=L.F_KEEP_TREE | L.F_SAMPLE_MOVES | L.F_RECORDS, seed=3, c_puct=1.0)
for _ in range(200): eng.step()
assert eng.counters()["overflow"] == 0
recs = eng.drain_records(); eng.close()
from alphazero_openspiel_b200.device_replay import DeviceReplay
from alphazero_openspiel_b200.replay import ExampleBatch
b = ExampleBatch.from_records(recs, "connect_four")
if len(b):
    d = DeviceReplay("connect_four", dev); d.append(b); f, pp, vv = d.remove_duplicates()
    x = d.boards(f[:min(8, len(f))]); assert x.shape[1:] == (4, 6, 7)
torch.cuda.synchronize()
print("sanitize target ok")
