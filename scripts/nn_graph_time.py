"""Time one FusedEvaluator forward (16,384 Connect Four boards) eagerly and as a replayed CUDA graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200.network import Net
from alphazero_openspiel_b200.nn_fused import FusedEvaluator
from alphazero_openspiel_b200 import engine as E
dev = torch.device("cuda:0")
shape, A = E.game_shape("connect_four")
net = Net(shape, A).eval()
B = int(os.environ.get("B", 16384))
fe = FusedEvaluator(net, B, dev)
fe.obs.copy_((torch.rand((B, shape[1], shape[2], 4), device=dev) > 0.5).to(torch.bfloat16))
def timeit(fn, n=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("eager  %.1f us" % timeit(fe))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fe(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fe()
    print("graph  %.1f us" % timeit(g.replay))
