"""Times az_nn_stem alone (CUDA events) for 16384 Connect Four boards; AZ_NN_DEBUG: 2 = no epilogue math/stores, 4 = no MMAs."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
B, H, W = 16384, 6, 7
obs = (torch.rand((B, H, W, 4), device=dev) > 0.5).to(torch.bfloat16)
u = torch.zeros((B, H + 1, W, 64), dtype=torch.bfloat16, device=dev)
w = (torch.randn((9, 2, 64, 8)) * 0.05).to(torch.bfloat16).to(dev)
b = torch.randn(64, device=dev); st8 = torch.rand(8, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    rc = lib.az_nn_stem(p(obs), p(w), p(b), p(st8), p(u), B, H, W, 0, st)
    assert rc == 0, lib.az_nn_last_error()
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("AZ_NN_DEBUG=%s stem %.1f us" % (os.environ.get("AZ_NN_DEBUG", "0"), e0.elapsed_time(e1) / 20 * 1e3))
