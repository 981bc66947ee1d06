"""Times az_nn_head_large alone (CUDA events); AZ_NN_HEAD_DEBUG selects experiment modes (1 = no softmax pass, 2 = no
epilogue, 4 = no MMAs).  Usage: head_microbench.py [boards] [H] [W]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W = int(sys.argv[3]) if len(sys.argv) > 3 else 8
A = H * W * 12
x = (torch.randn((B, (H + 1) * W * 64), device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn((A + 1, (H + 1) * W * 64), device=dev) * 0.03).to(torch.bfloat16)
b = torch.randn(A + 1, device=dev)
pri = torch.zeros((B, A), device=dev); val = torch.zeros((B,), device=dev)
n = int(lib.az_nn_head_large_scratch_bytes(B, A))
scratch = torch.zeros((n + 3) // 4, dtype=torch.int32, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    rc = lib.az_nn_head_large(p(x), p(w), p(b), p(pri), p(val), p(scratch), B, H, W, A, st)
    assert rc == 0, lib.az_nn_last_error()
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
fl = 2.0 * B * H * W * 64 * (A + 1)
print("AZ_NN_HEAD_DEBUG=%s boards %d %dx%d A=%d: %.1f us  (%.0f TFLOP/s issued)" % (os.environ.get("AZ_NN_HEAD_DEBUG", "0"), B, H, W, A, us, fl / us / 1e6))
