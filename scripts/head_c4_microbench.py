"""k_head (Connect Four FC head) time vs grid size."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
B, H, W, A = 16384, 6, 7, 7
x = torch.randn((B, H + 1, W, 64), device=dev).to(torch.bfloat16)
w = (torch.randn((8, H * W * 64), device=dev) * 0.03).to(torch.bfloat16)
b = torch.randn(8, device=dev); pri = torch.zeros((B, A), device=dev); val = torch.zeros((B,), device=dev)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
p = lambda t: C.c_void_p(t.data_ptr()); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for n_ctas in (0, 296, 222, 148, 111, 74):
    def run():
        assert lib.az_nn_head(p(x), p(w), p(b), p(pri), p(val), B, H, W, A, n_ctas, st) == 0
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        big.fill_(1)                      # flush L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("n_ctas %3d: median %.1f us (cold L2)  -> %.2f TB/s" % (n_ctas, ts[5], 88.1e6 / ts[5] / 1e6))
