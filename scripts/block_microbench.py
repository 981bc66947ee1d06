"""Times az_nn_block (fused residual block) against the two az_nn_conv3x3 launches it replaces, 16384 Connect Four boards."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import _lib as L
from alphazero_openspiel_b200.nn_fused import pack_conv3x3
lib = L.load()
dev = torch.device("cuda:0")
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 6, 7
t = torch.randn((B, H + 1, W, 64), device=dev).to(torch.bfloat16)
x = torch.randn((B, H + 1, W, 64), device=dev).to(torch.bfloat16)
u = torch.zeros((B, H + 1, W, 64), dtype=torch.bfloat16, device=dev)
to = torch.zeros((B, H + 1, W, 64), dtype=torch.bfloat16, device=dev)
w1 = pack_conv3x3(torch.randn(64, 64, 3, 3) * 0.05).to(dev); w2 = pack_conv3x3(torch.randn(64, 64, 3, 3) * 0.05).to(dev)
b1 = torch.randn(64, device=dev); b2 = torch.randn(64, device=dev); s2 = torch.rand(64, device=dev); t2 = torch.randn(64, device=dev)
p = lambda a: None if a is None else C.c_void_p(a.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def fused(out2):
    rc = lib.az_nn_block(p(t), p(w1), p(b1), p(w2), p(b2), p(x), p(to) if out2 else None, p(s2) if out2 else None,
                         p(t2) if out2 else None, B, H, W, 0, st)
    assert rc == 0, lib.az_nn_last_error()
def two(out2):
    rc = lib.az_nn_conv3x3(p(t), p(w1), p(b1), None, p(u), None, None, None, B, H, W, 1, 0, 0, st)
    rc |= lib.az_nn_conv3x3(p(u), p(w2), p(b2), p(x), p(x), p(to) if out2 else None, p(s2) if out2 else None,
                            p(t2) if out2 else None, B, H, W, 0, 1, 0, st)
    assert rc == 0, lib.az_nn_last_error()
for name, fn in [("fused block + out2", lambda: fused(True)), ("conv + conv+res+out2", lambda: two(True)),
                 ("fused block (last)", lambda: fused(False)), ("conv + conv+res", lambda: two(False))]:
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-24s %.1f us" % (name, e0.elapsed_time(e1) / 20 * 1e3))
