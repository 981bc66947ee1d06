"""Target for `ncu --set full`: two launches of each az_nn_conv3x3 kind (plain, +res, +res+out2) on 16,384 Connect Four
boards (the shapes of BASELINE configs[2])."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import _lib as L
from alphazero_openspiel_b200.nn_fused import pack_conv3x3
lib = L.load()
dev = torch.device("cuda:0")
B, H, W = 16384, 6, 7
x = torch.randn((B, H + 1, W, 64), device=dev).to(torch.bfloat16)
r = torch.randn((B, H + 1, W, 64), device=dev).to(torch.bfloat16)
o = torch.zeros((B, H + 1, W, 64), dtype=torch.bfloat16, device=dev)
o2 = torch.zeros((B, H + 1, W, 64), dtype=torch.bfloat16, device=dev)
w = pack_conv3x3(torch.randn(64, 64, 3, 3) * 0.05).to(dev)
b = torch.randn(64, device=dev); s2 = torch.rand(64, device=dev); t2 = torch.randn(64, device=dev)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, res, out2 in [("conv1-type", None, None), ("conv2+res", r, None), ("conv2+res+out2", r, o2)]:
    for _ in range(2):
        rc = lib.az_nn_conv3x3(p(x), p(w), p(b), p(res), p(o), p(out2), p(s2) if out2 is not None else None,
                               p(t2) if out2 is not None else None, B, H, W, 1 if res is None else 0, 0, 0, st)
        assert rc == 0, lib.az_nn_last_error()
torch.cuda.synchronize()
print("ok")
