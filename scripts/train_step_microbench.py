"""Where does the optimisation step of the 5x50 ResNet at batch 256 go?  Times fwd+bwd+Adam under a CUDA graph for a few
PyTorch settings (fp32 default, cudnn.benchmark, channels_last, TF32 matmul, bf16 autocast)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200.network import Net
dev = torch.device("cuda:0")
def run(name, channels_last=False, benchmark=False, autocast=False, tf32=False):
    torch.backends.cudnn.benchmark = benchmark
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.manual_seed(0)
    net = Net([3, 6, 7], 7, device=dev).to(dev).train()
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
    x = torch.rand(256, 4, 6, 7, device=dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    pr = torch.softmax(torch.rand(256, 7, device=dev), 1); vr = torch.rand(256, 1, device=dev)
    mse = torch.nn.MSELoss()
    def step():
        opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            p, v = net(x)
        p, v = p.float(), v.float()
        loss = mse(v, vr) - torch.sum(pr * torch.log(p)) / 256
        loss.backward()
        opt.step()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step()
    for _ in range(5): g.replay()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): g.replay()
    torch.cuda.synchronize()
    print("%-40s %.3f ms/step" % (name, (time.perf_counter() - t0) / 200 * 1e3))
run("fp32 default")
run("cudnn.benchmark", benchmark=True)
run("channels_last", channels_last=True)
run("channels_last + benchmark", channels_last=True, benchmark=True)
run("tf32 matmul + benchmark", benchmark=True, tf32=True)
run("bf16 autocast + channels_last + benchmark", channels_last=True, benchmark=True, autocast=True)
