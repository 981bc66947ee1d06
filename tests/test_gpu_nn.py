"""GPU numerics of the hand-written evaluator kernels (csrc/az_resnet.cu) against plain PyTorch fp32 references."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _nhwc(x):
    """[B,64,H,W] float -> bf16 [B,H+1,W,64], row H zero (the activation layout of az_resnet.cu)."""
    import torch
    B, Cc, H, W = x.shape
    out = torch.zeros((B, H + 1, W, Cc), dtype=torch.bfloat16, device=x.device)
    out[:, :H] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


@pytest.mark.parametrize("H,W,B", [(6, 7, 37), (6, 6, 300), (8, 8, 129), (6, 7, 4096), (5, 7, 1000), (3, 2, 50), (8, 8, 3000),
                                   (6, 7, 1), (16, 8, 5)])
def test_conv3x3_tcgen05_matches_torch(H, W, B):
    """out = conv3x3(in) + bias [-> LeakyReLU] [+ res], out2 = LeakyReLU(s2*out+t2) on [B,H+1,W,64] tensors; W = 8 has no
    pad column inside the SM (row wrap-around is masked in the epilogue), W < 8 has.
    Tolerance: inputs/weights are bf16-exact on both sides, accumulation fp32 -> only the final bf16 rounding differs
    (rel 2^-8) plus fp32 summation-order noise."""
    import torch
    import torch.nn.functional as F
    from alphazero_openspiel_b200 import _lib as L
    from alphazero_openspiel_b200.nn_fused import pack_conv3x3
    lib = L.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(H * 100 + W + B)
    x = torch.randn((B, 64, H, W), generator=g).to(dev).to(torch.bfloat16).float()
    w = (torch.randn((64, 64, 3, 3), generator=g) * 0.05).to(dev).to(torch.bfloat16).float()
    bias = torch.randn((64,), generator=g).to(dev)
    w[50:] = 0.0       # the conv computes the network's 50 filters from all 64 input channels;
    bias[50:] = 0.0    # output channels 50..63 are written as zeros (whatever the residual holds there)
    res = torch.randn((B, 64, H, W), generator=g).to(dev).to(torch.bfloat16).float()
    s2 = (torch.rand((64,), generator=g) + 0.5).to(dev)
    t2 = torch.randn((64,), generator=g).to(dev)
    xin, rin = _nhwc(x), _nhwc(res)
    wp = pack_conv3x3(w).to(dev)
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = F.conv2d(x, w, bias, padding=1)
    for lrelu, use_res, use_out2, inplace in [(0, 0, 0, 0), (1, 0, 0, 0), (0, 1, 1, 0), (0, 1, 0, 0), (0, 1, 1, 1)]:
        out = rin.clone() if inplace else torch.full((B, H + 1, W, 64), 7.0, dtype=torch.bfloat16, device=dev)
        out2 = torch.full((B, H + 1, W, 64), 7.0, dtype=torch.bfloat16, device=dev) if use_out2 else None
        rc = lib.az_nn_conv3x3(ptr(xin), ptr(wp), ptr(bias), ptr(out if inplace else rin) if use_res else None, ptr(out),
                               ptr(out2), ptr(s2) if use_out2 else None, ptr(t2) if use_out2 else None, B, H, W,
                               lrelu, use_res, 0, st)   # flags: the residual variants also run back to front
        assert rc == 0, lib.az_nn_last_error()
        torch.cuda.synchronize()
        want = ref
        if lrelu:
            want = F.leaky_relu(want)
        if use_res:
            want = want + res
        want = want.clone()
        want[:, 50:] = 0.0
        got = out[:, :H].float().permute(0, 3, 1, 2)
        assert float(out[:, H].float().abs().max()) == 0.0   # the pad row is (re)written as zeros
        err = (got - want).abs()
        tol = 2.0 ** -7 * want.abs() + 2e-2
        assert bool((err <= tol).all()), (lrelu, use_res, float(err.max()))
        if use_out2:
            want2 = F.leaky_relu(want * s2.view(1, -1, 1, 1) + t2.view(1, -1, 1, 1))
            want2[:, 50:] = 0.0
            err2 = (out2[:, :H].float().permute(0, 3, 1, 2) - want2).abs()
            assert float(out2[:, H].float().abs().max()) == 0.0
            assert bool((err2 <= 2.0 ** -6 * want2.abs() + 4e-2).all()), float(err2.max())


@pytest.mark.parametrize("game", ["connect_four", "breakthrough(rows=6,columns=6)", "breakthrough"])
def test_fused_evaluator_matches_fp32_reference(game):
    """Whole network on the tcgen05 path vs the oracle's fp32 RefNet (bf16 tolerance, stated) and vs the PyTorch
    bf16 evaluator (same precision class)."""
    import torch
    from oracle import ref_net
    from alphazero_openspiel_b200 import engine as E, _lib as L
    from alphazero_openspiel_b200.network import Net, BatchedEvaluator
    from alphazero_openspiel_b200.nn_fused import FusedEvaluator
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
    fe = FusedEvaluator(net, B, "cuda:0")
    p, v = fe.eval_batch(xbf)
    p, v = p.cpu(), v.cpu()
    assert torch.isfinite(p).all() and torch.isfinite(v).all()
    assert (p.sum(1) - 1).abs().max().item() < 1e-3
    # bf16 bounds = 2x the maxima observed on B200 (scripts/nn_error_probe.py, r02: |dp| 8.5e-4 / 2.9e-5 / 1.8e-5,
    # |dv| 3.6e-3 / 5.9e-3 / 4.2e-3 for Connect Four / Breakthrough 6x6 / 8x8; priors of the large action spaces are ~1/A)
    tol_p, tol_v = {"connect_four": (1.7e-3, 7.2e-3), "breakthrough(rows=6,columns=6)": (6e-5, 1.2e-2),
                    "breakthrough": (3.6e-5, 8.5e-3)}[game]
    assert (p - p_ref).abs().max().item() < tol_p
    assert (v - v_ref[:, 0]).abs().max().item() < tol_v
    assert (p.argmax(1) == p_ref.argmax(1)).float().mean().item() > 0.97
    pe, ve = BatchedEvaluator(net, B, "cuda:0").eval_batch(xbf)
    assert (p - pe.cpu()).abs().max().item() < 2 * tol_p and (v - ve.cpu()).abs().max().item() < 2 * tol_v
    # weights can be reloaded in place (new generation) and a second call is deterministic
    p2, v2 = fe.eval_batch(xbf)
    assert torch.equal(p2.cpu(), p) and torch.equal(v2.cpu(), v)
    # a differently sized batch gives the same per-board results (boards never see each other)
    p3, v3 = FusedEvaluator(net, 333, "cuda:0").eval_batch(xbf[:333])
    # (both FC heads accumulate a board's logits in a fixed order that does not depend on the batch; the softmax of the
    # large head combines per-chunk statistics whose chunk width depends on the batch size -> last-bit differences there)
    assert fe.fused_head
    assert torch.equal(v3.cpu(), v[:333])
    if fe.small_head:
        assert torch.equal(p3.cpu(), p[:333])
    else:
        assert (p3.cpu() - p[:333]).abs().max().item() < 1e-6


def test_fused_evaluator_is_batch_size_independent():
    """A board's priors / value do not depend on how many other boards share the launch (1, 2, 3 boards vs 64): tiles that
    are mostly out of range, single-CTA grids and the ragged tail of k_head."""
    import torch
    from alphazero_openspiel_b200 import engine as E
    from alphazero_openspiel_b200.network import Net
    from alphazero_openspiel_b200.nn_fused import FusedEvaluator
    shape, A = E.game_shape("connect_four")
    torch.manual_seed(3)
    net = Net(shape, A).eval()
    obs = (torch.rand((64, shape[1], shape[2], 4), device="cuda:0") > 0.5).to(torch.bfloat16)
    p64, v64 = FusedEvaluator(net, 64, "cuda:0").eval_batch(obs)
    for b in (1, 2, 3, 17):
        p, v = FusedEvaluator(net, b, "cuda:0").eval_batch(obs[:b])
        assert torch.equal(p, p64[:b]) and torch.equal(v, v64[:b]), b


# Observed on B200 (scripts/nn_error_probe.py) -- see the constants in the test: bounds are 2x the observed maxima.
def _golden_pins(which):
    import json
    import os
    import torch
    from alphazero_openspiel_b200.network import Net
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if which == "c4":
        g = json.load(open(os.path.join(G, "reference_golden.json")))["encoding_pins"]["c4"]
        game, shape, A, ck = "connect_four", [3, 6, 7], 7, "example_model_connect_four.pth"
        hists, p_ref, v_ref = g["histories"], np.array(g["p"]), np.array(g["v"])
    else:
        z = np.load(os.path.join(G, "bt6_pins.npz"))
        game, shape, A, ck = "breakthrough(rows=6,columns=6)", [3, 6, 6], 432, "example_model_breakthrough_6x6.pth"
        hists = [[int(a) for a in h if a >= 0] for h in z["histories"]]
        p_ref, v_ref = z["p"].astype(np.float64), z["v"].astype(np.float64)
    net = Net(shape, A).eval()
    net.load_state_dict(torch.load(os.path.join(G, ck), map_location="cpu", weights_only=True))
    return game, net, hists, p_ref, v_ref


@pytest.mark.parametrize("which,tol_p,tol_v", [("c4", 2.3e-2, 6e-2), ("bt6", 2.4e-2, 4.6e-2)])
def test_fused_evaluator_matches_reference_net_outputs_of_shipped_checkpoints(which, tol_p, tol_v):
    """The tcgen05 evaluator with the reference's SHIPPED checkpoints on the golden positions vs the priors / values the
    unmodified reference `Net` (network.py:48-64, fp32) produced for them (tests/golden/make_golden*.py): the bf16 bound
    is stated here and is <= 2x the maximum observed on B200 (scripts/nn_error_probe.py, r02: Connect Four |dp| 0.0113,
    |dv| 0.0297; Breakthrough 6x6 |dp| 0.0121, |dv| 0.0229; the trained networks are much sharper than random ones)."""
    from alphazero_openspiel_b200 import engine as E, _lib as L
    from alphazero_openspiel_b200.nn_fused import FusedEvaluator
    game, net, hists, p_ref, v_ref = _golden_pins(which)
    obs = E.game_replay(game, hists, L.OBS_BF16_NHWC)["obs"]
    fe = FusedEvaluator(net, len(hists), "cuda:0")
    assert fe.fused_head, "the FC head must run on the hand-written kernels for every supported game"
    p, v = fe.eval_batch(obs)
    p, v = p.cpu().numpy().astype(np.float64), v.cpu().numpy().astype(np.float64)
    assert np.isfinite(p).all() and np.isfinite(v).all()
    assert np.abs(p.sum(1) - 1).max() < 1e-3
    dp, dv = np.abs(p - p_ref).max(), np.abs(v - v_ref).max()
    assert dp < tol_p and dv < tol_v, (dp, dv)
    assert (p.argmax(1) == p_ref.argmax(1)).mean() >= 0.9


@pytest.mark.parametrize("H,W,A,B", [(6, 6, 432, 1000), (8, 8, 768, 300), (6, 6, 432, 64), (8, 8, 768, 4096), (6, 7, 7, 50),
                                     (5, 5, 300, 129), (8, 8, 768, 1), (6, 6, 432, 16384)])
def test_large_action_space_head_matches_torch(H, W, A, B):
    """az_nn_head_large (tcgen05 GEMM + softmax + tanh, fc1 of network.py:48,60-64): priors / values vs torch fp32 on the
    same bf16-exact inputs.  fp32 accumulation and fp32 logits on both sides -> only summation order differs: the stated
    bound is 1e-5 absolute + 1e-4 relative on the priors and 2e-5 on the values."""
    import torch
    from alphazero_openspiel_b200 import _lib as L
    lib = L.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(H * 1000 + A + B)
    K, KR = H * W * 64, (H + 1) * W * 64
    x = torch.zeros((B, KR), dtype=torch.bfloat16)
    x[:, :K] = (torch.randn((B, K), generator=g) * 0.5).to(torch.bfloat16)
    x[:, K:] = 3.0          # the pad row is never read by the head (it is zero in the real tensors)
    w = torch.zeros((A + 1, KR), dtype=torch.bfloat16)
    w[:, :K] = (torch.randn((A + 1, K), generator=g) * 0.03).to(torch.bfloat16)
    w[:, K:] = 5.0
    bias = torch.randn((A + 1,), generator=g)
    logits = x[:, :K].double() @ w[:, :K].double().t() + bias.double()
    p_ref = torch.softmax(logits[:, :A], dim=1)
    v_ref = torch.tanh(logits[:, A])
    xd, wd, bd = x.to(dev), w.to(dev), bias.to(dev)
    pri = torch.full((B, A), -1.0, dtype=torch.float32, device=dev)
    val = torch.full((B,), -9.0, dtype=torch.float32, device=dev)
    n = int(lib.az_nn_head_large_scratch_bytes(B, A))
    scratch = torch.zeros((n + 3) // 4, dtype=torch.int32, device=dev)
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    for rep in range(2):   # the second call checks that the tile counters reset themselves
        pri.fill_(-1.0)
        rc = lib.az_nn_head_large(ptr(xd), ptr(wd), ptr(bd), ptr(pri), ptr(val), ptr(scratch), B, H, W, A,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.az_nn_last_error()
        torch.cuda.synchronize()
        p, v = pri.cpu().double(), val.cpu().double()
        assert torch.isfinite(p).all() and (p >= 0).all()
        assert (p.sum(1) - 1).abs().max().item() < 1e-5
        err = (p - p_ref).abs()
        assert bool((err <= 1e-5 + 1e-4 * p_ref).all()), float(err.max())
        assert (v - v_ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("H,W,B", [(6, 7, 37), (6, 6, 300), (8, 8, 129), (6, 7, 4096), (5, 7, 1000), (3, 2, 50), (8, 8, 3000),
                                   (6, 7, 1), (16, 8, 5), (6, 7, 16), (6, 7, 17), (6, 7, 2500)])
def test_fused_residual_block_matches_torch(H, W, B):
    """az_nn_block (k_block): U = LeakyReLU(conv1(T) + b1) kept in shared memory, X <- conv2(U) + b2 + X in place,
    T' = LeakyReLU(s2*X + t2) (network.py:99-104 with the BatchNorms folded / carried as in FusedEvaluator) against torch fp32
    with U rounded to bf16 like the kernel does.  Tolerance: bf16-exact inputs and weights, fp32 accumulation -> the final
    bf16 rounding (rel 2^-8) plus the effect of single-ulp differences in U (stated bound: 2^-6 relative + 3e-2 absolute).
    Covers super-tile edges (16 boards), partial super-tiles, W = 8 (no pad column), one board, one CTA."""
    import torch
    import torch.nn.functional as F
    from alphazero_openspiel_b200 import _lib as L
    from alphazero_openspiel_b200.nn_fused import pack_conv3x3
    lib = L.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(7 * H + W + B)
    t = torch.randn((B, 64, H, W), generator=g).to(dev).to(torch.bfloat16).float()
    x = torch.randn((B, 64, H, W), generator=g).to(dev).to(torch.bfloat16).float()
    x[:, 50:] = 0.0                                   # layout invariant: the pad channels of the residual stream are zero
    w1 = (torch.randn((64, 64, 3, 3), generator=g) * 0.05).to(dev).to(torch.bfloat16).float()
    w2 = (torch.randn((64, 64, 3, 3), generator=g) * 0.05).to(dev).to(torch.bfloat16).float()
    w1[50:] = 0.0
    w2[50:] = 0.0
    w2[:, 50:] = 0.0                                  # conv2 sees the 50 channels of U
    b1 = torch.randn((64,), generator=g).to(dev)
    b2 = torch.randn((64,), generator=g).to(dev)
    b1[50:] = 0.0
    b2[50:] = 0.0
    s2 = (torch.rand((64,), generator=g) + 0.5).to(dev)
    t2 = torch.randn((64,), generator=g).to(dev)
    u = F.leaky_relu(F.conv2d(t, w1, b1, padding=1))
    u[:, 50:] = 0.0
    u = u.to(torch.bfloat16).float()
    want_x = F.conv2d(u, w2, b2, padding=1) + x
    want_x[:, 50:] = 0.0
    want_t = F.leaky_relu(want_x * s2.view(1, -1, 1, 1) + t2.view(1, -1, 1, 1))
    want_t[:, 50:] = 0.0
    ptr = lambda a: None if a is None else C.c_void_p(a.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p1, p2 = pack_conv3x3(w1).to(dev), pack_conv3x3(w2).to(dev)
    for with_out2 in (True, False):
        tin, xio = _nhwc(t), _nhwc(x)
        tout = torch.zeros((B, H + 1, W, 64), dtype=torch.bfloat16, device=dev) if with_out2 else None
        rc = lib.az_nn_block(ptr(tin), ptr(p1), ptr(b1), ptr(p2), ptr(b2), ptr(xio), ptr(tout), ptr(s2) if with_out2 else None,
                             ptr(t2) if with_out2 else None, B, H, W, 0, st)
        assert rc == 0, lib.az_nn_last_error()
        torch.cuda.synchronize()
        got = xio[:, :H].float().permute(0, 3, 1, 2)
        assert float(xio[:, H].float().abs().max()) == 0.0           # pad rows untouched (zero)
        err = (got - want_x).abs()
        assert bool((err <= 2.0 ** -6 * want_x.abs() + 3e-2).all()), (with_out2, float(err.max()))
        if with_out2:
            got_t = tout[:, :H].float().permute(0, 3, 1, 2)
            assert float(tout[:, H].float().abs().max()) == 0.0
            err_t = (got_t - want_t).abs()
            assert bool((err_t <= 2.0 ** -6 * want_t.abs() + 4e-2).all()), float(err_t.max())


@pytest.mark.parametrize("game", ["connect_four", "breakthrough"])
def test_evaluator_with_fused_blocks_matches_the_default_path(game, monkeypatch):
    """AZ_NN_BLOCK=1 (one k_block launch per residual block instead of two conv launches; opt-in, see az_resnet.cu): the
    whole evaluator agrees with the default path to bf16 rounding noise (both round the intermediate tensors to bf16 at
    the same places; only the order of fp32 operations inside the epilogues differs)."""
    import torch
    from alphazero_openspiel_b200 import engine as E, _lib as L
    from alphazero_openspiel_b200.network import Net
    from alphazero_openspiel_b200.nn_fused import FusedEvaluator
    shape, A = E.game_shape(game)
    torch.manual_seed(21)
    net = Net(shape, A).eval()
    B = 700
    hist, lens = E.game_random_playouts(game, B, seed=8, max_plies=30)
    xbf = E.game_replay_dev(game, hist, lens, L.OBS_BF16_NHWC)["obs"]
    monkeypatch.setenv("AZ_NN_BLOCK", "0")
    p0, v0 = FusedEvaluator(net, B, "cuda:0").eval_batch(xbf)
    monkeypatch.setenv("AZ_NN_BLOCK", "1")
    fe = FusedEvaluator(net, B, "cuda:0")
    assert fe.fused_blocks
    p1, v1 = fe.eval_batch(xbf)
    assert torch.isfinite(p1).all() and torch.isfinite(v1).all()
    assert (p1 - p0).abs().max().item() < 2e-3 and (v1 - v0).abs().max().item() < 1e-2
    p2, v2 = fe.eval_batch(xbf)                    # deterministic and re-entrant (the ring / barrier state is per launch)
    assert torch.equal(p1, p2) and torch.equal(v1, v2)
