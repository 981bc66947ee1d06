"""Round-2 overlap probe: does k_step of one game pool hide underneath the evaluator of another when it runs as a SMALL
persistent grid (az_set_step_ctas) that fits beside the evaluator's CTAs?  Two pools of B trees; times the evaluator alone,
k_step alone (full grid / small grids) and both concurrently on two streams.  Usage: overlap_probe2.py [B]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import engine as E, _lib as L
from alphazero_openspiel_b200.network import Net
from alphazero_openspiel_b200.nn_fused import FusedEvaluator
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
shape, A = E.game_shape("connect_four")
fe = FusedEvaluator(Net(shape, A).eval(), B, dev)
fe.obs.copy_((torch.rand((B, shape[1], shape[2], 4), device=dev) > 0.5).to(torch.bfloat16))
eng = E.Engine("connect_four", B, n_playouts=800, noise_mode=L.NOISE_DIRICHLET, eval_mode=L.EVAL_HASH,
               flags=L.F_KEEP_TREE | L.F_AUTO_RESTART | L.F_SAMPLE_MOVES | L.F_RANDOM_START, seed=1, max_sims_per_step=8,
               start_plies_mod=21)
obs = eng.new_obs(L.OBS_BF16_NHWC)
for _ in range(1500):
    eng.step(obs=obs)
torch.cuda.synchronize()
hi, lo = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
def run(n, do_nn, do_tree, tree_first):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = {"nn": [], "tree": []}
    ev_a, ev_b = torch.cuda.Event(), torch.cuda.Event()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        ev_a.record()
        hi.wait_event(ev_a); lo.wait_event(ev_a)
        def nn():
            if do_nn:
                with torch.cuda.stream(hi):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fe(); b.record(); per["nn"].append((a, b))
        def tree():
            if do_tree:
                with torch.cuda.stream(lo):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); eng.step(obs=obs); b.record(); per["tree"].append((a, b))
        if tree_first:
            tree(); nn()
        else:
            nn(); tree()
        ev_b.record(hi); torch.cuda.current_stream().wait_event(ev_b)
        ev_b.record(lo); torch.cuda.current_stream().wait_event(ev_b)
    e1.record(); torch.cuda.synchronize()
    extra = " ".join("%s %.1f" % (k, sum(a.elapsed_time(b) for a, b in v) / len(v) * 1e3) for k, v in per.items() if v)
    return "%.1f us  [%s]" % (e0.elapsed_time(e1) / n * 1e3, extra)
print("B = %d" % B)
print("nn alone                         %s" % run(100, True, False, False))
for ctas in (0, 148, 296, 444):
    eng.set_step_ctas(ctas)
    print("step_ctas %3d: tree alone        %s" % (ctas, run(100, False, True, False)))
    print("step_ctas %3d: both, tree first  %s" % (ctas, run(100, True, True, True)))
    print("step_ctas %3d: both, nn first    %s" % (ctas, run(100, True, True, False)))
