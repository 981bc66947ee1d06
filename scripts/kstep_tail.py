"""Where does k_step's time go?  Per-tree cycle counts of one steady-state launch (az_debug_timing)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import _lib as L
from alphazero_openspiel_b200.examplegenerator import SelfPlayRunner
from alphazero_openspiel_b200.network import Net
from alphazero_openspiel_b200.engine import game_shape
# usage: kstep_tail.py [game] [trees] [playouts] [sim cap] [cycle budget]
game = sys.argv[1] if len(sys.argv) > 1 else "connect_four"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
NP = int(sys.argv[3]) if len(sys.argv) > 3 else 800
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 8
budget = int(sys.argv[5]) if len(sys.argv) > 5 else 0
torch.manual_seed(0)
shape, A = game_shape(game)
net = Net(shape, A).eval()
r = SelfPlayRunner(net, game, "cuda:0", T, n_playouts=NP, seed=0xC4, random_start_mod=21, max_sims_per_step=cap,
                   step_cycle_budget=budget, use_graph=False)
print("== %s trees %d playouts %d cap %d budget %d" % (game, T, NP, cap, budget))
r.round(3000)
buf = np.zeros((T, 4), dtype=np.int64)
L.check(r.engine.lib.az_debug_timing(r.engine.h, buf.ctypes.data))   # arm
for it in range(3):
    r.round(1)
    torch.cuda.synchronize()
    L.check(r.engine.lib.az_debug_timing(r.engine.h, buf.ctypes.data))
    cyc, ph, sims, misc = buf[:, 0], buf[:, 1], buf[:, 2], buf[:, 3]
    consume, moved = misc // 2, misc % 2
    print("launch %d: cycles p50 %d p90 %d p99 %d max %d | consume p50 %d p99 %d" % (it, *np.percentile(cyc, [50, 90, 99, 100]), *np.percentile(consume, [50, 99])))
    for k in range(0, 17):
        m = sims == k
        if m.any(): print("   sims=%d: n=%5d  cycles mean %7.0f max %7d" % (k, m.sum(), cyc[m].mean(), cyc[m].max()))
    for name, m in [("moved", moved == 1), ("root-eval in", ph == 1), ("leaf-eval in", ph == 2), ("run in", ph == 4), ("begin in", ph == 6), ("compact in", ph == 7)]:
        if m.any(): print("   %-13s n=%5d  cycles mean %7.0f max %7d" % (name, m.sum(), cyc[m].mean(), cyc[m].max()))
    top = np.argsort(-cyc)[:8]
    print("   slowest:", [(int(cyc[t]), int(ph[t]), int(sims[t]), int(moved[t]), int(consume[t])) for t in top])
