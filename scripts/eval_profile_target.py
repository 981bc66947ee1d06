"""Target for ncu: the evaluator of BASELINE configs[2] (Connect Four, 16,384 boards) -- two evaluations (stem + 9 convs +
head each); profile the second one (`-s 11 -c 11` with `-k regex:k_conv8|k_head`).  Usage: eval_profile_target.py [game] [boards]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200 import engine as E, _lib as L
from alphazero_openspiel_b200.network import Net
from alphazero_openspiel_b200.nn_fused import FusedEvaluator
game = sys.argv[1] if len(sys.argv) > 1 else "connect_four"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
shape, A = E.game_shape(game)
torch.manual_seed(0)
net = Net(shape, A).eval()
hist, lens = E.game_random_playouts(game, B, seed=5, max_plies=20)
obs = E.game_replay_dev(game, hist, lens, L.OBS_BF16_NHWC)["obs"]
fe = FusedEvaluator(net, B, "cuda:0")
fe.obs.copy_(obs)
for _ in range(2):
    fe()
torch.cuda.synchronize()
print("ok")
