"""Strength numbers for the docs: shipped Connect Four checkpoint vs a uniform-random opponent through the whole stack."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200.evaluate import zero_vs_random
from alphazero_openspiel_b200.network import Net
ck = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "example_model_connect_four.pth")
net = Net([3, 6, 7], 7); net.load_state_dict(torch.load(ck, map_location="cpu", weights_only=True)); net.eval()
torch.manual_seed(0)
blank = Net([3, 6, 7], 7).eval()
for n_playouts in (100, 8, 3):
    print("%3d playouts, 256 pairs vs random: shipped %s   untrained %s" % (
        n_playouts, zero_vs_random(net, "connect_four", 256, n_playouts, seed=1),
        zero_vs_random(blank, "connect_four", 256, n_playouts, seed=1)))
