import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200.evaluate import zero_vs_random
from alphazero_openspiel_b200.network import Net
ck = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "example_model_connect_four.pth")
net = Net([3, 6, 7], 7); net.load_state_dict(torch.load(ck, map_location="cpu", weights_only=True)); net.eval()
print("shipped checkpoint, 100 playouts, 256 pairs vs random:", zero_vs_random(net, "connect_four", 256, 100, seed=1))
print("shipped checkpoint,  20 playouts, 256 pairs vs random:", zero_vs_random(net, "connect_four", 256, 20, seed=1))
torch.manual_seed(0)
print("untrained net,      100 playouts, 256 pairs vs random:", zero_vs_random(Net([3, 6, 7], 7).eval(), "connect_four", 256, 100, seed=1))
