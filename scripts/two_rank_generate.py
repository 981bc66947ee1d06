"""torchrun check (2+ GPUs): ExampleGenerator under torch.distributed -- every rank plays its share, records are
all-gathered over NCCL, every rank ends with the same list of games; Trainer broadcasts identical weights."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200.examplegenerator import ExampleGenerator
from alphazero_openspiel_b200.network import Net
from alphazero_openspiel_b200 import parallel

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = parallel.rank_world()
torch.manual_seed(rank)                      # different weights per rank ...
net = Net([3, 6, 7], 7).eval()
parallel.broadcast_weights(net, src=0, device=torch.device("cuda", local))   # ... until the NCCL broadcast
w = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).double().sum().item()
gen = ExampleGenerator(net, "connect_four", torch.device("cuda", local), n_playouts=40, n_trees=32, seed=5)
games = gen.generate_examples(50)
sig = sum(len(g) for g in games), sum(float(g[0][3]) for g in games)
out = [None] * world
dist.all_gather_object(out, (w, len(games), sig))
if rank == 0:
    assert all(o == out[0] for o in out), out
    assert out[0][1] == 50, out
    print("two_rank_generate ok:", out[0])
dist.destroy_process_group()
