"""CPU: the N>1 host path (sharding, weight broadcast, record gather) with world_size-2 gloo processes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alphazero_openspiel_b200 import parallel
    from alphazero_openspiel_b200.engine import record_dtype
    from alphazero_openspiel_b200.network import Net
    out = {}
    out["rank_world"] = parallel.rank_world()
    out["shard"] = parallel.shard(11)
    torch.manual_seed(rank)  # different weights per rank before the broadcast
    net = Net([3, 6, 7], 7)
    parallel.broadcast_weights(net, src=0)
    out["wsum"] = float(sum(p.double().sum() for p in net.parameters()))
    out["bn"] = float(net.resblock1.bn1.running_var.sum())
    dt = record_dtype(7, 120)
    recs = np.zeros((3 + 2 * rank,), dtype=dt)  # ragged: 3 and 5 records
    recs["tree"] = np.arange(len(recs))
    recs["ply"] = 100 * rank + np.arange(len(recs))
    recs["root_q"] = rank + 0.5
    allr = parallel.gather_records(recs)
    out["n"] = len(allr)
    out["trees"] = allr["tree"].tolist()
    out["plies"] = allr["ply"].tolist()
    out["q"] = allr["root_q"].tolist()
    empty = parallel.gather_records(np.zeros((0,), dtype=dt))
    out["empty"] = len(empty)
    # data-parallel gradient averaging (Trainer(ddp=True)): grads = rank + 1 on every parameter -> mean 1.5
    for p_ in net.parameters():
        p_.grad = torch.full_like(p_, float(rank + 1))
    parallel.allreduce_gradients(net)
    out["grad"] = sorted({float(p_.grad.flatten()[0]) for p_ in net.parameters()})
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0]["rank_world"] == (0, 2) and res[1]["rank_world"] == (1, 2)
    assert res[0]["shard"] == (0, 6) and res[1]["shard"] == (6, 11)
    assert res[0]["wsum"] == res[1]["wsum"] and res[0]["bn"] == res[1]["bn"]
    for r in (0, 1):
        assert res[r]["n"] == 8 and res[r]["empty"] == 0
        assert res[r]["trees"] == [0, 1, 2] + [(1 << 20) + i for i in range(5)]
        assert res[r]["plies"] == [0, 1, 2, 100, 101, 102, 103, 104]
        assert res[r]["q"] == [0.5] * 3 + [1.5] * 5
        assert res[r]["grad"] == [1.5]


def test_shard_covers_everything():
    from alphazero_openspiel_b200 import parallel
    for n in (0, 1, 7, 16384, 100001):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
