"""Trainer drop-in: CPU parity of remove_duplicates / net_step with the oracle's restatement of
train.py (and with the live reference's remove_duplicates when /root/reference is present); GPU end-to-end generations."""
import copy
import os
import sys

import numpy as np
import pytest
import torch


def _fake_buffer(rng, n=300, n_keys=60, A=7):
    flat = []
    for _ in range(n):
        k = int(rng.randint(n_keys))
        pol = rng.dirichlet(np.ones(A)).tolist() if rng.rand() > 0.1 else []
        flat.append([str(k), rng.rand(4, 6, 7), pol, float(rng.choice([-1.0, 0.0, 1.0]))])
    return flat


def test_remove_duplicates_matches_oracle_and_reference():
    from alphazero_openspiel_b200.train import Trainer
    from oracle import ref_port
    rng = np.random.RandomState(0)
    buf = _fake_buffer(rng)
    a = Trainer.remove_duplicates(copy.deepcopy(buf))
    b = ref_port.dedupe_examples(copy.deepcopy(buf))
    assert len(a) == len(b) <= 60
    for x, y in zip(a, b):
        assert x[0] == y[0] and x[2] == y[2] and x[3] == y[3] and np.array_equal(x[1], y[1])
    # in-place accumulation quirk: the first example of every key is the returned (mutated) object
    mine = copy.deepcopy(buf)
    out = Trainer.remove_duplicates(mine)
    firsts = {}
    for ex in mine:
        firsts.setdefault(ex[0], ex)
    assert all(o is firsts[o[0]] for o in out)
    if os.path.isdir("/root/reference"):
        from oracle import pyspiel_shim
        pyspiel_shim.install()
        if "/root/reference" not in sys.path:
            sys.path.insert(0, "/root/reference")
        os.makedirs("/tmp/az_logs/logs", exist_ok=True)
        import importlib
        ref_train = importlib.import_module("train")
        c = ref_train.Trainer.remove_duplicates(copy.deepcopy(buf))
        assert len(a) == len(c)
        for x, y in zip(a, c):
            assert x[0] == y[0] and x[2] == y[2] and x[3] == y[3]


def test_net_step_matches_oracle_restatement_on_cpu():
    """Same seed, same buffer, same initial weights: losses and updated weights are identical (fp32, CPU)."""
    from alphazero_openspiel_b200.train import Trainer
    from oracle import ref_port, ref_net
    rng = np.random.RandomState(1)
    buf = _fake_buffer(rng, n=400, n_keys=400)
    torch.manual_seed(5)
    tr = Trainer(device="cpu", batch_size=32, use_gpu=False)
    ref = ref_net.RefNet([3, 6, 7], 7)
    ref.load_state_dict(tr.current_net.state_dict())
    opt = torch.optim.Adam(ref.parameters(), lr=0.001, weight_decay=0.0001)
    tr.current_net.train()
    ref.train()
    for step in range(3):
        np.random.seed(100 + step)
        lp, lv = tr.net_step(buf)
        np.random.seed(100 + step)
        rp, rv = ref_port.train_step(ref, opt, buf, 32)
        assert float(lp) == float(rp) and float(lv) == float(rv)
    for (ka, va), (kb, vb) in zip(tr.current_net.state_dict().items(), ref.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    assert tr.it == 3


def test_update_buffer_size_schedule():
    from alphazero_openspiel_b200.train import Trainer
    tr = Trainer(device="cpu", use_gpu=False, n_games_per_generation=10, n_games_buffer_max=60)
    assert tr.n_games_buffer == 40
    sizes = []
    for g in range(1, 8):
        tr.generation = g
        tr.update_buffer_size()
        sizes.append(tr.n_games_buffer)
    assert sizes == [40, 50, 50, 60, 60, 60, 60]  # grows every second generation up to the cap (train.py:295-298)


@pytest.mark.gpu
def test_trainer_generations_end_to_end():
    """Two tiny generations on the GPU: games enter the FIFO buffer, the network trains, weights change."""
    from alphazero_openspiel_b200.train import Trainer
    torch.manual_seed(0)
    np.random.seed(0)
    tr = Trainer(n_games_per_generation=24, n_batches_per_generation=6, batch_size=32, n_playouts_train=20,
                 n_generations=2, backup="soft-Z")
    before = copy.deepcopy(tr.current_net.state_dict())
    tr.run(n_trees=24, seed=11)
    assert tr.generation == 2 and len(tr.buffer) == 48 and tr.it == 12
    assert all(len(g) >= 7 for g in tr.buffer)
    after = tr.current_net.state_dict()
    assert any(not torch.equal(before[k].cpu(), after[k].cpu()) for k in before if "weight" in k)
    assert all(torch.isfinite(v).all() for v in after.values())
    assert tr.last_generation_stats["overflow"] == 0 and not tr.current_net.training
