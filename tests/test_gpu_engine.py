"""GPU parity: the CUDA engine (through the C-ABI) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from tests import oracle_util as ou

pytestmark = pytest.mark.gpu

GAMES = ["connect_four", "breakthrough(rows=6,columns=6)", "breakthrough"]


@pytest.mark.parametrize("game", GAMES)
def test_game_dynamics_bit_exact(game):
    """legal sets / terminal / returns / observation planes vs the oracle on random playouts."""
    import torch
    from alphazero_openspiel_b200 import engine as E, _lib as L
    n, max_plies = 4096, 42 if game == "connect_four" else 160
    hist, lens = E.game_random_playouts(game, n, seed=11, max_plies=max_plies)
    hist_h, lens_h = hist.cpu().numpy(), lens.cpu().numpy()
    rng = np.random.RandomState(0)
    cut = np.array([rng.randint(0, l + 1) for l in lens_h], dtype=np.int32)  # random prefixes incl. full games
    cut[: n // 8] = lens_h[: n // 8]
    out = E.game_replay_dev(game, hist, torch.from_numpy(cut).to(hist.device), L.OBS_F32_NCHW)
    out_bf = E.game_replay_dev(game, hist, torch.from_numpy(cut).to(hist.device), L.OBS_BF16_NHWC)
    o = {k: v.cpu().numpy() for k, v in out.items()}
    obs_bf = out_bf["obs"].float().cpu().numpy()
    bb = o["bb"].view(np.uint64)
    n_term = 0
    for i in range(0, n, 4):
        ref = ou.replay(game, hist_h[i, :cut[i]])
        assert o["status"][i] == (1 if ref["terminal"] else 0)
        assert (int(bb[i, 0]), int(bb[i, 1])) == ref["bb"]
        assert o["return0"][i] == ref["returns0"]
        assert list(o["legal"][i, :o["n_legal"][i]]) == ref["legal"]
        if not ref["terminal"]:
            assert np.array_equal(o["obs"][i].astype(np.float64), ref["board"])
            assert np.array_equal(obs_bf[i].transpose(2, 0, 1).astype(np.float64), ref["board"])
        n_term += ref["terminal"]
    assert n_term > 50


def _run_engine_selfplay(game, n_trees, n_playouts, seed, noise, sample, keep, flags_extra=0, max_steps=200000,
                         max_sims_per_step=0, eval_mode=None, c_puct=2.5, step_cycle_budget=0):
    from alphazero_openspiel_b200 import engine as E, _lib as L
    flags = L.F_RECORDS | L.F_OFFPOLICY | flags_extra
    if keep:
        flags |= L.F_KEEP_TREE
    if sample:
        flags |= L.F_SAMPLE_MOVES
    eng = E.Engine(game, n_trees, n_playouts=n_playouts, noise_mode=noise, flags=flags, c_puct=c_puct,
                   eval_mode=L.EVAL_HASH if eval_mode is None else eval_mode,
                   seed=seed, max_sims_per_step=max_sims_per_step, step_cycle_budget=step_cycle_budget)
    steps = 0
    while True:
        for _ in range(64):
            eng.step()
        steps += 64
        ph = eng.phases()
        if int((ph != L.PH_IDLE).sum()) == 0:
            break
        assert steps < max_steps
    recs = eng.drain_records()
    ctr = eng.counters()
    eng.close()
    return recs, ctr


@pytest.mark.parametrize("game,n_trees,n_playouts", [("connect_four", 64, 100), ("connect_four", 16, 800),
                                                     ("breakthrough(rows=6,columns=6)", 16, 200),
                                                     ("breakthrough", 8, 120)])
@pytest.mark.parametrize("keep", [1, 0])
def test_selfplay_visit_counts_bit_exact(game, n_trees, n_playouts, keep):
    """Whole self-play games, counter-mode noise + sampling, hash evaluator: every ply's root visit counts, Q,
    A0C / off-policy targets, chosen action and position must equal the oracle's exactly."""
    from alphazero_openspiel_b200 import _lib as L
    seed = 1234
    recs, ctr = _run_engine_selfplay(game, n_trees, n_playouts, seed, L.NOISE_COUNTER, 1, keep)
    assert ctr["overflow"] == 0
    cfg = ou.selfplay_cfg(game, n_playouts, use_dirichlet=2, sample_moves=1, keep_tree=keep, seed=seed)
    tot = np.zeros(8, dtype=np.int64)
    for t in range(n_trees):
        mine = recs[recs["tree"] == t]
        plies = mine[mine["kind"] == 0]
        end = mine[mine["kind"] == 1]
        ref, ret, c = ou.selfplay_game(cfg, t)
        tot += np.array(c, dtype=np.int64)
        assert len(plies) == len(ref) and len(end) == 1
        order = np.argsort(plies["ply"])
        for r, g in zip(ref, plies[order]):
            assert g["ply"] == r["ply"] and g["n_legal"] == r["n_legal"]
            assert list(g["counts"][:r["n_legal"]]) == r["counts"], (t, r["ply"])
            assert g["action"] == r["action"]
            assert (int(g["bb"][0]), int(g["bb"][1])) == r["bb"]
            assert g["root_n"] == r["root_n"]
            assert g["root_q"] == r["root_q"]
            assert g["v_a0c"] == r["v_a0c"]
            assert g["v_offpolicy"] == r["v_offpolicy"]
        assert end[0]["root_q"] == ret[0]
    assert ctr["sims"] == tot[0] and ctr["depth"] == tot[1] and ctr["children"] == tot[2]
    assert ctr["expansions"] == tot[3] and ctr["legal"] == tot[4] and ctr["terminal"] == tot[5]
    assert ctr["root_evals"] == tot[6]


def test_no_dirichlet_argmax_and_sim_cap():
    """use_dirichlet=False + argmax moves; a per-step simulation cap (by count, max_sims_per_step, or by SM cycles,
    step_cycle_budget) must not change any result."""
    from alphazero_openspiel_b200 import _lib as L
    game, n_trees, n_playouts, seed = "connect_four", 32, 60, 7
    cfg = ou.selfplay_cfg(game, n_playouts, use_dirichlet=0, sample_moves=0, keep_tree=1, seed=seed)
    idle = {}
    for cap, budget in ((0, 0), (3, 0), (0, 1), (0, 30000), (8, 20000)):
        recs, ctr = _run_engine_selfplay(game, n_trees, n_playouts, seed, L.NOISE_NONE, 0, 1, max_sims_per_step=cap,
                                         step_cycle_budget=budget)
        idle[(cap, budget)] = ctr["idle_slots"]
        assert ctr["overflow"] == 0
        for t in range(0, n_trees, 8):
            plies = recs[(recs["tree"] == t) & (recs["kind"] == 0)]
            plies = plies[np.argsort(plies["ply"])]
            ref, ret, _ = ou.selfplay_game(cfg, t)
            assert len(plies) == len(ref)
            for r, g in zip(ref, plies):
                assert list(g["counts"][:r["n_legal"]]) == r["counts"] and g["action"] == r["action"]
                assert g["root_q"] == r["root_q"]
    assert idle[(0, 1)] > idle[(0, 0)]     # a one-cycle budget really stops trees after their first simulation


def test_full_size_pool_matches_oracle_on_sampled_trees():
    """BASELINE config 3 size (16,384 trees x 800 sims/move, tree reuse, random synthetic starts): trees are independent, so
    a sample of them must still equal the oracle ply by ply, and every recorded search must satisfy the size-independent
    invariants (visit counts sum to the root's visits gained, legal counts, arena never overflows)."""
    from alphazero_openspiel_b200 import engine as E, _lib as L
    game, n_trees, n_playouts, seed = "connect_four", 16384, 800, 0xC4
    flags = L.F_RECORDS | L.F_KEEP_TREE | L.F_SAMPLE_MOVES | L.F_RANDOM_START | L.F_AUTO_RESTART
    eng = E.Engine(game, n_trees, n_playouts=n_playouts, noise_mode=L.NOISE_COUNTER, eval_mode=L.EVAL_HASH, flags=flags,
                   seed=seed, start_plies_mod=21, max_sims_per_step=16, record_capacity=4 << 20)
    sample = list(range(0, 8)) + [4095, 8191, 16383]
    recs_all = []
    for chunk in range(60):  # up to 30,000 steps: until the sampled trees have finished their first game
        for _ in range(500):
            eng.step()
        recs_all.append(eng.drain_records())
        r = np.concatenate(recs_all)
        ends = r[(r["kind"] == 1) & (r["game_seq"] == 0)]
        if all((ends["tree"] == t).any() for t in sample):
            break
    recs = np.concatenate(recs_all)
    ctr = eng.counters()
    eng.close()
    assert ctr["overflow"] == 0 and ctr["moves"] > 3 * n_trees
    assert ctr["peak_nodes"] > 2 * 7 * n_playouts   # trees really carried several searches' worth of nodes across re-roots
    plies = recs[recs["kind"] == 0]
    assert len(plies) == ctr["moves"]
    # invariants over ALL records: a search adds exactly n_playouts visits below the root
    # (first search of a game: children sum == n_playouts; with reuse the root keeps inherited visits)
    csum = plies["counts"].sum(axis=1)
    # every playout of a Dirichlet-expanded root visits a child; a re-rooted node additionally carries the one visit
    # in which it was itself the expanded leaf (mcts.py:145-152)
    assert np.all((csum == plies["root_n"]) | (csum == plies["root_n"] - 1))
    assert np.all(plies["root_n"] >= n_playouts)
    assert np.all(plies["n_legal"] >= 1) and np.all(plies["n_legal"] <= 7)
    assert np.all(np.abs(plies["root_q"]) <= 1.0)
    cfg = ou.selfplay_cfg(game, n_playouts, use_dirichlet=2, sample_moves=1, keep_tree=1, seed=seed, start_mod=21)
    checked = 0
    for t in sample:
        mine = plies[(plies["tree"] == t) & (plies["game_seq"] == 0)]
        mine = mine[np.argsort(mine["ply"])]
        ref, ret, _ = ou.selfplay_game(cfg, t)
        assert len(mine) == len(ref)                      # the WHOLE first game, every ply, with tree reuse
        for r, g in zip(ref, mine):
            assert g["ply"] == r["ply"] and list(g["counts"][:r["n_legal"]]) == r["counts"] and g["action"] == r["action"]
            assert g["root_q"] == r["root_q"] and (int(g["bb"][0]), int(g["bb"][1])) == r["bb"]
            assert g["root_n"] == r["root_n"]
            checked += 1
        end = recs[(recs["kind"] == 1) & (recs["tree"] == t) & (recs["game_seq"] == 0)]
        assert len(end) == 1 and end[0]["root_q"] == ret[0]
    assert checked >= 100


def test_arena_overflow_is_counted_and_raised():
    """A too-small node arena must never corrupt memory or pass silently: the engine counts it, the host raises."""
    import torch
    from alphazero_openspiel_b200 import engine as E, _lib as L
    from alphazero_openspiel_b200.examplegenerator import ExampleGenerator
    from alphazero_openspiel_b200.network import Net
    eng = E.Engine("connect_four", 8, n_playouts=200, noise_mode=L.NOISE_COUNTER, eval_mode=L.EVAL_HASH,
                   flags=L.F_KEEP_TREE | L.F_SAMPLE_MOVES, seed=1, node_capacity=64)
    for _ in range(400):
        eng.step()
    assert eng.counters()["overflow"] > 0
    eng.close()
    net = Net([3, 6, 7], 7).eval()
    gen = ExampleGenerator(net, "connect_four", torch.device("cuda:0"), n_playouts=200, n_trees=8, node_capacity=64, seed=2)
    with pytest.raises(RuntimeError, match="overflow"):
        gen.generate_examples(8)


def test_device_dirichlet_noise_statistics():
    """AZ_NOISE_DIRICHLET (throughput mode): P = 0.75*p + 0.25*eta with eta ~ Dirichlet(0.3 * 1_7) at the initial position
    (uniform priors): recover eta from the root children and check simplex, mean 1/7 and the Dirichlet variance."""
    from alphazero_openspiel_b200 import engine as E, _lib as L
    n = 8192
    eng = E.Engine("connect_four", n, n_playouts=50, noise_mode=L.NOISE_DIRICHLET, eval_mode=L.EVAL_UNIFORM,
                   flags=L.F_KEEP_TREE, seed=77)
    eng.step()   # root-eval requests
    eng.step()   # consume: root expanded with noise, first simulation runs
    st = eng.root_stats(offpolicy=False)
    eng.close()
    assert np.all(st["n_children"] == 7)
    eta = (st["child_p"] - 0.75 / 7.0) / 0.25
    assert np.all(eta > -1e-12) and np.allclose(eta.sum(axis=1), 1.0, atol=1e-9)
    a, a0 = 0.3, 2.1
    assert np.allclose(eta.mean(axis=0), 1.0 / 7, atol=0.01)
    want_var = a * (a0 - a) / (a0 * a0 * (a0 + 1))
    assert np.allclose(eta.var(axis=0), want_var, rtol=0.08)
    # different trees draw different noise; the same seed reproduces the same noise
    assert len(np.unique(np.round(eta[:, 0], 9))) > 0.99 * n


@pytest.mark.parametrize("temperature", [1.0, 0.5])
def test_device_move_sampling_follows_visit_counts(temperature):
    """alphazerobot.py:78,84: moves are sampled with p ~ N^(1/T).  Uniform evaluator, no noise => every tree has the same
    root visit counts at ply 0, so the empirical move frequencies over 8,192 trees must match (chi-square)."""
    from alphazero_openspiel_b200 import engine as E, _lib as L
    n, n_playouts = 8192, 60
    eng = E.Engine("connect_four", n, n_playouts=n_playouts, noise_mode=L.NOISE_NONE, eval_mode=L.EVAL_UNIFORM,
                   flags=L.F_KEEP_TREE | L.F_SAMPLE_MOVES | L.F_RECORDS, seed=5, temperature=temperature)
    for _ in range(n_playouts + 4):
        eng.step()
    recs = eng.drain_records()
    eng.close()
    first = recs[(recs["kind"] == 0) & (recs["ply"] == 0)]
    assert len(first) == n
    counts = first["counts"][0, :7].astype(np.float64)
    assert np.all(first["counts"][:, :7] == counts)          # identical searches
    p = counts ** (1.0 / temperature)
    p /= p.sum()
    freq = np.bincount(first["action"], minlength=7).astype(np.float64)
    chi2 = float((((freq - n * p) ** 2) / (n * p)).sum())
    assert chi2 < 30.0, (chi2, freq, n * p)                  # 6 dof: P(chi2 > 30) ~ 4e-5


@pytest.mark.parametrize("eager", [0, 1])
def test_async_compaction_on_side_stream_is_bit_exact(eager):
    """AZ_F_ASYNC_COMPACT: the caller runs az_compact on a side stream between two az_step calls (what SelfPlayRunner does
    next to the evaluator).  Results must equal the oracle exactly, like the synchronous mode -- both when the kept subtree
    is compacted after every move (AZ_F_EAGER_COMPACT) and when the tree re-roots in place until its arena half fills up."""
    import torch
    from alphazero_openspiel_b200 import engine as E, _lib as L
    game, n_trees, n_playouts, seed = "connect_four", 64, 100, 4321
    flags = L.F_RECORDS | L.F_OFFPOLICY | L.F_KEEP_TREE | L.F_SAMPLE_MOVES | L.F_ASYNC_COMPACT | \
        (L.F_EAGER_COMPACT if eager else 0)
    eng = E.Engine(game, n_trees, n_playouts=n_playouts, noise_mode=L.NOISE_COUNTER, eval_mode=L.EVAL_HASH, flags=flags,
                   seed=seed)
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()
    for it in range(200000):
        eng.step()
        ev_fork.record(main)
        side.wait_event(ev_fork)
        eng.compact(side)
        ev_join.record(side)
        main.wait_event(ev_join)
        if it % 256 == 255 and int((eng.phases() != L.PH_IDLE).sum()) == 0:
            break
    recs = eng.drain_records()
    ctr = eng.counters()
    eng.close()
    assert ctr["overflow"] == 0 and ctr["compact_nodes"] > 0
    cfg = ou.selfplay_cfg(game, n_playouts, use_dirichlet=2, sample_moves=1, keep_tree=1, seed=seed)
    for t in range(0, n_trees, 4):
        plies = recs[(recs["tree"] == t) & (recs["kind"] == 0)]
        plies = plies[np.argsort(plies["ply"])]
        ref, _, _ = ou.selfplay_game(cfg, t)
        assert len(plies) == len(ref)
        for r, g in zip(ref, plies):
            assert list(g["counts"][:r["n_legal"]]) == r["counts"] and g["action"] == r["action"]
            assert g["root_q"] == r["root_q"] and g["v_offpolicy"] == r["v_offpolicy"] and g["root_n"] == r["root_n"]


@pytest.mark.parametrize("game,n_trees,n_playouts", [("connect_four", 256, 60), ("breakthrough(rows=6,columns=6)", 256, 24)])
def test_batched_external_evaluator_path_bit_exact(game, n_trees, n_playouts):
    """The path the benchmark runs: AZ_EVAL_EXTERNAL with fp32 `priors[tree*A + a]` / `values[tree]` rows for n_trees >> 1
    (az_engine.cu eval_prior / eval_value).  The host answers every request from the oracle's evaluator on the
    az_request_info bitboards; the records must equal the in-kernel AZ_EVAL_HASH run AND the C oracle ply by ply
    (visit counts, root N / Q, targets, actions, positions) -- north_star: "bit-exact given the same evaluator outputs"
    (mcts.py:146 policy_fn call, network.py:66-80 fp32 outputs widened to Python floats)."""
    import torch
    from oracle import cbind
    from alphazero_openspiel_b200 import engine as E, _lib as L
    seed = 2024
    flags = L.F_RECORDS | L.F_OFFPOLICY | L.F_KEEP_TREE | L.F_SAMPLE_MOVES
    eng = E.Engine(game, n_trees, n_playouts=n_playouts, noise_mode=L.NOISE_COUNTER, eval_mode=L.EVAL_EXTERNAL, flags=flags,
                   seed=seed)
    A = eng.num_actions
    shift = 2 if game == "connect_four" else 4
    olib = cbind.lib()
    pri_h = torch.zeros((n_trees, A), dtype=torch.float32).pin_memory()
    val_h = torch.zeros((n_trees,), dtype=torch.float32).pin_memory()
    pri_d = torch.zeros((n_trees, A), dtype=torch.float32, device=eng.device)
    val_d = torch.zeros((n_trees,), dtype=torch.float32, device=eng.device)
    eng.step()                       # produces the first requests (nothing to consume yet)
    answered = 0
    for it in range(200000):
        info = eng.request_info(max_depth=1)
        bb = np.ascontiguousarray(info["bb"])
        ply = np.ascontiguousarray(info["ply"])
        if int((ply >= 0).sum()) == 0 and int((eng.phases() != L.PH_IDLE).sum()) == 0:
            break                    # (a step without any request is possible: every live tree waits for k_compact)
        answered += int((ply >= 0).sum())
        # rows of trees without a request keep stale values on purpose: the engine must ignore them
        olib.oz_synth_eval_bb(A, n_trees, bb.ctypes.data, ply.ctypes.data, 1, seed, shift, pri_h.data_ptr(), val_h.data_ptr())
        pri_d.copy_(pri_h, non_blocking=True)
        val_d.copy_(val_h, non_blocking=True)
        eng.step(pri_d, val_d)
        torch.cuda.synchronize()     # the pinned rows are rewritten next iteration
    recs = eng.drain_records()
    ctr = eng.counters()
    eng.close()
    assert ctr["overflow"] == 0 and int((recs["kind"] == 1).sum()) == n_trees
    assert answered == ctr["expansions"] + ctr["root_evals"]      # one evaluator row per expansion / root evaluation
    recs_h, ctr_h = _run_engine_selfplay(game, n_trees, n_playouts, seed, L.NOISE_COUNTER, 1, 1)
    key = lambda r: np.lexsort((r["kind"], r["ply"], r["tree"]))  # noqa: E731
    a, b = recs[key(recs)], recs_h[key(recs_h)]
    assert len(a) == len(b)
    for f in ["tree", "ply", "action", "n_legal", "kind", "root_n", "bb", "root_q", "v_a0c", "v_offpolicy", "counts", "actions"]:
        assert np.array_equal(a[f], b[f]), f
    for k in ["sims", "depth", "children", "expansions", "legal", "terminal", "root_evals", "moves", "games"]:
        assert ctr[k] == ctr_h[k], k
    cfg = ou.selfplay_cfg(game, n_playouts, use_dirichlet=2, sample_moves=1, keep_tree=1, seed=seed)
    for t in range(n_trees):
        plies = recs[(recs["tree"] == t) & (recs["kind"] == 0)]
        plies = plies[np.argsort(plies["ply"])]
        ref, ret, _ = ou.selfplay_game(cfg, t)
        assert len(plies) == len(ref)
        for r, g in zip(ref, plies):
            assert list(g["counts"][:r["n_legal"]]) == r["counts"] and g["action"] == r["action"], (t, r["ply"])
            assert g["root_q"] == r["root_q"] and g["root_n"] == r["root_n"]
            assert g["v_a0c"] == r["v_a0c"] and g["v_offpolicy"] == r["v_offpolicy"]
            assert (int(g["bb"][0]), int(g["bb"][1])) == r["bb"]


@pytest.mark.parametrize("game,n_trees,n_playouts", [("connect_four", 64, 80), ("breakthrough(rows=6,columns=6)", 32, 40)])
def test_device_random_rollout_evaluator_bit_exact(game, n_trees, n_playouts):
    """AZ_EVAL_ROLLOUT = MCTS.random_rollout (mcts.py:205-223) in the kernel: priors of ones and the outcome of one random
    playout from the leaf for the player to move there.  Whole self-play games must equal the oracle running the same
    evaluator (oz_synth_eval kind 2) ply by ply -- this also pins the device playout loop (legal / apply / outcome)."""
    from alphazero_openspiel_b200 import _lib as L
    seed = 55
    recs, ctr = _run_engine_selfplay(game, n_trees, n_playouts, seed, L.NOISE_COUNTER, 1, 1, eval_mode=L.EVAL_ROLLOUT, c_puct=1.0)
    assert ctr["overflow"] == 0
    cfg = ou.selfplay_cfg(game, n_playouts, c_puct=1.0, use_dirichlet=2, sample_moves=1, keep_tree=1, seed=seed, eval_kind=2)
    for t in range(n_trees):
        plies = recs[(recs["tree"] == t) & (recs["kind"] == 0)]
        plies = plies[np.argsort(plies["ply"])]
        ref, ret, _ = ou.selfplay_game(cfg, t)
        assert len(plies) == len(ref)
        for r, g in zip(ref, plies):
            assert list(g["counts"][:r["n_legal"]]) == r["counts"] and g["action"] == r["action"], (t, r["ply"])
            assert g["root_q"] == r["root_q"] and g["root_n"] == r["root_n"] and g["v_offpolicy"] == r["v_offpolicy"]


def _first_search_counts(game, n_trees, n_playouts, leaves, seed=31):
    """Root visit counts of every tree's FIRST search (ply 0) with the hash evaluator, no root noise, argmax moves."""
    from alphazero_openspiel_b200 import engine as E, _lib as L
    flags = L.F_RECORDS | L.F_KEEP_TREE | (L.F_VIRTUAL_LOSS if leaves else 0)
    eng = E.Engine(game, n_trees, n_playouts=n_playouts, noise_mode=L.NOISE_NONE, eval_mode=L.EVAL_HASH, flags=flags,
                   seed=seed, leaves_per_tree=max(leaves, 1), eval_shift=0)
    steps = 0
    while True:
        for _ in range(32):
            eng.step()
        steps += 32
        recs = eng.drain_records()
        first = recs[(recs["kind"] == 0) & (recs["ply"] == 0)]
        if len(first) or steps > 100000:
            break
    ctr = eng.counters()
    eng.close()
    assert ctr["overflow"] == 0
    return first, steps, ctr


@pytest.mark.parametrize("game,n_playouts", [("connect_four", 400), ("breakthrough(rows=6,columns=6)", 200)])
def test_virtual_loss_mode_is_close_to_the_exact_search(game, n_playouts):
    """AZ_F_VIRTUAL_LOSS (K leaves in flight per tree) is a labelled NON-bit-exact throughput mode (the reference runs its
    playouts strictly in sequence, mcts.py:177-179).  It must finish a search in ~1/K of the evaluator round trips, run
    exactly n_playouts simulations, and its root visit distribution must stay close to the exact one: same evaluator, same
    position, total-variation distance of the normalised visit counts below the bound stated here, same most-visited move
    in most searches.  K = 0 (flag off) remains the bit-exact kernel (covered by the parity tests above)."""
    n_trees = 64   # the hash evaluator depends on the seed only: every tree runs the same search; trees differ by nothing
    exact, steps1, _ = _first_search_counts(game, n_trees, n_playouts, 0)
    assert len(exact) == n_trees
    a = exact["counts"][0].astype(np.float64)
    assert np.all(exact["counts"] == exact["counts"][0])
    for K, tv_bound in [(4, 0.12), (8, 0.2)]:
        vl, stepsK, ctr = _first_search_counts(game, n_trees, n_playouts, K)
        assert len(vl) == n_trees
        b = vl["counts"][0].astype(np.float64)
        assert np.all(vl["counts"] == vl["counts"][0])            # deterministic: same inputs, same schedule
        assert b.sum() == a.sum() == n_playouts - 1                # n_playouts simulations, the first one expands the root
        assert stepsK <= steps1 / K * 1.6 + 64                     # ~K leaves per round trip
        tv = 0.5 * np.abs(a / a.sum() - b / b.sum()).sum()
        assert tv < tv_bound, (K, tv, a, b)
        assert int(np.argmax(a)) == int(np.argmax(b)), (K, a, b)


def test_external_evaluator_rows_at_full_pool_size():
    """Size-independent property at BASELINE configs[2]'s pool size (16,384 trees, synthetic random starts, auto-restart,
    simulation cap): answering every request with fp32 rows computed by the oracle's evaluator from the az_request_info
    bitboards (the AZ_EVAL_EXTERNAL read path the benchmark uses) must reproduce the in-kernel AZ_EVAL_HASH engine
    exactly -- every record of the first moves of every tree and all counters."""
    import torch
    from oracle import cbind
    from alphazero_openspiel_b200 import engine as E, _lib as L
    n_trees, n_playouts, seed, steps = 16384, 100, 9, 330
    flags = L.F_RECORDS | L.F_KEEP_TREE | L.F_SAMPLE_MOVES | L.F_AUTO_RESTART | L.F_RANDOM_START
    olib = cbind.lib()
    out = {}
    for mode in (L.EVAL_HASH, L.EVAL_EXTERNAL):
        eng = E.Engine("connect_four", n_trees, n_playouts=n_playouts, noise_mode=L.NOISE_COUNTER, eval_mode=mode, flags=flags,
                       seed=seed, start_plies_mod=21, max_sims_per_step=8)
        if mode == L.EVAL_EXTERNAL:
            pri_h = torch.zeros((n_trees, 7), dtype=torch.float32).pin_memory()
            val_h = torch.zeros((n_trees,), dtype=torch.float32).pin_memory()
            pri_d = torch.zeros((n_trees, 7), dtype=torch.float32, device=eng.device)
            val_d = torch.zeros((n_trees,), dtype=torch.float32, device=eng.device)
            eng.step()
            for _ in range(steps - 1):
                info = eng.request_info(max_depth=1)
                bb, ply = np.ascontiguousarray(info["bb"]), np.ascontiguousarray(info["ply"])
                olib.oz_synth_eval_bb(7, n_trees, bb.ctypes.data, ply.ctypes.data, 1, seed, 2, pri_h.data_ptr(), val_h.data_ptr())
                pri_d.copy_(pri_h, non_blocking=True)
                val_d.copy_(val_h, non_blocking=True)
                eng.step(pri_d, val_d)
                torch.cuda.synchronize()
        else:
            for _ in range(steps):
                eng.step()
        out[mode] = (eng.drain_records(), eng.counters())
        eng.close()
    (ra, ca), (rb, cb) = out[L.EVAL_HASH], out[L.EVAL_EXTERNAL]
    assert ca == cb and ca["overflow"] == 0 and ca["moves"] > 2 * n_trees
    key = lambda r: np.lexsort((r["kind"], r["ply"], r["game_seq"], r["tree"]))  # noqa: E731
    a, b = ra[key(ra)], rb[key(rb)]
    assert len(a) == len(b) > 2 * n_trees
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f]), f
