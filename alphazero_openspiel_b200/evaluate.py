"""Batched strength checks, all games at once on the GPU:
  zero_vs_random   AlphaZeroBot against a uniform-random opponent   (`test_zero_vs_random`, game_utils.py:51-63)
  net_vs_random    NeuralNetBot (argmax of the network policy over the legal moves, no search; alphazerobot.py:96-118)
                   against a uniform-random opponent                 (`test_net_vs_random`, game_utils.py:100-112)

Every game is one manual-mode tree (AZ_F_MANUAL): on AlphaZero's turn the tree runs `n_playouts` simulations with the
batched evaluator (no Dirichlet noise, argmax move -- AlphaZeroBot with self_play=False, alphazerobot.py:81-86), the chosen
move and then the opponent's random move re-root the tree (alphazerobot.py:60-64).  Legal moves / terminal tests for the
random opponent come from the device game kernels (az_game_replay).  This is an end-to-end behavioural pin: with the
reference's shipped checkpoint a wrong observation encoding, action numbering or value sign shows up as a lost match.
"""
import numpy as np
import torch

from . import _lib as L
from .engine import Engine, game_replay
from .nn_fused import FusedEvaluator


@torch.no_grad()
def zero_vs_random(net, game_name, n_pairs, n_playouts=100, c_puct=2.5, device="cuda:0", seed=0, keep_search_tree=True):
    """n_pairs games with AlphaZero moving first and n_pairs with the random bot moving first.
    Returns (mean score as first player, mean score as second player), each in [-1, 1] from AlphaZero's point of view."""
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    n = 2 * n_pairs
    flags = L.F_MANUAL | (L.F_KEEP_TREE if keep_search_tree else 0)
    eng = Engine(game_name, n, n_playouts=n_playouts, c_puct=c_puct, noise_mode=L.NOISE_NONE, eval_mode=L.EVAL_EXTERNAL,
                 flags=flags, device=index)
    ev = FusedEvaluator(net, n, torch.device("cuda", index))
    rng = np.random.RandomState(seed)
    zero_player = np.array([0] * n_pairs + [1] * n_pairs)
    hist = [[] for _ in range(n)]
    score = np.zeros(n)
    alive = np.ones(n, dtype=bool)
    try:
        while alive.any():
            rep = game_replay(game_name, hist, L.OBS_NONE, device=index)
            status = rep["status"].cpu().numpy()
            ret0 = rep["return0"].cpu().numpy()
            n_legal = rep["n_legal"].cpu().numpy()
            legal = rep["legal"].cpu().numpy()
            for i in np.flatnonzero(alive & ((status & 1) == 1)):
                score[i] = ret0[i] if zero_player[i] == 0 else -ret0[i]
                alive[i] = False
            if not alive.any():
                break
            to_move = np.array([len(h) % 2 for h in hist])
            az_turn = alive & (to_move == zero_player)
            actions = np.full(n, -1, dtype=np.int32)
            if az_turn.any():
                if not keep_search_tree:
                    eng.command(reset_tree=az_turn.astype(np.int32))
                eng.command(begin=az_turn.astype(np.int32))
                first = True
                for _ in range(100000):
                    eng.step(None if first else ev.priors, None if first else ev.values, None, ev.obs, L.OBS_BF16_NHWC)
                    first = False
                    ev()
                    ph = eng.phases().cpu().numpy()
                    if not np.isin(ph[az_turn], (L.PH_ROOT_EVAL, L.PH_LEAF_EVAL, L.PH_RUN)).any():
                        break
                st = eng.root_stats(offpolicy=False)
                for i in np.flatnonzero(az_turn):
                    k = int(np.argmax(st["child_n"][i, :st["n_children"][i]]))  # first maximal visit count
                    actions[i] = st["child_action"][i, k]
            for i in np.flatnonzero(alive & ~az_turn):                          # uniform random opponent
                actions[i] = legal[i, rng.randint(n_legal[i])]
            eng.command(update_root=actions)
            for i in np.flatnonzero(alive):
                hist[i].append(int(actions[i]))
        if eng.counters()["overflow"]:
            raise RuntimeError("search arena overflow during evaluation")
    finally:
        eng.close()
    return float(score[:n_pairs].mean()), float(score[n_pairs:].mean())


@torch.no_grad()
def net_vs_random(net, game_name, n_pairs, device="cuda:0", seed=0):
    """`test_net_vs_random` for n_pairs games with the network moving first and n_pairs with the random bot moving first:
    the network side plays argmax(policy restricted to the legal moves), exactly NeuralNetBot.step.  Returns the two mean
    scores from the network's point of view."""
    from .alphazerobot import remove_illegal_actions
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    n = 2 * n_pairs
    ev = FusedEvaluator(net, n, torch.device("cuda", index))
    rng = np.random.RandomState(seed)
    net_player = np.array([0] * n_pairs + [1] * n_pairs)
    hist = [[] for _ in range(n)]
    score = np.zeros(n)
    alive = np.ones(n, dtype=bool)
    while alive.any():
        rep = game_replay(game_name, hist, L.OBS_BF16_NHWC, device=index)
        status = rep["status"].cpu().numpy()
        ret0 = rep["return0"].cpu().numpy()
        n_legal = rep["n_legal"].cpu().numpy()
        legal = rep["legal"].cpu().numpy()
        for i in np.flatnonzero(alive & ((status & 1) == 1)):
            score[i] = ret0[i] if net_player[i] == 0 else -ret0[i]
            alive[i] = False
        if not alive.any():
            break
        to_move = np.array([len(h) % 2 for h in hist])
        net_turn = alive & (to_move == net_player)
        if net_turn.any():
            priors, _ = ev.eval_batch(rep["obs"])
            priors = priors.double().cpu().numpy()
        for i in np.flatnonzero(alive):
            if net_turn[i]:
                acts = [int(a) for a in legal[i, :n_legal[i]]]
                a = int(np.argmax(remove_illegal_actions(priors[i].copy(), acts)))
            else:
                a = int(legal[i, rng.randint(n_legal[i])])
            hist[i].append(a)
    return float(score[:n_pairs].mean()), float(score[n_pairs:].mean())
