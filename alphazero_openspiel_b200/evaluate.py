"""Batched strength checks: the reference's match-ups (game_utils.py:51-145), all games at once on the GPU.

  zero_vs_random   AlphaZeroBot vs a uniform-random opponent            `test_zero_vs_random`  game_utils.py:51-63
  net_vs_random    NeuralNetBot (argmax of the policy, no search) vs random   `test_net_vs_random`   game_utils.py:100-112
  zero_vs_mcts     AlphaZeroBot vs a vanilla UCT bot with one random rollout  `test_zero_vs_mcts`    game_utils.py:66-80
  net_vs_mcts      NeuralNetBot vs the UCT bot                                `test_net_vs_mcts`     game_utils.py:83-97
  zero_vs_zero     two AlphaZeroBots (two nets / two settings), Dirichlet on  `test_zero_vs_zero`    game_utils.py:115-145

Every game is one manual-mode tree (AZ_F_MANUAL) per searching side.  An AlphaZero side runs `n_playouts` simulations with
the batched tcgen05 evaluator and plays the first maximal visit count (alphazerobot.py:81-86; or samples from the visit
counts with `use_probabilistic_actions`), keeping its tree across both players' moves (alphazerobot.py:53-64).  The UCT side
stands in for OpenSpiel's `MCTSBot(game, player, uct_c=1, max_simulations, RandomRolloutEvaluator(1))`, which is not part
of the reference: it is the reference's OWN search in its UCT mode (`use_puct=False`, mcts.py:80: inf for unvisited
children, else Q + c * P * sqrt(log N_parent / N)) with the `MCTS.random_rollout` evaluator (mcts.py:205-223: priors of ones,
value of one random playout), both on the device (AZ_F_UCT + AZ_EVAL_ROLLOUT), a fresh tree per move like MCTSBot, and the
most-visited child as its move.  Legal moves / terminal tests come from the device game kernels (az_game_replay).  These
are end-to-end behavioural pins: with the reference's shipped checkpoints a wrong observation encoding, action numbering
or value sign shows up as a lost match (tournament.py:19-27 claims > 99 % against the 200-simulation MCTS bot on 6x6
Breakthrough).
"""
import numpy as np
import torch

from . import _lib as L
from .engine import Engine, game_replay
from .nn_fused import FusedEvaluator


class _SearchSide:
    """One searching bot for n parallel games."""

    def __init__(self, kind, game_name, n, index, net=None, n_playouts=100, c_puct=2.5, use_dirichlet=False,
                 keep_search_tree=True, use_probabilistic_actions=False, temperature=1.0, seed=0, **_ignored):
        self.kind, self.n, self.keep = kind, n, keep_search_tree
        self.sample = bool(use_probabilistic_actions)
        self.temperature = float(temperature)
        self.rng = np.random.RandomState(seed + 17)
        if kind == "zero":
            flags = L.F_MANUAL | (L.F_KEEP_TREE if keep_search_tree else 0)
            self.eng = Engine(game_name, n, n_playouts=n_playouts, c_puct=c_puct, eval_mode=L.EVAL_EXTERNAL, flags=flags,
                              noise_mode=L.NOISE_DIRICHLET if use_dirichlet else L.NOISE_NONE, device=index, seed=seed)
            self.ev = FusedEvaluator(net, n, torch.device("cuda", index))
        elif kind == "uct":
            # vanilla UCT, one random rollout per expansion, fresh tree per move (see the module docstring)
            self.keep = False
            self.eng = Engine(game_name, n, n_playouts=n_playouts, c_puct=c_puct, eval_mode=L.EVAL_ROLLOUT,
                              flags=L.F_MANUAL | L.F_UCT, noise_mode=L.NOISE_NONE, device=index, seed=seed)
            self.ev = None
        else:
            raise ValueError(kind)

    def search(self, mask):
        """Run one search in the games of `mask`; -> int32 actions (-1 outside the mask)."""
        eng, ev = self.eng, self.ev
        m = mask.astype(np.int32)
        if not self.keep:
            eng.command(reset_tree=m * (2 if self.kind == "uct" else 1))   # 2: the root that carries the UCT formula
        eng.command(begin=m)
        first = True
        for _ in range(1000000):
            if ev is not None:
                eng.step(None if first else ev.priors, None if first else ev.values, None, ev.obs, L.OBS_BF16_NHWC)
                ev()
            else:
                eng.step()
            first = False
            ph = eng.phases().cpu().numpy()
            if (ph[mask] == L.PH_ERROR).any():
                raise RuntimeError("search arena overflow during evaluation")
            if not np.isin(ph[mask], (L.PH_ROOT_EVAL, L.PH_LEAF_EVAL, L.PH_RUN)).any():
                break
        st = eng.root_stats(offpolicy=False)
        actions = np.full(self.n, -1, dtype=np.int32)
        for i in np.flatnonzero(mask):
            counts = st["child_n"][i, :st["n_children"][i]].astype(np.float64)
            if self.sample and counts.sum() > 0:   # alphazerobot.py:78,84
                pr = counts ** (1.0 / self.temperature)
                k = int(self.rng.choice(len(pr), p=pr / pr.sum()))
            else:
                k = int(np.argmax(counts))          # first maximal visit count (alphazerobot.py:86)
            actions[i] = st["child_action"][i, k]
        return actions

    def advance(self, actions):
        """Both players' moves re-root the tree / advance the position (alphazerobot.py:60-64)."""
        self.eng.command(update_root=actions)

    def close(self):
        if self.eng.counters()["overflow"]:
            self.eng.close()
            raise RuntimeError("search arena overflow during evaluation")
        self.eng.close()


def _make_side(spec, game_name, n, index, seed):
    """spec: ("random",) | ("net", net) | ("zero", net, settings) | ("uct", settings)"""
    kind = spec[0]
    if kind == "random":
        return {"kind": kind, "rng": np.random.RandomState(seed)}
    if kind == "net":
        return {"kind": kind, "ev": FusedEvaluator(spec[1], n, torch.device("cuda", index))}
    if kind == "zero":
        return {"kind": kind, "bot": _SearchSide("zero", game_name, n, index, net=spec[1], seed=seed, **dict(spec[2]))}
    if kind == "uct":
        return {"kind": kind, "bot": _SearchSide("uct", game_name, n, index, seed=seed, **dict(spec[1]))}
    raise ValueError(kind)


@torch.no_grad()
def play_match(side_a, side_b, game_name, n_pairs, device="cuda:0", seed=0):
    """n_pairs games with side A moving first and n_pairs with side B moving first (`play_game`, game_utils.py:16-35, for
    2*n_pairs games at once).  Returns (mean score of A as first player, mean score of A as second player), in [-1, 1]."""
    from .alphazerobot import remove_illegal_actions
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    n = 2 * n_pairs
    sides = [_make_side(side_a, game_name, n, index, seed), _make_side(side_b, game_name, n, index, seed + 1)]
    a_player = np.array([0] * n_pairs + [1] * n_pairs)          # which seat side A has in game i
    hist = [[] for _ in range(n)]
    score = np.zeros(n)
    alive = np.ones(n, dtype=bool)
    try:
        while alive.any():
            need_obs = any(s["kind"] == "net" for s in sides)
            rep = game_replay(game_name, hist, L.OBS_BF16_NHWC if need_obs else L.OBS_NONE, device=index)
            status = rep["status"].cpu().numpy()
            ret0 = rep["return0"].cpu().numpy()
            n_legal = rep["n_legal"].cpu().numpy()
            legal = rep["legal"].cpu().numpy()
            for i in np.flatnonzero(alive & ((status & 1) == 1)):
                score[i] = ret0[i] if a_player[i] == 0 else -ret0[i]
                alive[i] = False
            if not alive.any():
                break
            to_move = np.array([len(h) % 2 for h in hist])
            actions = np.full(n, -1, dtype=np.int32)
            for which, side in enumerate(sides):
                seat = a_player if which == 0 else 1 - a_player
                turn = alive & (to_move == seat)
                if not turn.any():
                    continue
                if side["kind"] == "random":
                    for i in np.flatnonzero(turn):
                        actions[i] = legal[i, side["rng"].randint(n_legal[i])]
                elif side["kind"] == "net":      # NeuralNetBot.step (alphazerobot.py:96-118)
                    priors, _ = side["ev"].eval_batch(rep["obs"])
                    priors = priors.double().cpu().numpy()
                    for i in np.flatnonzero(turn):
                        acts = [int(a) for a in legal[i, :n_legal[i]]]
                        actions[i] = int(np.argmax(remove_illegal_actions(priors[i].copy(), acts)))
                else:
                    got = side["bot"].search(turn)
                    actions[turn] = got[turn]
            for side in sides:
                if "bot" in side:
                    side["bot"].advance(actions)
            for i in np.flatnonzero(alive):
                hist[i].append(int(actions[i]))
    finally:
        for side in sides:
            if "bot" in side:
                side["bot"].close()
    return float(score[:n_pairs].mean()), float(score[n_pairs:].mean())


def zero_vs_random(net, game_name, n_pairs, n_playouts=100, c_puct=2.5, device="cuda:0", seed=0, keep_search_tree=True):
    """`test_zero_vs_random` (game_utils.py:51-63) for 2*n_pairs games: AlphaZeroBot(use_dirichlet=False) vs uniform random.
    Returns the two mean scores from AlphaZero's point of view (as first / as second player)."""
    zero = ("zero", net, dict(n_playouts=n_playouts, c_puct=c_puct, use_dirichlet=False, keep_search_tree=keep_search_tree))
    return play_match(zero, ("random",), game_name, n_pairs, device=device, seed=seed)


def net_vs_random(net, game_name, n_pairs, device="cuda:0", seed=0):
    """`test_net_vs_random` (game_utils.py:100-112): argmax(policy restricted to the legal moves) vs uniform random."""
    return play_match(("net", net), ("random",), game_name, n_pairs, device=device, seed=seed)


def zero_vs_mcts(net, game_name, n_pairs, max_search_nodes, n_playouts=100, c_puct=2.5, uct_c=1.0, device="cuda:0", seed=0,
                 **settings):
    """`test_zero_vs_mcts` (game_utils.py:66-80): AlphaZeroBot(use_dirichlet=False, **settings) vs the UCT bot with
    `max_search_nodes` simulations of one random rollout each (tournament.py:19-27: 100 vs 200 on 6x6 Breakthrough)."""
    zs = dict(n_playouts=n_playouts, c_puct=c_puct, use_dirichlet=False)
    zs.update(settings)
    return play_match(("zero", net, zs), ("uct", dict(n_playouts=max_search_nodes, c_puct=uct_c)), game_name, n_pairs,
                      device=device, seed=seed)


def net_vs_mcts(net, game_name, n_pairs, max_search_nodes, uct_c=1.0, device="cuda:0", seed=0):
    """`test_net_vs_mcts` (game_utils.py:83-97): NeuralNetBot vs the UCT bot."""
    return play_match(("net", net), ("uct", dict(n_playouts=max_search_nodes, c_puct=uct_c)), game_name, n_pairs,
                      device=device, seed=seed)


def zero_vs_zero(net, game_name, n_pairs, net2=None, settings1=None, settings2=None, device="cuda:0", seed=0):
    """`test_zero_vs_zero` (game_utils.py:115-145): two AlphaZeroBots with Dirichlet noise on, their own settings
    (n_playouts, c_puct, use_probabilistic_actions, ...) and optionally their own networks; scores are bot 1's."""
    s1 = dict(use_dirichlet=True)
    s1.update(settings1 or {})
    s2 = dict(use_dirichlet=True)
    s2.update(settings2 or {})
    return play_match(("zero", net, s1), ("zero", net2 if net2 is not None else net, s2), game_name, n_pairs,
                      device=device, seed=seed)
