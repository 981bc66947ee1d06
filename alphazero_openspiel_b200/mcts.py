"""Drop-in for the reference's mcts.py: MCTS(policy_fn, num_distinct_actions, **kwargs) over the CUDA engine.

Same constructor keywords, methods and result semantics as mcts.py:92-203 (search / update_root /
get_normalized_visit_counts / .root), but the tree lives in HBM and select / expand / backup run in the
az_step kernel.  The evaluator hook is unchanged: `policy_fn(state) -> (prior_ps, leaf_value)` is called on
the host once per requested leaf (mcts.py:146) and once per search for the root (mcts.py:183); Dirichlet
noise comes from the global numpy RNG in the reference's call order (mcts.py:187), so results are
bit-identical to the reference for the same seed.  This single-tree, host-evaluator mode exists for API
parity; throughput comes from the batched path (examplegenerator.ExampleGenerator).

States must be OpenSpiel states of connect_four / breakthrough (anything with .history(), .clone(),
.apply_action(), .legal_actions(), .current_player() and .get_game() or a `game_name=` keyword).
"""
import numpy as np
import torch

from . import _lib as L
from .engine import Engine


class NodeView:
    """Read-only snapshot of a node's statistics (mcts.py:10-20 attribute names)."""

    def __init__(self, N, Q, P, children=None):
        self.N, self.Q, self.P = N, Q, P
        self.children = children if children is not None else {}
        self.parent = None

    def is_leaf(self):
        return self.children == {}

    def is_root(self):
        return self.parent is None


def _game_name_of(state, kwargs):
    name = kwargs.get("game_name")
    if name:
        return name
    get_game = getattr(state, "get_game", None)
    if get_game is None:
        raise TypeError("MCTS needs game_name=... or a state with get_game()")
    return str(get_game())


class MCTS:
    def __init__(self, policy_fn, num_distinct_actions, c_puct=2.5, n_playouts=100, use_dirichlet=True,
                 dirichlet_ratio=0.25, use_puct=True, **kwargs):
        self.policy_fn = policy_fn
        self.num_distinct_actions = num_distinct_actions
        self.c_puct = c_puct
        self.n_playouts = n_playouts
        self.use_dirichlet = use_dirichlet
        self.dirichlet_ratio = dirichlet_ratio
        self.use_puct = use_puct
        self.kwargs = kwargs
        self.device_index = int(kwargs.get("device_index", 0))
        self._eng = None
        self._hist = None       # action history the engine's root position corresponds to
        self._root_cache = None
        self._leaf_root_updated = False
        self.evaluations = 0

    # ---- engine plumbing
    def _engine(self, state):
        if self._eng is None:
            name = _game_name_of(state, self.kwargs)
            flags = L.F_MANUAL | L.F_PRIORS_F64 | L.F_KEEP_TREE | (0 if self.use_puct else L.F_UCT)
            self._eng = Engine(name, 1, n_playouts=self.n_playouts, c_puct=self.c_puct,
                               dirichlet_ratio=self.dirichlet_ratio,
                               noise_mode=L.NOISE_HOST if self.use_dirichlet else L.NOISE_NONE,
                               eval_mode=L.EVAL_EXTERNAL, flags=flags, device=self.device_index,
                               node_capacity=int(self.kwargs.get("node_capacity", 0)))
            if self._eng.num_actions != self.num_distinct_actions:
                raise ValueError("num_distinct_actions=%d does not match %s (%d)" %
                                 (self.num_distinct_actions, name, self._eng.num_actions))
            dev = self._eng.device
            self._priors = torch.zeros((1, self.num_distinct_actions), dtype=torch.float64, device=dev)
            self._values = torch.zeros((1,), dtype=torch.float64, device=dev)
            self._noise = torch.zeros((1, self._eng.max_children), dtype=torch.float64, device=dev)
            if self._leaf_root_updated:
                self._eng.command(reset_tree=[2])   # the root update_root built for the leaf root (carries use_puct)
        return self._eng

    def _sync_position(self, eng, state):
        hist = [int(a) for a in state.history()]
        if self._hist != hist:
            eng.set_positions([hist])
            self._hist = hist

    def search(self, state):
        """n_playouts simulations from `state`; returns normalised root visit counts (mcts.py:164-180)."""
        eng = self._engine(state)
        self._sync_position(eng, state)
        self._root_cache = None
        eng.command(begin=[1])
        have = False
        while True:
            eng.step(self._priors if have else None, self._values if have else None,
                     self._noise if self.use_dirichlet else None)
            st = eng.status()
            phase = int(st["phase"][0])
            if phase == L.PH_SEARCH_DONE:
                break
            if phase == L.PH_ERROR:
                raise RuntimeError("search tree overflow (node_capacity / depth)")
            if phase not in (L.PH_ROOT_EVAL, L.PH_LEAF_EVAL):
                have = False
                continue
            info = eng.request_info(max_depth=192)
            leaf = state.clone()
            for a in info["path"][0][:max(int(info["depth"][0]), 0)]:
                leaf.apply_action(int(a))
            prior_ps, leaf_value = self.policy_fn(leaf)
            self.evaluations += 1
            pri = np.asarray(prior_ps, dtype=np.float64).reshape(-1)
            self._priors.copy_(torch.from_numpy(pri).view(1, -1))
            self._values.fill_(float(leaf_value))
            if phase == L.PH_ROOT_EVAL:
                n_legal = int(st["req_legal"][0])
                eta = np.random.dirichlet(0.3 * np.ones(n_legal))  # mcts.py:187, after policy_fn as in the reference
                self._noise.zero_()
                self._noise[0, :n_legal] = torch.from_numpy(np.asarray(eta, dtype=np.float64))
            have = True
        if eng.counters()["overflow"]:
            raise RuntimeError("search tree overflow (node_capacity)")
        return self.get_normalized_visit_counts()

    def playout(self, state):
        """One simulation from the current root (mcts.py:126-153): select to a leaf applying the actions to `state` IN PLACE
        (pass a copy, as the reference asks), evaluate the leaf with policy_fn unless it is terminal, expand, back up.
        No root Dirichlet expansion -- that belongs to search() (mcts.py:174-175).  One az_command(begin = 2) + at most
        two az_step calls.  (`state` must be the position of the tree's root and non-terminal, as in MCTS.search.)"""
        eng = self._engine(state)
        self._sync_position(eng, state)
        self._root_cache = None
        eng.command(begin=[2])
        have = False
        for _ in range(4):
            eng.step(self._priors if have else None, self._values if have else None,
                     self._noise if self.use_dirichlet else None)
            st = eng.status()
            phase = int(st["phase"][0])
            if phase == L.PH_SEARCH_DONE:
                break
            if phase == L.PH_ERROR:
                raise RuntimeError("search tree overflow (node_capacity / depth)")
            assert phase == L.PH_LEAF_EVAL, phase
            info = eng.request_info(max_depth=192)
            for a in info["path"][0][:max(int(info["depth"][0]), 0)]:
                state.apply_action(int(a))                  # the caller's state walks to the leaf
            prior_ps, leaf_value = self.policy_fn(state)     # mcts.py:146
            self.evaluations += 1
            pri = np.asarray(prior_ps, dtype=np.float64).reshape(-1)
            self._priors.copy_(torch.from_numpy(pri).view(1, -1))
            self._values.fill_(float(leaf_value))
            have = True
        else:
            raise RuntimeError("playout did not finish")
        if not have:
            # terminal leaf (mcts.py:147-149): the engine backed it up itself; replay the path of that simulation
            info = eng.request_info(max_depth=192)
            for a in info["path"][0][:max(int(info["depth"][0]), 0)]:
                state.apply_action(int(a))
        if eng.counters()["overflow"]:
            raise RuntimeError("search tree overflow (node_capacity)")

    def _stats(self):
        if self._root_cache is None:
            if self._eng is None:
                self._root_cache = {"root_n": [0], "root_q": [0.0], "n_children": [0]}
            else:
                self._root_cache = self._eng.root_stats()
        return self._root_cache

    def root_child_visits(self):
        s = self._stats()
        visits = [0] * self.num_distinct_actions
        for k in range(int(s["n_children"][0])):
            visits[int(s["child_action"][0][k])] = int(s["child_n"][0][k])
        return visits

    def get_normalized_visit_counts(self):
        visits = self.root_child_visits()
        total = sum(visits)
        return [float(v) / total for v in visits]  # ZeroDivisionError when nothing was visited, like mcts.py:162

    @property
    def root(self):
        s = self._stats()
        kids = {}
        for k in range(int(s["n_children"][0])):
            kids[int(s["child_action"][0][k])] = NodeView(int(s["child_n"][0][k]), float(s["child_q"][0][k]),
                                                         float(s["child_p"][0][k]))
        node = NodeView(int(s["root_n"][0]), float(s["root_q"][0]), 0.0, kids)
        for c in kids.values():
            c.parent = node
        return node

    def value_targets(self):
        """(soft-Z, A0C, off-policy) targets of game_utils.py:172-194 for the current root."""
        s = self._stats()
        return -float(s["root_q"][0]), float(s["v_a0c"][0]), float(s["v_offpolicy"][0])

    def update_root(self, action):
        """Re-root at the child of `action`, or a fresh root when the root is a leaf (mcts.py:192-203)."""
        self._root_cache = None
        if self._eng is None:
            # untouched tree: the root is a leaf and is replaced by Node(None, 0.0, use_puct=self.use_puct) (mcts.py:199-200);
            # remembered until the engine exists
            self._leaf_root_updated = True
            return
        try:
            self._eng.command(update_root=[int(action)])
        except RuntimeError as e:
            raise KeyError(action) from e
        if self._hist is not None:
            self._hist = self._hist + [int(action)]

    def reset(self):
        """self.mcts = MCTS(...) (alphazerobot.py:66-68): fresh root, same engine."""
        self._root_cache = None
        self._leaf_root_updated = False
        if self._eng is not None:
            self._eng.command(reset_tree=[1])

    def random_rollout(self, state):
        """Stand-in evaluator of mcts.py:205-223: uniform priors and the outcome of one uniformly random playout from
        `state`, seen by the player to move there.  Host code on the caller's pyspiel state; draws from the global numpy
        RNG in the reference's order (one `choice` per ply), so searches driven by it reproduce the reference's."""
        playout = state.clone()
        mover = playout.current_player()
        while not playout.is_terminal():
            playout.apply_action(np.random.choice(playout.legal_actions()))
        return np.ones(self.num_distinct_actions), playout.player_return(mover)
