"""Drop-in for the reference's alphazerobot.py: AlphaZeroBot.step / restart over the CUDA engine.

  remove_illegal_actions(probs, legal)      alphazerobot.py:7-18
  AlphaZeroBot(game, player, policy_fn, self_play=False, keep_search_tree=True, **kwargs)
      .step(state) -> (policy, action)      alphazerobot.py:42-93
      .restart()                            (north star) fresh tree == constructing a new bot per game
      .mcts                                 the MCTS object callers read (.root, game_utils.py:174-194)
  NeuralNetBot(game, player, policy_fn).step(state) -> (policy, action)   alphazerobot.py:96-118 (no search, host only)

Host glue (masking, temperature, np.random.choice on the global numpy RNG) follows the reference expression
by expression so that policy targets and sampled actions are bit-identical for the same seed.
"""
import numpy as np

from .mcts import MCTS

try:  # real OpenSpiel when present; the bot is duck-typed like the reference otherwise (alphazerobot.py:27-28)
    import pyspiel as _pyspiel
    _BotBase = _pyspiel.Bot
    _GameType = _pyspiel.Game
except Exception:  # pragma: no cover - pyspiel absent
    _pyspiel = None
    _BotBase = object
    _GameType = None


def remove_illegal_actions(action_probabilities, legal_actions):
    mask = np.zeros(action_probabilities.shape, dtype=bool)
    mask[legal_actions] = True
    action_probabilities[~mask] = 0.0
    if np.sum(action_probabilities) > 1e-6:
        return action_probabilities / np.sum(action_probabilities)
    uniform = np.zeros(len(action_probabilities))
    uniform[legal_actions] = 1. / len(legal_actions)
    return uniform


class AlphaZeroBot(_BotBase):
    def __init__(self, game, player, policy_fn, self_play=False, keep_search_tree=True, **kwargs):
        if _GameType is not None and type(game) is _GameType:
            super().__init__(game, player)
        self.num_distinct_actions = game.num_distinct_actions()
        self.policy_fn = policy_fn
        self.kwargs = dict(kwargs)
        self.kwargs.setdefault("game_name", str(game))
        self.use_probabilistic_actions = self_play or bool(kwargs.get("use_probabilistic_actions"))
        self.use_random_actions = bool(kwargs.get("use_random_actions", False))
        self.num_probabilistic_actions = int(kwargs.get("num_probabilistic_actions", 1000))
        self.temperature = float(kwargs.get("temperature", 1.0))
        self.self_play = self_play
        self.keep_search_tree = keep_search_tree
        self.mcts = MCTS(self.policy_fn, self.num_distinct_actions, **self.kwargs)

    def restart(self):
        """Start a new game: forget the tree (the reference builds a new bot per game, game_utils.py:154)."""
        self.mcts.reset()

    def step(self, state):
        """One move.  (1) bring the tree to `state`: with tree reuse the moves played since the last search re-root it -- one
        move in self-play, where this bot plays both sides, two otherwise (alphazerobot.py:53-64); without reuse a new MCTS
        object is built, i.e. the tree is reset (alphazerobot.py:66-68).  (2) search on the device.  (3) host glue, kept
        expression by expression so that targets and sampled moves are bit-identical for a numpy seed: mask + renormalise
        the visit distribution (the returned policy, alphazerobot.py:75-78,89-93), temper it, then draw / argmax
        (alphazerobot.py:79-86)."""
        if self.keep_search_tree:
            history = state.history()
            if self.self_play:
                if history:
                    self.mcts.update_root(history[-1])
            elif len(history) >= 2:
                self.mcts.update_root(history[-2])
                self.mcts.update_root(history[-1])
        else:
            self.mcts.reset()

        visit_probs = np.array(self.mcts.search(state))
        legal_actions = state.legal_actions(state.current_player())
        visit_probs_legal = remove_illegal_actions(visit_probs, legal_actions)
        tempered = visit_probs_legal ** (1. / self.temperature)
        action_probabilities = tempered / sum(tempered)

        # exploration schedule: uniformly random or visit-proportional moves during the first num_probabilistic_actions
        # plies (global numpy RNG, one draw per move, after the search's Dirichlet draw: SURVEY A.10-A.11), greedy afterwards
        n_moves = len(state.history())
        if self.use_random_actions and n_moves < self.num_probabilistic_actions:
            action = np.random.choice(legal_actions)
        elif self.use_probabilistic_actions and n_moves < self.num_probabilistic_actions:
            action = np.random.choice(len(action_probabilities), p=action_probabilities)
        else:
            action = np.argmax(action_probabilities)
        policy = [(act, visit_probs_legal[act]) for act in legal_actions]
        return policy, action


class NeuralNetBot(_BotBase):
    """The network's policy without search (alphazerobot.py:96-118): mask the illegal moves, renormalise, play the first
    maximum.  Pure host code on the caller's pyspiel state."""

    def __init__(self, game, player, policy_fn):
        if _GameType is not None and type(game) is _GameType:
            super().__init__(game, player)
        self.policy_fn = policy_fn

    def step(self, state):
        priors, _value = self.policy_fn(state)
        legal = state.legal_actions(state.current_player())
        probs = remove_illegal_actions(np.array(priors), legal)
        action = np.argmax(probs)
        return [(a, probs[a]) for a in legal], action
