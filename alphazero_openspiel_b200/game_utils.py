"""Drop-in for the self-play driver of the reference's game_utils.py: play_game_self (game_utils.py:148-206), plus the
two-bot loop play_game (game_utils.py:16-35) for callers that run their own match-ups with AlphaZeroBot / NeuralNetBot.

One game, one AlphaZeroBot for both sides, host-side `policy_fn(state)`; emits per ply
[information_state string, state_to_board planes, dense policy list, value] with the value target chosen by
`backup` (on-policy / soft-Z / A0C / off-policy).  The three tree-derived targets are computed on the device
(az_root_stats) instead of walking Python Node objects.  Needs `pyspiel` for the State objects, exactly like
the reference; the batched, pyspiel-free path is examplegenerator.ExampleGenerator.
"""
import copy

from .alphazerobot import AlphaZeroBot
from .network import state_to_board


def play_game(game, player1, player2, generate_statistics=False):
    """One game between two bots (game_utils.py:16-35): player1 moves on even plies; returns player 0's return, and with
    generate_statistics the roots both bots held after every move."""
    stats = {"player1": [], "player2": []}
    state = game.new_initial_state()
    while not state.is_terminal():
        mover = player1 if len(state.history()) % 2 == 0 else player2
        _, action = mover.step(state)
        state.apply_action(action)
        if generate_statistics:
            stats["player1"].append({"root": copy.deepcopy(player1.mcts.root)})
            stats["player2"].append({"root": copy.deepcopy(player2.mcts.root)})
    return (state.returns()[0], stats) if generate_statistics else state.returns()[0]


def play_game_self(policy_fn, game_name, **kwargs):
    import pyspiel  # the caller's OpenSpiel, as in the reference (game_utils.py:2)
    game = pyspiel.load_game(game_name)
    state = game.new_initial_state()
    state_shape = game.information_state_normalized_vector_shape()
    n_actions = game.num_distinct_actions()
    kwargs = dict(kwargs)
    kwargs.setdefault("game_name", game_name)
    bot = AlphaZeroBot(game, 0, policy_fn, self_play=True, **kwargs)
    backup = str(kwargs.get("backup", "on-policy"))
    examples = []
    while not state.is_terminal():
        policy, action = bot.step(state)
        by_action = dict(policy)
        policy_list = [by_action.get(i, 0.0) for i in range(n_actions)]
        value = None
        if backup in ("soft-Z", "A0C", "off-policy"):
            soft_z, a0c, off_policy = bot.mcts.value_targets()
            value = {"soft-Z": soft_z, "A0C": a0c, "off-policy": off_policy}[backup]
        if backup in ("on-policy", "soft-Z", "A0C", "off-policy"):
            examples.append([state.information_state(), state_to_board(state, state_shape), policy_list, value])
        state.apply_action(action)
    if backup == "on-policy":
        reward = state.returns()[0]
        for ex in examples:
            ex[3] = reward
            reward *= -1
    return examples


# ---- the reference's match-up functions (game_utils.py:51-145) as names for ExampleGenerator.generate_tests(n, game_fn, m):
# the reference passes the function objects; here they are thin single-pair entry points over the batched harness
# (evaluate.py) with the reference's signatures: (policy_fn-or-net, max_search_nodes, game_name, **kwargs) -> (score1, score2, None)
def _net_of(policy_fn):
    net = getattr(policy_fn, "__self__", policy_fn)       # `net.predict` bound method or the net itself
    if not hasattr(net, "resblock1"):
        raise TypeError("the batched match-ups need the network (pass net or net.predict), not an arbitrary policy_fn")
    return net


def test_zero_vs_mcts(policy_fn, max_search_nodes, game_name, **kwargs):
    from . import evaluate
    s = {k: kwargs[k] for k in ("c_puct", "n_playouts", "temperature", "keep_search_tree") if k in kwargs}
    a, b = evaluate.zero_vs_mcts(_net_of(policy_fn), game_name, 1, max_search_nodes, **s)
    return a, b, None


def test_net_vs_mcts(policy_fn, max_search_nodes, game_name, **kwargs):
    from . import evaluate
    a, b = evaluate.net_vs_mcts(_net_of(policy_fn), game_name, 1, max_search_nodes)
    return a, b, None


def test_net_vs_random(policy_fn, game_name, **kwargs):
    from . import evaluate
    return evaluate.net_vs_random(_net_of(policy_fn), game_name, 1)


def test_zero_vs_random(policy_fn, game_name="connect_four", **kwargs):
    from . import evaluate
    a, b = evaluate.zero_vs_random(_net_of(policy_fn), game_name, 1)
    return a, b, None


def test_zero_vs_zero(policy_fn, max_search_nodes, game_name, policy_fn2=None, generate_statistics=False, **kwargs):
    from . import evaluate
    a, b = evaluate.zero_vs_zero(_net_of(policy_fn), game_name, 1, net2=_net_of(policy_fn2) if policy_fn2 else None,
                                 settings1=kwargs.get("settings1"), settings2=kwargs.get("settings2"))
    return a, b, {}


for _f in (test_zero_vs_mcts, test_net_vs_mcts, test_net_vs_random, test_zero_vs_random, test_zero_vs_zero):
    _f.__test__ = False     # reference API names, not pytest tests
