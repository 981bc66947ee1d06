"""B200-native AlphaZero self-play engine: drop-in for the self-play path of danielwillemsen/alphazero-openspiel.

Host-side mirror of the reference's Python API over the C-ABI CUDA engine (include/az_b200.h):
    mcts.MCTS, alphazerobot.AlphaZeroBot / remove_illegal_actions, game_utils.play_game_self,
    examplegenerator.ExampleGenerator, network.Net / state_to_board
"""
__version__ = "0.1.0"
