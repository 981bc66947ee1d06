"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (alphazero_openspiel_b200) never does, and fails loudly without its CUDA
library instead of falling back to anything here.

Contents
  az_oracle.{h,c}   plain-C restatement: games (OpenSpiel semantics) + mcts.py arithmetic
  cbind.py          ctypes binding of libaz_oracle.so
  pyspiel_shim.py   `pyspiel`-shaped Game/State (old API names) over the C games, so the reference's
                    own Python (mcts.py, alphazerobot.py, game_utils.py) runs unmodified in-container
  ref_port.py       pure-Python restatement of MCTS / AlphaZeroBot / play_game_self / ExampleGenerator
                    (flat arrays instead of Node objects) -- the CPU baseline that travels to the GPU box
"""
