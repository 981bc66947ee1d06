"""Times the public drop-in call ExampleGenerator.generate_examples end to end (device play + record drain + conversion to
the reference's Python example lists)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphazero_openspiel_b200.examplegenerator import ExampleGenerator
from alphazero_openspiel_b200.network import Net
torch.manual_seed(0)
net = Net([3, 6, 7], 7).eval()
for n_games, n_playouts in [(4096, 100), (16384, 100)]:
    gen = ExampleGenerator(net, "connect_four", torch.device("cuda:0"), n_playouts=n_playouts, c_puct=2.5, dirichlet_ratio=0.25,
                           temperature=1.0, backup="on-policy", seed=1)
    t0 = time.time()
    games = gen.generate_examples(n_games)
    dt = time.time() - t0
    plies = sum(len(g) for g in games)
    st = gen.last_stats
    print("n_games=%d n_playouts=%d: %.2f s total, %d plies, %.0f games/s, %.2f M sims/s end to end (device counters: %d sims, %d rounds)"
          % (n_games, n_playouts, dt, plies, len(games) / dt, st["sims"] / dt / 1e6, st["sims"], st["rounds"]))
    gen = ExampleGenerator(net, "connect_four", torch.device("cuda:0"), n_playouts=n_playouts, c_puct=2.5, dirichlet_ratio=0.25,
                           temperature=1.0, backup="on-policy", seed=1)
    t0 = time.time()
    batch = gen.generate_batch(n_games)
    dt = time.time() - t0
    print("   generate_batch (arrays, replay.ExampleBatch): %.2f s total, %d examples, %.0f games/s" % (dt, len(batch), batch.n_games / dt))
    t0 = time.time()
    first, pol, val = batch.remove_duplicates()
    print("   remove_duplicates on arrays: %d -> %d examples in %.2f s" % (len(batch), len(first), time.time() - t0))
