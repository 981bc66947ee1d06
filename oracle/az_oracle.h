/*
 * az_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the arithmetic on the self-play MCTS path of
 * danielwillemsen/alphazero-openspiel, used ONLY as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under alphazero_openspiel_b200/ may import, link or call this.
 *
 * What it restates (reference file:line, all under /root/reference):
 *   - Node.select / get_value      mcts.py:38-52, 68-80   (fp64 PUCT, first-max tie-break)
 *   - Node.expand                  mcts.py:54-66          (child per legal action, P overwrite on re-expand)
 *   - Node.update(_recursive)      mcts.py:82-89          (Q <- (N*Q+v)/(N+1), sign flip per level)
 *   - MCTS.playout                 mcts.py:126-153
 *   - MCTS.search                  mcts.py:164-180
 *   - MCTS.expand_root_dirichlet   mcts.py:182-190        (noise injected by the caller)
 *   - MCTS.update_root             mcts.py:192-203
 *   - value targets                game_utils.py:168-194  (soft-Z, A0C, off-policy)
 *
 * Game dynamics (Connect Four, Breakthrough RxC) are NOT in the reference: they
 * live in OpenSpiel (github.com/deepmind/open_spiel, un-vendored, un-pinned; API
 * names date it to Dec 2019 - Jan 2020).  oz_* game functions restate the published
 * connect_four.cc / breakthrough.cc semantics.
 * PARITY STATUS: the MCTS part is pinned against the reference's own mcts.py run
 * in-container (tests/test_oracle_vs_reference.py, tests/golden/); the game part is
 * "parity unpinned" against real pyspiel (absent, no network) and only indirectly
 * pinned through the shipped checkpoints (tests/golden/encoding_pins.npz).
 *
 * The oracle deliberately uses byte-per-cell boards and loops, not bitboards, so
 * that it is an independent implementation from the CUDA kernels it checks.
 */
#ifndef AZ_ORACLE_H
#define AZ_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OZ_GAME_CONNECT_FOUR 0
#define OZ_GAME_BREAKTHROUGH 1
#define OZ_MAX_CELLS 64
#define OZ_MAX_LEGAL 64
#define OZ_TERMINAL_PLAYER (-4) /* OpenSpiel kTerminalPlayerId */

typedef struct {
  int32_t game, rows, cols;
  int32_t player;    /* side to move, 0 or 1 (also at terminal: the side that WOULD move) */
  int32_t ply;
  int32_t winner;    /* -1 none yet, 0 / 1, 2 = draw */
  int32_t pieces[2]; /* breakthrough piece counts */
  int8_t cell[OZ_MAX_CELLS]; /* 0 empty, 1 player-0 piece, 2 player-1 piece; index row*cols+col */
} oz_state;

/* ---- games ---- */
void oz_init(oz_state* s, int game, int rows, int cols);
int oz_num_actions(const oz_state* s);
int oz_legal(const oz_state* s, int32_t* out); /* ascending; 0 when terminal */
int oz_apply(oz_state* s, int action);         /* 0 ok, -1 illegal (state untouched) */
int oz_terminal(const oz_state* s);
int oz_current_player(const oz_state* s);      /* OZ_TERMINAL_PLAYER when terminal */
void oz_returns(const oz_state* s, double out[2]);
/* OpenSpiel normalized vector, 3*rows*cols floats (plane order per game, App. B.2/B.3) */
void oz_normalized_vector(const oz_state* s, float* out);
/* network.py:9-18 state_to_board: (3+1, rows, cols) with current-player plane */
void oz_board(const oz_state* s, double* out);
/* canonical bitboards (bit index = cell index) used as the key of the synthetic evaluator */
void oz_bitboards(const oz_state* s, uint64_t out[2]);

/* ---- counter-based hash "RNG" + synthetic evaluator (shared definition with the CUDA engine) ---- */
uint64_t oz_mix64(uint64_t x);
uint64_t oz_counter(uint64_t seed, uint64_t tree, uint64_t game_seq, uint64_t ply, uint64_t idx, uint64_t stream);
/* kind 0: uniform 1/A, value 0.  kind 1: hash priors (1+r10)*2^-(10+shift), value on a 1/16 grid.
   kind 2: MCTS.random_rollout (mcts.py:205-223): priors = 1, value = one random playout with hash-stream move picks */
void oz_synth_eval(const oz_state* s, int kind, uint64_t seed, int shift, double* priors, double* value);

/* the same evaluator for n positions given as canonical bitboards + ply (fp32 out; rows with ply < 0 untouched) */
void oz_synth_eval_bb(int num_actions, int n, const uint64_t* bb, const int32_t* ply, int kind, uint64_t seed, int shift,
                      float* priors, float* values);

/* ---- MCTS ---- */
typedef void (*oz_eval_fn)(const oz_state* s, double* priors, double* value, void* user);
typedef struct oz_tree oz_tree;

oz_tree* oz_tree_new(int num_actions, double c_puct, int n_playouts, int use_dirichlet, double dirichlet_ratio);
void oz_tree_free(oz_tree* t);
void oz_tree_reset(oz_tree* t);
/* noise: L doubles indexed by position in the legal list (ignored unless use_dirichlet) */
void oz_tree_search(oz_tree* t, const oz_state* root, oz_eval_fn fn, void* user, const double* noise,
                    int64_t* counts_out /* num_actions */);
void oz_tree_update_root(oz_tree* t, int action);
int64_t oz_tree_root_n(const oz_tree* t);
double oz_tree_root_q(const oz_tree* t);
/* per-action root child stats; absent children: N=-1 */
void oz_tree_root_children(const oz_tree* t, int64_t* n_out, double* q_out, double* p_out);
double oz_tree_value_a0c(const oz_tree* t);
double oz_tree_value_offpolicy(const oz_tree* t);
/* measured means for the roofline: [sims, sum depth, sum children read, expansions,
   sum legal at expand, terminal sims, root evals, nodes] */
void oz_tree_counters(const oz_tree* t, uint64_t out[8]);

/* ---- whole self-play games with counter-mode noise/sampling (parity with the CUDA engine) ---- */
typedef struct {
  int32_t game, rows, cols;
  int32_t n_playouts;
  double c_puct, dirichlet_ratio;
  int32_t use_dirichlet;     /* 0 none, 2 counter-uniform noise */
  int32_t sample_moves;      /* 1: counter-mode proportional sampling while ply < num_prob; 0: argmax */
  int32_t num_probabilistic_actions;
  int32_t keep_tree;
  int32_t eval_kind, eval_shift;
  uint64_t seed;
  int32_t start_random_plies_mod; /* 0: initial position; else k = counter % mod random plies (bench synthetic starts) */
  int32_t max_plies;         /* stop after this many plies (0 = play to the end) */
} oz_selfplay_cfg;

typedef struct {
  int32_t tree, game_seq, ply, action, n_legal, player;
  uint64_t bb[2];
  double root_q, v_a0c, v_offpolicy;
  int64_t root_n;
  int32_t counts[OZ_MAX_LEGAL]; /* child visit counts in legal (ascending action) order */
} oz_ply_record;

/* plays one game for tree id `tree`, game_seq `game_seq`; returns number of plies recorded; returns[0..1] filled */
int oz_selfplay_game(const oz_selfplay_cfg* cfg, uint64_t tree, uint64_t game_seq, oz_ply_record* out, int max_out,
                     double returns_out[2], uint64_t counters_out[8]);
/* the synthetic start position used by bench / az_reset_random */
void oz_start_position(const oz_selfplay_cfg* cfg, uint64_t tree, uint64_t game_seq, oz_state* out);

#ifdef __cplusplus
}
#endif
#endif
