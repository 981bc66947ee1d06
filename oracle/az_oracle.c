/*
 * az_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See az_oracle.h.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  -ffp-contract=off is
 * mandatory: the reference's Python floats never fuse a multiply-add.
 */
#include "az_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================================
 * Games.  OpenSpiel connect_four.cc / breakthrough.cc semantics.
 * =================================================================================== */

/* breakthrough.cc direction tables: dirs 0-2 for player 0 (black, +row),
 * dirs 3-5 for player 1 (white, -row). */
static const int kDirRow[6] = {1, 1, 1, -1, -1, -1};
static const int kDirCol[6] = {-1, 0, 1, -1, 0, 1};

void oz_init(oz_state* s, int game, int rows, int cols) {
  memset(s, 0, sizeof(*s));
  s->game = game;
  s->winner = -1;
  if (game == OZ_GAME_CONNECT_FOUR) {
    s->rows = 6;
    s->cols = 7;
  } else {
    s->rows = rows;
    s->cols = cols;
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) {
        if (r < 2) s->cell[r * cols + c] = 1;              /* black = player 0, rows 0,1 */
        else if (r >= rows - 2) s->cell[r * cols + c] = 2; /* white = player 1 */
      }
    s->pieces[0] = s->pieces[1] = 2 * cols;
  }
}

int oz_num_actions(const oz_state* s) {
  return s->game == OZ_GAME_CONNECT_FOUR ? 7 : s->rows * s->cols * 6 * 2;
}

int oz_terminal(const oz_state* s) {
  if (s->winner >= 0) return 1;
  if (s->game == OZ_GAME_BREAKTHROUGH) return s->pieces[0] == 0 || s->pieces[1] == 0;
  return 0;
}

int oz_current_player(const oz_state* s) { return oz_terminal(s) ? OZ_TERMINAL_PLAYER : s->player; }

int oz_legal(const oz_state* s, int32_t* out) {
  int n = 0;
  if (oz_terminal(s)) return 0;
  if (s->game == OZ_GAME_CONNECT_FOUR) {
    for (int c = 0; c < 7; ++c)
      if (s->cell[5 * 7 + c] == 0) out[n++] = c; /* top row empty <=> column playable */
    return n;
  }
  const int R = s->rows, C = s->cols;
  const int8_t me = (int8_t)(s->player + 1), opp = (int8_t)(2 - s->player);
  const int d0 = s->player == 0 ? 0 : 3;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      if (s->cell[r * C + c] != me) continue;
      for (int d = d0; d < d0 + 3; ++d) {
        const int r2 = r + kDirRow[d], c2 = c + kDirCol[d];
        if (r2 < 0 || r2 >= R || c2 < 0 || c2 >= C) continue;
        const int8_t t = s->cell[r2 * C + c2];
        if (t == 0) out[n++] = ((r * C + c) * 6 + d) * 2;                          /* plain move */
        else if (t == opp && kDirCol[d] != 0) out[n++] = ((r * C + c) * 6 + d) * 2 + 1; /* diagonal capture */
      }
    }
  return n;
}

static int c4_line_through(const oz_state* s, int r, int c) {
  static const int dr[4] = {0, 1, 1, 1}, dc[4] = {1, 0, 1, -1};
  const int8_t v = s->cell[r * 7 + c];
  for (int k = 0; k < 4; ++k) {
    int run = 1;
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      int rr = r + sgn * dr[k], cc = c + sgn * dc[k];
      while (rr >= 0 && rr < 6 && cc >= 0 && cc < 7 && s->cell[rr * 7 + cc] == v) {
        ++run;
        rr += sgn * dr[k];
        cc += sgn * dc[k];
      }
    }
    if (run >= 4) return 1;
  }
  return 0;
}

int oz_apply(oz_state* s, int action) {
  if (oz_terminal(s)) return -1;
  if (s->game == OZ_GAME_CONNECT_FOUR) {
    if (action < 0 || action >= 7 || s->cell[5 * 7 + action] != 0) return -1;
    int r = 0;
    while (s->cell[r * 7 + action] != 0) ++r; /* lowest empty row, row 0 = bottom */
    s->cell[r * 7 + action] = (int8_t)(s->player + 1);
    s->ply++;
    if (c4_line_through(s, r, action)) s->winner = s->player;
    else if (s->ply == 42) s->winner = 2;
    s->player ^= 1;
    return 0;
  }
  const int R = s->rows, C = s->cols;
  if (action < 0 || action >= R * C * 12) return -1;
  const int cap = action & 1, d = (action >> 1) % 6, cellidx = (action >> 1) / 6;
  const int r = cellidx / C, c = cellidx % C;
  const int8_t me = (int8_t)(s->player + 1), opp = (int8_t)(2 - s->player);
  if (s->cell[cellidx] != me) return -1;
  if ((s->player == 0) != (d < 3)) return -1;
  const int r2 = r + kDirRow[d], c2 = c + kDirCol[d];
  if (r2 < 0 || r2 >= R || c2 < 0 || c2 >= C) return -1;
  const int8_t t = s->cell[r2 * C + c2];
  if (cap) {
    if (t != opp || kDirCol[d] == 0) return -1;
    s->pieces[1 - s->player]--;
  } else if (t != 0) {
    return -1;
  }
  s->cell[r2 * C + c2] = me;
  s->cell[cellidx] = 0;
  if (s->player == 0 && r2 == R - 1) s->winner = 0;
  else if (s->player == 1 && r2 == 0) s->winner = 1;
  s->player ^= 1;
  s->ply++;
  return 0;
}

void oz_returns(const oz_state* s, double out[2]) {
  out[0] = out[1] = 0.0;
  int w = s->winner;
  if (w < 0 && s->game == OZ_GAME_BREAKTHROUGH) {
    if (s->pieces[0] == 0) w = 1;
    else if (s->pieces[1] == 0) w = 0;
  }
  if (w == 0) { out[0] = 1.0; out[1] = -1.0; }
  else if (w == 1) { out[0] = -1.0; out[1] = 1.0; }
}

void oz_normalized_vector(const oz_state* s, float* out) {
  const int n = s->rows * s->cols;
  memset(out, 0, sizeof(float) * 3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    int plane;
    if (s->game == OZ_GAME_CONNECT_FOUR) {
      /* CellState {kEmpty=0, kNought=1 (player 1 'o'), kCross=2 (player 0 'x')}  (B.2) */
      plane = s->cell[i] == 0 ? 0 : (s->cell[i] == 2 ? 1 : 2);
    } else {
      /* plane 0 black (player 0), plane 1 white (player 1), plane 2 empty  (B.3) */
      plane = s->cell[i] == 1 ? 0 : (s->cell[i] == 2 ? 1 : 2);
    }
    out[plane * n + i] = 1.0f;
  }
}

void oz_board(const oz_state* s, double* out) {
  const int n = s->rows * s->cols;
  float tmp[3 * OZ_MAX_CELLS];
  oz_normalized_vector(s, tmp);
  const double cp = (double)oz_current_player(s);
  for (int i = 0; i < 3 * n; ++i) out[i] = (double)tmp[i];
  for (int i = 0; i < n; ++i) out[3 * n + i] = cp;
}

void oz_bitboards(const oz_state* s, uint64_t out[2]) {
  out[0] = out[1] = 0;
  const int n = s->rows * s->cols;
  for (int i = 0; i < n; ++i)
    if (s->cell[i]) out[s->cell[i] - 1] |= 1ULL << i;
}

/* ===================================================================================
 * Counter-based hash stream + synthetic evaluator (definition shared with the engine:
 * include/az_b200.h "Deterministic counter streams").
 * =================================================================================== */

uint64_t oz_mix64(uint64_t x) { /* splitmix64 finaliser */
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

uint64_t oz_counter(uint64_t seed, uint64_t tree, uint64_t game_seq, uint64_t ply, uint64_t idx, uint64_t stream) {
  uint64_t x = oz_mix64(seed ^ 0xA0761D3C5F2E8B49ULL);
  x = oz_mix64(x + tree);
  x = oz_mix64(x + game_seq);
  x = oz_mix64(x + ply);
  x = oz_mix64(x + idx);
  x = oz_mix64(x + stream);
  return x;
}

void oz_synth_eval(const oz_state* s, int kind, uint64_t seed, int shift, double* priors, double* value) {
  const int A = oz_num_actions(s);
  if (kind == 0) {
    for (int a = 0; a < A; ++a) priors[a] = 1.0 / (double)A;
    *value = 0.0;
    return;
  }
  uint64_t bb[2];
  oz_bitboards(s, bb);
  const uint64_t k = oz_mix64(bb[0] ^ oz_mix64(bb[1] ^ oz_mix64(seed + (uint64_t)s->player)));
  if (kind == 2) {
    /* MCTS.random_rollout (mcts.py:205-223): priors = ones, value = player_return(mover) of one uniformly random playout;
     * the "random" picks are the counter hash r_j = mix64(k + (j+1)*golden) % n_legal, a pure function of the position */
    for (int a = 0; a < A; ++a) priors[a] = 1.0;
    oz_state t = *s;
    const int mover = s->player;
    for (uint64_t j = 0; !oz_terminal(&t); ++j) {
      int32_t legal[OZ_MAX_LEGAL];
      const int n = oz_legal(&t, legal);
      const uint64_t r = oz_mix64(k + (j + 1) * 0x9E3779B97F4A7C15ULL);
      oz_apply(&t, legal[r % (uint64_t)n]);
    }
    double ret[2];
    oz_returns(&t, ret);
    *value = ret[mover];
    return;
  }
  const double scale = ldexp(1.0, -(10 + shift));
  for (int a = 0; a < A; ++a) {
    const uint64_t ha = oz_mix64(k + (uint64_t)(a + 1) * 0xD1B54A32D192ED03ULL);
    priors[a] = (double)(1 + (int)((ha >> 40) & 0x3FF)) * scale;
  }
  *value = (double)((int)((k >> 20) & 31) - 16) / 16.0;
}

/* Batched form over canonical bitboards (what az_request_info returns): row i gets the evaluator outputs of the position
 * (bb[2i], bb[2i+1], side to move = ply[i] & 1) as fp32 -- the values are small dyadic rationals, exact in fp32.  Rows with
 * ply[i] < 0 (no request pending) are left untouched.  Used by the AZ_EVAL_EXTERNAL parity test. */
void oz_synth_eval_bb(int num_actions, int n, const uint64_t* bb, const int32_t* ply, int kind, uint64_t seed, int shift,
                      float* priors, float* values) {
  const double scale = ldexp(1.0, -(10 + shift));
  for (int i = 0; i < n; ++i) {
    if (ply[i] < 0) continue;
    float* pr = priors + (size_t)i * (size_t)num_actions;
    if (kind == 0) {
      for (int a = 0; a < num_actions; ++a) pr[a] = (float)(1.0 / (double)num_actions);
      values[i] = 0.0f;
      continue;
    }
    const uint64_t k = oz_mix64(bb[2 * i] ^ oz_mix64(bb[2 * i + 1] ^ oz_mix64(seed + (uint64_t)(ply[i] & 1))));
    for (int a = 0; a < num_actions; ++a) {
      const uint64_t ha = oz_mix64(k + (uint64_t)(a + 1) * 0xD1B54A32D192ED03ULL);
      pr[a] = (float)((double)(1 + (int)((ha >> 40) & 0x3FF)) * scale);
    }
    values[i] = (float)((double)((int)((k >> 20) & 31) - 16) / 16.0);
  }
}

/* ===================================================================================
 * MCTS (mcts.py).  Flat node arrays; children of a node are one contiguous block in
 * legal-action (ascending) order == the reference's dict insertion order (A.2).
 * =================================================================================== */

struct oz_tree {
  int num_actions, n_playouts, use_dirichlet;
  double c_puct, dirichlet_ratio;
  int64_t cap, size;
  int64_t* N;
  double* Q;
  double* P;
  int64_t* parent;
  int64_t* first;  /* first child index, 0 = no children (node 0 / roots are never children) */
  int32_t* nchild;
  int32_t* action;
  int64_t root;
  uint64_t ctr[8];
};

static void tree_grow(oz_tree* t, int64_t need) {
  if (t->size + need <= t->cap) return;
  int64_t nc = t->cap * 2;
  while (nc < t->size + need) nc *= 2;
  t->N = (int64_t*)realloc(t->N, sizeof(int64_t) * (size_t)nc);
  t->Q = (double*)realloc(t->Q, sizeof(double) * (size_t)nc);
  t->P = (double*)realloc(t->P, sizeof(double) * (size_t)nc);
  t->parent = (int64_t*)realloc(t->parent, sizeof(int64_t) * (size_t)nc);
  t->first = (int64_t*)realloc(t->first, sizeof(int64_t) * (size_t)nc);
  t->nchild = (int32_t*)realloc(t->nchild, sizeof(int32_t) * (size_t)nc);
  t->action = (int32_t*)realloc(t->action, sizeof(int32_t) * (size_t)nc);
  t->cap = nc;
}

static int64_t tree_new_node(oz_tree* t, int64_t parent, double prior, int action) {
  tree_grow(t, 1);
  const int64_t i = t->size++;
  t->N[i] = 0;
  t->Q[i] = 0.0;
  t->P[i] = prior;
  t->parent[i] = parent;
  t->first[i] = 0;
  t->nchild[i] = 0;
  t->action[i] = action;
  return i;
}

void oz_tree_reset(oz_tree* t) {
  t->size = 0;
  t->root = tree_new_node(t, -1, 0.0, -1); /* mcts.py:122 Node(None, 0.0) */
}

oz_tree* oz_tree_new(int num_actions, double c_puct, int n_playouts, int use_dirichlet, double dirichlet_ratio) {
  oz_tree* t = (oz_tree*)calloc(1, sizeof(oz_tree));
  t->num_actions = num_actions;
  t->c_puct = c_puct;
  t->n_playouts = n_playouts;
  t->use_dirichlet = use_dirichlet;
  t->dirichlet_ratio = dirichlet_ratio;
  t->cap = 1024;
  t->N = (int64_t*)malloc(sizeof(int64_t) * 1024);
  t->Q = (double*)malloc(sizeof(double) * 1024);
  t->P = (double*)malloc(sizeof(double) * 1024);
  t->parent = (int64_t*)malloc(sizeof(int64_t) * 1024);
  t->first = (int64_t*)malloc(sizeof(int64_t) * 1024);
  t->nchild = (int32_t*)malloc(sizeof(int32_t) * 1024);
  t->action = (int32_t*)malloc(sizeof(int32_t) * 1024);
  oz_tree_reset(t);
  return t;
}

void oz_tree_free(oz_tree* t) {
  if (!t) return;
  free(t->N); free(t->Q); free(t->P); free(t->parent); free(t->first); free(t->nchild); free(t->action);
  free(t);
}

/* mcts.py:54-66 */
static void tree_expand(oz_tree* t, int64_t node, const double* priors, const int32_t* legal, int n_legal) {
  if (t->nchild[node] > 0) { /* children exist: same legal set, only P is overwritten (mcts.py:65-66) */
    for (int i = 0; i < n_legal; ++i) t->P[t->first[node] + i] = priors[legal[i]];
    return;
  }
  if (n_legal == 0) return;
  tree_grow(t, n_legal);
  t->first[node] = t->size;
  t->nchild[node] = n_legal;
  for (int i = 0; i < n_legal; ++i) tree_new_node(t, node, priors[legal[i]], legal[i]);
}

/* mcts.py:38-52 + 68-80: argmax_a Q + ((c*P)*sqrt(Np))/(N+1), first max wins */
static int64_t tree_select(const oz_tree* t, int64_t node) {
  const double sq = sqrt((double)t->N[node]);
  int64_t best = -1;
  double bestv = 0.0;
  for (int i = 0; i < t->nchild[node]; ++i) {
    const int64_t c = t->first[node] + i;
    const double u = ((t->c_puct * t->P[c]) * sq) / (double)(t->N[c] + 1);
    const double v = t->Q[c] + u;
    if (best < 0 || v > bestv) { best = c; bestv = v; }
  }
  return best;
}

/* mcts.py:82-89 */
static void tree_backup(oz_tree* t, int64_t node, double value) {
  while (node >= 0) {
    t->Q[node] = ((double)t->N[node] * t->Q[node] + value) / (double)(t->N[node] + 1);
    t->N[node] += 1;
    value = -value;
    node = t->parent[node];
  }
}

/* mcts.py:126-153 */
static void tree_playout(oz_tree* t, oz_state* st, oz_eval_fn fn, void* user, double* priors) {
  int64_t node = t->root;
  int current_player = oz_current_player(st);
  int depth = 0;
  while (t->nchild[node] > 0 && !oz_terminal(st)) {
    current_player = oz_current_player(st);
    t->ctr[2] += (uint64_t)t->nchild[node];
    node = tree_select(t, node);
    oz_apply(st, t->action[node]);
    ++depth;
  }
  double leaf_value;
  if (!oz_terminal(st)) {
    int32_t legal[OZ_MAX_LEGAL];
    fn(st, priors, &leaf_value, user);
    const int n = oz_legal(st, legal);
    tree_expand(t, node, priors, legal, n);
    t->ctr[3] += 1;
    t->ctr[4] += (uint64_t)n;
  } else {
    double ret[2];
    oz_returns(st, ret);
    leaf_value = -ret[current_player];
    t->ctr[5] += 1;
  }
  t->ctr[0] += 1;
  t->ctr[1] += (uint64_t)depth;
  tree_backup(t, node, -leaf_value);
}

/* mcts.py:164-190 */
void oz_tree_search(oz_tree* t, const oz_state* root, oz_eval_fn fn, void* user, const double* noise,
                    int64_t* counts_out) {
  double* priors = (double*)malloc(sizeof(double) * (size_t)t->num_actions);
  if (t->use_dirichlet) {
    double v;
    int32_t legal[OZ_MAX_LEGAL];
    fn(root, priors, &v, user);
    const int n = oz_legal(root, legal);
    const double keep = 1.0 - t->dirichlet_ratio;
    for (int a = 0; a < t->num_actions; ++a) priors[a] = keep * priors[a];
    for (int i = 0; i < n; ++i) priors[legal[i]] = priors[legal[i]] + 0.25 * (noise ? noise[i] : 0.0);
    tree_expand(t, t->root, priors, legal, n);
    t->ctr[6] += 1;
  }
  for (int i = 0; i < t->n_playouts; ++i) {
    oz_state copy = *root;
    tree_playout(t, &copy, fn, user, priors);
  }
  if (counts_out) {
    for (int a = 0; a < t->num_actions; ++a) counts_out[a] = 0;
    for (int i = 0; i < t->nchild[t->root]; ++i) {
      const int64_t c = t->first[t->root] + i;
      counts_out[t->action[c]] = t->N[c];
    }
  }
  free(priors);
}

/* mcts.py:192-203 */
void oz_tree_update_root(oz_tree* t, int action) {
  if (t->nchild[t->root] == 0) {
    t->root = tree_new_node(t, -1, 0.0, -1);
    return;
  }
  for (int i = 0; i < t->nchild[t->root]; ++i) {
    const int64_t c = t->first[t->root] + i;
    if (t->action[c] == action) {
      t->root = c;
      t->parent[c] = -1;
      return;
    }
  }
  abort(); /* KeyError in the reference */
}

int64_t oz_tree_root_n(const oz_tree* t) { return t->N[t->root]; }
double oz_tree_root_q(const oz_tree* t) { return t->Q[t->root]; }

void oz_tree_root_children(const oz_tree* t, int64_t* n_out, double* q_out, double* p_out) {
  for (int a = 0; a < t->num_actions; ++a) {
    n_out[a] = -1;
    q_out[a] = 0.0;
    p_out[a] = 0.0;
  }
  for (int i = 0; i < t->nchild[t->root]; ++i) {
    const int64_t c = t->first[t->root] + i;
    n_out[t->action[c]] = t->N[c];
    q_out[t->action[c]] = t->Q[c];
    p_out[t->action[c]] = t->P[c];
  }
}

/* game_utils.py:177-179 */
double oz_tree_value_a0c(const oz_tree* t) {
  double best = -99.0;
  int have = 0;
  for (int i = 0; i < t->nchild[t->root]; ++i) {
    const int64_t c = t->first[t->root] + i;
    const double v = t->N[c] > 0 ? t->Q[c] : -99.0;
    if (!have || v > best) { best = v; have = 1; }
  }
  return best;
}

/* game_utils.py:182-194 */
double oz_tree_value_offpolicy(const oz_tree* t) {
  int64_t node = t->root;
  double value = 0.0, mult = 1.0;
  while (t->nchild[node] > 0) {
    value = t->Q[node];
    int64_t best = -1;
    double bestv = 0.0;
    for (int i = 0; i < t->nchild[node]; ++i) {
      const int64_t c = t->first[node] + i;
      const double v = t->N[c] > 0 ? (double)t->N[c] + t->P[c] : -99.0;
      if (best < 0 || v > bestv) { best = c; bestv = v; }
    }
    node = best;
    mult *= -1.0;
  }
  if (t->N[node] > 0) {
    value = t->Q[node];
    mult *= -1.0;
  }
  return value * mult;
}

void oz_tree_counters(const oz_tree* t, uint64_t out[8]) {
  memcpy(out, t->ctr, sizeof(t->ctr));
  out[7] = (uint64_t)t->size;
}

/* ===================================================================================
 * Counter-mode self-play (alphazerobot.py:42-93 + game_utils.py:148-206 with the random
 * draws replaced by the engine's deterministic counter streams; include/az_b200.h).
 * =================================================================================== */

typedef struct {
  int kind, shift;
  uint64_t seed;
} synth_user;

static void synth_cb(const oz_state* s, double* priors, double* value, void* user) {
  const synth_user* u = (const synth_user*)user;
  oz_synth_eval(s, u->kind, u->seed, u->shift, priors, value);
}

void oz_start_position(const oz_selfplay_cfg* cfg, uint64_t tree, uint64_t game_seq, oz_state* out) {
  oz_init(out, cfg->game, cfg->rows, cfg->cols);
  if (cfg->start_random_plies_mod <= 0) return;
  const int k = (int)(oz_counter(cfg->seed, tree, game_seq, 0, 0, 4) % (uint64_t)cfg->start_random_plies_mod);
  for (uint64_t attempt = 0;; ++attempt) {
    oz_init(out, cfg->game, cfg->rows, cfg->cols);
    int ok = 1;
    for (int j = 0; j < k; ++j) {
      int32_t legal[OZ_MAX_LEGAL];
      const int n = oz_legal(out, legal);
      oz_apply(out, legal[oz_counter(cfg->seed, tree, game_seq, (uint64_t)j, attempt, 3) % (uint64_t)n]);
      if (oz_terminal(out)) { ok = 0; break; }
    }
    if (ok) return;
  }
}

int oz_selfplay_game(const oz_selfplay_cfg* cfg, uint64_t tree, uint64_t game_seq, oz_ply_record* out, int max_out,
                     double returns_out[2], uint64_t counters_out[8]) {
  oz_state st;
  oz_start_position(cfg, tree, game_seq, &st);
  const int A = oz_num_actions(&st);
  oz_tree* t = oz_tree_new(A, cfg->c_puct, cfg->n_playouts, cfg->use_dirichlet != 0, cfg->dirichlet_ratio);
  synth_user su = {cfg->eval_kind, cfg->eval_shift, cfg->seed};
  int64_t* counts = (int64_t*)malloc(sizeof(int64_t) * (size_t)A);
  int n_rec = 0, last_action = -1, first = 1;
  while (!oz_terminal(&st)) {
    if (cfg->max_plies > 0 && n_rec >= cfg->max_plies) break;
    /* alphazerobot.py:53-64 */
    if (cfg->keep_tree) {
      if (!first) oz_tree_update_root(t, last_action);
    } else {
      oz_tree_reset(t);
    }
    first = 0;
    int32_t legal[OZ_MAX_LEGAL];
    const int L = oz_legal(&st, legal);
    double noise[OZ_MAX_LEGAL];
    if (cfg->use_dirichlet) { /* counter-uniform noise: eta_i = u_i / sum(u), u_i in (0,1] */
      double sum = 0.0;
      for (int i = 0; i < L; ++i) {
        noise[i] = (double)((oz_counter(cfg->seed, tree, game_seq, (uint64_t)st.ply, (uint64_t)i, 1) >> 11) + 1) *
                   (1.0 / 9007199254740992.0);
        sum += noise[i];
      }
      for (int i = 0; i < L; ++i) noise[i] = noise[i] / sum;
    }
    oz_tree_search(t, &st, synth_cb, &su, noise, counts);
    /* move choice: proportional to visit counts (temperature 1) or first-max */
    int action = -1;
    int64_t total = 0;
    for (int i = 0; i < L; ++i) total += counts[legal[i]];
    if (cfg->sample_moves && st.ply < cfg->num_probabilistic_actions) {
      const uint64_t u32 = oz_counter(cfg->seed, tree, game_seq, (uint64_t)st.ply, 0, 2) >> 32;
      const int64_t r = (int64_t)((u32 * (uint64_t)total) >> 32);
      int64_t cum = 0;
      for (int i = 0; i < L; ++i) {
        cum += counts[legal[i]];
        if (cum > r) { action = legal[i]; break; }
      }
    } else {
      int64_t best = -1;
      for (int i = 0; i < L; ++i)
        if (counts[legal[i]] > best) { best = counts[legal[i]]; action = legal[i]; }
    }
    if (n_rec < max_out) {
      oz_ply_record* r = &out[n_rec];
      memset(r, 0, sizeof(*r));
      r->tree = (int32_t)tree;
      r->game_seq = (int32_t)game_seq;
      r->ply = st.ply;
      r->action = action;
      r->n_legal = L;
      r->player = st.player;
      oz_bitboards(&st, r->bb);
      r->root_q = oz_tree_root_q(t);
      r->root_n = oz_tree_root_n(t);
      r->v_a0c = oz_tree_value_a0c(t);
      r->v_offpolicy = oz_tree_value_offpolicy(t);
      for (int i = 0; i < L; ++i) r->counts[i] = (int32_t)counts[legal[i]];
    }
    ++n_rec;
    oz_apply(&st, action);
    last_action = action;
  }
  oz_returns(&st, returns_out);
  if (counters_out) oz_tree_counters(t, counters_out);
  free(counts);
  oz_tree_free(t);
  return n_rec;
}
