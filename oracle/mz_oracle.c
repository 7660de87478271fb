/*
 * mz_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A plain-C restatement of the arithmetic of the reference's search-and-target hot path
 * (JimOhman/model-based-rl), used only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the CHECKER for the CUDA kernels.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * file is pinned against outputs of the UNMODIFIED reference modules imported in the build
 * container (tests/golden/make_golden.py -> tests/golden/ .npz files, checked by
 * tests/test_oracle_golden.py).  It follows the reference as executed by CPython 3.12 +
 * numpy 2.3 + torch 2.11 on x86-64/glibc 2.39 (that matters: builtin sum() is Neumaier-compensated
 * since 3.12, and numpy 2 keeps python-float + np.float32 in float32).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).  All
 * search arithmetic is IEEE binary64 in the reference's operation order; exp/log are libm's, which
 * is what math.exp/math.log call.
 *
 * Layout: the tree is "edge centric".  The reference keeps statistics in child Node objects
 * (mcts.py:28-37); child Node (n, a) <-> edge record (n, a) here, the root Node <-> root_* scalars.
 * Nodes are numbered in expansion order: root = 0, the node expanded by simulation s = s + 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_A 64

/* ------------------------------------------------------------------------------------------ */
/* Batch-invariant fake network ("hashnet"): pure integer function of (state, action), so the   */
/* reference (B=1 calls), this oracle and the CUDA engine (B=G calls) see identical fp32 outputs */
/* (SURVEY.md section 4, item 2).  Mirrored in model-based-rl_b200/testing.py.                   */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

typedef struct {
  float value_scale, reward_scale, logit_scale;
  int32_t reward_density; /* reward != 0 iff (bits & 7) < reward_density */
} orc_hashnet;

static inline float hash_unit(uint64_t o) { /* exact: 24-bit integer times 2^-24, then 2u-1 */
  float u = (float)(o >> 40) * 0x1p-24f;
  return 2.0f * u - 1.0f;
}

uint64_t orc_hashnet_next(uint64_t state, int32_t action) {
  return splitmix64(state ^ ((uint64_t)(action + 1) * 0xD6E8FEB86659FD93ULL));
}

/* outputs of the node whose state is `state`: value, reward, logits[A] */
void orc_hashnet_outputs(const orc_hashnet* hn, uint64_t state, int32_t A, float* value,
                         float* reward, float* logits) {
  uint64_t o0 = splitmix64(state + 1ULL * 0xA0761D6478BD642FULL);
  uint64_t o1 = splitmix64(state + 2ULL * 0xA0761D6478BD642FULL);
  *value = hash_unit(o0) * hn->value_scale;
  *reward = ((int32_t)(o1 & 7) < hn->reward_density) ? hash_unit(o1) * hn->reward_scale : 0.0f;
  for (int a = 0; a < A; ++a) {
    uint64_t o = splitmix64(state + (uint64_t)(a + 3) * 0xA0761D6478BD642FULL);
    logits[a] = hash_unit(o) * hn->logit_scale;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* builtin sum() of python floats starting from int 0 (mcts.py:53 `sum(policy.values())`).      */
/* mode 1: CPython >= 3.12 (Neumaier, Python/bltinmodule.c), mode 0: older CPython (plain).      */
/* ------------------------------------------------------------------------------------------ */
double orc_py_sum(const double* x, int n, int mode) {
  if (n == 0) return 0.0;
  if (mode == 0) {
    double s = x[0]; /* int 0 + x0 == x0 */
    for (int i = 1; i < n; ++i) s += x[i];
    return s;
  }
  double f = x[0], c = 0.0;
  for (int i = 1; i < n; ++i) {
    double v = x[i];
    double t = f + v;
    if (fabs(f) >= fabs(v)) c += (f - t) + v;
    else c += (v - t) + f;
    f = t;
  }
  if (c != 0.0 && isfinite(c)) f += c;
  return f;
}

/* ------------------------------------------------------------------------------------------ */
/* Search                                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t S;            /* num_simulations        mcts.py:66 */
  int32_t A;            /* action_space           mcts.py:72 */
  int32_t two_players;  /*                        mcts.py:73 */
  int32_t sum_mode;     /* see orc_py_sum */
  double discount;      /*                        mcts.py:67 */
  double pb_c_base;     /*                        mcts.py:68 */
  double pb_c_init;     /*                        mcts.py:69 */
  double init_value_score; /*                     mcts.py:70 */
  double min_bound;     /* +inf when known_bounds[0] is None (mcts.py:9, 24) */
  double max_bound;     /* -inf when known_bounds[1] is None */
} orc_search_cfg;

typedef struct {
  /* per-game edge records [S+1][A] */
  double* prior;
  double* vsum;
  int32_t* visit;
  double* reward;
  int32_t* child; /* -1 unexpanded child, -2 no such child (illegal at the root), >=0 node id */
  double mn, mx;  /* MinMaxStats mcts.py:6-25 */
  double root_vsum;
  int32_t root_visit;
} orc_tree;

/* MinMaxStats.normalize mcts.py:16-21 */
static inline double mm_normalize(const orc_tree* t, double v) {
  if (t->mx > t->mn) return (v - t->mn) / (t->mx - t->mn);
  else if (t->mx == t->mn) return 1.0;
  return v;
}
/* MinMaxStats.update mcts.py:12-14 (python min/max keep the first argument on ties/NaN) */
static inline void mm_update(orc_tree* t, double v) {
  if (v < t->mn) t->mn = v;
  if (v > t->mx) t->mx = v;
}

/* Node.expand priors mcts.py:52-55: exp(logit)/sum over `actions` in order, python floats */
static void expand_priors(const orc_search_cfg* c, const float* logits, uint32_t legal_mask,
                          double* prior, int32_t* child, double* vsum, int32_t* visit,
                          double* reward) {
  double p[ORC_MAX_A];
  double dense[ORC_MAX_A];
  int n = 0;
  for (int a = 0; a < c->A; ++a) {
    if (c->A <= 32 && !((legal_mask >> a) & 1u)) continue;
    p[a] = exp((double)logits[a]);
    dense[n++] = p[a];
  }
  double s = orc_py_sum(dense, n, c->sum_mode);
  for (int a = 0; a < c->A; ++a) {
    vsum[a] = 0.0;
    visit[a] = 0;
    reward[a] = 0.0;
    if (c->A <= 32 && !((legal_mask >> a) & 1u)) {
      prior[a] = 0.0;
      child[a] = -2;
    } else {
      prior[a] = p[a] / s;
      child[a] = -1;
    }
  }
}

/* MCTS.ucb_score mcts.py:115-124 */
static inline double ucb_score(const orc_search_cfg* c, const orc_tree* t, int32_t N, double prior,
                               int32_t n, double vsum, double reward) {
  double pb_c = log(((double)N + c->pb_c_base + 1.0) / c->pb_c_base) + c->pb_c_init;
  pb_c *= sqrt((double)N) / (double)(n + 1);
  double prior_score = pb_c * prior;
  double value_score;
  if (n > 0) {
    double value = vsum / (double)n;
    if (c->two_players) value = -value;
    value_score = mm_normalize(t, reward + c->discount * value);
  } else {
    value_score = c->init_value_score;
  }
  return prior_score + value_score;
}

/*
 * Runs MCTS.run (mcts.py:78-102) for G independent games, including the root set-up the callers
 * do first (actors.py:139-143: root.expand over legal actions + add_exploration_noise).
 *
 * net_mode 0: hashnet; root_state[G] gives the root's integer state.
 * net_mode 1: replay; rec_value[G][S], rec_reward[G][S], rec_logits[G][S][A] are the outputs the
 *             network produced for simulation s of game g (recorded from the engine under test).
 *
 * Optional outputs may be NULL.
 */
int orc_search(const orc_search_cfg* c, int32_t G, const float* root_logits /*[G][A]*/,
               const uint32_t* legal_mask /*[G] or NULL*/, const double* noise /*[G][A] dense over legal, or NULL*/,
               double noise_frac, const int8_t* root_to_play /*[G] or NULL (=1)*/, int32_t net_mode,
               const orc_hashnet* hn, const uint64_t* root_state, const float* rec_value,
               const float* rec_reward, const float* rec_logits,
               /* outputs */
               int32_t* out_visits /*[G][A]*/, double* out_root_value /*[G]*/,
               double* out_root_vsum /*[G]*/, double* out_minmax /*[G][2]*/,
               int32_t* trace_parent /*[G][S]*/, int32_t* trace_action /*[G][S]*/,
               int32_t* trace_depth /*[G][S]*/, double* out_prior /*[G][S+1][A]*/,
               double* out_vsum /*[G][S+1][A]*/, int32_t* out_visit /*[G][S+1][A]*/,
               double* out_reward /*[G][S+1][A]*/, int32_t* out_child /*[G][S+1][A]*/) {
  const int S = c->S, A = c->A, NN = S + 1;
  if (A > ORC_MAX_A || A < 1 || S < 0) return -1;
  if (legal_mask && A > 32) return -1;
  orc_tree t;
  t.prior = (double*)malloc(sizeof(double) * NN * A);
  t.vsum = (double*)malloc(sizeof(double) * NN * A);
  t.visit = (int32_t*)malloc(sizeof(int32_t) * NN * A);
  t.reward = (double*)malloc(sizeof(double) * NN * A);
  t.child = (int32_t*)malloc(sizeof(int32_t) * NN * A);
  uint64_t* state = (uint64_t*)malloc(sizeof(uint64_t) * NN);
  int32_t* path_node = (int32_t*)malloc(sizeof(int32_t) * (NN + 1));
  int32_t* path_act = (int32_t*)malloc(sizeof(int32_t) * (NN + 1));
  float logits[ORC_MAX_A];

  for (int g = 0; g < G; ++g) {
    const uint32_t lm = legal_mask ? legal_mask[g] : 0xFFFFFFFFu;
    const int root_tp = root_to_play ? (int)root_to_play[g] : 1;
    /* root.expand (actors.py:142 -> mcts.py:47-55) */
    expand_priors(c, root_logits + (size_t)g * A, lm, t.prior, t.child, t.vsum, t.visit, t.reward);
    /* add_exploration_noise mcts.py:57-61: noise is dense over the root's children in order */
    if (noise) {
      int j = 0;
      for (int a = 0; a < A; ++a) {
        if (t.child[a] == -2) continue;
        double n = noise[(size_t)g * A + j++];
        t.prior[a] = t.prior[a] * (1.0 - noise_frac) + n * noise_frac;
      }
    }
    t.mn = c->min_bound; /* min_max_stats.reset mcts.py:79 */
    t.mx = c->max_bound;
    t.root_vsum = 0.0;
    t.root_visit = 0;
    if (net_mode == 0) state[0] = root_state[g];

    for (int s = 0; s < S; ++s) {
      int node = 0, depth = 0, tp = root_tp, N = t.root_visit, act = -1;
      path_node[0] = 0;
      for (;;) { /* while node.expanded(): select_child mcts.py:87-92, 104-113 */
        const double* pr = t.prior + (size_t)node * A;
        const double* vs = t.vsum + (size_t)node * A;
        const int32_t* vc = t.visit + (size_t)node * A;
        const double* rw = t.reward + (size_t)node * A;
        const int32_t* ch = t.child + (size_t)node * A;
        int best = -1;
        double best_score = 0.0;
        for (int a = 0; a < A; ++a) {
          if (ch[a] == -2) continue;
          double sc = (N == 0) ? pr[a] : ucb_score(c, &t, N, pr[a], vc[a], vs[a], rw[a]);
          /* max over (score, action, child) tuples: later action wins ties */
          if (best < 0 || sc > best_score || sc == best_score) {
            best = a;
            best_score = sc;
          }
        }
        act = best;
        depth++;
        path_act[depth] = act;
        if (c->two_players) tp = -tp;
        int chn = ch[act];
        if (chn < 0) break; /* child not expanded -> leaf */
        N = vc[act];
        node = chn;
        path_node[depth] = node;
      }
      const int parent = node, newn = s + 1;
      path_node[depth] = newn;
      if (trace_parent) trace_parent[(size_t)g * S + s] = parent;
      if (trace_action) trace_action[(size_t)g * S + s] = act;
      if (trace_depth) trace_depth[(size_t)g * S + s] = depth;

      float value_f, reward_f;
      if (net_mode == 0) { /* recurrent_inference(parent.hidden_state, [action]) mcts.py:96 */
        state[newn] = orc_hashnet_next(state[parent], act);
        orc_hashnet_outputs(hn, state[newn], A, &value_f, &reward_f, logits);
      } else {
        value_f = rec_value[(size_t)g * S + s];
        reward_f = rec_reward[(size_t)g * S + s];
        memcpy(logits, rec_logits + ((size_t)g * S + s) * A, sizeof(float) * A);
      }
      /* node.expand(network_output, to_play, range(A)) mcts.py:97, 47-55 */
      t.child[(size_t)parent * A + act] = newn;
      if (reward_f != 0.0f) t.reward[(size_t)parent * A + act] = (double)reward_f;
      expand_priors(c, logits, 0xFFFFFFFFu, t.prior + (size_t)newn * A, t.child + (size_t)newn * A,
                    t.vsum + (size_t)newn * A, t.visit + (size_t)newn * A,
                    t.reward + (size_t)newn * A);

      /* backpropagate(search_path, value.item(), to_play) mcts.py:126-143 */
      double value = (double)value_f;
      for (int k = depth; k >= 0; --k) {
        const int node_tp = c->two_players ? ((k & 1) ? -root_tp : root_tp) : root_tp;
        const int same = (node_tp == tp);
        double node_reward;
        if (k > 0) {
          size_t e = (size_t)path_node[k - 1] * A + path_act[k];
          t.vsum[e] += same ? value : -value;
          t.visit[e] += 1;
          node_reward = t.reward[e];
          double nv = t.vsum[e] / (double)t.visit[e];
          double new_q = c->two_players ? node_reward - c->discount * nv
                                        : node_reward + c->discount * nv;
          mm_update(&t, new_q);
        } else {
          t.root_vsum += same ? value : -value;
          t.root_visit += 1;
          node_reward = 0.0;
        }
        double r = (c->two_players && same) ? -node_reward : node_reward;
        value = r + c->discount * value;
      }
    }

    if (out_visits)
      for (int a = 0; a < A; ++a) out_visits[(size_t)g * A + a] = t.visit[a];
    if (out_root_value) /* Node.value mcts.py:42-45 */
      out_root_value[g] = t.root_visit == 0 ? 0.0 : t.root_vsum / (double)t.root_visit;
    if (out_root_vsum) out_root_vsum[g] = t.root_vsum;
    if (out_minmax) {
      out_minmax[2 * g] = t.mn;
      out_minmax[2 * g + 1] = t.mx;
    }
    size_t off = (size_t)g * NN * A;
    if (out_prior) memcpy(out_prior + off, t.prior, sizeof(double) * NN * A);
    if (out_vsum) memcpy(out_vsum + off, t.vsum, sizeof(double) * NN * A);
    if (out_visit) memcpy(out_visit + off, t.visit, sizeof(int32_t) * NN * A);
    if (out_reward) memcpy(out_reward + off, t.reward, sizeof(double) * NN * A);
    if (out_child) memcpy(out_child + off, t.child, sizeof(int32_t) * NN * A);
  }
  free(t.prior); free(t.vsum); free(t.visit); free(t.reward); free(t.child);
  free(state); free(path_node); free(path_act);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Config.select_action config.py:70-81 with host-supplied uniforms instead of np.random.         */
/*   T > 0 : p = visits**(1/T) / sum  (numpy float64, pairwise sum), np.random.choice(p=p) =      */
/*           searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right')                               */
/*   T == 0: uniform choice among argmax ties; deterministic tie-break = floor(u * n_ties)        */
/* `visits` is dense over the root's children (length n); returns an index into that list.       */
/* ------------------------------------------------------------------------------------------ */
static double np_pairwise_sum(const double* a, int n) { /* numpy DOUBLE_pairwise_sum, n <= 128 */
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i;
  for (i = 8; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}

int32_t orc_select_action(const int32_t* visits, int n, double temperature, double u) {
  if (n <= 0 || n > 128) return -1;
  if (temperature != 0.0) {
    double d[128], cdf[128];
    double inv_t = 1.0 / temperature;
    for (int i = 0; i < n; ++i) d[i] = pow((double)visits[i], inv_t);
    double s = np_pairwise_sum(d, n);
    for (int i = 0; i < n; ++i) d[i] = d[i] / s;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { /* cumsum is sequential */
      acc = (i == 0) ? d[0] : acc + d[i];
      cdf[i] = acc;
    }
    double last = cdf[n - 1];
    int idx = 0;
    for (int i = 0; i < n; ++i) { /* searchsorted side='right': count of cdf <= u */
      if (cdf[i] / last <= u) idx = i + 1;
    }
    if (idx >= n) idx = n - 1; /* cannot happen for u < 1 */
    return idx;
  }
  int32_t mx = visits[0];
  for (int i = 1; i < n; ++i) if (visits[i] > mx) mx = visits[i];
  int ties = 0;
  for (int i = 0; i < n; ++i) if (visits[i] == mx) ties++;
  int pick = (int)floor(u * (double)ties);
  if (pick >= ties) pick = ties - 1;
  for (int i = 0; i < n; ++i)
    if (visits[i] == mx && pick-- == 0) return i;
  return -1;
}

/* Game.store_search_statistics game.py:106-111: child_visits[a] = visit_a / sum_visits */
void orc_child_visits(const int32_t* visits, const int32_t* child /*root child[] or NULL*/, int A,
                      double* out) {
  long sum = 0;
  for (int a = 0; a < A; ++a) if (!child || child[a] != -2) sum += visits[a];
  for (int a = 0; a < A; ++a)
    out[a] = (!child || child[a] != -2) ? (double)visits[a] / (double)sum : 0.0;
}

/* ------------------------------------------------------------------------------------------ */
/* Targets: PrioritizedReplay.insert_target replay_buffer.py:165-198 for one sampled (chunk, step) */
/* chunk arrays: rewards[n_rewards] (python floats -> double), to_play[n_rewards],                */
/* root_values[n_values] double, child_visits[n_values][A] double.                                */
/* discounts_f32[K+T] = float32(discount**n) (replay_buffer.py:84), disc_pow_td = discount**T.    */
/* value arithmetic as executed under numpy 2: bootstrap (python float) + np.dot(f32,f32) happens */
/* in float32; np.dot's internal order is BLAS-defined -> accumulate exactly (double) and round.  */
/* ------------------------------------------------------------------------------------------ */
void orc_insert_target(const double* rewards, const int8_t* to_play, int n_rewards,
                       const double* root_values, const double* child_visits, int n_values, int A,
                       int K, int T, double disc_pow_td, const float* discounts_f32, int step,
                       float* t_rewards /*[K+1]*/, float* t_values /*[K+1]*/,
                       float* t_policies /*[K+1][A]*/) {
  const int end_index = n_values;
  for (int i = 0; i <= K; ++i) {
    const int ci = step + i;
    double last_reward = (ci > 0 && ci <= n_rewards) ? rewards[ci - 1] : 0.0;
    if (ci < end_index) {
      const int tp = to_play[ci];
      const int bi = ci + T;
      double boot = (bi < end_index) ? root_values[bi] * disc_pow_td : 0.0;
      int hi = bi < n_rewards ? bi : n_rewards;
      float value;
      if (hi > ci) {
        double acc = 0.0;
        for (int j = ci; j < hi; ++j) {
          float r = (float)rewards[j];
          if (to_play[j] != tp) r = -r;
          acc += (double)r * (double)discounts_f32[j - ci];
        }
        value = (float)boot + (float)acc; /* float32 + float32 (numpy 2 weak python scalar) */
      } else {
        value = (float)boot;
      }
      for (int a = 0; a < A; ++a) t_policies[i * A + a] = (float)child_visits[(size_t)ci * A + a];
      t_rewards[i] = (float)last_reward;
      t_values[i] = value;
    } else {
      for (int a = 0; a < A; ++a) t_policies[i * A + a] = 0.0f;
      t_rewards[i] = (float)last_reward;
      t_values[i] = 0.0f;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Scalar transforms, float32 in torch's op order (each op rounds to float32).                  */
/* ------------------------------------------------------------------------------------------ */
static inline float sgnf(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

/* Config.scalar_transform config.py:51-54: sign(x)*(sqrt(|x|+1)-1) + 0.001*x */
float orc_scalar_transform(float x) {
  volatile float a = fabsf(x) + 1.0f;
  volatile float b = sqrtf(a);
  volatile float c = b - 1.0f;
  volatile float d = sgnf(x) * c;
  volatile float e = 0.001f * x;
  return d + e;
}

/* h^-1 part of Config.inverse_transform config.py:31-32 */
float orc_inverse_scalar_transform(float v) {
  volatile float a = fabsf(v) + 1.0f;
  volatile float b = a + 0.001f;
  volatile float c = 0.004f * b; /* python 4*0.001 == 0.004 -> float32 scalar */
  volatile float d = 1.0f + c;
  volatile float e = sqrtf(d);
  volatile float f = e - 1.0f;
  volatile float g = f / 0.002f; /* python 2*0.001 == 0.002 */
  volatile float h = g * g;
  volatile float k = h - 1.0f;
  return sgnf(v) * k;
}

/* Config.scalar_to_support config.py:56-68 for one scalar -> size bins (zeroed by the callee) */
void orc_scalar_to_support(float x, int mn, int mx, float* out /*[mx-mn+1]*/) {
  int size = mx - mn + 1;
  for (int i = 0; i < size; ++i) out[i] = 0.0f;
  if (x < (float)mn) x = (float)mn;
  if (x > (float)mx) x = (float)mx;
  float lo = floorf(x), hi = ceilf(x);
  float p_high = x - lo;
  float p_low = 1.0f - p_high;
  out[(int)(hi - (float)mn)] = p_high; /* scatter high first ... */
  out[(int)(lo - (float)mn)] = p_low;  /* ... low second: integer x ends with 1.0 */
}

/* Config.inverse_transform config.py:27-33: softmax(logits) . support, then h^-1.
 * The expectation is evaluated in double (torch's vectorised float32 softmax order is not
 * reproducible); parity for this function is therefore a tolerance, see tests. */
float orc_inverse_transform(const float* logits, int mn, int mx, int no_target_transform) {
  int size = mx - mn + 1;
  double m = logits[0];
  for (int i = 1; i < size; ++i) if (logits[i] > m) m = logits[i];
  double den = 0.0, num = 0.0;
  for (int i = 0; i < size; ++i) {
    double e = exp((double)logits[i] - m);
    den += e;
    num += e * (double)(mn + i);
  }
  float v = (float)(num / den);
  return no_target_transform ? v : orc_inverse_scalar_transform(v);
}
