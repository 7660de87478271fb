// Sparse, node-parallel tree engine of the persistent search kernel (mz_fcsearch.cu, cluster size 4).
//
// What bounds the search is the length of ONE game's dependent instruction stream per simulation (a warp
// issues ~1 dependent instruction per 10 cycles), so this engine shortens that stream instead of packing
// more games per warp:
//
//  * select_child (mcts.py:104-113) only has to rank the EXPANDED children of a node plus its best
//    unexpanded child: every unexpanded child scores pb_c(N, 0) * prior + init_value_score, monotone in the
//    prior, so with the priors' maximum known the others cannot win.  (Exactness: the reference breaks score
//    ties towards the larger action.  The stand-in is the largest action among the maximal priors; a smaller
//    prior could only tie after rounding if it lies within 2^-50 relative of the maximum -- such nodes, and
//    every node when init_value_score != 0, are flagged at expansion and ranked densely over all actions.)
//  * the argmax of EVERY expanded node is evaluated each simulation, all nodes in parallel over the eight
//    lanes of a game (two items per node: its best unexpanded child and the edge that leads to it); the
//    descent (mcts.py:87-92) is then a pointer chase through the per-node winners in shared memory --
//    a handful of instructions per level instead of a full score evaluation per level.
//
// Arithmetic per score is the one of mz_tree.cu / the dense engine (IEEE binary64, reference operation order);
// results are bit-identical (tests/test_gpu_fcnet.py::test_fused_search_equals_per_launch_search).
//
// Per-game block in global memory (L2):   S1 = round_up(S + 1, 2) node slots, A2 = round_up(A, 2)
//   header 64 B : f64 min | f64 max
//   own[n]  16 B: f64 value_sum | f32 reward | u32 -
//   edge[n] 16 B: f64 prior of the edge parent(n) -> n | f64 q(n) = reward -/+ discount * value()
//   utop[n]  8 B: prior of the best unexpanded child of n
//   meta[n]  8 B: copy of the shared-memory node words at the end of the move (export / tests)
//   pri[n][A2]  : all priors of node n (Node.expand, mcts.py:52-55; the root's include the Dirichlet noise)
// Per-game shared memory: node word  N | action << 8 | uact << 16 (visit count, action from the parent, best
// unexpanded action with bit 6 = rank densely, 0xff = none), mask of actions that are not unexpanded children
// (expanded or illegal), parent id, and the per-node winners of the current simulation.
#pragma once

namespace fs2 {

constexpr int L = 8;               // lanes per game
constexpr int MAX_S = 63;          // items per lane: ceil((2 S + 1) / 8) <= 16
constexpr int MAX_ITEMS = 16;
constexpr int UACT_NONE = 0xff, UACT_DENSE = 0x40;

struct Smem {
  float* logit;        // [GP][A4]
  float* val;          // [GP]
  float* rew;          // [GP]
  double* mm;          // [GP][2]
  double* sp;          // [GP][25]  exp(logit) / reward scratch (aliases best_key: used in different phases)
  unsigned long long* best_key;  // [GP][S1]
  uint32_t* best_ca;   // [GP][S1]  (action << 8) | child id (0xff: the child is unexpanded)
  uint32_t* node;      // [GP][S1]  N | action << 8 | uact << 16
  uint32_t* xmask;     // [GP][S1]
  uint8_t* par;        // [GP][S1]
  uint8_t* path_n;     // [GP][PS]
  uint8_t* path_a;     // [GP][PS]
  uint8_t* depth;      // [GP]
  const unsigned long long* exp_tab;
  int ps, a4, s1;
};

struct Game {
  bool valid;
  int g, gl, sub;
  uint8_t* base;
};

struct Geo {  // byte offsets inside a game block
  int own, edge, utop, meta, pri, a2, s1;
  long long bytes;
};
__host__ __device__ inline Geo geo(int S, int A) {
  Geo g;
  g.s1 = (S + 2) & ~1;
  g.a2 = (A + 1) & ~1;
  g.own = 64;
  g.edge = g.own + 16 * g.s1;
  g.utop = g.edge + 16 * g.s1;
  g.meta = g.utop + 8 * g.s1;
  g.pri = g.meta + 8 * g.s1;
  g.bytes = ((long long)g.pri + (long long)g.s1 * 8 * g.a2 + 127) / 128 * 128;
  return g;
}

MZ_DEV double shfl8_xor_f64(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(MZ_FULL, lo, m, L);
  hi = __shfl_xor_sync(MZ_FULL, hi, m, L);
  return __hiloint2double(hi, lo);
}
MZ_DEV unsigned long long sortable(double x) {  // monotone map double -> u64 (no NaNs); > 0 for every finite x
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return b ^ ((unsigned long long)((long long)b >> 63) | 0x8000000000000000ull);
}

// Best unexpanded child over the priors this lane holds (pv[t] = prior of action 8 t + sub, available iff its
// bit is clear in `xmask`): returns across the eight lanes the maximal prior, the largest action that has it,
// and whether another available child with a LARGER action lies within 2^-50 relative below the maximum (then
// a rounded product could tie and the reference's tie-break would prefer it: the node is ranked densely).
template <int AL>
MZ_DEV void top_unexpanded(const double (&pv)[AL], uint32_t xmask, int A, int sub, bool force_dense, double& ptop,
                           int& uact) {
  double bp = -1.0;
  int ba = -1;
#pragma unroll
  for (int t = 0; t < AL; ++t) {
    const int a = L * t + sub;
    if (a < A && !((xmask >> a) & 1u) && pv[t] >= bp) {  // ascending actions: >= keeps the larger one
      bp = pv[t];
      ba = a;
    }
  }
#pragma unroll
  for (int m = 1; m < L; m <<= 1) {
    const double op = shfl8_xor_f64(bp, m);
    const int oa = __shfl_xor_sync(MZ_FULL, ba, m, L);
    if (oa >= 0 && (ba < 0 || op > bp || (op == bp && oa > ba))) {
      bp = op;
      ba = oa;
    }
  }
  bool close = false;
  if (ba >= 0) {
    const double thr = __dmul_rn(bp, 1.0 - 0x1p-50);
#pragma unroll
    for (int t = 0; t < AL; ++t) {
      const int a = L * t + sub;
      if (a < A && !((xmask >> a) & 1u) && a > ba && pv[t] < bp && pv[t] >= thr) close = true;
    }
  }
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = 0xffu << (lane & ~7u);
  const bool any_close = (__ballot_sync(MZ_FULL, close) & gmask) != 0u;
  ptop = bp;
  uact = ba < 0 ? UACT_NONE : (ba | ((any_close || force_dense || bp < 0x1p-900) ? UACT_DENSE : 0));
}

// Node.expand priors for the eight lanes of a game (see fs_prior_sum in mz_fcsearch.cu: same sum order)
template <int AL>
MZ_DEV double prior_sum(const FsParams& p, const Smem& sm, const Game& gm, const float* logits, uint32_t legal_bits,
                        double (&pexp)[AL]) {
  const int A = p.A;
  double* sp = sm.sp + gm.gl * SP_STRIDE;
#pragma unroll
  for (int t = 0; t < AL; ++t) {
    const int a = L * t + gm.sub;
    double e = 0.0;
    if (a < A && ((legal_bits >> a) & 1u)) e = fs_exp((double)logits[a], sm.exp_tab);
    pexp[t] = e;
    if (a < A) sp[a] = e;
  }
  __syncwarp();
  double f = 0.0, c = 0.0;
  bool first = true;
#pragma unroll 1
  for (int a = 0; a < A; ++a) {
    if (!((legal_bits >> a) & 1u)) continue;
    const double x = sp[a];
    if (first) {
      f = x;  // int 0 + x
      first = false;
    } else if (p.prior_sum_mode == 0) {
      f = __dadd_rn(f, x);
    } else {  // Neumaier step, CPython >= 3.12 Python/bltinmodule.c
      const double s = __dadd_rn(f, x);
      const bool big = fabs(f) >= fabs(x);
      const double hi = big ? f : x, lo = big ? x : f;
      c = __dadd_rn(c, __dadd_rn(__dsub_rn(hi, s), lo));
      f = s;
    }
  }
  if (p.prior_sum_mode != 0 && c != 0.0 && isfinite(c)) f = __dadd_rn(f, c);
  __syncwarp();
  return f;
}

template <int AL>
MZ_DEV void set_root(const FsParams& p, const Smem& sm, const Game& gm, const Geo& G) {
  const int A = p.A;
  uint32_t lm = 0u;
  if (gm.valid) {
    lm = p.legal ? p.legal[gm.g] : 0xffffffffu;
    if (A < 32) lm &= (1u << A) - 1u;
  }
  double pexp[AL];
  const double f = prior_sum<AL>(p, sm, gm, p.root_logits + (size_t)(gm.valid ? gm.g : 0) * A, lm, pexp);
  double pv[AL];
#pragma unroll
  for (int t = 0; t < AL; ++t) {
    const int a = L * t + gm.sub;
    const bool legal = a < A && ((lm >> a) & 1u);
    double prior = legal ? __ddiv_rn(pexp[t], f) : 0.0;
    if (p.noise && legal) {  // noise is dense over the root's children in action order (mcts.py:57-61)
      const int j = __popc(lm & ((1u << a) - 1u));
      const double nz = p.noise[(size_t)gm.g * A + j];
      prior = __dadd_rn(__dmul_rn(prior, __dsub_rn(1.0, p.noise_frac)), __dmul_rn(nz, p.noise_frac));
    }
    pv[t] = prior;
    if (gm.valid && a < A) *reinterpret_cast<double*>(gm.base + G.pri + 8 * a) = prior;
  }
  double ptop;
  int uact;
  top_unexpanded<AL>(pv, ~lm, A, gm.sub, p.init_score != 0.0, ptop, uact);
  if (gm.valid && gm.sub == 0) {
    sm.node[gm.gl * sm.s1] = 0u | (0xffu << 8) | ((uint32_t)uact << 16);
    sm.xmask[gm.gl * sm.s1] = ~lm;
    sm.par[gm.gl * sm.s1] = 0;
    sm.mm[2 * gm.gl] = p.min_bound;
    sm.mm[2 * gm.gl + 1] = p.max_bound;
    *reinterpret_cast<uint4*>(gm.base + G.own) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<double*>(gm.base + G.utop) = ptop;
  }
  if (!gm.valid) return;
  // hidden state of the root: float32 [50] -> bf16 [64] (zero padded), pool slot 0: one 16-byte block per lane
  const float* h = p.root_hidden + (size_t)gm.g * H;
  __nv_bfloat16* row = p.pool + (size_t)gm.g * (p.S + 1) * POOL_ROW;
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = 8 * gm.sub + 2 * j;
    w[j] = pack_bf16(k < H ? h[k] : 0.0f, k + 1 < H ? h[k + 1] : 0.0f);
  }
  *reinterpret_cast<uint4*>(row + 8 * gm.sub) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------------
// descent = [rank every expanded node's candidates in parallel] + [pointer chase from the root]
// ------------------------------------------------------------------------------------------------------
#define FS2_STAMP(slot_)                    \
  do {                                      \
    if (tl) tl[(slot_)] = clock64();        \
  } while (0)

MZ_DEV void descend(const FsParams& p, const Smem& sm, const Game& gm, const Geo& G, int sim, int& out_parent,
                    int& out_action, long long* tl) {
  const int A = p.A, SP1 = p.S + 1;
  const double init_score = p.init_score;
  const double mn = sm.mm[2 * gm.gl], mx = sm.mm[2 * gm.gl + 1];
  const double d = __dsub_rn(mx, mn);
  int mode = 0;
  double r = 0.0;
  if (mx > mn) {
    if (fs_divisor_ok(d)) {
      mode = 2;
      r = __drcp_rn(d);
    } else {
      mode = 3;
    }
  } else if (mx == mn) {
    mode = 1;
  }
  uint32_t* node = sm.node + gm.gl * sm.s1;
  const uint32_t* xmask = sm.xmask + gm.gl * sm.s1;
  const uint8_t* par = sm.par + gm.gl * sm.s1;
  unsigned long long* best_key = sm.best_key + gm.gl * sm.s1;
  uint32_t* best_ca = sm.best_ca + gm.gl * sm.s1;
  // nodes 0..sim exist; items: i <= sim -> the best unexpanded child of node i; i > sim -> the edge into node i - sim
  const int n_items = gm.valid ? 2 * sim + 1 : 0;
  for (int n = gm.sub; n <= sim; n += L) {
    best_key[n] = 0ull;
    best_ca[n] = 0u;
  }
  __syncwarp();
  FS2_STAMP(16);
  unsigned long long key[MAX_ITEMS];
  uint32_t pack[MAX_ITEMS];  // parent << 16 | action << 8 | child id (0xff = unexpanded)
  // ---- issue every global load of this lane's items first (one memory round trip) ----
  double ld_a[MAX_ITEMS], ld_b[MAX_ITEMS], ld_c[MAX_ITEMS];
#pragma unroll
  for (int it = 0; it < MAX_ITEMS; ++it) {
    const int i = L * it + gm.sub;
    ld_a[it] = ld_b[it] = ld_c[it] = 0.0;
    if (L * it < 2 * p.S + 1 && i < n_items) {
      if (i <= sim) {
        const uint32_t w = node[i];
        ld_a[it] = *reinterpret_cast<const double*>(gm.base + G.utop + 8 * i);
        ld_c[it] = __ldg(p.pb_c + (size_t)(w & 0xffu) * SP1);  // pb_c[N][0]
      } else {
        const int j = i - sim;
        const double2 e = *reinterpret_cast<const double2*>(gm.base + G.edge + 16 * j);
        ld_a[it] = e.x;
        ld_b[it] = e.y;
        const int np_ = (int)(node[par[j]] & 0xffu), nj = (int)(node[j] & 0xffu);
        ld_c[it] = __ldg(p.pb_c + (size_t)np_ * SP1 + nj);
      }
    }
  }
  FS2_STAMP(17);
#pragma unroll
  for (int it = 0; it < MAX_ITEMS; ++it) {
    const int i = L * it + gm.sub;
    key[it] = 0ull;
    pack[it] = 0u;
    if (L * it < 2 * p.S + 1 && i < n_items) {
      if (i <= sim) {  // best unexpanded child of node i
        const uint32_t w = node[i];
        const int N = (int)(w & 0xffu), ua = (int)((w >> 16) & 0xffu);
        if (ua != UACT_NONE) {
          double score;
          int action = ua & 0x3f;
          if (ua & UACT_DENSE) {  // near-tie among the priors (or init_value_score != 0): rank all of them
            const double pb_c = ld_c[it];
            const uint32_t xm = xmask[i];
            bool have = false;
            score = 0.0;
            for (int a = 0; a < A; ++a) {
              if ((xm >> a) & 1u) continue;
              const double pr = *reinterpret_cast<const double*>(gm.base + G.pri + ((size_t)i * G.a2 + a) * 8);
              const double s = N == 0 ? pr : __dadd_rn(__dmul_rn(pb_c, pr), init_score);
              if (!have || s >= score) {
                score = s;
                action = a;
                have = true;
              }
            }
          } else {
            // mcts.py:105-108 (N == 0: the root before its first visit ranks by prior) / ucb_score mcts.py:115-124
            score = N == 0 ? ld_a[it] : __dadd_rn(__dmul_rn(ld_c[it], ld_a[it]), init_score);
          }
          key[it] = sortable(score);
          pack[it] = ((uint32_t)i << 16) | ((uint32_t)action << 8) | 0xffu;
          atomicMax(&best_key[i], key[it]);
        }
      } else {  // the edge into node j (visited at least once)
        const int j = i - sim;
        const int pn = (int)par[j];
        const double q = ld_b[it];
        double value_score;
        if (mode == 2) {
          const double x = __dsub_rn(q, mn);
          value_score = fs_div_by_const(x, d, r);
          const unsigned hx = (unsigned)__double2hiint(x);  // x >= 0 because min <= q
          if (__builtin_expect(hx - 0x33700000u > 0x19000000u, 0)) value_score = (x == 0.0) ? 0.0 : __ddiv_rn(x, d);
        } else if (mode == 3) {
          value_score = __ddiv_rn(__dsub_rn(q, mn), d);
        } else if (mode == 1) {
          value_score = 1.0;
        } else {
          value_score = q;
        }
        const double score = __dadd_rn(__dmul_rn(ld_c[it], ld_a[it]), value_score);
        key[it] = sortable(score);
        pack[it] = ((uint32_t)pn << 16) | (((node[j] >> 8) & 0xffu) << 8) | (uint32_t)j;
        atomicMax(&best_key[pn], key[it]);
      }
    }
  }
  __syncwarp();
  FS2_STAMP(18);
  // ties -> larger action (mcts.py:106-112): among the items that reach a node's maximum the largest action wins
#pragma unroll
  for (int it = 0; it < MAX_ITEMS; ++it) {
    if (L * it < 2 * p.S + 1 && key[it] != 0ull) {
      const int pn = (int)(pack[it] >> 16);
      if (key[it] == best_key[pn]) atomicMax(&best_ca[pn], pack[it] & 0xffffu);
    }
  }
  __syncwarp();
  FS2_STAMP(19);
  // ---- pointer chase: while node.expanded(): node = the winner of select_child (mcts.py:87-92) ----
  uint8_t* pn_ = sm.path_n + gm.gl * sm.ps;
  uint8_t* pa_ = sm.path_a + gm.gl * sm.ps;
  int cur = 0, depth = 0, parent = 0, action = 0;
  bool done = !gm.valid;
  if (gm.valid && gm.sub == 0) pn_[0] = 0;
  while (__any_sync(MZ_FULL, !done)) {
    if (!done) {
      const uint32_t ca = best_ca[cur];
      const int a = (int)(ca >> 8), c = (int)(ca & 0xffu);
      if (gm.sub == 0) pa_[depth] = (uint8_t)a;
      ++depth;
      if (c == 0xff) {
        parent = cur;
        action = a;
        done = true;
      } else {
        cur = c;
        if (gm.sub == 0) pn_[depth] = (uint8_t)c;
      }
    }
  }
  FS2_STAMP(20);
  if (gm.valid && gm.sub == 0) {
    sm.depth[gm.gl] = (uint8_t)depth;
    if (p.trace_parent) p.trace_parent[(size_t)sim * p.G + gm.g] = parent;
    if (p.trace_action) p.trace_action[(size_t)sim * p.G + gm.g] = action;
    if (p.trace_depth) p.trace_depth[(size_t)sim * p.G + gm.g] = depth;
  }
  out_parent = parent;
  out_action = action;
}

// ------------------------------------------------------------------------------------------------------
// expand (mcts.py:47-55) + backpropagate (mcts.py:126-143) of simulation `sim`
// ------------------------------------------------------------------------------------------------------
template <int AL>
MZ_DEV void expand_backup(const FsParams& p, const Smem& sm, const Game& gm, const Geo& G, int sim, long long* tl) {
  const int A = p.A;
  const bool two = p.two_players != 0;
  const double disc = p.discount;
  const int newn = sim + 1;
  const float* logits = sm.logit + gm.gl * sm.a4;
  const float value_f = sm.val[gm.gl], reward_in = sm.rew[gm.gl];
  const float node_reward_new = (reward_in != 0.0f) ? reward_in : 0.0f;  // `if network_output.reward:`
  const uint8_t* pn_ = sm.path_n + gm.gl * sm.ps;
  const uint8_t* pa_ = sm.path_a + gm.gl * sm.ps;
  uint32_t* node = sm.node + gm.gl * sm.s1;
  uint32_t* xmask = sm.xmask + gm.gl * sm.s1;
  const int depth = gm.valid ? (int)sm.depth[gm.gl] : 0;
  const int parent = gm.valid ? (int)pn_[depth - 1] : 0, action = gm.valid ? (int)pa_[depth - 1] : 0;
  if (gm.valid) {
    if (gm.sub == 0) {
      if (p.rec_value) p.rec_value[(size_t)sim * p.G + gm.g] = value_f;
      if (p.rec_reward) p.rec_reward[(size_t)sim * p.G + gm.g] = reward_in;
    }
    if (p.rec_logits) {
#pragma unroll
      for (int t = 0; t < AL; ++t) {
        const int a = L * t + gm.sub;
        if (a < A) p.rec_logits[((size_t)sim * p.G + gm.g) * A + a] = logits[a];
      }
    }
  }
  int dmax = depth;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) dmax = max(dmax, __shfl_xor_sync(MZ_FULL, dmax, m));

  // loads that do not depend on the expansion arithmetic: the parent's priors (its next best unexpanded child and
  // the prior of the new edge) and the path nodes' own records (value sums, rewards)
  double ppv[AL];
#pragma unroll
  for (int t = 0; t < AL; ++t) {
    const int a = L * t + gm.sub;
    ppv[t] = 0.0;
    if (gm.valid && a < A) ppv[t] = *reinterpret_cast<const double*>(gm.base + G.pri + ((size_t)parent * G.a2 + a) * 8);
  }
  constexpr int MAXP = (MAX_S + 1 + L) / L;  // path positions per lane (depth <= S)
  uint4 own[MAXP];
#pragma unroll
  for (int m = 0; m < MAXP; ++m) {
    const int kk = L * m + gm.sub;
    own[m] = make_uint4(0u, 0u, 0u, 0u);
    if (L * m <= dmax && gm.valid && kk < depth) own[m] = ldg16(gm.base + G.own + 16 * (int)pn_[kk]);
  }

  FS2_STAMP(22);
  // ---- expand: priors of the new node (every action is legal below the root, mcts.py:72, 97) ----
  double pexp[AL];
  const uint32_t all = A < 32 ? (1u << A) - 1u : 0xffffffffu;
  const double f = prior_sum<AL>(p, sm, gm, logits, gm.valid ? all : 0u, pexp);
  FS2_STAMP(23);
  double pv[AL];
#pragma unroll
  for (int t = 0; t < AL; ++t) {
    const int a = L * t + gm.sub;
    pv[t] = (gm.valid && a < A) ? __ddiv_rn(pexp[t], f) : 0.0;
    if (gm.valid && a < A) *reinterpret_cast<double*>(gm.base + G.pri + ((size_t)newn * G.a2 + a) * 8) = pv[t];
  }
  FS2_STAMP(24);
  double ptop_new, ptop_par;
  int uact_new, uact_par;
  const bool force_dense = p.init_score != 0.0;
  top_unexpanded<AL>(pv, ~all, A, gm.sub, force_dense, ptop_new, uact_new);
  const uint32_t xm_par = gm.valid ? (xmask[parent] | (1u << action)) : 0xffffffffu;
  top_unexpanded<AL>(ppv, xm_par, A, gm.sub, force_dense, ptop_par, uact_par);
  // the prior of the new edge: the lane that holds the parent's prior of `action` hands it round
  double edge_prior = 0.0;
#pragma unroll
  for (int t = 0; t < AL; ++t) {
    const double v = __hiloint2double(__shfl_sync(MZ_FULL, __double2hiint(ppv[t]), action & (L - 1), L),
                                      __shfl_sync(MZ_FULL, __double2loint(ppv[t]), action & (L - 1), L));
    if (t == action / L) edge_prior = v;
  }
  __syncwarp();
  if (gm.valid && gm.sub == 0) {
    node[newn] = 0u | ((uint32_t)action << 8) | ((uint32_t)uact_new << 16);
    xmask[newn] = ~all;
    sm.par[gm.gl * sm.s1 + newn] = (uint8_t)parent;
    *reinterpret_cast<double*>(gm.base + G.utop + 8 * newn) = ptop_new;
    node[parent] = (node[parent] & 0xff00ffffu) | ((uint32_t)uact_par << 16);
    xmask[parent] = xm_par;
    if (uact_par != UACT_NONE) *reinterpret_cast<double*>(gm.base + G.utop + 8 * parent) = ptop_par;
  }
  __syncwarp();
  FS2_STAMP(25);

  // ---- backup.  Position k on the path is node path_n[k] (k < depth) or the new node (k == depth) ----
  float* scratch = reinterpret_cast<float*>(sm.sp + gm.gl * SP_STRIDE);  // [<= 50] rewards by position
  // (SP_STRIDE doubles = 50 floats: positions beyond 49 only exist for S > 49 -> chunked below)
  double value = (double)value_f;
  double lmin = INFINITY, lmax = -INFINITY;
  constexpr int CH = 48;  // positions per chunk (multiple of L, <= 50 scratch floats)
  for (int base = (dmax / CH) * CH; base >= 0; base -= CH) {
    const int m0 = base / L;
    // rewards of the chunk's positions -> scratch
#pragma unroll
    for (int m = 0; m < MAXP; ++m) {
      const int kk = L * m + gm.sub;
      if (m >= m0 && m < m0 + CH / L && L * m <= dmax && gm.valid && kk <= depth)
        scratch[kk - base] = kk < depth ? __uint_as_float(own[m].z) : node_reward_new;
    }
    __syncwarp();
    double myval[MAXP];
#pragma unroll
    for (int m = MAXP - 1; m >= 0; --m) {
      myval[m] = 0.0;
      if (m < m0 || m >= m0 + CH / L || L * m > dmax) continue;  // warp uniform
#pragma unroll
      for (int jj = L - 1; jj >= 0; --jj) {
        const int kk = L * m + jj;
        if (gm.valid && kk <= depth) {
          const float rj = scratch[kk - base];
          if (jj == gm.sub) myval[m] = value;
          // value = (-reward if two_players and node.to_play == to_play else reward) + discount * value
          const bool same = two ? (((depth - kk) & 1) == 0) : false;
          value = __dadd_rn((double)(same ? -rj : rj), __dmul_rn(disc, value));
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MAXP; ++m) {
      const int kk = L * m + gm.sub;
      if (m >= m0 && m < m0 + CH / L && L * m <= dmax && gm.valid && kk <= depth) {
        const bool same = two ? (((depth - kk) & 1) == 0) : true;
        const double vs0 = kk < depth ? u2d(own[m].x, own[m].y) : 0.0;
        const float rw = kk < depth ? __uint_as_float(own[m].z) : node_reward_new;
        const double nvs = __dadd_rn(vs0, same ? myval[m] : -myval[m]);
        const int nid = kk < depth ? (int)pn_[kk] : newn;
        const int nvc = (int)(node[nid] & 0xffu) + 1;
        node[nid] = (node[nid] & 0xffffff00u) | (uint32_t)nvc;
        *reinterpret_cast<uint4*>(gm.base + G.own + 16 * nid) =
            make_uint4((uint32_t)__double2loint(nvs), (uint32_t)__double2hiint(nvs), __float_as_uint(rw), 0u);
        if (kk > 0) {  // mcts.py:136-141
          const double dq = __dmul_rn(disc, __ddiv_rn(nvs, (double)nvc));
          const double new_q = two ? __dsub_rn((double)rw, dq) : __dadd_rn((double)rw, dq);
          lmin = fmin(lmin, new_q);
          lmax = fmax(lmax, new_q);
          if (kk == depth) *reinterpret_cast<double2*>(gm.base + G.edge + 16 * nid) = make_double2(edge_prior, new_q);
          else *reinterpret_cast<double*>(gm.base + G.edge + 16 * nid + 8) = new_q;
        }
      }
    }
    __syncwarp();
  }
  FS2_STAMP(26);
#pragma unroll
  for (int m = 1; m < L; m <<= 1) {
    lmin = fmin(lmin, shfl8_xor_f64(lmin, m));
    lmax = fmax(lmax, shfl8_xor_f64(lmax, m));
  }
  if (gm.valid && gm.sub == 0 && depth > 0) {  // MinMaxStats.update mcts.py:11-14
    if (lmin < sm.mm[2 * gm.gl]) sm.mm[2 * gm.gl] = lmin;
    if (lmax > sm.mm[2 * gm.gl + 1]) sm.mm[2 * gm.gl + 1] = lmax;
  }
  __syncwarp();
}

// game.py:106-111 + Node.value (mcts.py:42-45); also leaves the node words in the block for mz_fc_search_export
MZ_DEV void root_stats(const FsParams& p, const Smem& sm, const Game& gm, const Geo& G) {
  if (!gm.valid) return;
  const int A = p.A, S = p.S;
  const uint32_t* node = sm.node + gm.gl * sm.s1;
  const uint8_t* par = sm.par + gm.gl * sm.s1;
  const uint32_t lm = ~sm.xmask[gm.gl * sm.s1] | 0u;  // after the search the root's mask also has its expanded children
  (void)lm;
  uint32_t legal = p.legal ? p.legal[gm.g] : 0xffffffffu;
  if (A < 32) legal &= (1u << A) - 1u;
  for (int a = gm.sub; a < A; a += L) {
    if (p.visits) p.visits[(size_t)gm.g * A + a] = 0;
    if (p.child_visits) p.child_visits[(size_t)gm.g * A + a] = 0.0;
  }
  __syncwarp(0xffu << ((threadIdx.x & 31u) & ~7u));
  int sum = 0;
  for (int j = 1; j <= S; ++j)
    if (par[j] == 0) sum += (int)(node[j] & 0xffu);
  for (int j = 1 + gm.sub; j <= S; j += L) {
    if (par[j] == 0) {
      const int a = (int)((node[j] >> 8) & 0xffu), v = (int)(node[j] & 0xffu);
      if (p.visits) p.visits[(size_t)gm.g * A + a] = v;
      if (p.child_visits) p.child_visits[(size_t)gm.g * A + a] = __ddiv_rn((double)v, (double)sum);
    }
  }
  for (int n = gm.sub; n <= S; n += L)
    *reinterpret_cast<uint2*>(gm.base + G.meta + 8 * n) = make_uint2(node[n] | ((uint32_t)par[n] << 24), sm.xmask[gm.gl * sm.s1 + n]);
  if (gm.sub == 0) {
    const uint4 o = ldg16(gm.base + G.own);
    const int n = (int)(node[0] & 0xffu);
    if (p.root_value) p.root_value[gm.g] = n == 0 ? 0.0 : __ddiv_rn(u2d(o.x, o.y), (double)n);
    if (p.minmax) {
      p.minmax[2 * gm.g] = sm.mm[2 * gm.gl];
      p.minmax[2 * gm.g + 1] = sm.mm[2 * gm.gl + 1];
    }
    *reinterpret_cast<double2*>(gm.base) = make_double2(sm.mm[2 * gm.gl], sm.mm[2 * gm.gl + 1]);
    *reinterpret_cast<uint32_t*>(gm.base + 16) = legal;
  }
}

}  // namespace fs2
