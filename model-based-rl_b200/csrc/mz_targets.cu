// Replay-side kernels: fused target construction over the device-resident replay window, and the
// scalar transform / support kernels.
//
// mz_build_targets fuses, per sampled row: observation gather (+ optional normalisation), action
// slice + padding, n-step value target (td_steps window of sign-corrected, discounted rewards +
// discounted bootstrap root value), reward / policy targets, and optionally h(x) + two-hot support
// projection.  Reference: replay_buffer.py:124-198, learners.py:170-192, config.py:27-68.
#include <math.h>

#include "mz_common.cuh"
#include "mz_transforms.cuh"

namespace {

constexpr int kTgtWarps = 8;  // sampled rows per CTA
constexpr int kTgtThreads = kTgtWarps * 32;

// ClipRewardEnv.reward (wrappers.py:236-238): np.sign at the point the replay reads a stored reward
// (mz_window.clip_rewards); zeros of either sign give +0.0 like numpy, NaN stays NaN
MZ_DEV float tgt_reward(float r, int clip) {
  if (!clip) return r;
  return r > 0.0f ? 1.0f : (r < 0.0f ? -1.0f : (r == 0.0f ? 0.0f : r));
}

constexpr int kTgtFastK = 7;  // up to 8 unroll positions keep their sums in registers
constexpr int kDiscPad = 8;   // zeros in front of the widened discount table: reads at small negative indices are safe

// The interior of a row's window -- the elements m with 0 <= m - i < n_i for EVERY position i <= K, all of one
// player: no range or sign test there -- is summed on the FP64 tensor cores (see the kernel); this function adds the
// edges [0, lo) and [hi, win) with per-element tests: acc[i] = sum_j (+-)rewards[step+i+j] * discounts[j],
// j < min(T, len - (step+i)); the sign flips where to_play differs from the position's (replay_buffer.py:187-189).
// Exact float32 products accumulated in binary64; returns position `lane`'s edge sum.
template <int NP>
MZ_DEV double nstep_edges(const float* rw, const int8_t* tp, const double* s_disc, int K, int T, int step, int len,
                          int win, int lo, int hi, int lane, int clip) {
  double acc[NP];
  int tp_i[NP], n_i[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    acc[i] = 0.0;
    const bool in = i <= K && step + i < len;
    n_i[i] = in ? min(T, len - step - i) : 0;
    tp_i[i] = in ? tp[i] : 0;
  }
  for (int part = 0; part < 2; ++part) {
    const int m_end = part ? win : min(lo, win);
    for (int m = (part ? hi : 0) + lane; m < m_end; m += 32) {
      const float r = tgt_reward(rw[m], clip);
      const int t = tp[m];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int j = m - i;
        if (j >= 0 && j < n_i[i]) acc[i] += (double)(t != tp_i[i] ? -r : r) * s_disc[j];
      }
    }
  }
  double mine_acc = 0.0;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if (i <= K) {  // warp-uniform
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) acc[i] += shfl_xor_f64<32>(acc[i], m);
      if (lane == i) mine_acc = acc[i];
    }
  }
  return mine_acc;
}

// D (8 x 8, f64) += A (8 x 4, row major) * B (4 x 8, column major) on the FP64 tensor cores (DMMA.8x8x4).  Lane l holds
// A[l >> 2][l & 3], B[l & 3][l >> 2] and D[l >> 2][2 * (l & 3) + {0, 1}].
MZ_DEV void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// One warp per sampled row, kTgtWarps rows per CTA.  The row's reward and to_play windows (td_steps = 1000: 4 KB +
// 1 KB) are staged by the bulk-copy engine: one cp.async.bulk each over the 16-byte-aligned part of the window
// (start rounded down -- still inside the array --, end rounded down, the ragged tail of <= 3 rewards / <= 15 bytes
// with ordinary loads), completion on the warp's own mbarrier; the other global loads of the row (observation,
// bootstrap root values, child-visit rows) are issued before the wait, so a row costs two DRAM round trips
// (index arrays, then everything else); all stores are warp-contiguous.
// The n-step sums are a short convolution: value_i = sum_m rewards[m] * discounts[m - i] over the window.  For up
// to 8 unroll positions the CTA's eight rows go through the FP64 tensor cores together: D[i][row] += A[i][k] *
// B[k][row] with A[i][k] = discounts[m0 + k - i] (the same for every row: window-relative indices) and B[k][row] =
// the row's reward at element m0 + k (zero outside the row's interior), four window elements per DMMA.8x8x4, the
// steps dealt round-robin to the eight warps and the partial tiles added in shared memory.  One shared-memory load
// per lane and tensor-core step instead of seven per element and lane (the kernel was shared-memory-bandwidth bound:
// 13 wavefronts per 32 elements); the edges of the window (range / sign tests) stay with the row's own warp.
__global__ void __launch_bounds__(kTgtThreads)
build_targets_kernel(mz_window w, mz_target_cfg c, const int64_t* __restrict__ pos_arr,
                     const int64_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_len,
                     const int32_t* __restrict__ pad_actions, float* __restrict__ obs_out,
                     int32_t* __restrict__ actions_out, float* __restrict__ t_rewards,
                     float* __restrict__ t_values, float* __restrict__ t_policies,
                     float* __restrict__ value_support, float* __restrict__ reward_support,
                     int warp_smem_bytes, int bulk_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a warp without a row of its own (last CTA) repeats the batch's last row: same values to the same addresses, and
  // every warp stays for the CTA-wide tensor-core pass
  const int b = min(blockIdx.x * kTgtWarps + warp, c.batch - 1);
  const int K = c.num_unroll_steps, T = c.td_steps, A = w.num_actions, E = w.obs_elems;
  __shared__ int s_lo[kTgtWarps], s_hi[kTgtWarps], s_rwoff[kTgtWarps];
  __shared__ double s_part[kTgtWarps][64];
  // the discount table f32(discount ** n) (replay_buffer.py:84) is shared by the CTA's rows
  double* s_disc = reinterpret_cast<double*>(smem_raw) + kDiscPad;  // widened once per CTA, kDiscPad zeros in front
  for (int j = threadIdx.x; j < K + T + kDiscPad; j += kTgtThreads)
    s_disc[j - kDiscPad] = j < kDiscPad ? 0.0 : (double)c.discounts[j - kDiscPad];
  const int64_t pos = pos_arr[b];
  const int step = (int)(pos - chunk_start[b]);
  const int len = chunk_len[b];  // len(root_values) == len(rewards) == len(to_play)
  __syncthreads();

  unsigned char* rows_base = smem_raw + ((size_t)(K + T + kDiscPad) * sizeof(double) + 15) / 16 * 16;
  unsigned char* mine = rows_base + (size_t)warp * warp_smem_bytes;
  const int KT4 = (K + T + 4 + 3) & ~3, KT16 = (K + T + 16 + 15) & ~15;
  float* s_rew_buf = reinterpret_cast<float*>(mine);                   // [KT4] rewards from position pos & ~3 on
  int8_t* s_tp_buf = reinterpret_cast<int8_t*>(s_rew_buf + KT4);       // [KT16] to_play from position pos & ~15 on
  float* s_val = reinterpret_cast<float*>(s_tp_buf + KT16);            // [K+1] value targets
  float* s_lastr = s_val + (K + 1);                                    // [K+1] reward targets
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_lastr + (K + 1));      // the warp's mbarrier (8-byte aligned)

  // ---- loads -------------------------------------------------------------------------------------
  // reward / to_play window [step, min(step + K + T, len)): element j of the window = position pos + j
  const int win = max(0, min(K + T, len - step));
  const int rw_off = (int)(pos & 3), tp_off = (int)(pos & 15);
  float* s_rew = s_rew_buf + rw_off;   // s_rew[j] = rewards[pos + j] (raw: clipped where it is read)
  int8_t* s_tp = s_tp_buf + tp_off;    // s_tp[j] = to_play[pos + j]
  {
    const int64_t r_end = (pos + win) & ~(int64_t)3, t_end = (pos + win) & ~(int64_t)15;  // ends rounded down
    const uint32_t r_bytes = (bulk_ok && r_end > pos - rw_off) ? (uint32_t)(r_end - (pos - rw_off)) * 4u : 0u;
    const uint32_t t_bytes = (bulk_ok && t_end > pos - tp_off) ? (uint32_t)(t_end - (pos - tp_off)) : 0u;
    if (lane == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
      mbar_arrive_expect_tx(bar, r_bytes + t_bytes);
      if (r_bytes) bulk_copy_g2s(s_rew_buf, w.rewards + (pos - rw_off), r_bytes, bar);
      if (t_bytes) bulk_copy_g2s(s_tp_buf, w.to_play + (pos - tp_off), t_bytes, bar);
    }
    // ragged tails with ordinary loads (disjoint from what the engine writes)
    const int r_tail0 = r_bytes ? (int)(r_end - pos) : 0, t_tail0 = t_bytes ? (int)(t_end - pos) : 0;
    for (int j = r_tail0 + lane; j < win; j += 32) s_rew[j] = w.rewards[pos + j];
    for (int j = t_tail0 + lane; j < win; j += 32) s_tp[j] = w.to_play[pos + j];
  }
  const float prev_reward = (step > 0 && step <= len) ? tgt_reward(w.rewards[pos - 1], w.clip_rewards) : 0.0f;
  // bootstrap of position `lane`: root_values[step + lane + T] where it exists (replay_buffer.py:180-183)
  double root = 0.0;
  if (lane <= K && step + lane + T < len) root = w.root_values[pos + lane + T];
  // positions that still lie inside the chunk: step + i < len
  const int n_in = max(0, min(K + 1, len - step));

  // observation: np.float32(history.observations[step])  (replay_buffer.py:147)
  {
    float* dst = obs_out + (size_t)b * E;
    if ((E & 3) == 0) {
      float4* dst4 = reinterpret_cast<float4*>(dst);
      const float4* mn4 = reinterpret_cast<const float4*>(c.obs_min);
      const float4* rg4 = reinterpret_cast<const float4*>(c.obs_range);
      const bool vec_norm = c.normalize_obs && (((reinterpret_cast<uintptr_t>(c.obs_min) |
                                                  reinterpret_cast<uintptr_t>(c.obs_range)) & 15) == 0);
      for (int e = lane; e < (E >> 2); e += 32) {
        float4 v;
        if (w.obs_is_u8) {
          const uint32_t q = reinterpret_cast<const uint32_t*>(
              reinterpret_cast<const uint8_t*>(w.obs) + (size_t)pos * E)[e];
          v = make_float4((float)(q & 255u), (float)((q >> 8) & 255u), (float)((q >> 16) & 255u),
                          (float)(q >> 24));
        } else {
          v = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(w.obs) + (size_t)pos * E)[e];
        }
        if (c.normalize_obs) {
          float4 mn, rg;
          if (vec_norm) {
            mn = mn4[e];
            rg = rg4[e];
          } else {
            mn = make_float4(c.obs_min[4 * e], c.obs_min[4 * e + 1], c.obs_min[4 * e + 2], c.obs_min[4 * e + 3]);
            rg = make_float4(c.obs_range[4 * e], c.obs_range[4 * e + 1], c.obs_range[4 * e + 2],
                             c.obs_range[4 * e + 3]);
          }
          v.x = __fdiv_rn(__fsub_rn(v.x, mn.x), rg.x);
          v.y = __fdiv_rn(__fsub_rn(v.y, mn.y), rg.y);
          v.z = __fdiv_rn(__fsub_rn(v.z, mn.z), rg.z);
          v.w = __fdiv_rn(__fsub_rn(v.w, mn.w), rg.w);
        }
        dst4[e] = v;
      }
    } else {
      for (int e = lane; e < E; e += 32) {
        float v = w.obs_is_u8 ? (float)reinterpret_cast<const uint8_t*>(w.obs)[(size_t)pos * E + e]
                              : reinterpret_cast<const float*>(w.obs)[(size_t)pos * E + e];
        if (c.normalize_obs) v = __fdiv_rn(__fsub_rn(v, c.obs_min[e]), c.obs_range[e]);
        dst[e] = v;
      }
    }
  }
  // policy targets: child_visits[step + i] for the positions inside the chunk (one contiguous run of
  // n_in * A floats), zeros after the end (absorbing policy, replay_buffer.py:90)
  {
    float* pol = t_policies + (size_t)b * (K + 1) * A;
    const float* src = w.child_visits + (size_t)pos * A;
    const int n_real = n_in * A;
    for (int e = lane; e < (K + 1) * A; e += 32) pol[e] = e < n_real ? src[e] : 0.0f;
  }
  // actions: history.actions[step:step+K], padded with random actions (replay_buffer.py:149-152)
  {
    const int n_real = max(0, min(K, len - step));
    for (int k = lane; k < K; k += 32)
      actions_out[(size_t)b * K + k] = k < n_real ? w.actions[pos + k] : pad_actions[(size_t)b * K + (k - n_real)];
  }
  __syncwarp();       // the tails written by other lanes
  mbar_wait(bar, 0);  // the bulk copies have landed
  // does the whole window belong to one player?  The slack bytes around the window (up to 15 in front, up to 15
  // behind, inside the 16-byte groups the window touches) are overwritten with its first byte, then whole 16-byte
  // groups are compared against that byte, 16 per lane and step
  bool one_player = true;
  if (win > 0) {
    const int8_t first = s_tp[0];
    const int end16 = (tp_off + win + 15) & ~15;
    __syncwarp();
    if (lane < tp_off) s_tp_buf[lane] = first;
    if (tp_off + win + lane < end16) s_tp_buf[tp_off + win + lane] = first;
    __syncwarp();
    const uint32_t splat = 0x01010101u * (uint32_t)(uint8_t)first;
    uint32_t diff = 0;
    for (int q = lane; q < (end16 >> 4); q += 32) {
      const uint4 x4 = reinterpret_cast<const uint4*>(s_tp_buf)[q];
      diff |= (x4.x ^ splat) | (x4.y ^ splat) | (x4.z ^ splat) | (x4.w ^ splat);
    }
    one_player = diff == 0;
  }
  one_player = __all_sync(MZ_FULL, one_player);
  const int clip = w.clip_rewards;

  // ---- insert_target (replay_buffer.py:165-198) -----------------------------------------------------
  if (K <= kTgtFastK) {
    // the row's interior: window elements every position is in range of (replay_buffer.py:187-189 without tests)
    int lo_r = K, hi_r = win;
    for (int i = 0; i <= K; ++i) hi_r = min(hi_r, i + ((step + i < len) ? min(T, len - step - i) : 0));
    if (!one_player || hi_r <= lo_r) lo_r = hi_r = 0;  // everything through the edge loop
    if (lane == 0) {
      s_lo[warp] = lo_r;
      s_hi[warp] = hi_r;
      s_rwoff[warp] = rw_off;
    }
    __syncthreads();  // every row's window is staged, every interior known
    double dm = 0.0;
    {
      int lo_all = 1 << 30, hi_all = 0, lo_in = 0, hi_in = 1 << 30;  // union and intersection of the rows' interiors
#pragma unroll
      for (int r = 0; r < kTgtWarps; ++r) {
        if (s_hi[r] > s_lo[r]) {
          lo_all = min(lo_all, s_lo[r]);
          hi_all = max(hi_all, s_hi[r]);
        }
        lo_in = max(lo_in, s_lo[r]);
        hi_in = min(hi_in, s_hi[r]);
      }
      const int n = lane >> 2, k = lane & 3;  // B: element k of row n;  A: element k for position i = lane >> 2
      const float* rw_n = reinterpret_cast<const float*>(rows_base + (size_t)n * warp_smem_bytes) + s_rwoff[n];
      const int lo_n = s_lo[n], hi_n = s_hi[n];
      double c0 = 0.0, c1 = 0.0;
      const double* dk = s_disc - n;  // discounts[m0 + k - i], i = lane >> 2 (zeros in front of the table)
      int m = lo_all + 4 * warp + k;
      if (!clip && lo_in <= lo_all) {
        // steps that lie inside EVERY row's interior need no range test (lo_in / hi_in: intersection of the interiors)
        const int m_end = hi_in - 3 + k;  // m0 + 3 < hi_in
#pragma unroll 4
        for (; m < m_end; m += 4 * kTgtWarps) dmma_8x8x4(c0, c1, dk[m], (double)rw_n[m]);
      }
      for (; m - k < hi_all; m += 4 * kTgtWarps) {
        const double bv = (m >= lo_n && m < hi_n) ? (double)tgt_reward(rw_n[m], clip) : 0.0;
        dmma_8x8x4(c0, c1, dk[m], bv);
      }
      s_part[warp][2 * lane] = c0;
      s_part[warp][2 * lane + 1] = c1;
      __syncthreads();
      if (lane <= K) {  // position `lane` of this warp's row: D[lane][warp] of every warp's tile
#pragma unroll
        for (int r = 0; r < kTgtWarps; ++r) dm += s_part[r][2 * (4 * lane + (warp >> 1)) + (warp & 1)];
      }
    }
    double mine_acc = dm;
    if (hi_r > lo_r) {
      // the edges of an interior are a handful of elements: lane i walks those of position i itself, [i, lo) and
      // [hi, i + n_i), one player (no sign test) -- no warp reduction
      if (lane <= K && step + lane < len) {
        const int n_i = min(T, len - step - lane);
        for (int m = lane; m < lo_r; ++m) mine_acc += (double)tgt_reward(s_rew[m], clip) * s_disc[m - lane];
        for (int m = hi_r; m < lane + n_i; ++m) mine_acc += (double)tgt_reward(s_rew[m], clip) * s_disc[m - lane];
      }
    } else {
      mine_acc += K < 6 ? nstep_edges<6>(s_rew, s_tp, s_disc, K, T, step, len, win, lo_r, hi_r, lane, clip)
                        : nstep_edges<kTgtFastK + 1>(s_rew, s_tp, s_disc, K, T, step, len, win, lo_r, hi_r, lane, clip);
    }
    if (lane <= K) {
      const int ci = step + lane;
      float value = 0.0f;
      if (ci < len) {
        const double boot = (ci + T < len) ? __dmul_rn(root, c.disc_pow_td) : 0.0;
        value = __fadd_rn((float)boot, (float)mine_acc);  // numpy 2: python float + np.float32 -> float32
      }
      float last_reward = 0.0f;
      if (ci > 0 && ci <= len) last_reward = (lane > 0) ? tgt_reward(s_rew[lane - 1], clip) : prev_reward;
      s_val[lane] = value;
      s_lastr[lane] = last_reward;
    }
  } else {
    for (int i = 0; i <= K; ++i) {
      const int ci = step + i;
      float value = 0.0f;
      if (ci < len) {
        const double root_i = (ci + T < len && lane == 0) ? w.root_values[pos + i + T] : 0.0;
        const int tp = s_tp[i];
        const int n = min(T, len - ci);
        double acc = 0.0;
        for (int j = lane; j < n; j += 32) {
          float r = tgt_reward(s_rew[i + j], clip);
          if (s_tp[i + j] != tp) r = -r;
          acc += (double)r * s_disc[j];
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc += shfl_xor_f64<32>(acc, m);
        const double boot = (ci + T < len) ? __dmul_rn(root_i, c.disc_pow_td) : 0.0;
        value = __fadd_rn((float)boot, (float)acc);
      }
      if (lane == 0) {
        float last_reward = 0.0f;
        if (ci > 0 && ci <= len) last_reward = (i > 0) ? tgt_reward(s_rew[i - 1], clip) : prev_reward;
        s_val[i] = value;
        s_lastr[i] = last_reward;
      }
    }
  }
  __syncwarp();
  for (int i = lane; i <= K; i += 32) {
    t_values[(size_t)b * (K + 1) + i] = s_val[i];
    t_rewards[(size_t)b * (K + 1) + i] = s_lastr[i];
  }
  if (!c.fuse_supports) return;
  // learners.py:186-192: h(x) then two-hot projection of values and rewards: the row's supports are zero-filled
  // with 8-byte stores, then every position scatters its two bins (high first: the low bin wins on integers)
  const int vb = c.value_max - c.value_min + 1, rb = c.reward_max - c.reward_min + 1;
  float* vs_row = value_support + (size_t)b * (K + 1) * vb;
  float* rs_row = reward_support + (size_t)b * (K + 1) * rb;
  auto zero_fill = [&](float* dst, int n) {
    if (((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
      for (int e = lane; e < (n >> 1); e += 32) reinterpret_cast<float2*>(dst)[e] = make_float2(0.0f, 0.0f);
    } else {
      for (int e = lane; e < n; e += 32) dst[e] = 0.0f;
    }
  };
  zero_fill(vs_row, (K + 1) * vb);
  zero_fill(rs_row, (K + 1) * rb);
  __syncwarp();  // orders the fills before the scatters of the other lanes
  for (int i = lane; i <= K; i += 32) {
    const float v = s_val[i], r = s_lastr[i];
    const MzTwoHot tv = mz_two_hot(c.no_target_transform ? v : mz_scalar_transform_f(v), c.value_min, c.value_max);
    const MzTwoHot tr = mz_two_hot(c.no_target_transform ? r : mz_scalar_transform_f(r), c.reward_min, c.reward_max);
    vs_row[i * vb + tv.hi] = tv.p_hi;
    vs_row[i * vb + tv.lo] = tv.p_lo;
    rs_row[i * rb + tr.hi] = tr.p_hi;
    rs_row[i * rb + tr.lo] = tr.p_lo;
  }
}

// ---- lane-per-position variant ---------------------------------------------------------------------
// For the learner's usual shapes (K + 1 <= 16 unroll positions, td_steps <= 64) a warp owns R = 32 / (K + 1)
// consecutive sampled rows and lane (r, i) owns unroll position i of row r: its n-step sum is a short serial
// loop over the (L1-resident) reward window, the bootstrap / reward / two-hot arithmetic runs once per lane
// with no idle lanes, and the outputs of the warp's rows -- contiguous in every output array -- leave as
// warp-wide runs.  The two-hot supports are zero-filled with 8-byte stores and the two non-zero bins of every
// position scattered afterwards (high bin first: the low bin wins on integers, config.py:64-67).
// 284 warp instructions per row (ncu, C3 shape, both warp roles) against ~1000 for the warp-per-row kernel above.
constexpr int kRowsMaxPos = 16;
constexpr int kRowsMaxTd = 64;
constexpr int kRowsGroup = 4;  // rows whose loads are in flight together

__global__ void __launch_bounds__(kTgtThreads, 4)
build_targets_rows_kernel(mz_window w, mz_target_cfg c, const int64_t* __restrict__ pos_arr,
                          const int64_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_len,
                          const int32_t* __restrict__ pad_actions, float* __restrict__ obs_out,
                          int32_t* __restrict__ actions_out, float* __restrict__ t_rewards,
                          float* __restrict__ t_values, float* __restrict__ t_policies,
                          float* __restrict__ value_support, float* __restrict__ reward_support) {
  __shared__ int64_t s_pos[kTgtWarps][kRowsMaxPos];
  __shared__ int32_t s_rem[kTgtWarps][kRowsMaxPos];  // len - step of the row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = c.num_unroll_steps, T = c.td_steps, A = w.num_actions, E = w.obs_elems;
  const int NP = K + 1, R = 32 / NP;
  // two warps share a group of R rows: role 0 moves observations / policies / actions, role 1 builds the
  // value / reward targets and their supports.  The two dependent-load chains run side by side.
  const int role = warp & 1;
  const int row0 = (blockIdx.x * (kTgtWarps / 2) + (warp >> 1)) * R;
  if (row0 >= c.batch) return;
  const int nrows = min(R, c.batch - row0);
  const int r = lane / NP, i = lane - r * NP;
  const bool active = r < nrows;
  const int b = row0 + r;
  int64_t pos = 0;
  int step = 0, len = 0;
  if (active) {
    pos = pos_arr[b];
    step = (int)(pos - chunk_start[b]);
    len = chunk_len[b];  // len(root_values) == len(rewards) == len(to_play)
    if (i == 0) {
      s_pos[warp][r] = pos;
      s_rem[warp][r] = len - step;
    }
  }
  __syncwarp();

  // ---- observations: np.float32(history.observations[step]) (replay_buffer.py:147).  Rows go in groups of
  // kRowsGroup: all loads of a group are issued before its first store (one DRAM round trip per group)
  if (role == 0) {
    const bool vec = (E & 3) == 0 && ((reinterpret_cast<uintptr_t>(w.obs) | reinterpret_cast<uintptr_t>(obs_out)) & 15) == 0 &&
                     (!c.normalize_obs || ((reinterpret_cast<uintptr_t>(c.obs_min) | reinterpret_cast<uintptr_t>(c.obs_range)) & 15) == 0);
    const uint8_t* obs_u8 = reinterpret_cast<const uint8_t*>(w.obs);
    const float* obs_f32 = reinterpret_cast<const float*>(w.obs);
    for (int rr0 = 0; rr0 < nrows; rr0 += kRowsGroup) {
      int64_t p[kRowsGroup];
#pragma unroll
      for (int u = 0; u < kRowsGroup; ++u) p[u] = s_pos[warp][min(rr0 + u, nrows - 1)];
      if (vec && w.obs_is_u8) {
        for (int e = lane; e < (E >> 2); e += 32) {
          uint32_t q[kRowsGroup];
#pragma unroll
          for (int u = 0; u < kRowsGroup; ++u) q[u] = __ldg(reinterpret_cast<const uint32_t*>(obs_u8 + (size_t)p[u] * E) + e);
#pragma unroll
          for (int u = 0; u < kRowsGroup; ++u) {
            if (rr0 + u >= nrows) break;
            float4 v = make_float4((float)(q[u] & 255u), (float)((q[u] >> 8) & 255u), (float)((q[u] >> 16) & 255u),
                                   (float)(q[u] >> 24));
            if (c.normalize_obs) {
              const float4 mn = reinterpret_cast<const float4*>(c.obs_min)[e];
              const float4 rg = reinterpret_cast<const float4*>(c.obs_range)[e];
              v.x = __fdiv_rn(__fsub_rn(v.x, mn.x), rg.x);
              v.y = __fdiv_rn(__fsub_rn(v.y, mn.y), rg.y);
              v.z = __fdiv_rn(__fsub_rn(v.z, mn.z), rg.z);
              v.w = __fdiv_rn(__fsub_rn(v.w, mn.w), rg.w);
            }
            reinterpret_cast<float4*>(obs_out + (size_t)(row0 + rr0 + u) * E)[e] = v;
          }
        }
      } else if (vec) {
        for (int e = lane; e < (E >> 2); e += 32) {
          float4 q[kRowsGroup];
#pragma unroll
          for (int u = 0; u < kRowsGroup; ++u) q[u] = __ldg(reinterpret_cast<const float4*>(obs_f32 + (size_t)p[u] * E) + e);
#pragma unroll
          for (int u = 0; u < kRowsGroup; ++u) {
            if (rr0 + u >= nrows) break;
            float4 v = q[u];
            if (c.normalize_obs) {
              const float4 mn = reinterpret_cast<const float4*>(c.obs_min)[e];
              const float4 rg = reinterpret_cast<const float4*>(c.obs_range)[e];
              v.x = __fdiv_rn(__fsub_rn(v.x, mn.x), rg.x);
              v.y = __fdiv_rn(__fsub_rn(v.y, mn.y), rg.y);
              v.z = __fdiv_rn(__fsub_rn(v.z, mn.z), rg.z);
              v.w = __fdiv_rn(__fsub_rn(v.w, mn.w), rg.w);
            }
            reinterpret_cast<float4*>(obs_out + (size_t)(row0 + rr0 + u) * E)[e] = v;
          }
        }
      } else {
        for (int e = lane; e < E; e += 32) {
          float q[kRowsGroup];
#pragma unroll
          for (int u = 0; u < kRowsGroup; ++u)
            q[u] = w.obs_is_u8 ? (float)__ldg(obs_u8 + (size_t)p[u] * E + e) : __ldg(obs_f32 + (size_t)p[u] * E + e);
#pragma unroll
          for (int u = 0; u < kRowsGroup; ++u) {
            if (rr0 + u >= nrows) break;
            float v = q[u];
            if (c.normalize_obs) v = __fdiv_rn(__fsub_rn(v, c.obs_min[e]), c.obs_range[e]);
            obs_out[(size_t)(row0 + rr0 + u) * E + e] = v;
          }
        }
      }
    }
  }
  // ---- policy targets: child_visits of the positions inside the chunk, zeros after its end
  // (replay_buffer.py:90, 184, 196); the (K+1) * A floats of a row are one contiguous run on both sides
  for (int rr0 = 0; rr0 < nrows && role == 0; rr0 += kRowsGroup) {
    const float* src[kRowsGroup];
    int n_real[kRowsGroup];
#pragma unroll
    for (int u = 0; u < kRowsGroup; ++u) {
      const int rr = min(rr0 + u, nrows - 1);
      src[u] = w.child_visits + (size_t)s_pos[warp][rr] * A;
      n_real[u] = max(0, min(NP, s_rem[warp][rr])) * A;
    }
    for (int e = lane; e < NP * A; e += 32) {
      float q[kRowsGroup];
#pragma unroll
      for (int u = 0; u < kRowsGroup; ++u) q[u] = e < n_real[u] ? __ldg(src[u] + e) : 0.0f;
#pragma unroll
      for (int u = 0; u < kRowsGroup; ++u) {
        if (rr0 + u >= nrows) break;
        t_policies[(size_t)(row0 + rr0 + u) * NP * A + e] = q[u];
      }
    }
  }
  // ---- actions: history.actions[step:step+K], padded with random actions (replay_buffer.py:149-152)
  if (role == 0 && active && i < K) {
    const int n_real = max(0, min(K, len - step));
    actions_out[(size_t)b * K + i] = i < n_real ? w.actions[pos + i] : pad_actions[(size_t)b * K + (i - n_real)];
  }

  if (role == 0) return;
  // ---- insert_target (replay_buffer.py:165-198) for position i of row r
  const int ci = step + i;
  float value = 0.0f, last_reward = 0.0f;
  if (active) {
    if (ci < len) {
      const double root = (ci + T < len) ? w.root_values[pos + i + T] : 0.0;  // replay_buffer.py:180-183
      const int n = min(T, len - ci);
      const float* rw = w.rewards + pos + i;
      const int8_t* tp = w.to_play + pos + i;
      const int tp0 = __ldg(tp);
      double acc = 0.0;  // exact products of float32 pairs, accumulated in binary64
      for (int j0 = 0; j0 < n; j0 += 8) {  // eight window elements in flight
        float x[8];
        int t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = min(j0 + u, n - 1);
          x[u] = tgt_reward(__ldg(rw + j), w.clip_rewards);
          t[u] = __ldg(tp + j);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (j0 + u < n)  // replay_buffer.py:187-189
            acc += (double)(t[u] != tp0 ? -x[u] : x[u]) * (double)__ldg(c.discounts + j0 + u);
        }
      }
      const double boot = (ci + T < len) ? __dmul_rn(root, c.disc_pow_td) : 0.0;
      value = __fadd_rn((float)boot, (float)acc);  // numpy 2: python float + np.float32 -> float32
    }
    if (ci > 0 && ci <= len) last_reward = tgt_reward(w.rewards[pos + i - 1], w.clip_rewards);
    t_values[(size_t)b * NP + i] = value;
    t_rewards[(size_t)b * NP + i] = last_reward;
  }
  if (!c.fuse_supports) return;

  // ---- learners.py:186-192: h(x) then two-hot projection of values and rewards
  const int vb = c.value_max - c.value_min + 1, rb = c.reward_max - c.reward_min + 1;
  auto zero_fill = [&](float* dst, int n) {
    if (((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
      for (int e = lane; e < (n >> 1); e += 32) reinterpret_cast<float2*>(dst)[e] = make_float2(0.0f, 0.0f);
    } else {
      for (int e = lane; e < n; e += 32) dst[e] = 0.0f;
    }
  };
  zero_fill(value_support + (size_t)row0 * NP * vb, nrows * NP * vb);
  zero_fill(reward_support + (size_t)row0 * NP * rb, nrows * NP * rb);
  __syncwarp();  // orders the fills before the scatters of the other lanes
  if (active) {
    const MzTwoHot tv = mz_two_hot(c.no_target_transform ? value : mz_scalar_transform_f(value), c.value_min, c.value_max);
    const MzTwoHot tr = mz_two_hot(c.no_target_transform ? last_reward : mz_scalar_transform_f(last_reward), c.reward_min,
                                   c.reward_max);
    float* pv = value_support + ((size_t)b * NP + i) * vb;
    pv[tv.hi] = tv.p_hi;
    pv[tv.lo] = tv.p_lo;
    float* pr = reward_support + ((size_t)b * NP + i) * rb;
    pr[tr.hi] = tr.p_hi;
    pr[tr.lo] = tr.p_lo;
  }
}

// ---- TMA-staged variant ------------------------------------------------------------------------------
// The learner's shapes again (K + 1 <= 16, td_steps <= 64), with the window staged through shared memory by the
// bulk-copy engine.  A CTA owns kTmaRows consecutive sampled rows and keeps a shared-memory image of EVERY output
// run of those rows (observations, actions, reward / value / policy targets, both supports: contiguous in each
// output array, a multiple of 16 bytes because kTmaRows is a multiple of four):
//   1. one lane per row issues cp.async.bulk (global -> shared, mbarrier complete_tx) for the row's observation and
//      its run of child-visit rows -- the latter straight into the policy image, float32 observations straight into
//      the observation image -- so the bulky reads of all rows are in flight at once without holding registers;
//   2. meanwhile all threads fetch the small windows (rewards from pos - 1, to_play, bootstrap root values, actions)
//      with one independent load per element, zero the supports and the policy tails behind the chunk's end;
//   3. one thread per (row, unroll position) does insert_target's arithmetic out of shared memory (same operation
//      order as the lane-per-position kernel: bit-identical outputs) and scatters the two-hot bins;
//   4. bytes -> float32 (+ normalisation) after the mbarrier flips, fence.proxy.async, and ONE thread hands the seven
//      images to the bulk-copy engine (shared -> global) and waits until they have been read.
// ~45 warp instructions per row against 284, no store instruction on the output path.  A last, partial CTA (or
// unaligned output pointers) writes its images with ordinary stores.
constexpr int kTmaMaxThreads = 256;

struct TmaPlan {  // byte offsets into dynamic shared memory, every one a multiple of 16
  int obs_out, pol, act, val, rew, vs, rs, obs_in, rw, tp, root, pos, bar;
  int obs_bulk, cv_bulk, out_bulk;
  int rows;       // rows per CTA: 4, 8, 16 or 32 (every output run of a CTA is then a multiple of 16 bytes)
  int lpr_shift;  // log2(threads / rows): the lanes that serve one row
};

MZ_DEV void bulk_copy_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
// 4- / 8-byte asynchronous copies global -> shared (LDGSTS): nothing waits on the loaded value, so a thread's
// copies of several windows overlap; !valid writes zeros without reading
MZ_DEV void cp_async4(void* dst_smem, const void* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
MZ_DEV void cp_async8(void* dst_smem, const void* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
MZ_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
MZ_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
MZ_DEV void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
MZ_DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kTmaMaxThreads)
build_targets_tma_kernel(mz_window w, mz_target_cfg c, const int64_t* __restrict__ pos_arr,
                         const int64_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_len,
                         const int32_t* __restrict__ pad_actions, float* __restrict__ obs_out,
                         int32_t* __restrict__ actions_out, float* __restrict__ t_rewards,
                         float* __restrict__ t_values, float* __restrict__ t_policies,
                         float* __restrict__ value_support, float* __restrict__ reward_support, TmaPlan pl) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int tid = threadIdx.x, nthr = blockDim.x, RB = pl.rows;
  const int K = c.num_unroll_steps, T = c.td_steps, A = w.num_actions, E = w.obs_elems;
  const int NP = K + 1, KT = K + T, RW = KT + 1;  // RW: rewards from pos - 1 on
  const int row0 = blockIdx.x * RB;
  const int nrows = min(RB, c.batch - row0);
  const int vb = c.value_max - c.value_min + 1, rb = c.reward_max - c.reward_min + 1;
  float* s_obs = reinterpret_cast<float*>(sm + pl.obs_out);
  float* s_pol = reinterpret_cast<float*>(sm + pl.pol);
  int32_t* s_act = reinterpret_cast<int32_t*>(sm + pl.act);
  float* s_val = reinterpret_cast<float*>(sm + pl.val);
  float* s_rew = reinterpret_cast<float*>(sm + pl.rew);
  float* s_vs = reinterpret_cast<float*>(sm + pl.vs);
  float* s_rs = reinterpret_cast<float*>(sm + pl.rs);
  uint8_t* s_obs_in = sm + pl.obs_in;
  float* s_rw = reinterpret_cast<float*>(sm + pl.rw);
  int8_t* s_tp = reinterpret_cast<int8_t*>(sm + pl.tp);
  double* s_root = reinterpret_cast<double*>(sm + pl.root);
  int32_t* s_step = reinterpret_cast<int32_t*>(sm + pl.pos);
  int32_t* s_len = s_step + RB;
  int32_t* s_tpoff = s_len + RB;
  const int TPS = (KT + 6) & ~3;  // to_play bytes per row: the window plus its 4-byte alignment slack
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + pl.bar);

  // thread (r, q0): lane q0 of the LPR = blockDim / rows lanes that serve row r (both powers of two)
  const int LPR = 1 << pl.lpr_shift;
  const int r = tid >> pl.lpr_shift, q0 = tid & (LPR - 1);
  if (tid == 0) {
    mbar_init(bar, nrows);  // one arrive.expect_tx per row
    mbar_fence_init();
  }
  int64_t pos = 0;
  int step = 0, len = 0;
  if (r < nrows) {
    pos = pos_arr[row0 + r];
    step = (int)(pos - chunk_start[row0 + r]);
    len = chunk_len[row0 + r];  // len(root_values) == len(rewards) == len(to_play)
    if (q0 == 0) {
      s_step[r] = step;
      s_len[r] = len;
      s_tpoff[r] = (int)(pos & 3);
    }
  }
  __syncthreads();  // the mbarrier is initialised

  // ---- 1. bulk loads: the first lane of a row sends for its observation and its run of child-visit rows
  const int rem = len - step;  // positions from `step` to the chunk's end
  const int n_pol = max(0, min(NP, rem)) * A;
  if (r < nrows && q0 == 0) {
    const uint32_t obs_row_bytes = (uint32_t)E * (w.obs_is_u8 ? 1u : 4u);
    const uint32_t cv_bytes = (uint32_t)n_pol * 4u;
    mbar_arrive_expect_tx(bar, (pl.obs_bulk ? obs_row_bytes : 0u) + (pl.cv_bulk ? cv_bytes : 0u));
    if (pl.obs_bulk) {
      void* dst = w.obs_is_u8 ? static_cast<void*>(s_obs_in + (size_t)r * E) : static_cast<void*>(s_obs + (size_t)r * E);
      bulk_copy_g2s(dst, reinterpret_cast<const uint8_t*>(w.obs) + (size_t)pos * obs_row_bytes, obs_row_bytes, bar);
    }
    if (pl.cv_bulk && cv_bytes) bulk_copy_g2s(s_pol + (size_t)r * NP * A, w.child_visits + (size_t)pos * A, cv_bytes, bar);
  }

  // ---- 2. small windows: asynchronous 4- / 8-byte copies (no thread waits on a loaded value, so the windows of a
  // row are one round trip), to_play bytes through registers; zero fills
  if (r < nrows) {
    {  // to_play: the 4-byte words that cover bytes [pos, pos + min(K + T, rem)); byte j of the window then sits at
       // s_tp[r * TPS + (pos & 3) + j].  The last word is read only as far as the window reaches.
      const int off = (int)(pos & 3), nbytes = off + max(0, min(KT, rem));
      const int8_t* base = w.to_play + (pos - off);
      for (int q = q0; q * 4 < off + KT; q += LPR) {
        // a whole word of the array -- also for the window's last, partial word while that word still lies inside
        // the chunk (bytes behind the window are read but never used) --, or nothing (zero fill)
        const int have = max(0, min(4, nbytes - q * 4));
        if (have == 0 || q * 4 + 4 <= off + rem) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(s_tp + r * TPS + q * 4)),
                       "l"(have > 0 ? base + q * 4 : w.to_play), "r"(have > 0 ? 4 : 0)
                       : "memory");
        } else {  // the chunk (and possibly the array) ends inside this word: byte loads
          for (int bb = 0; bb < 4; ++bb) s_tp[r * TPS + q * 4 + bb] = bb < have ? __ldg(base + q * 4 + bb) : (int8_t)0;
        }
      }
    }
    for (int q = q0; q < RW; q += LPR) {  // raw rewards pos - 1 .. pos + K + T - 1 (clipped where they are read)
      const int ci = step + q - 1;
      const bool ok = ci >= 0 && ci < len;
      cp_async4(s_rw + r * RW + q, ok ? w.rewards + pos + q - 1 : w.rewards, ok);
    }
    const int n_act = max(0, min(K, rem));
    for (int i = q0; i < NP; i += LPR) {
      const bool ok = i + T < rem;  // replay_buffer.py:180-183
      cp_async8(s_root + r * NP + i, ok ? w.root_values + pos + i + T : w.root_values, ok);
      if (i < K)  // replay_buffer.py:149-152
        cp_async4(s_act + r * K + i, i < n_act ? w.actions + pos + i : pad_actions + (size_t)(row0 + r) * K + (i - n_act), true);
    }
    if (pl.cv_bulk) {  // absorbing policy behind the chunk's end
      for (int q = n_pol + q0; q < NP * A; q += LPR) s_pol[r * NP * A + q] = 0.0f;
    } else {
      for (int q = q0; q < NP * A; q += LPR) 
        cp_async4(s_pol + r * NP * A + q, q < n_pol ? w.child_visits + (size_t)pos * A + q : w.child_visits, q < n_pol);
    }
  }
  if (c.fuse_supports) {  // both images are multiples of 16 bytes (rows % 4 == 0) and adjacent: one fill
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* z = reinterpret_cast<float4*>(s_vs) + tid;
    float4* const zend = reinterpret_cast<float4*>(s_vs) + ((RB * NP * (vb + rb)) >> 2);
    for (; z + 3 * nthr < zend; z += 4 * nthr) {  // four stores per trip
      z[0] = z4;
      z[nthr] = z4;
      z[2 * nthr] = z4;
      z[3 * nthr] = z4;
    }
    for (; z < zend; z += nthr) *z = z4;
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- 3. insert_target (replay_buffer.py:165-198) for position i of row ri, then learners.py:186-192
  for (int it = tid; it < nrows * NP; it += nthr) {
    const int ri = it / NP, i = it - ri * NP;
    const int len_i = s_len[ri], ci = s_step[ri] + i;
    float value = 0.0f, last_reward = 0.0f;
    if (ci < len_i) {
      const int n = min(T, len_i - ci);
      const float* rw = s_rw + ri * RW + 1 + i;
      const int8_t* tp = s_tp + ri * TPS + s_tpoff[ri] + i;
      const int tp0 = tp[0];
      double acc = 0.0;  // exact products of float32 pairs, accumulated in binary64, j ascending
      for (int j = 0; j < n; ++j) {  // replay_buffer.py:187-189
        const float x = tgt_reward(rw[j], w.clip_rewards);
        acc += (double)(tp[j] != tp0 ? -x : x) * (double)__ldg(c.discounts + j);
      }
      const double boot = (ci + T < len_i) ? __dmul_rn(s_root[it], c.disc_pow_td) : 0.0;
      value = __fadd_rn((float)boot, (float)acc);  // numpy 2: python float + np.float32 -> float32
    }
    if (ci > 0 && ci <= len_i) last_reward = tgt_reward(s_rw[ri * RW + i], w.clip_rewards);
    s_val[it] = value;
    s_rew[it] = last_reward;
    if (c.fuse_supports) {
      const MzTwoHot tv = mz_two_hot(c.no_target_transform ? value : mz_scalar_transform_f(value), c.value_min, c.value_max);
      const MzTwoHot tr = mz_two_hot(c.no_target_transform ? last_reward : mz_scalar_transform_f(last_reward), c.reward_min,
                                     c.reward_max);
      float* pv = s_vs + (size_t)it * vb;
      pv[tv.hi] = tv.p_hi;
      pv[tv.lo] = tv.p_lo;  // the low bin wins on integers (config.py:64-67)
      float* pr = s_rs + (size_t)it * rb;
      pr[tr.hi] = tr.p_hi;
      pr[tr.lo] = tr.p_lo;
    }
  }

  // ---- 4. observations: np.float32(history.observations[step]) (replay_buffer.py:147)
  mbar_wait(bar, 0);
  if (r < nrows) {
    float* so = s_obs + (size_t)r * E;
    if (w.obs_is_u8 && pl.obs_bulk) {  // E % 16 == 0
      const uint32_t* in = reinterpret_cast<const uint32_t*>(s_obs_in + (size_t)r * E);
      for (int e = q0; e < (E >> 2); e += LPR) {
        const uint32_t q = in[e];
        float4 v = make_float4((float)(q & 255u), (float)((q >> 8) & 255u), (float)((q >> 16) & 255u), (float)(q >> 24));
        if (c.normalize_obs) {
          const int k = e * 4;
          v.x = __fdiv_rn(__fsub_rn(v.x, c.obs_min[k]), c.obs_range[k]);
          v.y = __fdiv_rn(__fsub_rn(v.y, c.obs_min[k + 1]), c.obs_range[k + 1]);
          v.z = __fdiv_rn(__fsub_rn(v.z, c.obs_min[k + 2]), c.obs_range[k + 2]);
          v.w = __fdiv_rn(__fsub_rn(v.w, c.obs_min[k + 3]), c.obs_range[k + 3]);
        }
        reinterpret_cast<float4*>(so)[e] = v;
      }
    } else if (w.obs_is_u8) {
      const uint8_t* in = reinterpret_cast<const uint8_t*>(w.obs) + (size_t)pos * E;
      for (int k = q0; k < E; k += LPR) {
        float v = (float)__ldg(in + k);
        if (c.normalize_obs) v = __fdiv_rn(__fsub_rn(v, c.obs_min[k]), c.obs_range[k]);
        so[k] = v;
      }
    } else if (!pl.obs_bulk || c.normalize_obs) {
      const float* in = reinterpret_cast<const float*>(w.obs) + (size_t)pos * E;
      for (int k = q0; k < E; k += LPR) {
        float v = pl.obs_bulk ? so[k] : __ldg(in + k);
        if (c.normalize_obs) v = __fdiv_rn(__fsub_rn(v, c.obs_min[k]), c.obs_range[k]);
        so[k] = v;
      }
    }
  }

  // ---- 5. the images leave
  if (pl.out_bulk && nrows == RB) {
    fence_proxy_async_smem();  // generic-proxy writes of this thread -> visible to the bulk-copy engine
    __syncthreads();
    if (tid == 0) {
      bulk_copy_s2g(obs_out + (size_t)row0 * E, s_obs, (uint32_t)(RB * E) * 4u);
      bulk_copy_s2g(t_policies + (size_t)row0 * NP * A, s_pol, (uint32_t)(RB * NP * A) * 4u);
      if (K > 0) bulk_copy_s2g(actions_out + (size_t)row0 * K, s_act, (uint32_t)(RB * K) * 4u);
      bulk_copy_s2g(t_values + (size_t)row0 * NP, s_val, (uint32_t)(RB * NP) * 4u);
      bulk_copy_s2g(t_rewards + (size_t)row0 * NP, s_rew, (uint32_t)(RB * NP) * 4u);
      if (c.fuse_supports) {
        bulk_copy_s2g(value_support + (size_t)row0 * NP * vb, s_vs, (uint32_t)(RB * NP * vb) * 4u);
        bulk_copy_s2g(reward_support + (size_t)row0 * NP * rb, s_rs, (uint32_t)(RB * NP * rb) * 4u);
      }
      bulk_commit();
      bulk_wait_read_all();  // shared memory must stay until the engine has read it
    }
    return;
  }
  __syncthreads();
  auto flush = [&](float* dst, const float* src, int n) {
    for (int e = tid; e < n; e += nthr) dst[e] = src[e];
  };
  flush(obs_out + (size_t)row0 * E, s_obs, nrows * E);
  flush(t_policies + (size_t)row0 * NP * A, s_pol, nrows * NP * A);
  flush(reinterpret_cast<float*>(actions_out) + (size_t)row0 * K, reinterpret_cast<const float*>(s_act), nrows * K);
  flush(t_values + (size_t)row0 * NP, s_val, nrows * NP);
  flush(t_rewards + (size_t)row0 * NP, s_rew, nrows * NP);
  if (c.fuse_supports) {
    flush(value_support + (size_t)row0 * NP * vb, s_vs, nrows * NP * vb);
    flush(reward_support + (size_t)row0 * NP * rb, s_rs, nrows * NP * rb);
  }
}

__global__ void scalar_transform_kernel(long long n, const float* __restrict__ x,
                                        float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = mz_scalar_transform_f(x[i]);
}

__global__ void scalar_to_support_kernel(long long n, float* __restrict__ x, int mn, int mx,
                                         int clamp_in_place, float* __restrict__ support) {
  const int bins = mx - mn + 1;
  const long long total = n * bins;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / bins;
    const int j = (int)(idx % bins);
    const MzTwoHot th = mz_two_hot(x[i], mn, mx);
    support[idx] = j == th.lo ? th.p_lo : (j == th.hi ? th.p_hi : 0.0f);
  }
}

__global__ void clamp_kernel(long long n, float* __restrict__ x, int mn, int mx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    x[i] = fminf(fmaxf(x[i], (float)mn), (float)mx);
}

__global__ void support_to_scalar_kernel(long long n, const float* __restrict__ logits, int mn,
                                         int mx, int no_tt, float* __restrict__ out) {
  const int bins = mx - mn + 1;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const float v = mz_support_to_scalar_warp(logits + row * bins, bins, mn, no_tt, lane);
    if (lane == 0) out[row] = v;
  }
}

// ---- self-play -> replay hand-off on the device ------------------------------------------------------
// One trajectory step of every game written straight into the replay window (Game.apply +
// store_search_statistics, game.py:79-115): a warp per game, coalesced row copies.
__global__ void window_append_kernel(mz_window w, int G, const int64_t* __restrict__ dst_pos, const void* __restrict__ obs,
                                     const int32_t* __restrict__ actions, const float* __restrict__ rewards,
                                     const int8_t* __restrict__ to_play, const double* __restrict__ root_values,
                                     const double* __restrict__ child_visits) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= G) return;
  const int64_t pos = dst_pos[g];
  if (pos < 0) return;
  const int E = w.obs_elems, A = w.num_actions;
  if (w.obs_is_u8) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(obs) + (size_t)g * E;
    uint8_t* dst = reinterpret_cast<uint8_t*>(const_cast<void*>(w.obs)) + (size_t)pos * E;
    if ((E & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0) {
      for (int e = lane; e < (E >> 2); e += 32) reinterpret_cast<uint32_t*>(dst)[e] = reinterpret_cast<const uint32_t*>(src)[e];
    } else {
      for (int e = lane; e < E; e += 32) dst[e] = src[e];
    }
  } else {
    const float* src = reinterpret_cast<const float*>(obs) + (size_t)g * E;
    float* dst = reinterpret_cast<float*>(const_cast<void*>(w.obs)) + (size_t)pos * E;
    for (int e = lane; e < E; e += 32) dst[e] = src[e];
  }
  float* cv = const_cast<float*>(w.child_visits) + (size_t)pos * A;
  for (int a = lane; a < A; a += 32) cv[a] = (float)child_visits[(size_t)g * A + a];  // float32(history.child_visits)
  if (lane == 0) {
    const_cast<int32_t*>(w.actions)[pos] = actions[g];
    const_cast<float*>(w.rewards)[pos] = rewards[g];
    const_cast<int8_t*>(w.to_play)[pos] = to_play[g];
    const_cast<double*>(w.root_values)[pos] = root_values[g];
  }
}

// (src, dst, n) runs of window positions: the overlap a running game's next chunk starts with (actors.py:160-166)
__global__ void window_copy_kernel(mz_window w, int count, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                   const int32_t* __restrict__ n) {
  const int r = blockIdx.x;
  if (r >= count) return;
  const int E = w.obs_elems, A = w.num_actions;
  const int64_t s0 = src[r], d0 = dst[r];
  const int len = n[r];
  const size_t obs_bytes = (size_t)E * (w.obs_is_u8 ? 1 : 4);
  const uint8_t* os = reinterpret_cast<const uint8_t*>(w.obs) + (size_t)s0 * obs_bytes;
  uint8_t* od = reinterpret_cast<uint8_t*>(const_cast<void*>(w.obs)) + (size_t)d0 * obs_bytes;
  for (size_t i = threadIdx.x; i < (size_t)len * obs_bytes; i += blockDim.x) od[i] = os[i];
  float* cvd = const_cast<float*>(w.child_visits) + (size_t)d0 * A;
  for (int i = threadIdx.x; i < len * A; i += blockDim.x) cvd[i] = w.child_visits[(size_t)s0 * A + i];
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const_cast<int32_t*>(w.actions)[d0 + i] = w.actions[s0 + i];
    const_cast<float*>(w.rewards)[d0 + i] = w.rewards[s0 + i];
    const_cast<int8_t*>(w.to_play)[d0 + i] = w.to_play[s0 + i];
    const_cast<double*>(w.root_values)[d0 + i] = w.root_values[s0 + i];
  }
}

// padding actions of sample_batch (np.random.randint(action_space), replay_buffer.py:151) drawn on the device:
// splitmix64 of (seed, index), reduced to [0, A) by multiply-shift
__global__ void pad_actions_kernel(int n, int A, unsigned long long seed, int32_t* __restrict__ pads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  pads[i] = (int32_t)(((z >> 32) * (unsigned long long)A) >> 32);
}

int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int mz_sumtree_sample_mt(const double* tree, int64_t max_capacity, int32_t n, const double* u01,
                                    int32_t u01_is_mt_words, const int64_t* slot_pos, const int64_t* slot_start,
                                    const int32_t* slot_len, int64_t num_memories, double beta, int64_t* tree_idx,
                                    double* priority, int64_t* pos, int64_t* chunk_start, int32_t* chunk_len,
                                    double* is_weights, void* stream);
int g_tma_rows = 8, g_tma_threads = 128;
// 0: by shape, 1: always the warp-per-row kernel, 2: lane-per-position kernel where the shape allows it (never the
// TMA-staged one), 3: TMA-staged kernel where the shape allows it (tests run all on the same inputs)
int g_targets_kernel = 0;

extern "C" {

int mz_debug_set_targets_kernel(int32_t which) {
  g_targets_kernel = which;
  return MZ_OK;
}

int mz_debug_set_targets_tma(int32_t rows_per_cta, int32_t threads) {
  if (rows_per_cta < 4 || rows_per_cta > 32 || (rows_per_cta & (rows_per_cta - 1)) || threads < 32 ||
      threads > kTmaMaxThreads || (threads & (threads - 1)) || threads < rows_per_cta)
    return MZ_ERR_BAD_ARG;
  g_tma_rows = rows_per_cta;
  g_tma_threads = threads;
  return MZ_OK;
}

int mz_scalar_transform(int64_t n, const float* x, float* out, void* stream) {
  if (n < 0 || (n > 0 && (!x || !out))) return MZ_ERR_BAD_ARG;
  if (n == 0) return MZ_OK;
  scalar_transform_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x, out);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_scalar_to_support(int64_t n, float* x, int32_t mn, int32_t mx, int32_t clamp_in_place,
                         float* support, void* stream) {
  if (n < 0 || mx < mn || (n > 0 && (!x || !support))) return MZ_ERR_BAD_ARG;
  if (n == 0) return MZ_OK;
  const int bins = mx - mn + 1;
  scalar_to_support_kernel<<<grid_for(n * bins, 256), 256, 0, (cudaStream_t)stream>>>(
      n, x, mn, mx, clamp_in_place, support);
  MZ_LAUNCH_CHECK();
  if (clamp_in_place) {  // x.clamp_(min, max) config.py:57 (after the projection read x)
    clamp_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x, mn, mx);
    MZ_LAUNCH_CHECK();
  }
  return MZ_OK;
}

int mz_support_to_scalar(int64_t n, const float* logits, int32_t mn, int32_t mx, int32_t no_tt,
                         float* out, void* stream) {
  if (n < 0 || mx < mn || (n > 0 && (!logits || !out))) return MZ_ERR_BAD_ARG;
  if (n == 0) return MZ_OK;
  support_to_scalar_kernel<<<grid_for(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(n, logits, mn, mx,
                                                                                    no_tt, out);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_window_append(const mz_window* w, int32_t num_games, const int64_t* dst_pos, const void* obs,
                     const int32_t* actions, const float* rewards, const int8_t* to_play, const double* root_values,
                     const double* child_visits, void* stream) {
  if (!w || num_games < 1 || !dst_pos || !obs || !actions || !rewards || !to_play || !root_values || !child_visits)
    return MZ_ERR_BAD_ARG;
  if (!w->obs || !w->actions || !w->rewards || !w->to_play || !w->root_values || !w->child_visits) return MZ_ERR_BAD_ARG;
  window_append_kernel<<<(num_games + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*w, num_games, dst_pos, obs, actions, rewards,
                                                                            to_play, root_values, child_visits);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_window_copy(const mz_window* w, int32_t count, const int64_t* src, const int64_t* dst, const int32_t* n,
                   void* stream) {
  if (!w || count < 0 || (count > 0 && (!src || !dst || !n))) return MZ_ERR_BAD_ARG;
  if (count == 0) return MZ_OK;
  window_copy_kernel<<<count, 128, 0, (cudaStream_t)stream>>>(*w, count, src, dst, n);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_build_targets(const mz_window* w, const mz_target_cfg* c, const int64_t* pos,
                     const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* pad_actions,
                     float* obs_out, int32_t* actions_out, float* t_rewards, float* t_values,
                     float* t_policies, float* value_support, float* reward_support, void* stream) {
  if (!w || !c || !pos || !chunk_start || !chunk_len || !obs_out || !actions_out || !t_rewards ||
      !t_values || !t_policies)
    return MZ_ERR_BAD_ARG;
  if (c->batch < 1 || c->num_unroll_steps < 0 || c->num_unroll_steps > kTgtThreads ||
      c->td_steps < 1 || !c->discounts)
    return MZ_ERR_BAD_ARG;
  if (c->num_unroll_steps > 0 && !pad_actions) return MZ_ERR_BAD_ARG;
  if (c->fuse_supports && (!value_support || !reward_support)) return MZ_ERR_BAD_ARG;
  if (c->normalize_obs && (!c->obs_min || !c->obs_range)) return MZ_ERR_BAD_ARG;
  if (!w->obs || !w->actions || !w->rewards || !w->to_play || !w->root_values || !w->child_visits)
    return MZ_ERR_BAD_ARG;
  const int K1 = c->num_unroll_steps + 1, KT = c->num_unroll_steps + c->td_steps;
  if (K1 <= kRowsMaxPos && c->td_steps <= kRowsMaxTd && g_targets_kernel != 1 && g_targets_kernel != 2) {
    // TMA-staged kernel: shared-memory images of `rows` rows of every output
    const int A = w->num_actions, E = w->obs_elems;
    const int vb = c->fuse_supports ? c->value_max - c->value_min + 1 : 0;
    const int rb = c->fuse_supports ? c->reward_max - c->reward_min + 1 : 0;
    TmaPlan pl;
    const int kTmaRows = g_tma_rows, kTmaThreads = g_tma_threads;
    pl.rows = kTmaRows;
    pl.lpr_shift = 0;
    while ((kTmaRows << (pl.lpr_shift + 1)) <= kTmaThreads) ++pl.lpr_shift;
    size_t off = 0;
    auto take = [&](size_t bytes) {
      const size_t o = off;
      off += (bytes + 15) / 16 * 16;
      return (int)o;
    };
    pl.obs_out = take((size_t)kTmaRows * E * 4);
    pl.pol = take((size_t)kTmaRows * K1 * A * 4);
    pl.act = take((size_t)kTmaRows * c->num_unroll_steps * 4);
    pl.val = take((size_t)kTmaRows * K1 * 4);
    pl.rew = take((size_t)kTmaRows * K1 * 4);
    pl.vs = take((size_t)kTmaRows * K1 * vb * 4);
    pl.rs = take((size_t)kTmaRows * K1 * rb * 4);
    pl.obs_in = take(w->obs_is_u8 ? (size_t)kTmaRows * E : 0);
    pl.rw = take((size_t)kTmaRows * (KT + 1) * 4);
    pl.tp = take((size_t)kTmaRows * ((KT + 6) & ~3));
    pl.root = take((size_t)kTmaRows * K1 * 8);
    pl.pos = take((size_t)kTmaRows * 16);
    pl.bar = take(8);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const size_t obs_row_bytes = (size_t)E * (w->obs_is_u8 ? 1 : 4);
    pl.obs_bulk = (obs_row_bytes % 16 == 0) && al16(w->obs);
    pl.cv_bulk = (A % 4 == 0) && al16(w->child_visits);
    pl.out_bulk = al16(obs_out) && al16(actions_out) && al16(t_rewards) && al16(t_values) && al16(t_policies) &&
                  al16(value_support) && al16(reward_support);
    if (off <= 96 * 1024 && (reinterpret_cast<uintptr_t>(w->to_play) & 3) == 0) {  // several CTAs per SM, so that loads, arithmetic and stores of different CTAs overlap
      static bool tma_attr_set = false;
      if (!tma_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(build_targets_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return (int)e;
        tma_attr_set = true;
      }
      build_targets_tma_kernel<<<(c->batch + kTmaRows - 1) / kTmaRows, kTmaThreads, off, (cudaStream_t)stream>>>(
          *w, *c, pos, chunk_start, chunk_len, pad_actions, obs_out, actions_out, t_rewards, t_values, t_policies,
          value_support, reward_support, pl);
      MZ_LAUNCH_CHECK();
      return MZ_OK;
    }
  }
  if (K1 <= kRowsMaxPos && c->td_steps <= kRowsMaxTd && g_targets_kernel != 1 && g_targets_kernel != 3) {
    const int rows_per_cta = (kTgtWarps / 2) * (32 / K1);
    build_targets_rows_kernel<<<(c->batch + rows_per_cta - 1) / rows_per_cta, kTgtThreads, 0, (cudaStream_t)stream>>>(
        *w, *c, pos, chunk_start, chunk_len, pad_actions, obs_out, actions_out, t_rewards, t_values, t_policies,
        value_support, reward_support);
    MZ_LAUNCH_CHECK();
    return MZ_OK;
  }
  // per warp: rewards [KT4] f32, to_play [KT16], values / rewards [K+1] f32 each, mbarrier
  const int KT4 = (KT + 4 + 3) & ~3, KT16 = (KT + 16 + 15) & ~15;
  const int warp_smem = (int)((sizeof(float) * (KT4 + 2 * K1) + KT16 + 8 + 15) / 16 * 16);
  const size_t smem = (size_t)warp_smem * kTgtWarps + (sizeof(double) * (KT + kDiscPad) + 15) / 16 * 16;  // + discount table
  const int bulk_ok = ((reinterpret_cast<uintptr_t>(w->rewards) | reinterpret_cast<uintptr_t>(w->to_play)) & 15) == 0;
  if (smem > 200 * 1024) return MZ_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(build_targets_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  build_targets_kernel<<<(c->batch + kTgtWarps - 1) / kTgtWarps, kTgtThreads, smem, (cudaStream_t)stream>>>(
      *w, *c, pos, chunk_start, chunk_len, pad_actions, obs_out, actions_out, t_rewards, t_values,
      t_policies, value_support, reward_support, warp_smem, bulk_ok);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_replay_sample_targets(const double* tree, int64_t max_capacity, const double* u01, int32_t u01_is_mt_words,
                             const int64_t* slot_pos, const int64_t* slot_start, const int32_t* slot_len,
                             int64_t num_memories, double beta, int64_t* tree_idx, double* priority, int64_t* pos,
                             int64_t* chunk_start, int32_t* chunk_len, double* is_weights, const mz_window* w,
                             const mz_target_cfg* c, int32_t* pad_actions, uint64_t pad_seed, float* obs_out,
                             int32_t* actions_out, float* t_rewards, float* t_values, float* t_policies,
                             float* value_support, float* reward_support, void* stream) {
  if (!c || !w || !slot_pos) return MZ_ERR_BAD_ARG;
  const int rc = mz_sumtree_sample_mt(tree, max_capacity, c->batch, u01, u01_is_mt_words, slot_pos, slot_start, slot_len,
                                      num_memories, beta, tree_idx, priority, pos, chunk_start, chunk_len, is_weights,
                                      stream);
  if (rc != MZ_OK) return rc;
  if (pad_seed && c->num_unroll_steps > 0) {
    if (!pad_actions) return MZ_ERR_BAD_ARG;
    const int n = c->batch * c->num_unroll_steps;
    pad_actions_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, w->num_actions, pad_seed, pad_actions);
    MZ_LAUNCH_CHECK();
  }
  return mz_build_targets(w, c, pos, chunk_start, chunk_len, pad_actions, obs_out, actions_out, t_rewards, t_values,
                          t_policies, value_support, reward_support, stream);
}

}  // extern "C"
