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

constexpr int kTgtThreads = 128;
constexpr int kTgtWarps = kTgtThreads / 32;

__global__ void __launch_bounds__(kTgtThreads)
build_targets_kernel(mz_window w, mz_target_cfg c, const int64_t* __restrict__ pos_arr,
                     const int64_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_len,
                     const int32_t* __restrict__ pad_actions, float* __restrict__ obs_out,
                     int32_t* __restrict__ actions_out, float* __restrict__ t_rewards,
                     float* __restrict__ t_values, float* __restrict__ t_policies,
                     float* __restrict__ value_support, float* __restrict__ reward_support) {
  extern __shared__ unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int K = c.num_unroll_steps, T = c.td_steps, A = w.num_actions;
  const int64_t pos = pos_arr[b];
  const int step = (int)(pos - chunk_start[b]);
  const int len = chunk_len[b];  // len(root_values) == len(rewards) == len(to_play)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // stage the reward / to_play window [step, min(step + K + T, len)) in shared memory
  const int win = max(0, min(K + T, len - step));
  float* s_rew = reinterpret_cast<float*>(smem_raw);
  int8_t* s_tp = reinterpret_cast<int8_t*>(s_rew + (K + T));
  float* s_val = reinterpret_cast<float*>(s_tp + ((K + T + 3) / 4) * 4);  // [K+1] values
  float* s_lastr = s_val + (K + 1);                                       // [K+1] rewards
  for (int j = threadIdx.x; j < win; j += kTgtThreads) {
    s_rew[j] = w.rewards[pos + j];
    s_tp[j] = w.to_play[pos + j];
  }
  __syncthreads();

  // observation: np.float32(history.observations[step])  (replay_buffer.py:147)
  {
    const int E = w.obs_elems;
    float* dst = obs_out + (size_t)b * E;
    if (w.obs_is_u8) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(w.obs) + (size_t)pos * E;
      for (int e = threadIdx.x; e < E; e += kTgtThreads) {
        float v = (float)src[e];
        if (c.normalize_obs) v = __fdiv_rn(__fsub_rn(v, c.obs_min[e]), c.obs_range[e]);
        dst[e] = v;
      }
    } else {
      const float* src = reinterpret_cast<const float*>(w.obs) + (size_t)pos * E;
      for (int e = threadIdx.x; e < E; e += kTgtThreads) {
        float v = src[e];
        if (c.normalize_obs) v = __fdiv_rn(__fsub_rn(v, c.obs_min[e]), c.obs_range[e]);
        dst[e] = v;
      }
    }
  }
  // actions: history.actions[step:step+K], padded with random actions (replay_buffer.py:149-152)
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    const int n_real = max(0, min(K, len - step));
    actions_out[(size_t)b * K + k] =
        k < n_real ? w.actions[pos + k] : pad_actions[(size_t)b * K + (k - n_real)];
  }

  // insert_target (replay_buffer.py:165-198): one warp per unroll position
  for (int i = warp; i <= K; i += kTgtWarps) {
    const int ci = step + i;
    float last_reward = 0.0f;
    if (ci > 0 && ci <= len) last_reward = (i > 0) ? s_rew[i - 1] : w.rewards[pos - 1];
    float value = 0.0f;
    float* pol = t_policies + ((size_t)b * (K + 1) + i) * A;
    if (ci < len) {
      const int tp = s_tp[i];
      const int n = min(T, len - ci);
      double acc = 0.0;  // exact products of float32 pairs, accumulated in binary64
      for (int j = lane; j < n; j += 32) {
        float r = s_rew[i + j];
        if (s_tp[i + j] != tp) r = -r;
        acc += (double)r * (double)c.discounts[j];
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) acc += shfl_xor_f64<32>(acc, m);
      const double boot = (ci + T < len) ? __dmul_rn(w.root_values[pos + i + T], c.disc_pow_td) : 0.0;
      value = __fadd_rn((float)boot, (float)acc);  // numpy 2: python float + np.float32 -> float32
      for (int a = lane; a < A; a += 32) pol[a] = w.child_visits[(size_t)(pos + i) * A + a];
    } else {
      for (int a = lane; a < A; a += 32) pol[a] = 0.0f;  // absorbing policy (replay_buffer.py:90)
    }
    if (lane == 0) {
      t_rewards[(size_t)b * (K + 1) + i] = last_reward;
      t_values[(size_t)b * (K + 1) + i] = value;
      s_val[i] = value;
      s_lastr[i] = last_reward;
    }
  }
  if (!c.fuse_supports) return;
  __syncthreads();
  // learners.py:186-192: h(x) then two-hot projection of values and rewards
  const int vb = c.value_max - c.value_min + 1, rb = c.reward_max - c.reward_min + 1;
  for (int idx = threadIdx.x; idx < (K + 1) * vb; idx += kTgtThreads) {
    const int i = idx / vb, j = idx % vb;
    float x = s_val[i];
    if (!c.no_target_transform) x = mz_scalar_transform_f(x);
    const MzTwoHot th = mz_two_hot(x, c.value_min, c.value_max);
    value_support[((size_t)b * (K + 1) + i) * vb + j] = j == th.lo ? th.p_lo : (j == th.hi ? th.p_hi : 0.0f);
  }
  for (int idx = threadIdx.x; idx < (K + 1) * rb; idx += kTgtThreads) {
    const int i = idx / rb, j = idx % rb;
    float x = s_lastr[i];
    if (!c.no_target_transform) x = mz_scalar_transform_f(x);
    const MzTwoHot th = mz_two_hot(x, c.reward_min, c.reward_max);
    reward_support[((size_t)b * (K + 1) + i) * rb + j] = j == th.lo ? th.p_lo : (j == th.hi ? th.p_hi : 0.0f);
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

int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

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
  const int KT = c->num_unroll_steps + c->td_steps;
  const size_t smem = sizeof(float) * KT + ((KT + 3) / 4) * 4 + sizeof(float) * 2 * (c->num_unroll_steps + 1);
  if (smem > 200 * 1024) return MZ_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(build_targets_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  build_targets_kernel<<<c->batch, kTgtThreads, smem, (cudaStream_t)stream>>>(
      *w, *c, pos, chunk_start, chunk_len, pad_actions, obs_out, actions_out, t_rewards, t_values,
      t_policies, value_support, reward_support);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

}  // extern "C"
