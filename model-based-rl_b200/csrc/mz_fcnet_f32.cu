// FCNetwork inference, float32 on CUDA cores: the reference-precision path used for parity tests
// and as the numerical baseline of the bf16 tensor-core kernel (mz_fcnet_tc.cu).
//
// One CTA evaluates the whole network for kRows rows, activations never leave shared memory:
//   recurrent: [h | onehot(a)] -> reward MLP, transition MLP -> LayerNorm+ReLU -> value MLP,
//              policy MLP -> softmax-expectation + h^-1 on value and reward.
// Reference: networks.py:26-34, 55-174; config.py:27-33 (JimOhman/model-based-rl).
#include <math.h>

#include "mz_common.cuh"
#include "mz_transforms.cuh"

namespace {

constexpr int kRows = 8;
constexpr int kThreadsFc = 256;
constexpr int kWarps = kThreadsFc / 32;
constexpr int H = MZ_FC_HIDDEN;
constexpr int W = MZ_FC_WIDTH;

// out[r][j] = relu(b1[j] + sum_k x[r][k] * w1t[k][j] (+ w1t[H + act[r]][j])), j < 512
MZ_DEV void dense_first(const float* __restrict__ w1t, const float* __restrict__ b1,
                        const float* x, int ldx, int K, const int* act, float* out, int ldo) {
  for (int j = threadIdx.x; j < W; j += kThreadsFc) {
    float acc[kRows];
    const float b = b1[j];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(&w1t[(size_t)k * W + j]);
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = fmaf(w, x[r * ldx + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float v = acc[r];
      if (act) v += __ldg(&w1t[(size_t)(K + act[r]) * W + j]);  // one-hot column (attach_action)
      out[r * ldo + j] = fmaxf(v + b, 0.0f);
    }
  }
}

// out[r][o] = b2[o] + sum_k in[r][k] * w2[o][k], o < N2 ; one warp per output unit
MZ_DEV void dense_second(const float* __restrict__ w2, const float* __restrict__ b2, const float* in,
                         int ldi, int N2, float* out, int ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < N2; o += kWarps) {
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
    for (int k = lane; k < W; k += 32) {
      const float w = __ldg(&w2[(size_t)o * W + k]);
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = fmaf(w, in[r * ldi + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float v = acc[r];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(MZ_FULL, v, m);
      if (lane == 0) out[r * ldo + o] = v + b2[o];
    }
  }
}

// relu(LayerNorm(x)) over H features, one warp per row (networks.py:149, 164)
MZ_DEV void layernorm_relu(const float* __restrict__ ln_w, const float* __restrict__ ln_b, float* x,
                           int ldx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kWarps) {
    float* row = x + r * ldx;
    float s = 0.0f;
    for (int i = lane; i < H; i += 32) s += row[i];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(MZ_FULL, s, m);
    const float mean = s / (float)H;
    float q = 0.0f;
    for (int i = lane; i < H; i += 32) {
      const float d = row[i] - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) q += __shfl_xor_sync(MZ_FULL, q, m);
    const float rstd = 1.0f / sqrtf(q / (float)H + 1e-5f);
    for (int i = lane; i < H; i += 32)
      row[i] = fmaxf((row[i] - mean) * rstd * ln_w[i] + ln_b[i], 0.0f);
  }
}

struct FcSmem {
  float* x;    // [kRows][ldx] inputs
  float* a;    // [kRows][2*W]  first-layer activations of two heads
  float* h;    // [kRows][64]   hidden state
  float* o;    // [kRows][128]  second-layer outputs (support logits / policy)
  int* act;    // [kRows]
};

MZ_DEV FcSmem carve(float* base, int ldx) {
  FcSmem s;
  s.x = base;
  s.a = s.x + kRows * ldx;
  s.h = s.a + kRows * 2 * W;
  s.o = s.h + kRows * 64;
  s.act = reinterpret_cast<int*>(s.o + kRows * 128);
  return s;
}

// prediction (networks.py:151-157) from s.h; writes value [B], logits [B][A]
MZ_DEV void prediction(const mz_fc_weights& w, const FcSmem& s, int row0, int batch, float* value,
                       float* logits) {
  dense_first(w.val_w1, w.val_b1, s.h, 64, H, nullptr, s.a, 2 * W);
  dense_first(w.pol_w1, w.pol_b1, s.h, 64, H, nullptr, s.a + W, 2 * W);
  __syncthreads();
  dense_second(w.val_w2, w.val_b2, s.a, 2 * W, w.value_bins, s.o, 128);
  dense_second(w.pol_w2, w.pol_b2, s.a + W, 2 * W, w.num_actions, s.o + 64, 128);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kWarps) {
    const int b = row0 + r;
    if (b >= batch) continue;
    const float v = w.no_support ? s.o[r * 128]  // networks.py:153: the raw output of a one-unit head
                                 : mz_support_to_scalar_warp(s.o + r * 128, w.value_bins, w.value_min,
                                                             w.no_target_transform, lane);
    if (lane == 0) value[b] = v;
    for (int a = lane; a < w.num_actions; a += 32)
      logits[(size_t)b * w.num_actions + a] = s.o[r * 128 + 64 + a];
  }
}

__global__ void __launch_bounds__(kThreadsFc)
fc_recurrent_f32_kernel(mz_fc_weights w, int batch, const float* __restrict__ hidden_in,
                        long long in_row_stride, const int32_t* __restrict__ in_index,
                        const int32_t* __restrict__ actions, float* __restrict__ hidden_out,
                        long long out_row_stride, long long out_offset, float* __restrict__ value,
                        float* __restrict__ reward, float* __restrict__ logits) {
  extern __shared__ float smem[];
  const FcSmem s = carve(smem, 64);
  const int row0 = blockIdx.x * kRows;
  for (int i = threadIdx.x; i < kRows * H; i += kThreadsFc) {
    const int r = i / H, k = i % H, b = min(row0 + r, batch - 1);
    const float* src = hidden_in + (size_t)b * in_row_stride + (in_index ? (size_t)in_index[b] * H : 0);
    s.x[r * 64 + k] = src[k];
  }
  if (threadIdx.x < kRows) s.act[threadIdx.x] = actions[min(row0 + (int)threadIdx.x, batch - 1)];
  __syncthreads();
  // dynamics (networks.py:159-165)
  dense_first(w.rew_w1, w.rew_b1, s.x, 64, H, s.act, s.a, 2 * W);
  dense_first(w.dyn_w1, w.dyn_b1, s.x, 64, H, s.act, s.a + W, 2 * W);
  __syncthreads();
  dense_second(w.rew_w2, w.rew_b2, s.a, 2 * W, w.reward_bins, s.o, 128);
  dense_second(w.dyn_w2, w.dyn_b2, s.a + W, 2 * W, H, s.h, 64);
  __syncthreads();
  layernorm_relu(w.ln_w, w.ln_b, s.h, 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kWarps) {
    const int b = row0 + r;
    if (b >= batch) continue;
    const float rv = w.no_support ? s.o[r * 128]  // networks.py:161
                                  : mz_support_to_scalar_warp(s.o + r * 128, w.reward_bins, w.reward_min,
                                                              w.no_target_transform, lane);
    if (lane == 0) reward[b] = rv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRows * H; i += kThreadsFc) {
    const int r = i / H, k = i % H, b = row0 + r;
    if (b < batch) hidden_out[(size_t)b * out_row_stride + out_offset + k] = s.h[r * 64 + k];
  }
  prediction(w, s, row0, batch, value, logits);
}

__global__ void __launch_bounds__(kThreadsFc)
fc_initial_f32_kernel(mz_fc_weights w, int batch, const float* __restrict__ obs,
                      float* __restrict__ hidden, long long hidden_stride, float* __restrict__ value,
                      float* __restrict__ logits, int ldx) {
  extern __shared__ float smem[];
  const FcSmem s = carve(smem, ldx);
  const int row0 = blockIdx.x * kRows;
  const int K = w.obs_dim;
  for (int i = threadIdx.x; i < kRows * K; i += kThreadsFc) {
    const int r = i / K, k = i % K, b = min(row0 + r, batch - 1);
    s.x[r * ldx + k] = obs[(size_t)b * K + k];
  }
  __syncthreads();
  // representation (networks.py:146-149)
  dense_first(w.rep_w1, w.rep_b1, s.x, ldx, K, nullptr, s.a, 2 * W);
  __syncthreads();
  dense_second(w.rep_w2, w.rep_b2, s.a, 2 * W, H, s.h, 64);
  __syncthreads();
  layernorm_relu(w.ln_w, w.ln_b, s.h, 64);
  __syncthreads();
  for (int i = threadIdx.x; i < kRows * H; i += kThreadsFc) {
    const int r = i / H, k = i % H, b = row0 + r;
    if (b < batch) hidden[(size_t)b * hidden_stride + k] = s.h[r * 64 + k];
  }
  prediction(w, s, row0, batch, value, logits);
}

size_t fc_smem_bytes(int ldx) {
  return sizeof(float) * (size_t)(kRows * ldx + kRows * 2 * W + kRows * 64 + kRows * 128) +
         sizeof(int) * kRows;
}

int check_weights(const mz_fc_weights* w, bool need_rep) {
  if (!w) return MZ_ERR_BAD_ARG;
  if (w->num_actions < 1 || w->num_actions > 64 || w->value_bins < 1 || w->value_bins > 64 ||
      w->reward_bins < 1 || w->reward_bins > 64)
    return MZ_ERR_UNSUPPORTED;
  if (w->no_support && (w->value_bins != 1 || w->reward_bins != 1)) return MZ_ERR_BAD_ARG;
  if (!w->dyn_w1 || !w->rew_w1 || !w->val_w1 || !w->pol_w1 || !w->ln_w || !w->ln_b)
    return MZ_ERR_BAD_ARG;
  if (need_rep && (!w->rep_w1 || !w->rep_w2 || w->obs_dim < 1)) return MZ_ERR_BAD_ARG;
  return MZ_OK;
}

}  // namespace

// Observation ingest: the learner / actor normalisation `(obs - obs_min) / obs_range` (actors.py:127-129,
// learners.py:170-171; float32 numpy arithmetic) for byte observations, on the device, so that only the
// bytes cross PCIe.  One rounding per operation like numpy.
__global__ void obs_normalize_u8_kernel(long long n, int obs_dim, const uint8_t* __restrict__ in,
                                        const float* __restrict__ obs_min, const float* __restrict__ obs_range,
                                        float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % obs_dim);
    const float x = (float)in[i];
    out[i] = __fdiv_rn(__fsub_rn(x, obs_min ? obs_min[k] : 0.0f), obs_range ? obs_range[k] : 255.0f);
  }
}

extern "C" {

int mz_obs_normalize_u8(int64_t rows, int32_t obs_dim, const uint8_t* obs, const float* obs_min,
                        const float* obs_range, float* out, void* stream) {
  if (rows < 1 || obs_dim < 1 || !obs || !out) return MZ_ERR_BAD_ARG;
  const long long n = (long long)rows * obs_dim;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  obs_normalize_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, obs_dim, obs, obs_min, obs_range, out);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_fc_initial_f32(const mz_fc_weights* w, int32_t batch, const float* obs, float* hidden,
                      int64_t hidden_stride, float* value, float* logits, void* stream) {
  int rc = check_weights(w, true);
  if (rc) return rc;
  if (batch < 1 || !obs || !hidden || !value || !logits) return MZ_ERR_BAD_ARG;
  const int ldx = (w->obs_dim + 3) / 4 * 4 + 1;
  const size_t smem = fc_smem_bytes(ldx);
  if (smem > 200 * 1024) return MZ_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(fc_initial_f32_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = (batch + kRows - 1) / kRows;
  fc_initial_f32_kernel<<<grid, kThreadsFc, smem, (cudaStream_t)stream>>>(
      *w, batch, obs, hidden, hidden_stride, value, logits, ldx);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

int mz_fc_recurrent_f32(const mz_fc_weights* w, int32_t batch, const float* hidden_in,
                        int64_t in_row_stride, const int32_t* in_index, const int32_t* actions,
                        float* hidden_out, int64_t out_row_stride, int64_t out_offset, float* value,
                        float* reward, float* logits, void* stream) {
  int rc = check_weights(w, false);
  if (rc) return rc;
  if (batch < 1 || !hidden_in || !actions || !hidden_out || !value || !reward || !logits)
    return MZ_ERR_BAD_ARG;
  const size_t smem = fc_smem_bytes(64);
  cudaError_t e = cudaFuncSetAttribute(fc_recurrent_f32_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = (batch + kRows - 1) / kRows;
  fc_recurrent_f32_kernel<<<grid, kThreadsFc, smem, (cudaStream_t)stream>>>(
      *w, batch, hidden_in, in_row_stride, in_index, actions, hidden_out, out_row_stride, out_offset,
      value, reward, logits);
  MZ_LAUNCH_CHECK();
  return MZ_OK;
}

}  // extern "C"
