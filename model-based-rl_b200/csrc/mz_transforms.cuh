// Scalar transforms and supports (config.py:27-68) as device functions.  float32, one rounding per
// torch op in the reference's op order; explicit *_rn intrinsics so nvcc never contracts to FMA.
#pragma once
#include "mz_common.cuh"

MZ_DEV float mz_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// Config.scalar_transform config.py:51-54: sign(x) * (sqrt(|x| + 1) - 1) + 0.001 * x
MZ_DEV float mz_scalar_transform_f(float x) {
  const float c = __fsub_rn(__fsqrt_rn(__fadd_rn(fabsf(x), 1.0f)), 1.0f);
  return __fadd_rn(__fmul_rn(mz_sign(x), c), __fmul_rn(0.001f, x));
}

// h^-1, config.py:32: sign(v) * (((sqrt(1 + 4*0.001*(|v| + 1 + 0.001)) - 1) / (2*0.001))**2 - 1)
MZ_DEV float mz_inverse_scalar_transform_f(float v) {
  const float b = __fadd_rn(__fadd_rn(fabsf(v), 1.0f), 0.001f);
  const float d = __fadd_rn(1.0f, __fmul_rn(0.004f, b));
  const float g = __fdiv_rn(__fsub_rn(__fsqrt_rn(d), 1.0f), 0.002f);
  return __fmul_rn(mz_sign(v), __fsub_rn(__fmul_rn(g, g), 1.0f));
}

// Config.inverse_transform config.py:27-33 for one row held in memory; all 32 lanes of a warp call.
// softmax(logits) . support in float32 (max-subtracted like torch.softmax), then h^-1.
MZ_DEV float mz_support_to_scalar_warp(const float* logits, int bins, int support_min,
                                       int no_target_transform, int lane) {
  float m = -INFINITY;
  for (int i = lane; i < bins; i += 32) m = fmaxf(m, logits[i]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(MZ_FULL, m, s));
  float den = 0.0f;
  for (int i = lane; i < bins; i += 32) den += expf(logits[i] - m);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) den += __shfl_xor_sync(MZ_FULL, den, s);
  float num = 0.0f;
  for (int i = lane; i < bins; i += 32)
    num += (float)(support_min + i) * __fdiv_rn(expf(logits[i] - m), den);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) num += __shfl_xor_sync(MZ_FULL, num, s);
  return no_target_transform ? num : mz_inverse_scalar_transform_f(num);
}

// Config.scalar_to_support config.py:56-68 for one scalar: returns clamped x and the two bins.
struct MzTwoHot {
  float x;       // clamped
  int lo, hi;    // bin indices (floor - min, ceil - min)
  float p_lo, p_hi;
};
MZ_DEV MzTwoHot mz_two_hot(float x, int support_min, int support_max) {
  MzTwoHot t;
  x = fminf(fmaxf(x, (float)support_min), (float)support_max);
  const float fl = floorf(x), ce = ceilf(x);
  t.x = x;
  t.p_hi = __fsub_rn(x, fl);
  t.p_lo = __fsub_rn(1.0f, t.p_hi);
  t.lo = (int)fl - support_min;
  t.hi = (int)ce - support_min;
  return t;  // writer stores p_hi at hi first, then p_lo at lo (integer x ends with 1.0)
}
