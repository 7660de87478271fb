// exp() used for Node.expand priors (mcts.py:52: math.exp(logit) -> libm exp on the host).
// Bit-compatible with glibc's exp on every float32-valued argument in the useful range, so that
// priors -- and everything downstream -- match the reference run on the host exactly.
#pragma once
#include "mz_common.cuh"

__device__ const unsigned long long mz_exp_tab[256] = {
#include "mz_exp_table.inc"
};

#define MZ_EXP_FN MZ_DEV
#define MZ_FMA(a, b, c) __fma_rn((a), (b), (c))
#define MZ_MUL(a, b) __dmul_rn((a), (b))
#define MZ_ADD(a, b) __dadd_rn((a), (b))
#define MZ_SUB(a, b) __dsub_rn((a), (b))
#define MZ_EXP_TAB mz_exp_tab
#define MZ_ASU(x) ((unsigned long long)__double_as_longlong(x))
#define MZ_ASD(u) __longlong_as_double((long long)(u))
#include "mz_exp_algo.h"

MZ_DEV double mz_exp(double x) {
  const double ax = fabs(x);
  if (ax >= 0x1p-54 && ax < 512.0) return mz_exp_core(x);
  if (ax < 0x1p-54) return __dadd_rn(1.0, x);  // glibc: tiny |x| -> 1.0 + x
  return exp(x);  // |x| >= 512, inf, nan: outside any meaningful logit; CUDA's exp saturates the same way
}
