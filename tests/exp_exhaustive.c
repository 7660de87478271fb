/* CPU-side check of model-based-rl_b200/csrc/mz_exp_algo.h (the algorithm the device uses for
 * math.exp) against the host libm, bit for bit, over float32-valued inputs.
 *   exp_exhaustive <stride>     stride 1 = all 2^32 float bit patterns (about a minute per core)
 * Prints "<mismatches> <checked>".  Test infrastructure only. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const unsigned long long mz_exp_tab[256] = {
#include "mz_exp_table.inc"
};
static inline unsigned long long asu(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
static inline double asd(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
#define MZ_EXP_FN static inline
#define MZ_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define MZ_MUL(a, b) ((a) * (b))
#define MZ_ADD(a, b) ((a) + (b))
#define MZ_SUB(a, b) ((a) - (b))
#define MZ_EXP_TAB mz_exp_tab
#define MZ_ASU(x) asu(x)
#define MZ_ASD(u) asd(u)
#include "mz_exp_algo.h"

static double mz_exp(double x) {
  const double ax = fabs(x);
  if (ax >= 0x1p-54 && ax < 512.0) return mz_exp_core(x);
  if (ax < 0x1p-54) return 1.0 + x;
  return exp(x);
}

int main(int argc, char** argv) {
  uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
  uint64_t lo = argc > 2 ? strtoull(argv[2], 0, 10) : 0, hi = argc > 3 ? strtoull(argv[3], 0, 10) : (1ull << 32);
  long bad = 0, n = 0;
  for (uint64_t uu = lo; uu < hi; uu += stride) {
    uint32_t u = (uint32_t)uu;
    float xf;
    memcpy(&xf, &u, 4);
    if (!(fabsf(xf) < 512.0f)) continue; /* also skips NaN */
    double x = xf;
    n++;
    if (asu(mz_exp(x)) != asu(exp(x))) {
      if (bad < 5) fprintf(stderr, "x=%a mine=%a libm=%a\n", x, mz_exp(x), exp(x));
      bad++;
    }
  }
  printf("%ld %ld\n", bad, n);
  return 0;
}
