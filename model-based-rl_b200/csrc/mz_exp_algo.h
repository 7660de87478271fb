// exp(x) for binary64, bit-compatible with glibc >= 2.28 (x86-64, FMA variant) on float32-valued
// arguments -- the function behind math.exp() in Node.expand (mcts.py:52).
//
// Algorithm (Szabolcs Nagy's exp from ARM optimized-routines, as shipped in glibc
// sysdeps/ieee754/dbl-64/e_exp.c; restated here from its published description):
//   x = k * ln2/128 + r,  exp(x) = 2^(k/128) * exp(r) ~= scale + scale * (tail + r + r^2 p(r))
// with every multiply-add that GCC contracts in the FMA build written as an explicit fma.  The
// contraction pattern and the constants were validated by comparing this exact code against the
// libm of the build image for ALL 1,056,964,608 float32 inputs with 2^-54 <= |x| < 512
// (0 mismatches; tests/exp_exhaustive.c, result recorded in DESIGN.md) and it is re-checked on a
// strided sample by tests/test_exp_exact.py and, on the device, by tests/test_gpu_exp.py.
//
// Shared between CUDA (mz_exp.cuh) and plain C (the CPU-side test); the includer defines
//   MZ_EXP_FN   function qualifiers
//   MZ_FMA(a,b,c), MZ_MUL(a,b), MZ_ADD(a,b), MZ_SUB(a,b)   correctly rounded, never contracted
//   MZ_EXP_TAB  the 256-entry uint64 table (mz_exp_table.inc)
//   MZ_ASU(double)->uint64, MZ_ASD(uint64)->double
#ifndef MZ_EXP_ALGO_H_
#define MZ_EXP_ALGO_H_

// valid for 2^-54 <= |x| < 512 (the caller handles the rest)
MZ_EXP_FN double mz_exp_core(double x) {
  const double InvLn2N = 0x1.71547652b82fep7;     // 128 / ln 2
  const double Shift = 0x1.8p52;
  const double NegLn2hiN = -0x1.62e42fefa0000p-8;  // -ln2/128, high part (exact product with k)
  const double NegLn2loN = -0x1.cf79abc9e3b3ap-47;
  const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3;
  const double C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
  const double z = MZ_MUL(InvLn2N, x);
  double kd = MZ_ADD(z, Shift);                    // round to nearest integer in the low bits
  const unsigned long long ki = MZ_ASU(kd);
  kd = MZ_SUB(kd, Shift);
  const double r = MZ_FMA(kd, NegLn2loN, MZ_FMA(kd, NegLn2hiN, x));
  const unsigned long long idx = 2 * (ki % 128), top = ki << 45;
  const double tail = MZ_ASD(MZ_EXP_TAB[idx]);
  const unsigned long long sbits = MZ_EXP_TAB[idx + 1] + top;
  const double r2 = MZ_MUL(r, r);
  const double a = MZ_FMA(r, C3, C2), b = MZ_FMA(r, C5, C4);
  const double tmp = MZ_FMA(MZ_MUL(r2, r2), b, MZ_FMA(r2, a, MZ_ADD(tail, r)));
  const double scale = MZ_ASD(sbits);
  return MZ_FMA(scale, tmp, scale);
}
#endif
