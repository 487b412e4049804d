// Straight-line fp64 exp / log for the pair kernels (sm_100a fp64 pipe: 64 lanes/clk/SM, so every DFMA counts).
//
// The reference's serial code calls libm pow/exp per pair (reaxc_nonbonded_sw64.c:137-170) and its slave-core
// kernels replace them by ~1e-6 polynomial versions (p_expd/p_powd).  Here:
//   exp_b(x) : x = (16 k + j) ln2/16 + r, |r| <= ln2/32;  2^(j/16) from a 16-entry table, expm1(r) by a degree-7
//              Taylor polynomial in Estrin form, 2^k by an exponent-field add.            11 DP ops, <= 2 ulp
//   log_b(x) : x = 2^e m, m in [1,2);  the top 5 mantissa bits pick rc ~ 1/m (20 significant bits, so
//              t = m rc - 1 is EXACT in one fma) and -log(rc);  log1p(t), |t| < 2^-6, degree 9.   11 DP ops,
//              absolute error <= 1 ulp of the result + 2e-17 (NOT relative near x = 1: use it only under exp/pow)
//   The sub-tables are 128 / 256 bytes, so the 32 lookups of a warp hit one or two shared-memory rows whatever the
//   indices are (a 64/128-entry version with shorter polynomials ran at the same speed; this one cannot bank-conflict).
//   rcbrt_b(x): fp32 seed + one fourth-order correction step (below).
// No branches, no special cases: arguments must be finite, exp_b needs |x| <= 700, log_b needs a normal x > 0.
// The table (640 bytes) is passed in by the caller (shared memory on the device).  The same source compiles for the
// host (tests/test_fast_math.py builds it with g++ and checks it against libm), and because every operation is an
// explicit fma/add/mul the host and device results are bit-identical (the cube root's fp32 seed aside, whose error the
// correction step removes: the test perturbs the seed).
#pragma once
#include <cmath>
#include <cstring>

#include "rxb_math_tables.h"

#if defined(__CUDACC__)
#define RXB_HD __host__ __device__ __forceinline__
#else
#define RXB_HD inline
#endif

namespace rxb {
namespace fm {

constexpr int kExpTabN = 16, kLogTabN = 32;
constexpr int kTabDoubles = kExpTabN + 2 * kLogTabN;   // 2^(j/16), then rc_i, then -log rc_i

RXB_HD int hi32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
RXB_HD int lo32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL);
#endif
}
RXB_HD double mk(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  long long b = ((long long)hi << 32) | (unsigned int)lo; double x; std::memcpy(&x, &b, 8); return x;
#endif
}

// Polynomial and reduction constants.  On the device they live in __constant__ memory so that a DFMA takes them as a
// c[bank][offset] operand; as literals most of them do not fit the 32-bit-high immediate form and cost two extra move
// instructions per use (ncu: 26 % of the nonbonded kernel's issued instructions were such moves).
enum Coef { C_16_LN2, C_NLN2_16_HI, C_NLN2_16_LO, C_1_6, C_1_120, C_1_24, C_1_5040, C_1_720, C_I2D, C_LN2, C_1_3, C_1_5, C_1_7,
            C_N1_6, C_1_9, C_14_81, C_2_9, C_NUM };
#define RXB_FM_COEF_INIT                                                                                              \
  { k16oLn2, -kLn2o16Hi, -kLn2o16Lo, 1.0 / 6.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 5040.0, 1.0 / 720.0, 4503601774854144.0, kLn2, \
    1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 9.0, 14.0 / 81.0, 2.0 / 9.0 }
#if defined(__CUDACC__)
static __constant__ double c_fm_coef[C_NUM] = RXB_FM_COEF_INIT;
#endif
static const double h_fm_coef[C_NUM] = RXB_FM_COEF_INIT;
#if defined(__CUDA_ARCH__)
#define RXB_FMC(i) c_fm_coef[i]
#else
#define RXB_FMC(i) h_fm_coef[i]
#endif

RXB_HD double exp_b(double x, const double* __restrict__ tab) {
  constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52: the low word of x*16/ln2 + magic is round(x*16/ln2)
  const double kf = fma(x, RXB_FMC(C_16_LN2), kMagic);
  const int k = lo32(kf);
  const double kd = kf - kMagic;
  double r = fma(kd, RXB_FMC(C_NLN2_16_HI), x);
  r = fma(kd, RXB_FMC(C_NLN2_16_LO), r);
  const double r2 = r * r;
  const double a = fma(r, RXB_FMC(C_1_6), 0.5);
  const double b = fma(r, RXB_FMC(C_1_120), RXB_FMC(C_1_24));
  const double c = fma(r, RXB_FMC(C_1_5040), RXB_FMC(C_1_720));
  const double q = fma(r2, fma(r2, c, b), a);
  const double p = fma(r2, q, r);            // expm1(r), |r| <= 0.0217: next term r^8/40320 < 2e-18
  const double T = tab[k & (kExpTabN - 1)];
  const double y = fma(T, p, T);             // in (0.97, 2.05): a normal number, so 2^(k>>4) is an exponent-field add
  return mk(hi32(y) + ((k >> 4) << 20), lo32(y));
}

RXB_HD double log_b(double x, const double* __restrict__ tab) {
  const int hi = hi32(x);
  const int idx = (hi >> 15) & (kLogTabN - 1);
  const double m = mk((hi & 0x000fffff) | 0x3ff00000, lo32(x));
  // e as a double without I2F: 2^52 + 2^31 + e, minus 2^52 + 2^31
  const double ed = mk(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000) - RXB_FMC(C_I2D);
  const double rc = tab[kExpTabN + idx], lc = tab[kExpTabN + kLogTabN + idx];
  const double t = fma(m, rc, -1.0);         // exact; |t| <= 0.0157
  const double t2 = t * t;
  const double a = fma(t, RXB_FMC(C_1_3), -0.5);
  const double b = fma(t, RXB_FMC(C_1_5), -0.25);
  const double c = fma(t, RXB_FMC(C_1_7), RXB_FMC(C_N1_6));
  const double d = fma(t, RXB_FMC(C_1_9), -0.125);
  const double p = fma(t2, fma(t2, fma(t2, d, c), b), a);   // next term t^10/10 < 1e-19
  return fma(ed, RXB_FMC(C_LN2), lc) + fma(t2, p, t);
}

// x^(-1/3) for x in the fp32 normal range (here r^3 + shielding, 0.1 .. 1e4): fp32 seed (2 MUFU ops on the device,
// relative error ~1e-6), then ONE fourth-order step y(1 + e/3 + 2e^2/9 + 14e^3/81), e = 1 - x y^3 (error ~e^4/8).
// 7 DP ops instead of the ~28 instructions of rcbrt(); <= 2e-16 relative for any seed within 1e-4.
RXB_HD double rcbrt_seeded(double x, double y) {
  const double y3 = (y * y) * y;
  const double e = fma(-x, y3, 1.0);
  const double p = fma(e, fma(e, RXB_FMC(C_14_81), RXB_FMC(C_2_9)), RXB_FMC(C_1_3));
  return fma(y * e, p, y);
}
RXB_HD double rcbrt_b(double x) {
#if defined(__CUDA_ARCH__)
  const double y = (double)__powf(__double2float_rn(x), -0.33333334f);
#else
  const double y = (double)powf((float)x, -0.33333334f);
#endif
  return rcbrt_seeded(x, y);
}

// host copy of the table (tests)
inline const double* host_tables() {
  static const double t[kTabDoubles] = RXB_FM_TAB_INIT;
  return t;
}

#if defined(__CUDACC__)
// device copy (one per translation unit, 640 bytes); kernels stage it in shared memory, 128-byte aligned
__device__ const double d_fm_tab[kTabDoubles] = RXB_FM_TAB_INIT;
__device__ __forceinline__ double exp_c(double x, const double* tab) { return exp_b(fmin(fmax(x, -700.0), 700.0), tab); }
#endif

}  // namespace fm
}  // namespace rxb
