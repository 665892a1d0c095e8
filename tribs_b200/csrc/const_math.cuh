// const_math.cuh — exp / log for the water-balance sweeps with every coefficient in the constant
// bank (TRIBS_CONST_MATH).
//
// Why: 69 % of the instructions the pipelined sweep executes sit in its out-of-line exp / log / pow
// routines, and a third of those are UMOV / IMAD.MOV pairs that materialise the 64-bit polynomial
// coefficients of the CUDA library routines as immediates (SASS census in
// profiles/r01_const_math_sass.md). The sweep is bound by instruction delivery, so the cheapest
// instruction is the one that is not there: a DFMA can take its coefficient straight from c[3][..].
//
// The algorithms are the textbook ones (argument reduction by ln 2 in two pieces + Taylor
// polynomial for exp; the fdlibm decomposition log x = k ln 2 + log(1+f), f/(2+f) series with the
// seven Remez coefficients Lg1..Lg7, for log), written with explicit fma() so that nvcc -fmad=false
// and g++ -ffp-contract=off produce the same IEEE operations: the functions are bit-identical on the
// host and on the device, and tests/test_const_math.py bounds their error against glibc (<= 1 ulp
// for exp on |x| <= 700, <= 1 ulp for log on normal arguments). Anything outside the fast domain
// (NaN, Inf, zero, negative, subnormal, overflow) goes to the library routine.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace tribs {
namespace cm {

// slot indices into the coefficient table
enum : int {
  kLog2e = 0, kLn2Hi, kLn2Lo, kShift,                    // exp: reduction
  kE2, kE3, kE4, kE5, kE6, kE7, kE8, kE9, kE10, kE11, kE12, kE13,   // exp: 1/n!
  kLg1, kLg2, kLg3, kLg4, kLg5, kLg6, kLg7, kLLn2Hi, kLLn2Lo,       // log
  kNumConst
};

#define TRIBS_CM_TABLE                                                                                          \
  {1.4426950408889634074, 6.93147180369123816490e-01, 1.90821492927058770002e-10, 6755399441055744.0,           \
   1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800,      \
   1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,                                                         \
   6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,      \
   1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,                                \
   6.93147180369123816490e-01, 1.90821492927058770002e-10}

#if defined(__CUDACC__)
// not const-qualified on purpose: the compiler must read the bank, it cannot fold the values back
// into immediates
__constant__ double d_table[kNumConst] = TRIBS_CM_TABLE;
#endif
static const double h_table[kNumConst] = TRIBS_CM_TABLE;

#if defined(__CUDA_ARCH__)
#define TRIBS_CM_C(i) (::tribs::cm::d_table[i])
#define TRIBS_CM_FN __device__ __forceinline__
TRIBS_CM_FN int32_t hi32(double x) { return __double2hiint(x); }
TRIBS_CM_FN int32_t lo32(double x) { return __double2loint(x); }
TRIBS_CM_FN double from_hilo(int32_t hi, int32_t lo) { return __hiloint2double(hi, lo); }
#else
#define TRIBS_CM_C(i) (::tribs::cm::h_table[i])
#define TRIBS_CM_FN inline
TRIBS_CM_FN int32_t hi32(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
TRIBS_CM_FN int32_t lo32(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u & 0xffffffffu); }
TRIBS_CM_FN double from_hilo(int32_t hi, int32_t lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x; memcpy(&x, &u, 8); return x;
}
#endif

// the library routines for arguments outside the fast domains: out of line, so the fast paths stay small
#if defined(__CUDA_ARCH__)
__device__ __noinline__ double exp_cold(double x) { return exp(x); }
__device__ __noinline__ double log_cold(double x) { return log(x); }
__device__ __noinline__ double pow_cold(double x, double y) { return pow(x, y); }
#else
inline double exp_cold(double x) { return exp(x); }
inline double log_cold(double x) { return log(x); }
inline double pow_cold(double x, double y) { return pow(x, y); }
#endif

// e^x. Fast domain |x| <= 700 (result normal, 2^k applied by an exponent-field add).
TRIBS_CM_FN double exp_c(double x) {
  if (!(fabs(x) <= 700.0)) return exp_cold(x);
  const double t = fma(x, TRIBS_CM_C(kLog2e), TRIBS_CM_C(kShift));   // round(x / ln 2) in the low word
  const int32_t k = lo32(t);
  const double kf = t - TRIBS_CM_C(kShift);
  double r = fma(kf, -TRIBS_CM_C(kLn2Hi), x);
  r = fma(kf, -TRIBS_CM_C(kLn2Lo), r);                               // |r| <= 0.3466
  double p = TRIBS_CM_C(kE13);
  p = fma(p, r, TRIBS_CM_C(kE12)); p = fma(p, r, TRIBS_CM_C(kE11)); p = fma(p, r, TRIBS_CM_C(kE10));
  p = fma(p, r, TRIBS_CM_C(kE9)); p = fma(p, r, TRIBS_CM_C(kE8)); p = fma(p, r, TRIBS_CM_C(kE7));
  p = fma(p, r, TRIBS_CM_C(kE6)); p = fma(p, r, TRIBS_CM_C(kE5)); p = fma(p, r, TRIBS_CM_C(kE4));
  p = fma(p, r, TRIBS_CM_C(kE3)); p = fma(p, r, TRIBS_CM_C(kE2));
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);                                                // in [0.70, 1.42]
  return from_hilo(hi32(p) + (k << 20), lo32(p));
}

// ln x. Fast domain: positive normal finite x.
TRIBS_CM_FN double log_c(double x) {
  if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return log_cold(x);
  int32_t hx = hi32(x);
  int32_t k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int32_t up = (hx + 0x95f64) & 0x100000;                      // mantissa above sqrt(2): halve it
  k += up >> 20;
  const double m = from_hilo(hx | (up ^ 0x3ff00000), lo32(x));       // in [sqrt(1/2), sqrt(2))
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double dk = (double)k;
  const double z = s * s, w = z * z;
  const double t1 = w * fma(w, fma(w, TRIBS_CM_C(kLg6), TRIBS_CM_C(kLg4)), TRIBS_CM_C(kLg2));
  const double t2 = z * fma(w, fma(w, fma(w, TRIBS_CM_C(kLg7), TRIBS_CM_C(kLg5)), TRIBS_CM_C(kLg3)), TRIBS_CM_C(kLg1));
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  return dk * TRIBS_CM_C(kLLn2Hi) - ((hfsq - fma(s, hfsq + R, dk * TRIBS_CM_C(kLLn2Lo))) - f);
}

// x^y as exp(y ln x) for positive bases (relative error ~ |y ln x| * 1e-16); other bases keep the
// library's semantics (0^y, NaN for negative bases with fractional exponents)
TRIBS_CM_FN double pow_c(double x, double y) { return (x > 0.0) ? exp_c(y * log_c(x)) : pow_cold(x, y); }

}  // namespace cm
}  // namespace tribs
