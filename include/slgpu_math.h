/*
 * slgpu_math.h -- deterministic fp64 exp / log / sincos(2*pi*u) for the step kernels.
 *
 * The reference calls libm's std::exp / std::log (src/infer/chainarray.cpp:41,64, sampler.cpp:160,
 * adaptive.cpp:100-139) and libstdc++'s normal_distribution (sampler.cpp:83-85).  CUDA's libdevice and
 * glibc differ in the last bit, and the tier regression amplifies such differences, so the engine and
 * the CPU oracle both use the functions below instead: straight-line code made only of
 * + - * / and explicit fma(), every one of which is correctly rounded on both sides (host:
 * -ffp-contract=off, device: -fmad=false).  The results are therefore bit-identical on CPU and GPU,
 * and within ~1 ulp of libm (tests/test_math.py checks both).
 *
 *   slm_exp(x)              e^x; +inf above 709.78, 0 below -708.39 (no subnormal results), NaN -> NaN
 *   slm_log(x)              natural log; x<0 -> NaN, 0 -> -inf, handles subnormals, inf, NaN
 *   slm_sincos2pi(u,&s,&c)  sin(2 pi u), cos(2 pi u) for u in [0,1]
 *
 * Plain C++ (no CUDA types); usable from host code and, under nvcc, from device code.
 */
#ifndef SLGPU_MATH_H
#define SLGPU_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __CUDACC__
#define SLM_FN __host__ __device__ __forceinline__
#else
#define SLM_FN static inline
#endif

SLM_FN uint64_t slm_bits(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, sizeof u);
  return u;
#endif
}

SLM_FN double slm_from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, sizeof x);
  return x;
#endif
}

/* e^x: x = k ln2 + r, |r| <= 0.35; Taylor to r^14 by Horner with fma; scale by 2^k */
SLM_FN double slm_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return slm_from_bits(0x7ff0000000000000ull);
  if (x < -708.3964185322641) return 0.0;
  const double t = x * 1.4426950408889634074;
  const long long k = (long long)(t + (t < 0.0 ? -0.5 : 0.5));
  const double kd = (double)k;
  /* ln2 = hi + lo, hi has 21 trailing zero bits so kd*hi is exact for |k| <= 1024 */
  double r = fma(-kd, 6.93147180369123816490e-01, x);
  r = fma(-kd, 1.90821492927058770002e-10, r);
  double p = 1.0 / 87178291200.0;           /* 1/14! */
  p = fma(p, r, 1.0 / 6227020800.0);        /* 1/13! */
  p = fma(p, r, 1.0 / 479001600.0);         /* 1/12! */
  p = fma(p, r, 1.0 / 39916800.0);          /* 1/11! */
  p = fma(p, r, 1.0 / 3628800.0);           /* 1/10! */
  p = fma(p, r, 1.0 / 362880.0);            /* 1/9!  */
  p = fma(p, r, 1.0 / 40320.0);             /* 1/8!  */
  p = fma(p, r, 1.0 / 5040.0);              /* 1/7!  */
  p = fma(p, r, 1.0 / 720.0);               /* 1/6!  */
  p = fma(p, r, 1.0 / 120.0);               /* 1/5!  */
  p = fma(p, r, 1.0 / 24.0);                /* 1/4!  */
  p = fma(p, r, 1.0 / 6.0);                 /* 1/3!  */
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  if (k > 1023) return (p * slm_from_bits((uint64_t)(1023 + 1023) << 52)) * 2.0;
  return p * slm_from_bits((uint64_t)(k + 1023) << 52);
}

/* log x: x = 2^e m, m in [sqrt(1/2), sqrt 2]; f = m-1, s = f/(2+f), log m = f - f^2/2 + s (f^2/2 + R(s^2)),
 * R(z) = sum_{k>=1} 2/(2k+1) z^k (the atanh series), result = e ln2_hi - ((hfsq - (s (hfsq+R) + e ln2_lo)) - f) */
SLM_FN double slm_log(double x) {
  uint64_t b = slm_bits(x);
  int e = 0;
  if (!(x > 0.0) || b >= 0x7ff0000000000000ull || b < 0x0010000000000000ull) {
    if (x != x) return x;
    if (x < 0.0) return slm_from_bits(0x7ff8000000000000ull);
    if (x == 0.0) return slm_from_bits(0xfff0000000000000ull);
    if (b == 0x7ff0000000000000ull) return x;
    x = x * 18014398509481984.0; /* subnormal: scale by 2^54 */
    b = slm_bits(x);
    e = -54;
  }
  e += (int)(b >> 52) - 1023;
  double m = slm_from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
  if (m > 1.4142135623730951) {
    m = m * 0.5;
    e += 1;
  }
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  double R = 2.0 / 25.0;
  R = fma(R, z, 2.0 / 23.0);
  R = fma(R, z, 2.0 / 21.0);
  R = fma(R, z, 2.0 / 19.0);
  R = fma(R, z, 2.0 / 17.0);
  R = fma(R, z, 2.0 / 15.0);
  R = fma(R, z, 2.0 / 13.0);
  R = fma(R, z, 2.0 / 11.0);
  R = fma(R, z, 2.0 / 9.0);
  R = fma(R, z, 2.0 / 7.0);
  R = fma(R, z, 2.0 / 5.0);
  R = fma(R, z, 2.0 / 3.0);
  R = R * z;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)e;
  return dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
}

/* sin and cos of 2*pi*u: nearest quarter turn q, r = u - q/4 in [-1/8, 1/8] (exact), y = 2 pi r, Taylor */
SLM_FN void slm_sincos2pi(double u, double* sn, double* cs) {
  const int q = (int)(4.0 * u + 0.5);
  const double r = u - 0.25 * (double)q;
  const double y = r * 6.283185307179586476925286766559;
  const double y2 = y * y;
  double ps = 1.0 / 355687428096000.0;           /*  1/17! */
  ps = fma(ps, y2, -1.0 / 1307674368000.0);      /* -1/15! */
  ps = fma(ps, y2, 1.0 / 6227020800.0);          /*  1/13! */
  ps = fma(ps, y2, -1.0 / 39916800.0);           /* -1/11! */
  ps = fma(ps, y2, 1.0 / 362880.0);              /*  1/9!  */
  ps = fma(ps, y2, -1.0 / 5040.0);               /* -1/7!  */
  ps = fma(ps, y2, 1.0 / 120.0);                 /*  1/5!  */
  ps = fma(ps, y2, -1.0 / 6.0);                  /* -1/3!  */
  const double s = fma(y * y2, ps, y);
  double pc = 1.0 / 6402373705728000.0;          /*  1/18! */
  pc = fma(pc, y2, -1.0 / 20922789888000.0);     /* -1/16! */
  pc = fma(pc, y2, 1.0 / 87178291200.0);         /*  1/14! */
  pc = fma(pc, y2, -1.0 / 479001600.0);          /* -1/12! */
  pc = fma(pc, y2, 1.0 / 3628800.0);             /*  1/10! */
  pc = fma(pc, y2, -1.0 / 40320.0);              /* -1/8!  */
  pc = fma(pc, y2, 1.0 / 720.0);                 /*  1/6!  */
  pc = fma(pc, y2, -1.0 / 24.0);                 /* -1/4!  */
  pc = fma(pc, y2, 0.5);
  const double c = fma(-y2, pc, 1.0);
  switch (q & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}

#endif
