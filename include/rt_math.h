/* rt_math.h — transcendental functions with ONE definition for gcc and nvcc.
 *
 * Why: the reference calls pow_f32 / sin_f32 / cos_f32 / atan2_f32 / asin_f32
 * from its absent stdlib (driver.c:99-100,119-123,204-214,238-240;
 * common.h:82-92).  glibc and CUDA's libdevice differ from each other by ulps,
 * and one ulp at a lobe-pick or a texel boundary sends a whole path elsewhere.
 * Every function here is built only from IEEE-754 +,-,*,/ and sqrt in binary64
 * plus integer bit moves and EXPLICIT fused multiply-adds (RT_FMA), so — compiled
 * without automatic FMA contraction (gcc -ffp-contract=off, nvcc -fmad=false) —
 * the CPU oracle and the sm_100a kernels produce bit-identical results.  Accuracy is < 1 ulp of binary32
 * (tests/test_rt_math.py checks against libm), i.e. the same contract a libm
 * gives; the reference's own libm is unpinned (Codin is not in the tree).
 */
#ifndef RT_MATH_H
#define RT_MATH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
/* the f64 bodies stay out of line on the device: inlined at every call site they
 * pushed the render kernel to 99 KB of SASS, past the instruction cache */
#define RT_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define RT_HD static inline
#define RT_HD_NOINLINE static inline
#endif

/* Fused multiply-add, EXPLICIT and identical on both sides (one IEEE rounding):
 * DFMA on the device, vfmadd on the host (oracle built with -mfma; without the
 * instruction __builtin_fma falls back to libm's exact software fma).  Automatic
 * contraction stays off everywhere else. */
#if defined(__CUDA_ARCH__)
#define RT_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define RT_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

#define RT_PI      3.14159265358979323846
#define RT_PI_F64  RT_PI

RT_HD double rt_u64_as_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { uint64_t u; double d; } c; c.u = u; return c.d;
#endif
}
RT_HD uint64_t rt_f64_as_u64(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  union { uint64_t u; double d; } c; c.d = d; return c.u;
#endif
}
RT_HD uint32_t rt_f32_as_u32(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { uint32_t u; float f; } c; c.f = f; return c.u;
#endif
}
RT_HD float rt_u32_as_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}

/* floor for |v| < 2^51, exact, no libm. */
RT_HD double rt_floor_f64(double v) {
  double t = (double)(int64_t)v;      /* trunc toward zero */
  return (t > v) ? t - 1.0 : t;
}

/* log2 of a finite positive double: exponent + 2*atanh((m-1)/(m+1))/ln2,
 * m in [sqrt(1/2), sqrt(2)].  Truncation error < 2e-12 relative. */
RT_HD double rt_log2_pos(double d) {
  uint64_t bits = rt_f64_as_u64(d);
  int      e    = (int)((bits >> 52) & 0x7ff) - 1023;
  double   m    = rt_u64_as_f64((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
  if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
  double s  = (m - 1.0) / (m + 1.0);
  double s2 = s * s;
  double p  = 1.0 / 15.0;
  p = RT_FMA(p, s2, 1.0 / 13.0);
  p = RT_FMA(p, s2, 1.0 / 11.0);
  p = RT_FMA(p, s2, 1.0 / 9.0);
  p = RT_FMA(p, s2, 1.0 / 7.0);
  p = RT_FMA(p, s2, 1.0 / 5.0);
  p = RT_FMA(p, s2, 1.0 / 3.0);
  p = RT_FMA(p, s2, 1.0);
  /* 2/ln(2) */
  return RT_FMA(s * p, 2.8853900817779268, (double)e);
}

/* 2^t for |t| <= 1000 → double (caller clamps). */
RT_HD double rt_exp2_f64(double t) {
  double k = rt_floor_f64(t + 0.5);
  double f = (t - k) * 0.6931471805599453;   /* |f| <= 0.3466 */
  double p = 1.0 / 39916800.0;
  p = RT_FMA(p, f, 1.0 / 3628800.0);
  p = RT_FMA(p, f, 1.0 / 362880.0);
  p = RT_FMA(p, f, 1.0 / 40320.0);
  p = RT_FMA(p, f, 1.0 / 5040.0);
  p = RT_FMA(p, f, 1.0 / 720.0);
  p = RT_FMA(p, f, 1.0 / 120.0);
  p = RT_FMA(p, f, 1.0 / 24.0);
  p = RT_FMA(p, f, 1.0 / 6.0);
  p = RT_FMA(p, f, 0.5);
  p = RT_FMA(p, f, 1.0);
  p = RT_FMA(p, f, 1.0);
  int64_t ki = (int64_t)k;
  return p * rt_u64_as_f64((uint64_t)(ki + 1023) << 52);
}

/* powf with C99 special cases for the inputs the path can produce. */
RT_HD_NOINLINE float rt_powf(float x, float y) {
  if (y == 0.0f) return 1.0f;
  if (x != x || y != y) return x + y;
  float sign = 1.0f;
  if (x < 0.0f) {
    float yi = (float)(int32_t)y;
    if (!(y > -16777216.0f && y < 16777216.0f)) yi = y;   /* huge |y| is an even integer */
    else if (yi != y) return rt_u32_as_f32(0x7fc00000u);  /* negative base, fractional power */
    else if (((int32_t)y) & 1) sign = -1.0f;
    x = -x;
  }
  if (x == 0.0f) return (y > 0.0f) ? 0.0f : rt_u32_as_f32(0x7f800000u);
  if (x == rt_u32_as_f32(0x7f800000u)) return (y > 0.0f) ? sign * x : 0.0f;
  double t = (double)y * rt_log2_pos((double)x);
  if (t >  300.0) return sign * rt_u32_as_f32(0x7f800000u);
  if (t < -300.0) return sign * 0.0f;
  return sign * (float)rt_exp2_f64(t);
}

/* sin and cos of a float angle, |x| < 1e6.  Cody–Waite reduction by pi/2 in
 * binary64 (the high part has 33 significant bits, so q*hi is exact). */
RT_HD void rt_sincos_reduce(float x, double *r, int *quadrant) {
  double d = (double)x;
  double q = rt_floor_f64(d * 0.6366197723675814 + 0.5);
  double a = RT_FMA(-q, 1.5707963267341256, d);      /* pi/2 high: q*hi is exact */
  a = RT_FMA(-q, 6.077100506506192e-11, a);          /* pi/2 low  */
  *r = a;
  *quadrant = (int)((int64_t)q & 3);
}
RT_HD double rt_sin_poly(double r) {
  double r2 = r * r;
  double p = 1.0 / 6227020800.0;
  p = RT_FMA(p, r2, -1.0 / 39916800.0);
  p = RT_FMA(p, r2, 1.0 / 362880.0);
  p = RT_FMA(p, r2, -1.0 / 5040.0);
  p = RT_FMA(p, r2, 1.0 / 120.0);
  p = RT_FMA(p, r2, -1.0 / 6.0);
  p = RT_FMA(p, r2, 1.0);
  return r * p;
}
RT_HD double rt_cos_poly(double r) {
  double r2 = r * r;
  double p = 1.0 / 87178291200.0;
  p = RT_FMA(p, r2, -1.0 / 479001600.0);
  p = RT_FMA(p, r2, 1.0 / 3628800.0);
  p = RT_FMA(p, r2, -1.0 / 40320.0);
  p = RT_FMA(p, r2, 1.0 / 720.0);
  p = RT_FMA(p, r2, -1.0 / 24.0);
  p = RT_FMA(p, r2, 0.5);
  return RT_FMA(-r2, p, 1.0);
}
RT_HD_NOINLINE float rt_sinf(float x) {
  double r; int q;
  rt_sincos_reduce(x, &r, &q);
  double v = (q & 1) ? rt_cos_poly(r) : rt_sin_poly(r);
  return (float)((q & 2) ? -v : v);
}
RT_HD_NOINLINE float rt_cosf(float x) {
  double r; int q;
  rt_sincos_reduce(x, &r, &q);
  double v = (q & 1) ? rt_sin_poly(r) : rt_cos_poly(r);
  return (float)(((q + 1) & 2) ? -v : v);
}

/* atan2 in binary64 on finite inputs; one division. */
RT_HD_NOINLINE double rt_atan2_f64(double y, double x) {
  int neg_y = (int)(rt_f64_as_u64(y) >> 63);
  int neg_x = (int)(rt_f64_as_u64(x) >> 63);
  double ay = neg_y ? -y : y;
  double ax = neg_x ? -x : x;
  int    swap = ay > ax;
  double mn = swap ? ax : ay;
  double mx = swap ? ay : ax;
  double a;
  if (mx == 0.0) {
    a = 0.0;
  } else {
    double base = 0.0, z;
    if (mn > 0.41421356237309503 * mx) { z = (mn - mx) / (mn + mx); base = 0.78539816339744831; }
    else                               { z = mn / mx; }
    double z2 = z * z;
    double p = 1.0 / 27.0;
    p = RT_FMA(-p, z2, 1.0 / 25.0);
    p = RT_FMA(-p, z2, 1.0 / 23.0);
    p = RT_FMA(-p, z2, 1.0 / 21.0);
    p = RT_FMA(-p, z2, 1.0 / 19.0);
    p = RT_FMA(-p, z2, 1.0 / 17.0);
    p = RT_FMA(-p, z2, 1.0 / 15.0);
    p = RT_FMA(-p, z2, 1.0 / 13.0);
    p = RT_FMA(-p, z2, 1.0 / 11.0);
    p = RT_FMA(-p, z2, 1.0 / 9.0);
    p = RT_FMA(-p, z2, 1.0 / 7.0);
    p = RT_FMA(-p, z2, 1.0 / 5.0);
    p = RT_FMA(-p, z2, 1.0 / 3.0);
    p = RT_FMA(-p, z2, 1.0);
    a = RT_FMA(z, p, base);
  }
  if (swap)  a = 1.57079632679489662 - a;
  if (neg_x) a = RT_PI - a;
  return neg_y ? -a : a;
}
RT_HD float rt_atan2f(float y, float x) {
  return (float)rt_atan2_f64((double)y, (double)x);
}

#if defined(__CUDA_ARCH__)
#define RT_SQRT_F64(v) __dsqrt_rn(v)
#define RT_SQRT_F32(v) __fsqrt_rn(v)
#else
#define RT_SQRT_F64(v) __builtin_sqrt(v)
#define RT_SQRT_F32(v) __builtin_sqrtf(v)
#endif

/* asin on [-1,1]; NaN outside like libm. */
RT_HD float rt_asinf(float x) {
  double d = (double)x;
  double c = (1.0 - d) * (1.0 + d);
  if (c < 0.0 || x != x) return rt_u32_as_f32(0x7fc00000u);
  return (float)rt_atan2_f64(d, RT_SQRT_F64(c));
}

#endif
