/* oracle_vec.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Scalar vector helpers standing in for Codin's linalg.h, which the reference
 * includes but does not ship (reference common.h:3-4).  The operation ORDER in
 * each helper is part of the oracle's definition because it fixes the f32
 * rounding; where Codin's choice cannot be known the assumption is marked
 * UNPINNED (see DESIGN.md "Unpinned Codin semantics").
 */
#ifndef ORACLE_VEC_H
#define ORACLE_VEC_H

#include "rt_base.h"
#include "rt_math.h"

static inline Vec3 v3(f32 x, f32 y, f32 z) { Vec3 v; v.x = x; v.y = y; v.z = z; return v; }
static inline Vec3 v3_splat(f32 s)          { return v3(s, s, s); }
static inline Vec3 v3_add(Vec3 a, Vec3 b)   { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline Vec3 v3_sub(Vec3 a, Vec3 b)   { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline Vec3 v3_mul(Vec3 a, Vec3 b)   { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline Vec3 v3_scale(Vec3 a, f32 s)  { return v3(a.x * s, a.y * s, a.z * s); }
/* UNPINNED: left-to-right sum. */
static inline f32  v3_dot(Vec3 a, Vec3 b)   { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline Vec3 v3_cross(Vec3 a, Vec3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* UNPINNED: scale by 1/sqrt(len2), the form the author uses at common.h:39. */
static inline Vec3 v3_normalize(Vec3 a) {
  f32 inv = 1.0f / RT_SQRT_F32(v3_dot(a, a));
  return v3_scale(a, inv);
}
/* UNPINNED: a*(1-t) + b*t, the form of the author's own lerp_f32 (driver.c:283-285). */
static inline f32  f32_lerp(f32 a, f32 b, f32 t) { return a * (1 - t) + b * t; }
static inline Vec3 v3_lerp(Vec3 a, Vec3 b, f32 t) {
  return v3(f32_lerp(a.x, b.x, t), f32_lerp(a.y, b.y, t), f32_lerp(a.z, b.z, t));
}
/* UNPINNED: I - N*(2*dot(I,N)); called as reflect(-V, H) at driver.c:324. */
static inline Vec3 v3_reflect(Vec3 i, Vec3 n) {
  return v3_sub(i, v3_scale(n, 2.0f * v3_dot(i, n)));
}
static inline f32 f32_min(f32 a, f32 b) { return a < b ? a : b; }
static inline f32 f32_max(f32 a, f32 b) { return a > b ? a : b; }
static inline f32 f32_abs(f32 a)        { return a < 0 ? -a : a; }
static inline f32 f32_clamp(f32 x, f32 lo, f32 hi) { return x < lo ? lo : (x > hi ? hi : x); }

#endif
