/* rt_fastdiv.h — unsigned division by a launch-time constant as multiply-high + shift.
 *
 * The primary trace kernel turns a path id into (pixel, sample) and a job tile into (x, y): three divisions by
 * numbers that are fixed per launch (samples per chunk, tiles per row, chunks per row).  A 32-bit division by a
 * register is ~20 instructions on the device; with the divisor's magic pair it is two.
 *
 *   q = umulhi(n, mul) >> shift      exact for every n <= n_max
 *
 * rt_fastdiv_make picks the smallest shift for which mul = ceil(2^(32+shift) / d) fits 32 bits and
 * n_max * (mul * d - 2^(32+shift)) < 2^(32+shift) — the classic round-up method's exactness condition.  When no
 * such pair exists (d == 1, or a divisor too awkward for the range) mul is 0 and rt_fastdiv divides.
 */
#ifndef RT_FASTDIV_H
#define RT_FASTDIV_H

#include <stdint.h>

typedef struct { uint32_t mul, shift; } RT_FastDiv;

static inline RT_FastDiv rt_fastdiv_make(uint32_t d, uint32_t n_max) {
  RT_FastDiv f = { 0u, 0u };
  if (d < 2u) return f;
  for (uint32_t k = 0; k < 31u; k++) {
    const uint64_t two_k = (uint64_t)1 << (32u + k);
    const uint64_t m = (two_k + d - 1u) / d;
    if (m >> 32) break;
    const uint64_t err = m * d - two_k;                       /* < d */
    if (err == 0 || (uint64_t)n_max < (two_k + err - 1u) / err) {   /* n_max * err < 2^(32+k) */
      f.mul = (uint32_t)m; f.shift = k;
      return f;
    }
  }
  return f;
}

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned rt_fastdiv(unsigned n, const RT_FastDiv f, unsigned d) {
  return f.mul ? __umulhi(n, f.mul) >> f.shift : n / d;
}
#else
static inline uint32_t rt_fastdiv(uint32_t n, const RT_FastDiv f, uint32_t d) {
  return f.mul ? (uint32_t)(((uint64_t)n * f.mul) >> 32) >> f.shift : n / d;
}
#endif

#endif
