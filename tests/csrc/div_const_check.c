/* Exhaustive check of the three-instruction division by 1.055f used by the kernels' sRGB decode (rt_device.cuh,
 * div_1p055): q = x*r; e = fma(-d, q, x); q' = fma(e, r, q) must equal the IEEE quotient x / 1.055f for every
 * binary32 x the decode can see (x = c + 0.055f with c in [0, 1]) — checked on a wider interval. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void) {
  const float d = 1.055f, r = 1.0f / 1.055f;
  float lo = 0.01f, hi = 4.0f;
  uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
  unsigned long bad = 0, n = 0;
  for (uint32_t u = a; u <= b; u++) {
    float x; memcpy(&x, &u, 4);
    volatile float q0 = x * r;
    float e = fmaf(-d, q0, x);
    float q = fmaf(e, r, q0);
    volatile float ref = x / d;
    if (q != ref) { if (bad < 5) printf("x=%a got %a want %a\n", x, q, ref); bad++; }
    n++;
  }
  printf("%lu values, %lu mismatches\n", n, bad);
  puts(bad ? "FAIL" : "ok");
  return bad != 0;
}
