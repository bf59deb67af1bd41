/* The samplers turn a texel byte b into (float)b / 255.999f (driver.c:69-88).  The kernels do it in two instructions:
 * PRMT drops the byte into the mantissa of 2^23 (the float 8388608 + b, exact), one fused multiply-add computes
 * (2^23 + b) * k - 2^23 * k with k = RN(1 / 255.999f): the sum is b * k exactly, rounded once.  This checks, for all 256
 * bytes, that it equals the IEEE quotient. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void) {
  const float k = 1.0f / 255.999f, c = -8388608.0f * k;
  int bad = 0;
  for (uint32_t b = 0; b < 256; b++) {
    uint32_t bits = 0x4B000000u | b;
    float x; memcpy(&x, &bits, 4);
    if (x != 8388608.0f + (float)b) bad++;
    volatile float ref = (float)b / 255.999f;
    float got = fmaf(x, k, c);
    volatile float mul = (float)b * k;
    if (got != ref || mul != ref) { printf("b=%u got %a mul %a want %a\n", b, got, mul, ref); bad++; }
  }
  puts(bad ? "FAIL" : "ok");
  return bad != 0;
}
