/* rt_seed.h — the per-(pixel, sample) RNG seeding rule ("parity mode").
 *
 * The reference's BSDF randomness is one thread-local stream that starts at 0
 * and runs across every pixel a thread happens to pull (common.h:13-24,
 * driver.c:119-120,237-238,303), which no parallel renderer
 * can reproduce.  Parity mode instead sets random_state to rt_path_seed(...)
 * immediately before each cast_ray (the call at raytracer.c:695); the oracle
 * and the CUDA kernels share this one definition, which makes results
 * independent of tiling, thread count, GPU count and spp split.
 * The generator itself (rand_u32 / rand_f32) is restated literally.
 */
#ifndef RT_SEED_H
#define RT_SEED_H

#include <stdint.h>
#include "rt_math.h"

RT_HD uint32_t rt_path_seed(uint32_t pixel_index, uint32_t sample_index, uint32_t user_seed) {
  uint32_t h = (pixel_index * 0x9E3779B1u) ^ (sample_index * 0x85EBCA77u) ^ user_seed;
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}

/* lightmap_bake (reference raytracer.c:722-784) draws its hemisphere directions from a second generator copy
 * (raytracer.c's, common.h:13,30-42); per-(texel, sample) seeding gives that stream its own salt.  A normal that
 * no direction satisfies (zero vector, NaN) would loop forever in the reference: both sides stop after this many draws. */
#define RT_LIGHTMAP_DIR_SALT 0xD1B54A32u
#define RT_LIGHTMAP_MAX_TRIES 64

/* reference common.h:15-20 — the output is fed back as the state. */
RT_HD uint32_t rt_rand_u32(uint32_t *random_state) {
  uint32_t state = *random_state * 747796405u + 2891336453u;
  uint32_t word  = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  *random_state  = (word >> 22u) ^ word;
  return *random_state;
}

/* reference common.h:22-24 — u / (f32)U32_MAX, i.e. u * 2^-32 with the u32→f32
 * conversion rounding to nearest; range [0,1] inclusive. */
RT_HD float rt_rand_f32(uint32_t *random_state) {
  return (float)rt_rand_u32(random_state) / 4294967296.0f;
}

#endif
