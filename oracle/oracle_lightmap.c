/* oracle_lightmap.c — TEST INFRASTRUCTURE.  CPU restatement of the reference's lightmap_bake
 * (raytracer.c:722-784) with the two seeding rules of this repo:
 *
 *   ORACLE_SEED_REFERENCE   the two thread-local generator copies run as the reference runs them: one sequential
 *                           stream for the hemisphere directions (raytracer.c's copy, common.h:13,30-42 -> :767)
 *                           and one for the BSDF (driver.c's copy), continued from the states handed in;
 *   ORACLE_SEED_PER_SAMPLE  both are re-seeded before every sample from (texel, sample): directions from
 *                           rt_path_seed(texel, sample, user_seed ^ RT_LIGHTMAP_DIR_SALT), the BSDF from
 *                           rt_path_seed(texel, sample, user_seed) — the rule the GPU kernels share, which makes
 *                           every sample independent.
 *
 * Deviations from the reference, applied to the GPU path alike: texels outside the lightmap are
 * skipped (the reference writes out of bounds); the final f32 -> u8 conversion (raytracer.c:777-779 stores the
 * UN-SCALED float) saturates to [0, 255] and maps NaN to 0 where C leaves it undefined.
 */
#include <math.h>
#include <string.h>

#include "oracle.h"
#include "oracle_vec.h"
#include "rt_seed.h"

/* common.h:26-28 */
static inline f32 rand_range(u32 *state, f32 lo, f32 hi) { return rt_rand_f32(state) * (hi - lo) + lo; }

/* common.h:30-42: rejection from the cube, then scale by 1.0 / sqrt_f32(lensq) — a DOUBLE division narrowed to f32 */
static Vec3 rand_vec3(u32 *state) {
  for (;;) {
    Vec3 p;
    p.x = rand_range(state, -1, 1);
    p.y = rand_range(state, -1, 1);
    p.z = rand_range(state, -1, 1);
    f32 lensq = v3_dot(p, p);
    if (RT_EPSILON < lensq && lensq <= 1) return v3_scale(p, (f32)(1.0 / (f64)RT_SQRT_F32(lensq)));
  }
}

static inline u8 store_u8(f32 v) {
  if (!(v == v)) return 0;
  if (v <= 0) return 0;
  if (v >= 255) return 255;
  return (u8)v;
}

Color3 oracle_cast_ray(Scene const *scene, Ray ray, isize max_bounces);

void oracle_lightmap_bake(Image const *lightmap, Scene const *scene, isize samples, Oracle_Lightmap_Options *opt) {
  u32 dir_state = opt->dir_state;
  u32 *shader_state = oracle_shader_random_state();
  if (opt->seed_mode == ORACLE_SEED_REFERENCE) *shader_state = opt->shader_state;
  isize W = lightmap->width, H = lightmap->height;
  opt->texels_written = 0;

  for (isize i = 0; i < scene->triangles.len; i++) {
    Triangle_AOS aos = scene->triangles.aos[i];
    /* raytracer.c:728-731: float products truncated to i32 */
    i32 min_x = (i32)(f32_min(aos.tex_coords_a.x, f32_min(aos.tex_coords_b.x, aos.tex_coords_c.x)) * (f32)W);
    i32 max_x = (i32)(f32_max(aos.tex_coords_a.x, f32_max(aos.tex_coords_b.x, aos.tex_coords_c.x)) * (f32)W);
    i32 min_y = (i32)(f32_min(aos.tex_coords_a.y, f32_min(aos.tex_coords_b.y, aos.tex_coords_c.y)) * (f32)H);
    i32 max_y = (i32)(f32_max(aos.tex_coords_a.y, f32_max(aos.tex_coords_b.y, aos.tex_coords_c.y)) * (f32)H);

    f32 p0x = aos.tex_coords_a.x * (f32)W, p0y = aos.tex_coords_a.y * (f32)H;
    f32 p1x = aos.tex_coords_b.x * (f32)W, p1y = aos.tex_coords_b.y * (f32)H;
    f32 p2x = aos.tex_coords_c.x * (f32)W, p2y = aos.tex_coords_c.y * (f32)H;
    f32 denom = (p1y - p2y) * (p0x - p2x) + (p2x - p1x) * (p0y - p2y);

    for (isize y = min_y; y < (isize)max_y + 1; y++) {
      for (isize x = min_x; x < (isize)max_x + 1; x++) {
        f32 px = (f32)x, py = (f32)y;
        f32 w0 = ((p1y - p2y) * (px - p2x) + (p2x - p1x) * (py - p2y)) / denom;
        f32 w1 = ((p2y - p0y) * (px - p2x) + (p0x - p2x) * (py - p2y)) / denom;
        f32 w2 = 1.0f - w0 - w1;
        if (!(w0 >= -RT_EPSILON && w1 >= -RT_EPSILON && w2 >= -RT_EPSILON)) continue;
        /* deviation: the reference writes texels outside the image (a UV of exactly 1.0 gives x == W).  They are
         * not stored here; in reference seed mode their samples are still drawn, so the two streams stay in step */
        bool in_range = !(x < 0 || y < 0 || x >= W || y >= H);
        if (!in_range && opt->seed_mode != ORACLE_SEED_REFERENCE) continue;

        Vec3 position = v3(scene->triangles.x[0][i] * w0 + scene->triangles.x[1][i] * w1 + scene->triangles.x[2][i] * w2,
                           scene->triangles.y[0][i] * w0 + scene->triangles.y[1][i] * w1 + scene->triangles.y[2][i] * w2,
                           scene->triangles.z[0][i] * w0 + scene->triangles.z[1][i] * w1 + scene->triangles.z[2][i] * w2);
        Vec3 normal = v3(aos.normal_a.x * w0 + aos.normal_b.x * w1 + aos.normal_c.x * w2,
                         aos.normal_a.y * w0 + aos.normal_b.y * w1 + aos.normal_c.y * w2,
                         aos.normal_a.z * w0 + aos.normal_b.z * w1 + aos.normal_c.z * w2);
        Vec3 accumulated = v3(0, 0, 0);
        Ray r;
        r.position = v3_add(position, v3_scale(normal, RT_EPSILON));
        r.direction = v3(0, 0, 0);
        u32 texel = (u32)(x + y * W);
        bool lit = false;
        for (isize s = 0; s < samples; s++) {
          if (opt->seed_mode == ORACLE_SEED_PER_SAMPLE) {
            dir_state = rt_path_seed(texel, (u32)s, opt->user_seed ^ RT_LIGHTMAP_DIR_SALT);
            *shader_state = rt_path_seed(texel, (u32)s, opt->user_seed);
          }
          f32 cosine;
          /* raytracer.c:764-772.  A zero normal would loop forever in the reference; give up after 64 draws */
          int tries = 0;
          for (;;) {
            Vec3 d = rand_vec3(&dir_state);
            cosine = v3_dot(d, normal);
            if (cosine > 0) { r.direction = d; break; }
            if (++tries >= RT_LIGHTMAP_MAX_TRIES) { cosine = 0; break; }
          }
          if (!(cosine > 0)) continue;
          lit = true;
          accumulated = v3_add(accumulated, v3_scale(oracle_cast_ray(scene, r, 8), cosine));
        }
        (void)lit;
        if (!in_range) continue;
        f32 out[3] = { accumulated.x / (f32)samples, accumulated.y / (f32)samples, accumulated.z / (f32)samples };
        u8 *dst = lightmap->pixels.data + (x + y * lightmap->stride) * lightmap->components;
        dst[0] = store_u8(out[0]); dst[1] = store_u8(out[1]); dst[2] = store_u8(out[2]);
        if (opt->values) { f32 *v = opt->values + 3 * (size_t)texel; v[0] = out[0]; v[1] = out[1]; v[2] = out[2]; }
        if (opt->owner) opt->owner[texel] = (i32)i;
        opt->texels_written++;
      }
    }
  }
  opt->dir_state = dir_state;
  opt->shader_state = *shader_state;
}
