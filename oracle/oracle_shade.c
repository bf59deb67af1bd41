/* oracle_shade.c — TEST INFRASTRUCTURE.  Restatement of the reference's
 * material, texture and environment callbacks (reference driver.c:49-104,
 * 118-183, 200-409; common.h:82-88).
 *
 * The reference mixes f32 and f64 through unsuffixed literals and a double
 * PI (e.g. driver.c:119,214,220,238,241-246,263,313); those promotions are
 * kept literally so the rounding sequence is the reference's.  UNPINNED: that
 * Codin's PI is a double-typed literal.
 *
 * Deviation: asin's argument is clamped to [-1,1] (driver.c:100 would feed a
 * NaN texture coordinate into an integer cast, which is UB).
 */
#include <math.h>
#include <string.h>

#include "oracle.h"
#include "oracle_vec.h"
#include "rt_seed.h"

#define PI RT_PI

/* common.h:13 — one generator per thread in the shader's translation unit */
static _Thread_local u32 random_state = 0;
u32 *oracle_shader_random_state(void) { return &random_state; }
static inline f32 rand_f32(void) { return rt_rand_f32(&random_state); }

/* common.h:82-88 */
static inline Color3 decode_srgb(Color3 c) {
  return v3(rt_powf((c.x + 0.055f) / 1.055f, 2.4f),
            rt_powf((c.y + 0.055f) / 1.055f, 2.4f),
            rt_powf((c.z + 0.055f) / 1.055f, 2.4f));
}

static inline Color3 texel(Image const *img, isize u, isize v) {
  u8 const *p = img->pixels.data + img->components * (u + img->stride * v);
  return v3(p[0] / 255.999f, p[1] / 255.999f, p[2] / 255.999f);
}

/* driver.c:49-93 */
Color3 oracle_sample_texture_bilinear(Image const *img, Vec2 uv) {
  if (uv.x < 0) uv.x += (f32)(-(i32)uv.x + 1);
  if (uv.y < 0) uv.y += (f32)(-(i32)uv.y + 1);
  uv.x = uv.x - floorf(uv.x);
  uv.y = uv.y - floorf(uv.y);
  f32 px = uv.x * (f32)img->width;
  f32 py = uv.y * (f32)img->height;
  isize u = (isize)px, v = (isize)py;
  f32 a = px - (f32)u, b = py - (f32)v;
  isize u2 = (u + 1 < img->width)  ? u + 1 : u;
  isize v2 = (v + 1 < img->height) ? v + 1 : v;
  Color3 top = v3_lerp(texel(img, u, v),  texel(img, u2, v),  a);
  Color3 bot = v3_lerp(texel(img, u, v2), texel(img, u2, v2), a);
  return v3_lerp(top, bot, b);
}

/* driver.c:95-104 */
Color3 oracle_sample_background(rawptr image, Vec3 dir) {
  f32 inv_pi     = (f32)(1.0f / PI);
  f32 inv_two_pi = (f32)(1.0f / (2.0f * PI));
  f32 u = 0.5f + rt_atan2f(dir.z, dir.x) * inv_two_pi;
  f32 v = 0.5f - rt_asinf(f32_clamp(dir.y, -1.0f, 1.0f)) * inv_pi;
  Vec2 uv; uv.x = u; uv.y = v;
  return decode_srgb(oracle_sample_texture_bilinear((Image const *)image, uv));
}

/* driver.c:118-127 */
static inline Vec3 cosine_hemisphere(void) {
  f32 angle  = (f32)(rand_f32() * 2 * PI);
  f32 radius = RT_SQRT_F32(rand_f32());
  Vec3 d;
  d.x = rt_sinf(angle) * radius;
  d.y = rt_cosf(angle) * radius;
  d.z = RT_SQRT_F32(1 - radius * radius);
  return d;
}

/* driver.c:129-153 */
static inline Vec3 perturb_normal(Image const *normal_map, f32 strength, Shader_Input const *in) {
  if (!normal_map) return in->normal;
  Vec3 s = oracle_sample_texture_bilinear(normal_map, in->tex_coords);
  s = v3_add(v3_scale(s, 2.0f), v3(-1, -1, -1));
  s.y *= -1;
  Vec3 t = in->tangent, b = in->bitangent, n = in->normal;
  f32  k = strength;
  return v3_normalize(v3(k * (s.x * t.x + s.y * b.x + s.z * n.x) + n.x * (1 - k),
                         k * (s.x * t.y + s.y * b.y + s.z * n.y) + n.y * (1 - k),
                         k * (s.x * t.z + s.y * b.z + s.z * n.z) + n.z * (1 - k)));
}

/* driver.c:155-164 */
static inline void shading_frame(Vec3 view, Vec3 n, Vec3 *t, Vec3 *b) {
  if (f32_abs(v3_dot(n, view)) < 0.9999f)             *t = v3_normalize(v3_cross(n, view));
  else if (f32_abs(v3_dot(n, v3(0, 1, 0))) < 0.9999f) *t = v3_normalize(v3_cross(n, v3(0, 1, 0)));
  else                                                *t = v3_normalize(v3_cross(n, v3(1, 0, 0)));
  *b = v3_cross(n, *t);
}

/* driver.c:200-202 */
static inline f32 luma(Color3 c) { return v3_dot(c, v3(0.2126f, 0.7152f, 0.0722f)); }

/* driver.c:166-183 */
static inline Vec3 sheen_term(f32 sheen, Color3 base, f32 sheen_tint, f32 h_dot_l) {
  if (sheen <= 0.0f) return v3_splat(0);
  f32 lum = v3_dot(v3(0.3f, 0.6f, 1.0f), base);
  Vec3 tint = (lum > 0.0f) ? v3_scale(base, 1.0f / lum) : v3(1, 1, 1);
  f32 m = 1 - h_dot_l;
  f32 w = m * m * m * m * m;
  return v3_scale(v3_lerp(v3(1, 1, 1), tint, sheen_tint), sheen * w);
}

/* driver.c:204-210 */
static inline f32 schlick1(f32 f0, f32 f90, f32 c) { return f0 + (f90 - f0) * rt_powf(1 - c, 5); }
static inline Vec3 schlick3(Color3 f0, f32 f90, f32 c) {
  return v3_add(f0, v3_scale(v3_sub(v3_splat(f90), f0), rt_powf(1 - c, 5)));
}

/* driver.c:212-215 (the author's GGX variant: alpha = roughness^2 squared again) */
static inline f32 ggx_d(f32 roughness, f32 n_h, f32 k) {
  f32 a2 = roughness * roughness;
  return (f32)(a2 / (PI * rt_powf((n_h * n_h) * (a2 * a2 - 1) + 1, k)));
}

/* driver.c:217-221 */
static inline f32 smith_g1(f32 n_v, f32 alpha2) {
  f32 a = alpha2 * alpha2;
  f32 b = n_v * n_v;
  return (f32)((2.0 * n_v) / (n_v + RT_SQRT_F32(a + b - a * b)));
}

/* driver.c:230-250 */
static inline Vec3 sample_vndf(Vec3 V, f32 ax, f32 ay) {
  Vec3 Vh = v3_normalize(v3(ax * V.x, ay * V.y, V.z));
  f32 lensq = Vh.x * Vh.x + Vh.y * Vh.y;
  Vec3 T1 = lensq > 0 ? v3_scale(v3(-Vh.y, Vh.x, 0), 1.0f / RT_SQRT_F32(lensq)) : v3(1, 0, 0);
  Vec3 T2 = v3_cross(Vh, T1);

  f32 r   = RT_SQRT_F32(rand_f32());
  f32 phi = (f32)(2.0 * PI * rand_f32());
  f32 t1  = r * rt_cosf(phi);
  f32 t2  = r * rt_sinf(phi);
  f32 s   = (f32)(0.5 * (1.0 + Vh.z));
  t2      = (f32)((1.0 - s) * RT_SQRT_F32((f32)(1.0 - t1 * t1)) + s * t2);

  f64 rem = 1.0 - t1 * t1 - t2 * t2;
  f32 h = RT_SQRT_F32((f32)(0.0 > rem ? 0.0 : rem));             /* max(0.0, .) as a macro */
  Vec3 Nh = v3_add(v3_add(v3_scale(T1, t1), v3_scale(T2, t2)), v3_scale(Vh, h));
  return v3_normalize(v3(ax * Nh.x, ay * Nh.y, (0.0 > Nh.z ? 0.0f : Nh.z)));
}

typedef struct {
  f32    roughness, metalness, sheen, sheen_tint, aniso2;
  Color3 base;
} Lobe_Params;

/* driver.c:287-348 */
static Vec4 sample_bsdf(Lobe_Params const *p, Vec3 wi, Vec3 *wo) {
  f32  alpha_x = f32_lerp(p->roughness * p->roughness, 1, p->aniso2);
  f32  alpha_y = p->roughness * p->roughness;
  Vec3 h = sample_vndf(wi, alpha_x, alpha_y);

  Color3 f0 = v3_lerp(v3_splat(0.04f), p->base, p->metalness);
  f32 f90 = f32_min(1.0f, (1.0f / 0.04f) * luma(f0));          /* driver.c:273-276 */
  Color3 F = schlick3(f0, f90, v3_dot(wi, h));

  f32 w_diff = 1 - p->metalness;
  f32 w_spec = luma(F);
  f32 inv_w  = 1 / (w_diff + w_spec);
  w_diff *= inv_w;
  w_spec *= inv_w;

  Vec4 out; out.x = out.y = out.z = out.w = 0;
  if (rand_f32() < w_diff) {
    *wo = cosine_hemisphere();
    h = v3_normalize(v3_add(*wo, wi));
    f32 n_l = wo->z, n_v = wi.z;
    if (n_l <= 0 || n_v <= 0) return out;
    f32 l_h = v3_dot(*wo, h);
    f32 pdf = (f32)(n_l / PI);
    /* driver.c:258-264 */
    f32 fd90 = 0.5f + 2 * p->roughness * l_h * l_h;
    f32 fa = schlick1(1.0f, fd90, n_l);
    f32 fb = schlick1(1.0f, fd90, n_v);
    Color3 diff = v3_scale(p->base, (f32)(fa * fb / PI));
    diff = v3_mul(diff, v3_sub(v3(1, 1, 1), F));
    diff = v3_add(diff, sheen_term(p->sheen, p->base, p->sheen_tint, l_h));
    out.x = diff.x * n_l; out.y = diff.y * n_l; out.z = diff.z * n_l;
    out.w = w_diff * pdf;
  } else {
    *wo = v3_reflect(v3_scale(wi, -1), h);
    f32 n_l = wo->z, n_v = wi.z;
    if (n_l <= 0 || n_v <= 0) return out;
    n_l = f32_max(n_l, 0.001f);
    n_v = f32_max(n_v, 0.001f);
    f32 n_h = f32_min(h.z, 0.99f);
    /* driver.c:252-256 */
    f32 a2  = p->roughness * p->roughness;
    f32 pdf = (ggx_d(p->roughness, n_h, 2) * smith_g1(n_v, a2)) / f32_max(0.00001f, 4.0f * n_v);
    /* driver.c:266-271, 223-228 */
    f32 D = ggx_d(p->roughness, n_h, 2);
    f32 G = smith_g1(n_v, a2) * smith_g1(n_l, a2);
    Vec3 spec = v3_scale(F, D * G / (4 * n_l * n_v));
    out.x = spec.x * n_l; out.y = spec.y * n_l; out.z = spec.z * n_l;
    out.w = w_spec * pdf;
  }
  *wo = v3_normalize(*wo);
  return out;
}

/* driver.c:350-409 */
void oracle_disney_shader_proc(rawptr data, Shader_Input const *in, Shader_Output *out) {
  PBR_Shader_Data const *mat = (PBR_Shader_Data const *)data;
  Vec3 n = perturb_normal(mat->texture_normal, mat->normal_map_strength, in);

  Color3 base = mat->base_color;
  if (mat->texture_albedo)
    base = v3_mul(base, decode_srgb(oracle_sample_texture_bilinear(mat->texture_albedo, in->tex_coords)));

  f32 roughness = mat->roughness, metalness = mat->metalness;
  if (mat->texture_metal_roughness) {
    Vec3 mr = oracle_sample_texture_bilinear(mat->texture_metal_roughness, in->tex_coords);
    roughness *= mr.y;
    metalness *= mr.z;
  }
  roughness = f32_clamp(roughness, 0.001f, 1);
  if (metalness > 0.9f) metalness = 0.9f;
  metalness /= 0.9f;

  Color3 glow = mat->emission;
  if (mat->texture_emission)
    glow = v3_mul(glow, decode_srgb(oracle_sample_texture_bilinear(mat->texture_emission, in->tex_coords)));
  out->emission = glow;

  Vec3 t, b;
  shading_frame(in->direction, n, &t, &b);

  Lobe_Params p;
  p.roughness  = roughness;
  p.metalness  = metalness;
  p.base       = base;
  p.sheen      = mat->sheen;
  p.sheen_tint = mat->sheen_tint;
  p.aniso2     = mat->anisotropic_strength * mat->anisotropic_strength;

  /* world_to_tangent = transpose(from_basis(t, b, n)): rows are t, b, n.
   * UNPINNED: matrix_3x3_from_basis puts t, b, n in columns. */
  Vec3 minus_d = v3_scale(in->direction, -1);
  Vec3 wi = v3(v3_dot(t, minus_d), v3_dot(b, minus_d), v3_dot(n, minus_d));
  Vec3 wo;
  Vec4 f = sample_bsdf(&p, wi, &wo);

  out->direction = v3(t.x * wo.x + b.x * wo.y + n.x * wo.z,
                      t.y * wo.x + b.y * wo.y + n.y * wo.z,
                      t.z * wo.x + b.z * wo.y + n.z * wo.z);
  if (f.w > 0) out->tint = v3(f.x / f.w, f.y / f.w, f.z / f.w);
  else         out->terminate = true;
}
