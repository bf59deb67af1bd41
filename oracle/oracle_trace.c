/* oracle_trace.c — TEST INFRASTRUCTURE.  Restatement of the reference's
 * path-tracing core: 8-wide slab and Möller–Trumbore kernels, ordered 8-ary
 * traversal, cast_ray, the chunk scheduler, camera and film
 * (reference raytracer.c:15-32,84-230,443-558,582-720).
 *
 * Build: gcc -O2 -mavx2 -mno-fma -ffp-contract=off.  No FMA anywhere: every
 * product and sum rounds separately, which is what the sm_100a kernels do
 * (-fmad=false), so primary-hit slots compare bit-exactly.
 *
 * Deviations from the reference:
 *   - exact 1/sqrt for primary-ray normalisation unless approx_rsqrt is set
 *     (raytracer.c:663 uses the vendor-specific _mm256_rsqrt_ps);
 *   - optional per-(pixel,sample) seeding (rt_seed.h);
 *   - NaN radiance stores as 0 (the (u8) cast at raytracer.c:714 is UB for NaN).
 */
#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "oracle_vec.h"
#include "rt_seed.h"

typedef struct {
  f32    distance;
  Vec3   normal, normal_geo, point, tangent, bitangent;
  Vec2   tex_coords;
  Shader shader;
  i32    slot;
} Hit;

typedef struct {
  u64 c[8];
} Counters;

static _Thread_local Counters *tl_counters;

/* ---- raytracer.c:15-32: horizontal min with lowest-lane argmin; lanes that
 * are <= eps or NaN count as +inf. ---- */
static inline f32 lane_min(__m256 v, f32 eps, i32 *lane) {
  __m256 inf   = _mm256_set1_ps(INFINITY);
  __m256 keep  = _mm256_cmp_ps(v, _mm256_set1_ps(eps), _CMP_GT_OQ);
  __m256 clean = _mm256_blendv_ps(inf, v, keep);
  __m256 m = _mm256_min_ps(clean, _mm256_permute_ps(clean, 0xB1));
  m = _mm256_min_ps(m, _mm256_permute_ps(m, 0x4E));
  m = _mm256_min_ps(m, _mm256_permute2f128_ps(m, m, 0x01));
  int eq = _mm256_movemask_ps(_mm256_cmp_ps(clean, m, _CMP_EQ_OQ));
  *lane = __builtin_ctz((unsigned)eq);
  return _mm256_cvtss_f32(m);
}

#define SUB(a, b) _mm256_sub_ps(a, b)
#define MUL(a, b) _mm256_mul_ps(a, b)
#define ADD(a, b) _mm256_add_ps(a, b)

typedef struct { __m256 x, y, z; } V8;

static inline V8 v8_cross(V8 a, V8 b) {
  V8 r;
  r.x = SUB(MUL(a.y, b.z), MUL(a.z, b.y));
  r.y = SUB(MUL(a.z, b.x), MUL(a.x, b.z));
  r.z = SUB(MUL(a.x, b.y), MUL(a.y, b.x));
  return r;
}
static inline __m256 v8_dot(V8 a, V8 b) {
  return ADD(ADD(MUL(a.x, b.x), MUL(a.y, b.y)), MUL(a.z, b.z));
}

/* ---- raytracer.c:84-188 ---- */
static inline bool leaf_test(Ray const *ray, Triangles const *tris, isize offset, Hit *hit) {
  V8 d = { _mm256_set1_ps(ray->direction.x), _mm256_set1_ps(ray->direction.y), _mm256_set1_ps(ray->direction.z) };
  V8 o = { _mm256_set1_ps(ray->position.x),  _mm256_set1_ps(ray->position.y),  _mm256_set1_ps(ray->position.z) };
  V8 p0 = { _mm256_load_ps(tris->x[0] + offset), _mm256_load_ps(tris->y[0] + offset), _mm256_load_ps(tris->z[0] + offset) };
  V8 p1 = { _mm256_load_ps(tris->x[1] + offset), _mm256_load_ps(tris->y[1] + offset), _mm256_load_ps(tris->z[1] + offset) };
  V8 p2 = { _mm256_load_ps(tris->x[2] + offset), _mm256_load_ps(tris->y[2] + offset), _mm256_load_ps(tris->z[2] + offset) };

  V8 e1 = { SUB(p1.x, p0.x), SUB(p1.y, p0.y), SUB(p1.z, p0.z) };
  V8 e2 = { SUB(p2.x, p0.x), SUB(p2.y, p0.y), SUB(p2.z, p0.z) };

  V8     pvec    = v8_cross(d, e2);
  __m256 det     = v8_dot(e1, pvec);
  __m256 inv_det = _mm256_div_ps(_mm256_set1_ps(1.0f), det);
  V8     tvec    = { SUB(o.x, p0.x), SUB(o.y, p0.y), SUB(o.z, p0.z) };
  V8     qvec    = v8_cross(tvec, e1);

  __m256 u = MUL(inv_det, v8_dot(tvec, pvec));
  __m256 v = MUL(inv_det, v8_dot(d, qvec));
  __m256 t = MUL(inv_det, v8_dot(e2, qvec));

  __m256 lo  = _mm256_set1_ps(-RT_EPSILON);
  __m256 hi  = _mm256_set1_ps(1 + RT_EPSILON);
  __m256 out = _mm256_or_ps(_mm256_cmp_ps(u, lo, _CMP_LT_OQ), _mm256_cmp_ps(u, hi, _CMP_GT_OQ));
  out = _mm256_or_ps(out, _mm256_or_ps(_mm256_cmp_ps(v, lo, _CMP_LT_OQ), _mm256_cmp_ps(ADD(u, v), hi, _CMP_GT_OQ)));
  out = _mm256_or_ps(out, _mm256_cmp_ps(t, _mm256_set1_ps(RT_EPSILON), _CMP_LT_OQ));

  __m256 dist = _mm256_blendv_ps(t, _mm256_set1_ps(INFINITY), out);

  i32 lane;
  f32 nearest = lane_min(dist, 0, &lane);
  if (!(nearest < hit->distance)) return false;

  f32 us[8] __attribute__((aligned(32))), vs[8] __attribute__((aligned(32)));
  _mm256_store_ps(us, u);
  _mm256_store_ps(vs, v);

  isize s  = lane + offset;
  f32   w1 = us[lane], w2 = vs[lane];
  f32   w0 = 1 - w1 - w2;
  Triangle_AOS const *rec = &tris->aos[s];

  hit->distance = nearest;
  hit->slot     = (i32)s;
  hit->point    = v3_add(ray->position, v3_scale(ray->direction, nearest));
  hit->normal   = v3(rec->normal_a.x * w0 + rec->normal_b.x * w1 + rec->normal_c.x * w2,
                     rec->normal_a.y * w0 + rec->normal_b.y * w1 + rec->normal_c.y * w2,
                     rec->normal_a.z * w0 + rec->normal_b.z * w1 + rec->normal_c.z * w2);
  hit->tex_coords.x = rec->tex_coords_a.x * w0 + rec->tex_coords_b.x * w1 + rec->tex_coords_c.x * w2;
  hit->tex_coords.y = rec->tex_coords_a.y * w0 + rec->tex_coords_b.y * w1 + rec->tex_coords_c.y * w2;
  hit->shader     = rec->shader;
  hit->normal_geo = rec->normal;
  hit->tangent    = rec->tangent;
  hit->bitangent  = rec->bitangent;
  return true;
}

/* ---- raytracer.c:190-230.  _mm256_min_ps/max_ps return the SECOND operand
 * when either is NaN; the nesting below keeps that order. ---- */
static inline void child_box_test(Ray const *ray, f32 t_min, f32 t_max, BVH_Node const *node, f32 *out8) {
  __m256 ix = _mm256_set1_ps(1.0f / ray->direction.x);
  __m256 iy = _mm256_set1_ps(1.0f / ray->direction.y);
  __m256 iz = _mm256_set1_ps(1.0f / ray->direction.z);
  __m256 ox = _mm256_set1_ps(ray->position.x);
  __m256 oy = _mm256_set1_ps(ray->position.y);
  __m256 oz = _mm256_set1_ps(ray->position.z);

  __m256 ax = MUL(SUB(_mm256_load_ps(node->mins[0]), ox), ix);
  __m256 ay = MUL(SUB(_mm256_load_ps(node->mins[1]), oy), iy);
  __m256 az = MUL(SUB(_mm256_load_ps(node->mins[2]), oz), iz);
  __m256 bx = MUL(SUB(_mm256_load_ps(node->maxs[0]), ox), ix);
  __m256 by = MUL(SUB(_mm256_load_ps(node->maxs[1]), oy), iy);
  __m256 bz = MUL(SUB(_mm256_load_ps(node->maxs[2]), oz), iz);

  __m256 nx = _mm256_min_ps(ax, bx), ny = _mm256_min_ps(ay, by), nz = _mm256_min_ps(az, bz);
  __m256 fx = _mm256_max_ps(ax, bx), fy = _mm256_max_ps(ay, by), fz = _mm256_max_ps(az, bz);

  __m256 enter = _mm256_max_ps(_mm256_set1_ps(t_min), _mm256_max_ps(nx, _mm256_max_ps(ny, nz)));
  __m256 leave = _mm256_min_ps(_mm256_set1_ps(t_max), _mm256_min_ps(fx, _mm256_min_ps(fy, fz)));

  __m256 miss = _mm256_cmp_ps(enter, leave, _CMP_GE_OQ);
  _mm256_store_ps(out8, _mm256_blendv_ps(enter, _mm256_set1_ps(INFINITY), miss));
}

/* ---- raytracer.c:443-483 ---- */
static void visit_node(Ray const *ray, Scene const *scene, isize index, Hit *hit, isize depth) {
  f32 entry[8] __attribute__((aligned(32)));
  child_box_test(ray, RT_EPSILON, hit->distance, &scene->bvh.nodes.data[index], entry);
  if (tl_counters) tl_counters->c[ORACLE_CTR_NODES]++;

  for (int round = 0; round < RT_SIMD_WIDTH; round++) {
    f32 best = hit->distance;
    int pick = -1;
    for (int j = 0; j < RT_SIMD_WIDTH; j++) {
      if (entry[j] < best) { best = entry[j]; pick = j; }
    }
    if (pick < 0 || best >= hit->distance) return;

    isize child = RT_SIMD_WIDTH * index + 1 + pick;
    if (depth == 1) {
      if (tl_counters) tl_counters->c[ORACLE_CTR_LEAVES]++;
      bool took = leaf_test(ray, &scene->triangles, (child - scene->bvh.last_row_offset) * RT_SIMD_WIDTH, hit);
      if (took && tl_counters) tl_counters->c[ORACLE_CTR_ACCEPTS]++;
    } else {
      visit_node(ray, scene, child, hit, depth - 1);
    }
    entry[pick] = INFINITY;
  }
}

static inline void closest_hit(Ray const *ray, Scene const *scene, Hit *hit) {
  if (tl_counters) tl_counters->c[ORACLE_CTR_RAYS]++;
  visit_node(ray, scene, 0, hit, scene->bvh.depth);
}

i32 oracle_trace_ray(Scene const *scene, Ray ray, f32 *t_out) {
  Hit hit;
  memset(&hit, 0, sizeof hit);
  hit.distance = INFINITY;
  hit.slot = -1;
  closest_hit(&ray, scene, &hit);
  if (t_out) *t_out = hit.distance;
  return hit.distance != INFINITY ? hit.slot : -1;
}

/* ---- raytracer.c:505-558 ---- */
static Color3 trace_path(Scene const *scene, Ray ray, isize max_bounces, i32 *primary_slot) {
  Color3 tint     = v3(1, 1, 1);
  Color3 emission = v3(0, 0, 0);
  if (primary_slot) *primary_slot = -1;

  for (isize bounce = 0; bounce < max_bounces; bounce++) {
    Hit hit;
    memset(&hit, 0, sizeof hit);
    hit.distance = INFINITY;
    hit.slot = -1;
    closest_hit(&ray, scene, &hit);
    if (bounce == 0 && primary_slot && hit.distance != INFINITY) *primary_slot = hit.slot;

    if (hit.distance == INFINITY) {
      if (tl_counters) tl_counters->c[ORACLE_CTR_MISSES]++;
      Color3 sky = scene->background.proc(scene->background.data, ray.direction);
      return v3_add(v3_mul(sky, tint), emission);
    }

    if (v3_dot(hit.normal_geo, ray.direction) > 0 || v3_dot(hit.normal, ray.direction) > 0) {
      if (tl_counters) tl_counters->c[ORACLE_CTR_PASSTHROUGH]++;
      ray.position = v3_add(hit.point, v3_scale(ray.direction, RT_EPSILON));
      continue;
    }

    Shader_Input in;
    memset(&in, 0, sizeof in);
    in.direction  = ray.direction;
    in.normal     = v3_normalize(hit.normal);
    in.normal_geo = hit.normal_geo;
    in.tangent    = hit.tangent;
    in.bitangent  = hit.bitangent;
    in.position   = hit.point;
    in.tex_coords = hit.tex_coords;
    Shader_Output out;
    memset(&out, 0, sizeof out);

    if (tl_counters) tl_counters->c[ORACLE_CTR_SHADES]++;
    hit.shader.proc(hit.shader.data, &in, &out);

    emission = v3_add(emission, v3_mul(out.emission, tint));
    if (out.terminate) break;

    ray.direction = out.direction;
    tint          = v3_mul(tint, out.tint);

    f32 bias = (0.5f - (f32)(v3_dot(hit.normal_geo, out.direction) < 0)) * 2.0f * RT_EPSILON;
    ray.position = v3_add(hit.point, v3_scale(hit.normal_geo, bias));
  }
  return emission;
}

Color3 oracle_cast_ray(Scene const *scene, Ray ray, isize max_bounces) {
  i32 slot;
  return trace_path(scene, ray, max_bounces, &slot);
}

/* ---- raytracer.c:582-594, one lane of hash12x8 ---- */
static inline f32 fractf(f32 v) { return v - floorf(v); }
f32 oracle_hash12(f32 px, f32 py) {
  f32 a = fractf(px * 0.1031f);
  f32 b = fractf(py * 0.1031f);
  f32 c = fractf(px * 0.1031f);
  f32 k = 33.33f;
  f32 d = a * (b + k) + b * (c + k) + c * (a + k);
  return fractf((a + b + d * 2.0f) * (c + d));
}

/* common.h:90-92 with pow from rt_math.h */
static inline f32 encode_srgb(f32 c) {
  return (c <= 0.0031308f) ? (12.92f * c) : (1.055f * rt_powf(c, 1.0f / 2.4f) - 0.055f);
}

/* raytracer.c:700-716 */
static inline void resolve_pixel(Color3 sum, f32 inv_samples, u8 *dst) {
  for (int c = 0; c < 3; c++) {
    f32 v = sum.data[c] * inv_samples;
    if (v != v) v = 0;                       /* NaN: defined here, UB in the reference */
    v = f32_clamp(v, 0, 1);
    v = encode_srgb(v);
    v = v * 255.999f;
    dst[c] = (u8)v;
  }
}

void oracle_resolve(f32 const *accum, isize n_pixels, isize samples, u8 *rgb_out) {
  f32 inv_samples = 1.0f / (f32)samples;
  for (isize i = 0; i < n_pixels; i++) {
    resolve_pixel(v3(accum[3 * i], accum[3 * i + 1], accum[3 * i + 2]), inv_samples, rgb_out + 3 * i);
  }
}

typedef struct {
  Rendering_Context *ctx;
  Oracle_Options    *opt;
  Counters           counters;
  pthread_t          thread;
} Worker;

/* ---- raytracer.c:596-720 ---- */
static void *render_worker(void *arg) {
  Worker            *w   = arg;
  Rendering_Context *ctx = w->ctx;
  Oracle_Options    *opt = w->opt;
  tl_counters = &w->counters;
  u32 *rng = oracle_shader_random_state();
  *rng = 0;                                          /* driver.c's copy starts at 0 */

  Scene const *scene = ctx->scene;
  Camera cam = scene->camera;
  isize width = ctx->image.width, height = ctx->image.height;
  isize samples = ctx->samples, max_bounces = ctx->max_bounces;
  isize s_begin = opt->sample_begin, s_end = opt->sample_end ? opt->sample_end : samples;
  isize chunks_x = (width  + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  isize chunks_y = (height + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  isize n_chunks = chunks_x * chunks_y;

  /* view_matrix * (0,0,0,1) */
  Vec3 eye = v3(cam.view_matrix.rows[0][3], cam.view_matrix.rows[1][3], cam.view_matrix.rows[2][3]);

  f32 inv_samples = 1.0f / (f32)samples;
  f32 inv_width   = 1.0f / (f32)width;
  f32 inv_height  = 1.0f / (f32)height;
  f32 aspect      = (f32)width / (f32)height;
  f32 (*m)[4] = cam.view_matrix.rows;

  for (;;) {
    isize c = atomic_fetch_add(&ctx->_current_chunk, 1);
    if (c >= n_chunks) break;
    isize x0 = (c % chunks_x) * RT_CHUNK_SIZE, y0 = (c / chunks_x) * RT_CHUNK_SIZE;

    for (isize y = y0; y < y0 + RT_CHUNK_SIZE && y < height; y++) {
      for (isize x = x0; x < x0 + RT_CHUNK_SIZE && x < width; x++) {
        Color3 sum = v3(0, 0, 0);
        isize  pixel = x + y * width;
        if (opt->pixel_mask && !opt->pixel_mask[pixel]) continue;

        for (isize batch = 0; batch < (samples + 7) / 8; batch++) {
          if (batch * 8 + 8 <= s_begin || batch * 8 >= s_end) continue;
          __m256 lane_s = _mm256_add_ps(_mm256_setr_ps(0, 1, 2, 3, 4, 5, 6, 7), _mm256_set1_ps((f32)(batch * 8)));
          f32 sidx[8] __attribute__((aligned(32)));
          _mm256_store_ps(sidx, lane_s);
          /* rand_a and rand_b are the same expression (raytracer.c:644-651) */
          f32 jit[8] __attribute__((aligned(32)));
          for (int l = 0; l < 8; l++) jit[l] = oracle_hash12((f32)x * 50.0f + sidx[l], (f32)y);
          __m256 j8 = _mm256_load_ps(jit);

          /* ((f32)x + rand - 0.5) is evaluated in double in the reference only
           * if vectors promoted, which GCC vector types do not: 0.5 converts to
           * f32 (raytracer.c:654-655). */
          __m256 ux = SUB(MUL(MUL(SUB(ADD(_mm256_set1_ps((f32)x), j8), _mm256_set1_ps(0.5f)), _mm256_set1_ps(2.0f)), _mm256_set1_ps(inv_width)),  _mm256_set1_ps(1.0f));
          __m256 uy = SUB(MUL(MUL(SUB(ADD(_mm256_set1_ps((f32)y), j8), _mm256_set1_ps(0.5f)), _mm256_set1_ps(2.0f)), _mm256_set1_ps(inv_height)), _mm256_set1_ps(1.0f));

          __m256 cx = MUL(ux, _mm256_set1_ps(aspect));
          __m256 cy = _mm256_xor_ps(uy, _mm256_set1_ps(-0.0f));   /* unary minus */
          __m256 cz = _mm256_set1_ps(-cam.focal_length);

          __m256 len2 = ADD(ADD(MUL(cx, cx), MUL(cy, cy)), MUL(cz, cz));
          __m256 inv_len = opt->approx_rsqrt ? _mm256_rsqrt_ps(len2)
                                             : _mm256_div_ps(_mm256_set1_ps(1.0f), _mm256_sqrt_ps(len2));

          __m256 wx = ADD(ADD(MUL(_mm256_set1_ps(m[0][0]), cx), MUL(_mm256_set1_ps(m[0][1]), cy)), MUL(_mm256_set1_ps(m[0][2]), cz));
          __m256 wy = ADD(ADD(MUL(_mm256_set1_ps(m[1][0]), cx), MUL(_mm256_set1_ps(m[1][1]), cy)), MUL(_mm256_set1_ps(m[1][2]), cz));
          __m256 wz = ADD(ADD(MUL(_mm256_set1_ps(m[2][0]), cx), MUL(_mm256_set1_ps(m[2][1]), cy)), MUL(_mm256_set1_ps(m[2][2]), cz));
          f32 dx[8] __attribute__((aligned(32))), dy[8] __attribute__((aligned(32))), dz[8] __attribute__((aligned(32)));
          _mm256_store_ps(dx, MUL(wx, inv_len));
          _mm256_store_ps(dy, MUL(wy, inv_len));
          _mm256_store_ps(dz, MUL(wz, inv_len));

          for (int l = 0; l < 8; l++) {
            isize s = batch * 8 + l;
            if (s >= samples) break;
            if (s < s_begin || s >= s_end) continue;
            if (opt->seed_mode == ORACLE_SEED_PER_SAMPLE) *rng = rt_path_seed((u32)pixel, (u32)s, opt->user_seed);
            Ray r;
            r.position  = eye;
            r.direction = v3(dx[l], dy[l], dz[l]);
            i32 slot;
            Color3 radiance = trace_path(scene, r, max_bounces, &slot);
            sum = v3_add(sum, radiance);
            w->counters.c[ORACLE_CTR_SAMPLES]++;
            if (opt->hit_ids && s == s_begin) opt->hit_ids[pixel] = slot;
            if (opt->per_sample) {
              f32 *dst = opt->per_sample + ((size_t)pixel * (size_t)(s_end - s_begin) + (size_t)(s - s_begin)) * 3;
              dst[0] = radiance.x; dst[1] = radiance.y; dst[2] = radiance.z;
            }
          }
        }

        if (opt->accum) {
          opt->accum[3 * pixel + 0] = sum.x;
          opt->accum[3 * pixel + 1] = sum.y;
          opt->accum[3 * pixel + 2] = sum.z;
        }
        if (ctx->image.pixels.data) {
          resolve_pixel(sum, inv_samples, ctx->image.pixels.data + ctx->image.components * (x + y * ctx->image.stride));
        }
      }
    }
  }
  tl_counters = NULL;
  atomic_fetch_add(&ctx->n_threads, -1);
  return NULL;
}

void oracle_render(Rendering_Context *ctx, Oracle_Options *opt, i32 n_threads) {
  if (n_threads < 1) n_threads = 1;
  Worker *workers = calloc((size_t)n_threads, sizeof *workers);
  atomic_store(&ctx->n_threads, n_threads);
  atomic_store(&ctx->_current_chunk, 0);
  for (i32 i = 0; i < n_threads; i++) {
    workers[i].ctx = ctx;
    workers[i].opt = opt;
  }
  for (i32 i = 1; i < n_threads; i++) pthread_create(&workers[i].thread, NULL, render_worker, &workers[i]);
  render_worker(&workers[0]);
  for (i32 i = 1; i < n_threads; i++) pthread_join(workers[i].thread, NULL);
  memset(opt->counters, 0, sizeof opt->counters);
  for (i32 i = 0; i < n_threads; i++)
    for (int k = 0; k < 8; k++) opt->counters[k] += workers[i].counters.c[k];
  free(workers);
}
