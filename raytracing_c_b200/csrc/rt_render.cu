// rt_render.cu — the path-tracing kernels for sm_100a.
//
// Replaces render_thread_proc's pixel loop and everything under it (reference
// raytracer.c:443-558 traversal + cast_ray, :582-594 jitter hash, :596-720 chunk
// loop, camera, film) — see DESIGN.md for the mapping.
//
// Execution model (one persistent kernel per sample slice):
//   * Every lane owns one PATH and one PIXEL JOB.  A job is "all samples
//     [sample_begin, sample_end) of one pixel", pulled from a global atomic
//     counter in 8x4-pixel tile order (the GPU form of the reference's atomic
//     32x32 chunk queue, raytracer.c:619-627), so the 32 lanes of a warp always
//     work on one 8x4 tile and their primary rays stay coherent.  The lane sums
//     its samples in sample order in registers and writes the pixel once, so the
//     f32 sum is bit-identical to the sequential CPU loop.  When a path ends the
//     lane starts the next sample (or pulls the next pixel) — lanes never idle
//     on finished paths.
//   * TRAVERSAL is one thread per ray (trace_ray): the reference's AVX2 8-wide box
//     and triangle tests become eight unrolled scalar tests per lane, so a warp
//     tests 256 boxes per node step with no cross-lane traffic.  The walk is a flat
//     while-while loop (no recursion): internal nodes until a leaf is found, then
//     the warp runs the Möller–Trumbore body together.  Entry distances of levels
//     with untried children are parked in shared memory ([level][2][thread]
//     float4, conflict-free), a bitmask of such levels lets a pop jump straight
//     to the next useful ancestor.  Node and leaf rows are read as 16-byte
//     vectors (12 per node, 18 per leaf), warp-uniform for coherent rays.
//   * SHADING is per lane (rt_shade.cuh), all 32 lanes of the warp active.
// Arithmetic is IEEE f32 without FMA contraction (-fmad=false) in the reference's
// operation order: primary-hit slots and radiance equal the CPU oracle bit for bit.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "rt_device.cuh"
#include "rt_shade.cuh"
#include "rt_kernels.h"

#define RT_BLOCK 256
#ifndef RT_MIN_BLOCKS
#define RT_MIN_BLOCKS 3
#endif

// raytracer.c:582-594, one lane of hash12x8
__device__ __forceinline__ float fract1(float v) { return v - floorf(v); }
__device__ __forceinline__ float hash12(float px, float py) {
  float a = fract1(px * 0.1031f);
  float b = fract1(py * 0.1031f);
  float c = fract1(px * 0.1031f);
  float k = 33.33f;
  float d = a * (b + k) + b * (c + k) + c * (a + k);
  return fract1((a + b + d * 2.0f) * (c + d));
}

// raytracer.c:190-230 for ONE child box.
// `regular` = all three reciprocal direction components are finite.  Then no product
// below can be NaN (finite * finite), so MINPS/MAXPS' "second operand when unordered"
// rule never fires and the hardware FMNMX gives the same value (a zero's sign can
// differ, but `enter` is >= EPS and `leave` is only compared).  Otherwise the exact
// operand-order selects are used (0 * inf lanes, raytracer.c:212-225).
template <bool REGULAR>
__device__ __forceinline__ float child_entry(float lox, float loy, float loz, float hix, float hiy, float hiz,
                                             float ox, float oy, float oz, float ix, float iy, float iz, float t_max) {
  float ax = (lox - ox) * ix;
  float ay = (loy - oy) * iy;
  float az = (loz - oz) * iz;
  float bx = (hix - ox) * ix;
  float by = (hiy - oy) * iy;
  float bz = (hiz - oz) * iz;
  float enter, leave;
  if (REGULAR) {
    enter = fmaxf(RT_EPS, fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz))));
    leave = fminf(t_max,  fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz))));
  } else {
    float nx = sel_min(ax, bx), ny = sel_min(ay, by), nz = sel_min(az, bz);
    float fx = sel_max(ax, bx), fy = sel_max(ay, by), fz = sel_max(az, bz);
    enter = sel_max(RT_EPS, sel_max(nx, sel_max(ny, nz)));
    leave = sel_min(t_max,  sel_min(fx, sel_min(fy, fz)));
  }
  return (enter >= leave) ? CUDART_INF_F : enter;
}

// raytracer.c:190-230, ray_aabbs_hit_8: the eight children of one node, entry distance or +inf.
// A node is six 32-byte rows (min x/y/z, max x/y/z; child j in column j): twelve 16-byte loads,
// warp-uniform for coherent rays.
template <bool REGULAR>
__device__ __forceinline__ void node_entries(const float4 *__restrict__ n4, float ox, float oy, float oz,
                                             float ix, float iy, float iz, float t_max, float (&e)[8]) {
  #pragma unroll
  for (int h = 0; h < 2; h++) {
    float4 lx = __ldg(n4 + 0 + h), ly = __ldg(n4 + 2 + h), lz = __ldg(n4 + 4 + h);
    float4 hx = __ldg(n4 + 6 + h), hy = __ldg(n4 + 8 + h), hz = __ldg(n4 + 10 + h);
    e[4 * h + 0] = child_entry<REGULAR>(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, ox, oy, oz, ix, iy, iz, t_max);
    e[4 * h + 1] = child_entry<REGULAR>(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, ox, oy, oz, ix, iy, iz, t_max);
    e[4 * h + 2] = child_entry<REGULAR>(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, ox, oy, oz, ix, iy, iz, t_max);
    e[4 * h + 3] = child_entry<REGULAR>(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, ox, oy, oz, ix, iy, iz, t_max);
  }
}

// fminf ignores NaN operands: the minimum of the ordered entries (NaN only if all eight are NaN)
__device__ __forceinline__ float min8(const float (&e)[8]) {
  return fminf(fminf(fminf(e[0], e[1]), fminf(e[2], e[3])), fminf(fminf(e[4], e[5]), fminf(e[6], e[7])));
}

// per-thread slice of the level store: levels[(level - 1) * 2 + half][tid]
#define RT_LEVELS(level, half) levels[(((level) - 1) * 2 + (half)) * RT_BLOCK]

// raytracer.c:443-503: closest hit of one ray, one thread per ray; a WARP-COLLECTIVE call
// (all 32 lanes enter; lanes without a ray pass active = false).
// The reference recursion (8 entry distances per level on the C stack, up to 8 selection
// rounds per node) is a flat loop here: the tree is a complete 8-ary heap, parent = (n-1)>>3;
// the current node's entry distances live in registers, those of ancestors that still hold
// untried candidates in shared memory, and `pending` (bit = level) lets a pop jump straight
// to the nearest such ancestor.  Visit order and every compare are the reference's, so the
// closest hit — ties included — is the same triangle slot.
// Shape: while-while with explicit reconvergence — lanes walk internal nodes until each
// holds a leaf (or is finished), the warp syncs, then runs the triangle loop together.
__device__ __forceinline__ void trace_ray(const SceneDev &sc, float4 *levels, bool active, float ox, float oy, float oz,
                                          float dx, float dy, float dz,
                                          float &hit_t, float &hit_u, float &hit_v, int &hit_slot,
                                          unsigned &c_nodes, unsigned &c_leaves, unsigned &c_accepts) {
  const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;         // raytracer.c:198-202
  // a zero direction component makes 1/d infinite and 0 * inf NaN: only then the slab test
  // needs the exact MINPS/MAXPS operand-order rule (see child_entry)
  const bool regular = (fabsf(ix) < CUDART_INF_F) & (fabsf(iy) < CUDART_INF_F) & (fabsf(iz) < CUDART_INF_F);
  int      node = 0, level = sc.depth;                                 // raytracer.c:501
  unsigned pending = 0;
  bool     need_box = true, done = !active;
  int      leaf = -1;
  float    e[8];
  hit_t = CUDART_INF_F; hit_u = 0; hit_v = 0; hit_slot = -1;

  for (;;) {
    while (!done && leaf < 0) {
      if (need_box) {
        const float4 *n4 = (const float4 *)(sc.nodes + (size_t)node * 48);
        if (regular) node_entries<true >(n4, ox, oy, oz, ix, iy, iz, hit_t, e);
        else         node_entries<false>(n4, ox, oy, oz, ix, iy, iz, hit_t, e);
        need_box = false;
        c_nodes++;
      }
      // raytracer.c:459-472: nearest untried child strictly below the current hit, lowest index on ties
      const float best = min8(e);
      if (!(best < hit_t)) {
        if (pending == 0) { done = true; break; }
        const int up = __ffs(pending) - 1;
        pending &= pending - 1;
        for (; level < up; level++) node = (node - 1) >> 3;
        float4 a = RT_LEVELS(level, 0), b = RT_LEVELS(level, 1);
        e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
        continue;
      }
      int pick = 7;
      #pragma unroll
      for (int j = 6; j >= 0; j--) pick = (e[j] == best) ? j : pick;
      #pragma unroll
      for (int j = 0; j < 8; j++) e[j] = (j == pick) ? CUDART_INF_F : e[j];       // raytracer.c:481
      const int child = 8 * node + 1 + pick;
      if (level == 1) { leaf = child - sc.n_internal; break; }
      if (min8(e) < hit_t) {                         // other candidates remain: remember this level
        RT_LEVELS(level, 0) = make_float4(e[0], e[1], e[2], e[3]);
        RT_LEVELS(level, 1) = make_float4(e[4], e[5], e[6], e[7]);
        pending |= 1u << level;
      }
      node = child;
      level -= 1;
      need_box = true;
    }
    __syncwarp();
    if (__all_sync(0xffffffffu, done)) return;

    // ---- raytracer.c:84-188: the eight triangles of the leaf, three 16-byte loads each:
    // p0 and the edges e1 = p1 - p0, e2 = p2 - p0 (the same f32 subtractions raytracer.c:116-122
    // does per ray, done once at upload).  Strict <, ascending j: the lowest lane wins a tie
    // inside the leaf and an earlier leaf wins across leaves (raytracer.c:15-32 with eps 0, :159).
    if (leaf >= 0) {
      const float4 *tp = sc.tri_pos + (size_t)leaf * 24;
      const float t_before = hit_t;
      c_leaves++;
      #pragma unroll 1
      for (int j = 0; j < 8; j++) {
        const float4 A = __ldg(tp + 3 * j), B = __ldg(tp + 3 * j + 1), C = __ldg(tp + 3 * j + 2);
        const float e1x = A.w, e1y = B.x, e1z = B.y, e2x = B.z, e2y = B.w, e2z = C.x;
        float pvx = dy * e2z - dz * e2y, pvy = dz * e2x - dx * e2z, pvz = dx * e2y - dy * e2x;
        float det = e1x * pvx + e1y * pvy + e1z * pvz;
        float inv_det = 1.0f / det;
        float tvx = ox - A.x, tvy = oy - A.y, tvz = oz - A.z;
        float u = inv_det * (tvx * pvx + tvy * pvy + tvz * pvz);
        // the reject mask is an OR (raytracer.c:137-152): a triangle that fails on u fails whatever v and t are
        if ((u < -RT_EPS) | (u > 1 + RT_EPS)) continue;
        float qvx = tvy * e1z - tvz * e1y, qvy = tvz * e1x - tvx * e1z, qvz = tvx * e1y - tvy * e1x;
        float v = inv_det * (dx * qvx + dy * qvy + dz * qvz);
        float t = inv_det * (e2x * qvx + e2y * qvy + e2z * qvz);
        bool miss = (v < -RT_EPS) | (u + v > 1 + RT_EPS) | (t < RT_EPS);
        // t <= 0 and NaN count as +inf (min_f32x8 with eps 0); NaN also fails the ordered compare
        if (!miss && t > 0.0f && t < hit_t) { hit_t = t; hit_u = u; hit_v = v; hit_slot = leaf * 8 + j; }
      }
      if (hit_t < t_before) c_accepts++;
      leaf = -1;
    }
    __syncwarp();
  }
}

struct Shared {
  float texel_lut[256];
};

__global__ void __launch_bounds__(RT_BLOCK, RT_MIN_BLOCKS)
rt_render_kernel(const __grid_constant__ RenderParams P) {
  __shared__ Shared sh;
  extern __shared__ float4 level_store[];          // [depth][2][RT_BLOCK] entry distances of pending levels

  const SceneDev &sc = P.scene;
  const int tid   = threadIdx.x;
  const int lane  = tid & 31;
  float4 *levels = level_store + tid;

  // u8 -> f32 texel table: the same IEEE division the reference does per tap
  // (driver.c:69-88), done once per block instead of 12 times per bilinear fetch
  sh.texel_lut[tid] = (float)tid / 255.999f;
  __syncthreads();

  const int   W = P.width, H = P.height;
  const int   tiles_x = (W + 7) >> 3, tiles_y = (H + 3) >> 2;
  const unsigned total_jobs = (unsigned)(tiles_x * tiles_y) * 32u;
  const float inv_w = 1.0f / (float)W, inv_h = 1.0f / (float)H;
  const float aspect = (float)W / (float)H;
  const V3    eye = mk3(sc.view[0][3], sc.view[1][3], sc.view[2][3]);
  const int   n_samples = P.sample_end - P.sample_begin;

  // job state
  int  pixel = -1, px = 0, py = 0, s = 0;
  V3   sum = mk3(0, 0, 0);
  bool alive = n_samples > 0;
  // path state
  bool has_path = false;
  V3   o = eye, d = mk3(0, 0, -1), tint = mk3(1, 1, 1), emis = mk3(0, 0, 0);
  int  bounce = 0;
  uint32_t rng = 0;
  unsigned c_rays = 0, c_nodes = 0, c_leaves = 0, c_accepts = 0, c_shades = 0, c_misses = 0, c_pass = 0, c_samples = 0;

  for (;;) {
    // ---------------------------------------------------------------- regenerate
    if (!has_path && alive) {
      if (pixel >= 0 && s >= P.sample_end) {
        P.accum[3 * pixel + 0] = sum.x;
        P.accum[3 * pixel + 1] = sum.y;
        P.accum[3 * pixel + 2] = sum.z;
        pixel = -1;
      }
      while (pixel < 0) {
        unsigned job = atomicAdd(P.job_counter, 1u);
        if (job >= total_jobs) { alive = false; break; }
        unsigned tile = job >> 5, in_tile = job & 31u;
        px = (int)(tile % (unsigned)tiles_x) * 8 + (int)(in_tile & 7u);
        py = (int)(tile / (unsigned)tiles_x) * 4 + (int)(in_tile >> 3);
        if (px < W && py < H) {
          pixel = py * W + px;
          s = P.sample_begin;
          sum = P.accumulate ? mk3(P.accum[3 * pixel], P.accum[3 * pixel + 1], P.accum[3 * pixel + 2]) : mk3(0, 0, 0);
        }
      }
      if (pixel >= 0) {
        // raytracer.c:644-677; rand_a == rand_b; exact 1/sqrt instead of rsqrt_ps
        float jit = hash12((float)px * 50.0f + (float)s, (float)py);
        float ux = ((float)px + jit - 0.5f) * 2.0f * inv_w - 1.0f;
        float uy = ((float)py + jit - 0.5f) * 2.0f * inv_h - 1.0f;
        float cx = ux * aspect, cy = -uy, cz = -sc.focal_length;
        float inv_len = 1.0f / __fsqrt_rn(cx * cx + cy * cy + cz * cz);
        d = mk3((sc.view[0][0] * cx + sc.view[0][1] * cy + sc.view[0][2] * cz) * inv_len,
                (sc.view[1][0] * cx + sc.view[1][1] * cy + sc.view[1][2] * cz) * inv_len,
                (sc.view[2][0] * cx + sc.view[2][1] * cy + sc.view[2][2] * cz) * inv_len);
        o = eye;
        tint = mk3(1, 1, 1);
        emis = mk3(0, 0, 0);
        bounce = 0;
        rng = rt_path_seed((uint32_t)pixel, (uint32_t)s, P.user_seed);
        has_path = true;
      }
    }
    if (__ballot_sync(0xffffffffu, has_path) == 0) break;

    // --------------------------------------------------------------------- trace
    float hit_t = CUDART_INF_F, hit_u = 0, hit_v = 0;
    int   slot = -1;
    if (has_path) c_rays++;
    trace_ray(sc, levels, has_path, o.x, o.y, o.z, d.x, d.y, d.z, hit_t, hit_u, hit_v, slot, c_nodes, c_leaves, c_accepts);

    // --------------------------------------------------------------------- shade
    if (has_path) {
      bool done = false;
      V3 radiance = mk3(0, 0, 0);

      if (bounce == 0 && s == P.sample_begin && P.hit_ids) P.hit_ids[pixel] = slot;

      if (slot < 0) {
        // raytracer.c:554
        c_misses++;
        radiance = add3(mul3(environment(sc, sh.texel_lut, d), tint), emis);
        done = true;
      } else {
        const float4 *rec = sc.tri_rec + (size_t)slot * 7;
        float4 r0 = __ldg(rec + 0), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
        V3 ng = mk3(r0.x, r0.y, r0.z);
        V3 na = mk3(r0.w, r1.x, r1.y), nb = mk3(r1.z, r1.w, r2.x), nc = mk3(r2.y, r2.z, r2.w);
        float w1 = hit_u, w2 = hit_v, w0 = 1 - w1 - w2;           // raytracer.c:164-177
        V3 point  = add3(o, scale3(d, hit_t));
        V3 normal = mk3(na.x * w0 + nb.x * w1 + nc.x * w2,
                        na.y * w0 + nb.y * w1 + nc.y * w2,
                        na.z * w0 + nb.z * w1 + nc.z * w2);
        if (dot3(ng, d) > 0 || dot3(normal, d) > 0) {
          // raytracer.c:516-522: back face — step through, the bounce is consumed
          c_pass++;
          o = add3(point, scale3(d, RT_EPS));
        } else {
          float4 r3 = __ldg(rec + 3), r4 = __ldg(rec + 4), r5 = __ldg(rec + 5), r6 = __ldg(rec + 6);
          ShadeIn in;
          in.dir = d;
          in.normal = normalize3(normal);
          in.normal_geo = ng;
          in.tangent   = mk3(r3.x, r3.y, r3.z);
          in.bitangent = mk3(r3.w, r4.x, r4.y);
          in.u = r4.z * w0 + r5.x * w1 + r5.z * w2;
          in.v = r4.w * w0 + r5.y * w1 + r5.w * w2;
          ShadeOut out;
          c_shades++;
          shade_pbr(sc, sh.texel_lut, __float_as_int(r6.x), in, rng, out);
          emis = add3(emis, mul3(out.emission, tint));          // raytracer.c:537
          if (out.terminate) {
            radiance = emis;
            done = true;
          } else {
            d = out.dir;
            tint = mul3(tint, out.tint);
            float bias = (0.5f - (float)(dot3(ng, out.dir) < 0)) * 2.0f * RT_EPS;   // raytracer.c:551
            o = add3(point, scale3(ng, bias));
          }
        }
        bounce++;
        if (!done && bounce >= P.max_bounces) { radiance = emis; done = true; }      // raytracer.c:557
      }

      if (done) {
        sum = add3(sum, radiance);
        if (P.per_sample) {
          float *dst = P.per_sample + ((size_t)pixel * (size_t)n_samples + (size_t)(s - P.sample_begin)) * 3;
          dst[0] = radiance.x; dst[1] = radiance.y; dst[2] = radiance.z;
        }
        c_samples++;
        s++;
        has_path = false;
      }
    }
  }

  if (P.counters) {
    unsigned vals[8] = { c_rays, c_nodes, c_leaves, c_accepts, c_shades, c_misses, c_pass, c_samples };
    #pragma unroll
    for (int k = 0; k < 8; k++) {
      unsigned v = vals[k];
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (lane == 0 && v) atomicAdd(&P.counters[k], (unsigned long long)v);
    }
  }
}

// ----------------------------------------------------------------------- film
// raytracer.c:700-716 + common.h:90-92: /spp, clamp, sRGB OETF, *255.999, truncate.
__global__ void rt_resolve_kernel(const float *__restrict__ accum, int width, int height, float inv_samples,
                                  unsigned char *__restrict__ pixels, int stride, int components) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width * height) return;
  int x = i % width, y = i / width;
  unsigned char *dst = pixels + (size_t)components * (size_t)(x + y * stride);
  #pragma unroll
  for (int c = 0; c < 3; c++) {
    float v = accum[3 * i + c] * inv_samples;
    if (v != v) v = 0;
    v = clamp1(v, 0, 1);
    v = (v <= 0.0031308f) ? (12.92f * v) : (1.055f * rt_powf(v, 1.0f / 2.4f) - 0.055f);
    v = v * 255.999f;
    dst[c] = (unsigned char)v;
  }
}

// ------------------------------------------------------------------- launchers
static int g_blocks_per_sm = 0;
static size_t g_level_bytes = 0;

int rt_launch_render(const RenderParams &p, int sm_count, cudaStream_t stream) {
  // level store: [depth][2][RT_BLOCK] float4 of dynamic shared memory (see trace_ray)
  const size_t level_bytes = (size_t)(p.scene.depth > 0 ? p.scene.depth : 1) * 2 * RT_BLOCK * sizeof(float4);
  if (g_blocks_per_sm == 0 || level_bytes != g_level_bytes) {
    int n = 0;
    cudaFuncSetAttribute(rt_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_MAX_DEPTH * 2 * RT_BLOCK * 16);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rt_render_kernel, RT_BLOCK, level_bytes) != cudaSuccess || n < 1) n = 1;
    g_blocks_per_sm = n;
    g_level_bytes = level_bytes;
  }
  long long jobs = (long long)((p.width + 7) / 8) * ((p.height + 3) / 4) * 32;
  long long want = (jobs + RT_BLOCK - 1) / RT_BLOCK;
  long long grid = (long long)sm_count * g_blocks_per_sm;       // persistent: one wave
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  cudaMemsetAsync(p.job_counter, 0, sizeof(unsigned int), stream);
  rt_render_kernel<<<(unsigned)grid, RT_BLOCK, level_bytes, stream>>>(p);
  return (int)cudaGetLastError();
}

int rt_launch_resolve(const float *accum, int width, int height, int samples, unsigned char *pixels,
                      int stride, int components, cudaStream_t stream) {
  int n = width * height;
  float inv_samples = 1.0f / (float)samples;
  rt_resolve_kernel<<<(n + 255) / 256, 256, 0, stream>>>(accum, width, height, inv_samples, pixels, stride, components);
  return (int)cudaGetLastError();
}

int rt_render_blocks_per_sm(void) { return g_blocks_per_sm; }
