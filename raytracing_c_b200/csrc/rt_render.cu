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
//     32x32 chunk queue, raytracer.c:619-627).  The lane sums its samples in
//     sample order in registers and writes the pixel once, so the f32 sum is
//     bit-identical to the sequential CPU loop.  When a path ends the lane starts
//     the next sample (or pulls the next pixel) — lanes never idle on finished paths.
//   * TRAVERSAL is cooperative: the reference's AVX2 design tests 8 child boxes /
//     8 triangles per instruction; here an OCTET (8 adjacent lanes) does the same,
//     one child box or one triangle per lane, with __shfl_xor min-reductions in
//     place of the horizontal min.  The warp's (up to) 32 rays sit in a shared-
//     memory mailbox; each of the 4 octets pulls the next ray from a per-warp
//     bitmask when it finishes one, so octets stay busy until the mailbox drains.
//     The walk is a flat state machine (no recursion, no stack): the tree is a
//     complete 8-ary heap, parent = (n-1)>>3; per-level entry distances live in
//     shared memory and a bitmask of levels that still hold untried children lets
//     a pop jump straight to the next useful ancestor.  Box tests / selection run
//     in one uniform loop body; leaf tests are batched behind it so octets of a
//     warp execute the long Möller–Trumbore body together instead of serialising
//     it against box tests.  Node rows are 32-byte sectors, so an octet's 6 loads
//     per box test and 9 per leaf are fully coalesced.
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
#ifndef RT_BATCH_LEAF
#define RT_BATCH_LEAF 1
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
__device__ __forceinline__ float child_entry(const float *__restrict__ row, float ox, float oy, float oz,
                                             float ix, float iy, float iz, float t_max, bool regular) {
  float ax = (__ldg(row +  0) - ox) * ix;
  float ay = (__ldg(row +  8) - oy) * iy;
  float az = (__ldg(row + 16) - oz) * iz;
  float bx = (__ldg(row + 24) - ox) * ix;
  float by = (__ldg(row + 32) - oy) * iy;
  float bz = (__ldg(row + 40) - oz) * iz;
  float enter, leave;
  if (regular) {
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

__device__ __forceinline__ unsigned octet_min(unsigned omask, unsigned key) {
  key = min(key, __shfl_xor_sync(omask, key, 1));
  key = min(key, __shfl_xor_sync(omask, key, 2));
  key = min(key, __shfl_xor_sync(omask, key, 4));
  return key;
}

struct Shared {
  float ox[RT_BLOCK], oy[RT_BLOCK], oz[RT_BLOCK], dx[RT_BLOCK], dy[RT_BLOCK], dz[RT_BLOCK];   // in:  rays
  float t[RT_BLOCK], u[RT_BLOCK], v[RT_BLOCK];                                                 // out: closest hits
  int   slot[RT_BLOCK];
  float level_entry[RT_BLOCK / 8][RT_MAX_DEPTH + 1][8];
  float texel_lut[256];
  unsigned todo[RT_BLOCK / 32];
};

__global__ void __launch_bounds__(RT_BLOCK, RT_MIN_BLOCKS)
rt_render_kernel(const __grid_constant__ RenderParams P) {
  __shared__ Shared sh;

  const SceneDev &sc = P.scene;
  const int tid   = threadIdx.x;
  const int lane  = tid & 31;
  const int l8    = tid & 7;
  const int wbase = tid & ~31;
  const unsigned omask = 0xffu << (lane & 24);
  float *my_levels = &sh.level_entry[tid >> 3][0][l8];
  volatile unsigned *todo = &sh.todo[tid >> 5];

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
  // counters (octet leader counts traversal work, every lane counts its paths)
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
    const unsigned active = __ballot_sync(0xffffffffu, has_path);
    if (active == 0) break;

    // --------------------------------------------------------------------- trace
    if (has_path) {
      sh.ox[tid] = o.x; sh.oy[tid] = o.y; sh.oz[tid] = o.z;
      sh.dx[tid] = d.x; sh.dy[tid] = d.y; sh.dz[tid] = d.z;
      c_rays++;
    }
    if (lane == 0) *todo = active;
    __syncwarp();
    {
      int   cur = -1, node = 0, level = 0, hit_slot = -1, leaf = -1;
      unsigned pending = 0;               // levels (bit = level) that still hold untried children
      bool  need_box = false, regular = true, finished = false;
      float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, ix = 0, iy = 0, iz = 0;
      float hit_t = CUDART_INF_F, hit_u = 0, hit_v = 0, entry = CUDART_INF_F;

      for (;;) {
        // ---- uniform part: fetch / box test / select / pop, until this octet holds a leaf to test
        while (!finished && leaf < 0) {
          if (cur < 0) {
            int got = -1;
            if (l8 == 0) {                               // the octet leader claims the next ray of the warp
              unsigned seen = *todo;
              while (seen) {
                unsigned bit = seen & (0u - seen);
                unsigned prev = atomicCAS((unsigned *)todo, seen, seen ^ bit);
                if (prev == seen) { got = __ffs(bit) - 1; break; }
                seen = prev;
              }
            }
            got = __shfl_sync(omask, got, lane & 24);
            if (got < 0) { finished = true; break; }
            cur = wbase + got;
            ox = sh.ox[cur]; oy = sh.oy[cur]; oz = sh.oz[cur];
            dx = sh.dx[cur]; dy = sh.dy[cur]; dz = sh.dz[cur];
            ix = 1.0f / dx; iy = 1.0f / dy; iz = 1.0f / dz;         // raytracer.c:198-202
            regular = (fabsf(ix) < CUDART_INF_F) & (fabsf(iy) < CUDART_INF_F) & (fabsf(iz) < CUDART_INF_F);
            hit_t = CUDART_INF_F; hit_slot = -1; hit_u = 0; hit_v = 0;
            node = 0; level = sc.depth; pending = 0;                   // raytracer.c:501
            need_box = true;
          }
          if (need_box) {
            entry = child_entry(sc.nodes + (size_t)node * 48 + l8, ox, oy, oz, ix, iy, iz, hit_t, regular);
            need_box = false;
            c_nodes++;
          }
          // raytracer.c:459-472: nearest unvisited child strictly below the current hit.
          // entry is >= EPS, +inf or NaN: its u32 order equals its f32 order
          const unsigned hit_bits = __float_as_uint(hit_t);
          const unsigned key  = __float_as_uint(entry);
          const unsigned best = octet_min(omask, key);
          if (best >= hit_bits) {
            // nothing left under this node: jump to the nearest ancestor with untried children
            if (pending == 0) {
              if (l8 == 0) { sh.t[cur] = hit_t; sh.u[cur] = hit_u; sh.v[cur] = hit_v; sh.slot[cur] = hit_slot; }
              cur = -1;
              continue;
            }
            const int up = __ffs(pending) - 1;
            pending &= pending - 1;
            for (; level < up; level++) node = (node - 1) >> 3;
            entry = my_levels[level * 8];
            continue;
          }
          const unsigned below = (__ballot_sync(omask, key < hit_bits) >> (lane & 24)) & 0xffu;
          const unsigned who   = (__ballot_sync(omask, key == best) >> (lane & 24)) & 0xffu;
          const int pick = __ffs(who) - 1;                 // lowest index on ties
          if (l8 == pick) entry = CUDART_INF_F;            // raytracer.c:481
          const int child = 8 * node + 1 + pick;
          if (level == 1) {
            leaf = child - sc.n_internal;
          } else {
            if (below & (below - 1)) {                     // more than one candidate: remember this level
              my_levels[level * 8] = entry;
              pending |= 1u << level;
            }
            node = child;
            level -= 1;
            need_box = true;
          }
#if !RT_BATCH_LEAF
          break;
#endif
        }
        if (finished) break;
        if (leaf < 0) continue;

        // ---- raytracer.c:84-188: eight Möller–Trumbore tests, one per lane.
        // leaf rows hold p0 and the edges e1 = p1 - p0, e2 = p2 - p0 (the same f32
        // subtractions raytracer.c:116-122 does per ray, done once at upload)
        {
          const float *lp = sc.leaf_pos + (size_t)leaf * 72 + l8;
          float p0x = __ldg(lp +  0), p0y = __ldg(lp +  8), p0z = __ldg(lp + 16);
          float e1x = __ldg(lp + 24), e1y = __ldg(lp + 32), e1z = __ldg(lp + 40);
          float e2x = __ldg(lp + 48), e2y = __ldg(lp + 56), e2z = __ldg(lp + 64);
          c_leaves++;
          float pvx = dy * e2z - dz * e2y, pvy = dz * e2x - dx * e2z, pvz = dx * e2y - dy * e2x;
          float det = e1x * pvx + e1y * pvy + e1z * pvz;
          float inv_det = 1.0f / det;
          float tvx = ox - p0x, tvy = oy - p0y, tvz = oz - p0z;
          float qvx = tvy * e1z - tvz * e1y, qvy = tvz * e1x - tvx * e1z, qvz = tvx * e1y - tvy * e1x;
          float u = inv_det * (tvx * pvx + tvy * pvy + tvz * pvz);
          float v = inv_det * (dx * qvx + dy * qvy + dz * qvz);
          float t = inv_det * (e2x * qvx + e2y * qvy + e2z * qvz);
          bool miss = (u < -RT_EPS) | (u > 1 + RT_EPS) | (v < -RT_EPS) | (u + v > 1 + RT_EPS) | (t < RT_EPS);
          // raytracer.c:15-32 with eps 0: lanes <= 0 or NaN count as +inf
          float cand = (!miss && t > 0.0f) ? t : CUDART_INF_F;
          unsigned ck = __float_as_uint(cand);
          unsigned cbest = octet_min(omask, ck);
          if (cbest < __float_as_uint(hit_t)) {          // strict <, raytracer.c:159
            unsigned cw = (__ballot_sync(omask, ck == cbest) >> (lane & 24)) & 0xffu;
            int w = (lane & 24) + __ffs(cw) - 1;
            hit_t = __uint_as_float(cbest);
            hit_slot = leaf * 8 + (w & 7);
            hit_u = __shfl_sync(omask, u, w);
            hit_v = __shfl_sync(omask, v, w);
            c_accepts++;
          }
          leaf = -1;
        }
      }
    }
    __syncwarp();

    // --------------------------------------------------------------------- shade
    if (has_path) {
      const int slot = sh.slot[tid];
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
        float w1 = sh.u[tid], w2 = sh.v[tid], w0 = 1 - w1 - w2;           // raytracer.c:164-177
        V3 point  = add3(o, scale3(d, sh.t[tid]));
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
      if ((k >= 1 && k <= 3) && l8 != 0) v = 0;           // traversal work is per octet
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

int rt_launch_render(const RenderParams &p, int sm_count, cudaStream_t stream) {
  if (g_blocks_per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rt_render_kernel, RT_BLOCK, 0) != cudaSuccess || n < 1) n = 1;
    g_blocks_per_sm = n;
  }
  long long jobs = (long long)((p.width + 7) / 8) * ((p.height + 3) / 4) * 32;
  long long want = (jobs + RT_BLOCK - 1) / RT_BLOCK;
  long long grid = (long long)sm_count * g_blocks_per_sm;       // persistent: one wave
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  cudaMemsetAsync(p.job_counter, 0, sizeof(unsigned int), stream);
  rt_render_kernel<<<(unsigned)grid, RT_BLOCK, 0, stream>>>(p);
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
