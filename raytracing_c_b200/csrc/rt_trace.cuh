// rt_trace.cuh — closest-hit BVH traversal, one thread per ray.
//
// Replaces reference raytracer.c:190-230 (ray_aabbs_hit_8), :84-188
// (ray_triangles_hit_8), :15-32 (min_f32x8) and :443-503 (ray_bvh_node_hit /
// ray_scene_hit).  The reference's AVX2 8-wide box and triangle tests become
// eight scalar tests per lane, so a warp tests 256 boxes per node step with no
// cross-lane traffic.  All arithmetic is IEEE f32 in the reference's operation
// order without FMA contraction (-fmad=false): visit order, every compare and the
// closest hit (ties included) are the reference's.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>

#include "rt_device.cuh"

#ifndef RT_BLOCK
#define RT_BLOCK 256
#endif

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

