/* scene.h — scene data model shared by the C host, the GPU library and the
 * oracle.  Mirrors the reference's scene.h:10-119 (same type and field names,
 * same buffer layouts) without the Codin dependency.
 *
 * Buffer contracts (what "BVH layout bit-exact" means):
 *   - BVH_Node (192 B, 32-B aligned): the boxes OF ITS 8 CHILDREN, stored
 *     mins.x[8] mins.y[8] mins.z[8] maxs.x[8] maxs.y[8] maxs.z[8]
 *     (reference scene.h:73-76, common.h:50-52).  Child j of node i is node
 *     8*i+1+j (reference scene.c:398, raytracer.c:474); the tree is a complete
 *     implicit heap with `depth` internal levels; the virtual row after the
 *     last internal one are leaves of 8 triangle slots each.
 *   - Triangles: one allocation of N*(9*4+sizeof(Triangle_AOS)) bytes holding
 *     nine f32[N] arrays x0 x1 x2 y0 y1 y2 z0 z1 z2 and then Triangle_AOS[N]
 *     (reference scene.h:54-63, scene.c:78-99).  Unused slots are all-zero.
 */
#ifndef RT_SCENE_H
#define RT_SCENE_H

#include "rt_base.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { Vec3 min, max; } AABB;

typedef struct {
  Matrix_4x4 view_matrix;
  f32        fov, focal_length;
} Camera;

typedef struct {
  Vec3 direction, normal, normal_geo, tangent, bitangent, position;
  Vec2 tex_coords;
} Shader_Input;

typedef struct {
  Vec3   direction;
  Color3 tint, emission;
  bool   terminate;
} Shader_Output;

typedef void (*Shader_Proc)(rawptr, Shader_Input const *, Shader_Output *);

typedef struct {
  rawptr      data;
  Shader_Proc proc;
} Shader;

typedef struct {
  Vec3   positions [3];
  Vec3   normals   [3];
  Vec2   tex_coords[3];
  Shader shader;
} Triangle;

typedef struct { Triangle *data; isize len; } Triangle_Slice;

typedef struct {
  Vec3   normal, normal_a, normal_b, normal_c;
  Vec3   tangent, bitangent;
  Vec2   tex_coords_a, tex_coords_b, tex_coords_c;
  Shader shader;
} Triangle_AOS;

typedef struct {
  f32          *x[3];
  f32          *y[3];
  f32          *z[3];
  Triangle_AOS *aos;
  i32           len;
} Triangles;

#define TRIANGLES_ALLOCATION_SIZE(N) \
  ((size_t)(N) * (sizeof(f32) * 9 + sizeof(Triangle_AOS)))

typedef Color3 (*Background_Proc)(rawptr, Vec3 direction);

typedef struct {
  Background_Proc proc;
  rawptr          data;
} Background;

typedef struct {
  f32 mins[3][RT_SIMD_WIDTH];
  f32 maxs[3][RT_SIMD_WIDTH];
} __attribute__((aligned(32))) BVH_Node;

typedef int BVH_Index;

typedef struct {
  struct { BVH_Node *data; isize len; } nodes;
  isize depth;
  isize last_row_offset;
} BVH;

typedef struct {
  BVH        bvh;
  Camera     camera;
  Triangles  triangles;
  Background background;
} Scene;

/* 8^depth (reference scene.h:103-109). */
static inline isize bvh_n_leaf_nodes(isize depth) {
  isize n = 1;
  while (depth-- > 0) n *= RT_SIMD_WIDTH;
  return n;
}

/* sum_{k<depth} 8^k (reference scene.h:111-119). */
static inline isize bvh_n_internal_nodes(isize depth) {
  isize total = 0, row = 1;
  while (depth-- > 0) { total += row; row *= RT_SIMD_WIDTH; }
  return total;
}

/* Host BVH build + triangle packing (reference scene.c:416-426).  The
 * reference takes a Codin `Allocator`; here buffers come from aligned_alloc and
 * are released by scene_destroy.  Implemented in
 * raytracing_c_b200/host/scene_build.c. */
void scene_init(Scene *scene, Triangle_Slice src_triangles);
void scene_destroy(Scene *scene);

#ifdef __cplusplus
}
#endif
#endif
