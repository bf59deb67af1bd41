/* oracle_scene.c — TEST INFRASTRUCTURE.  Restatement of the reference's BVH
 * builder and triangle packing (reference scene.c:78-242,311-426).
 *
 * Conscious deviations (both documented in DESIGN.md):
 *   Q1  the `len <= 8` early-out (scene.c:318-321) is taken only at leaf level;
 *       at an internal level it would compute a negative slot offset
 *       (fov_test.obj: 72 = 64 + 8).  Falling through to the generic loop gives
 *       split = 0 and one finished slice in child slot 0.
 *   Q2  depth is forced >= 1 (quad.obj has 2 triangles => depth 0 => zero
 *       nodes, but raytracer.c:451 always reads node 0).
 * UNPINNED: Codin's sort_slice_by (scene.c:206) — taken to be a stable sort.
 * The caller's triangle array is copied first (the reference sorts in place).
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <math.h>

#include "oracle.h"
#include "oracle_vec.h"

/* scene.c:224-233 */
static isize required_depth(isize n_triangles) {
  isize groups = (n_triangles + RT_SIMD_WIDTH - 1) / RT_SIMD_WIDTH;
  isize cap = 1, depth = 0;
  while (cap < groups) { cap *= RT_SIMD_WIDTH; depth += 1; }
  return depth;
}

/* scene.c:235-242 */
static isize partition_point(isize n_triangles, isize per_child) {
  isize taken = 0, remaining = n_triangles;
  while (taken < n_triangles / 2 && remaining > per_child) {
    taken     += per_child;
    remaining -= per_child;
  }
  return taken;
}

/* scene.c:177-188 */
static AABB triangle_bounds(Triangle const *t) {
  AABB box;
  for (int a = 0; a < 3; a++) {
    f32 p0 = t->positions[0].data[a], p1 = t->positions[1].data[a], p2 = t->positions[2].data[a];
    box.min.data[a] = f32_min(p0, f32_min(p1, p2)) - RT_EPSILON;
    box.max.data[a] = f32_max(p0, f32_max(p1, p2)) + RT_EPSILON;
  }
  return box;
}

/* scene.c:190-201 (incl. the zero box for an empty slice) */
static AABB slice_bounds(Triangle const *tris, isize n) {
  AABB acc;
  memset(&acc, 0, sizeof acc);
  for (isize i = 0; i < n; i++) {
    AABB b = triangle_bounds(&tris[i]);
    if (i == 0) acc = b;
    for (int a = 0; a < 3; a++) {
      acc.min.data[a] = f32_min(acc.min.data[a], b.min.data[a]);
      acc.max.data[a] = f32_max(acc.max.data[a], b.max.data[a]);
    }
  }
  return acc;
}

/* scene.c:157-162 */
static f32 surface_area(AABB const *b) {
  f32 x = b->max.x - b->min.x;
  f32 y = b->max.y - b->min.y;
  f32 z = b->max.z - b->min.z;
  return 2.0f * (x * y + y * z + z * x);
}

/* scene.c:213-219: key = p0[a] + p1[a] + p2[a], compared with strict < */
static f32 centroid_key(Triangle const *t, int axis) {
  return t->positions[0].data[axis] + t->positions[1].data[axis] + t->positions[2].data[axis];
}

/* Stable top-down merge sort of Triangle records (stands in for sort_slice_by). */
static void merge_sort(Triangle *a, Triangle *tmp, isize n, int axis) {
  if (n < 2) return;
  isize h = n / 2;
  merge_sort(a, tmp, h, axis);
  merge_sort(a + h, tmp, n - h, axis);
  isize i = 0, j = h, k = 0;
  while (i < h && j < n) {
    if (centroid_key(&a[j], axis) < centroid_key(&a[i], axis)) tmp[k++] = a[j++];
    else                                                       tmp[k++] = a[i++];
  }
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof *a);
}

typedef struct {
  Scene    *scene;
  Triangle *scratch;
} Build;

/* scene.c:105-155 */
static void insert_leaf(Triangles *dst, Triangle const *src, isize n, isize offset) {
  assert(offset >= 0 && n + offset <= dst->len);
  for (isize i = 0; i < n; i++) {
    Triangle const *t = &src[i];
    isize s = offset + i;
    for (int v = 0; v < 3; v++) {
      dst->x[v][s] = t->positions[v].x;
      dst->y[v][s] = t->positions[v].y;
      dst->z[v][s] = t->positions[v].z;
    }
    Vec3 e1 = v3_sub(t->positions[1], t->positions[0]);
    Vec3 e2 = v3_sub(t->positions[2], t->positions[0]);
    f32 du1 = t->tex_coords[1].x - t->tex_coords[0].x, dv1 = t->tex_coords[1].y - t->tex_coords[0].y;
    f32 du2 = t->tex_coords[2].x - t->tex_coords[0].x, dv2 = t->tex_coords[2].y - t->tex_coords[0].y;
    f32 d = du1 * dv2 - du2 * dv1;
    if (f32_abs(d) < 0.0001f) d = (d < 0) ? -0.0001f : 0.0001f;
    f32 inv_d = 1.0f / d;

    Triangle_AOS rec;
    memset(&rec, 0, sizeof rec);
    rec.tangent   = v3_normalize(v3_scale(v3_sub(v3_scale(e1, dv2), v3_scale(e2, dv1)), inv_d));
    rec.bitangent = v3_normalize(v3_scale(v3_sub(v3_scale(e2, du1), v3_scale(e1, du2)), inv_d));
    rec.normal    = v3_normalize(v3_cross(e1, e2));
    rec.normal_a  = t->normals[0];
    rec.normal_b  = t->normals[1];
    rec.normal_c  = t->normals[2];
    rec.tex_coords_a = t->tex_coords[0];
    rec.tex_coords_b = t->tex_coords[1];
    rec.tex_coords_c = t->tex_coords[2];
    rec.shader    = t->shader;
    dst->aos[s] = rec;
  }
}

/* scene.c:311-414 (sequential; the task queue at :262-309 only reorders
 * independent subtrees, results are schedule-independent). */
static void build_node(Build *b, Triangle *tris, isize n, isize depth, isize index) {
  Scene *scene = b->scene;
  if (depth == 0) {                                   /* Q1: leaf level only */
    assert(n <= RT_SIMD_WIDTH);
    insert_leaf(&scene->triangles, tris, n, (index - scene->bvh.last_row_offset) * RT_SIMD_WIDTH);
    return;
  }

  isize per_child = bvh_n_leaf_nodes(depth);
  assert(per_child * RT_SIMD_WIDTH >= n);

  struct { Triangle *p; isize n; } pending[RT_SIMD_WIDTH], done[RT_SIMD_WIDTH];
  isize n_pending = 0, n_done = 0;
  pending[n_pending].p = tris; pending[n_pending].n = n; n_pending++;

  while (n_pending) {
    n_pending--;
    Triangle *sp = pending[n_pending].p;
    isize     sn = pending[n_pending].n;
    isize  split = partition_point(sn, per_child);

    f32 best_area = INFINITY;
    int best_axis = 0;
    for (int axis = 0; axis < 3; axis++) {
      merge_sort(sp, b->scratch, sn, axis);
      AABB l = slice_bounds(sp, split);
      AABB r = slice_bounds(sp + split, sn - split);
      f32 area = surface_area(&l) + surface_area(&r);
      if (area <= best_area) { best_area = area; best_axis = axis; }
    }
    if (best_axis != 2) merge_sort(sp, b->scratch, sn, best_axis);

    isize ln = split, rn = sn - split;
    if (ln > per_child)      { assert(n_pending < RT_SIMD_WIDTH); pending[n_pending].p = sp; pending[n_pending].n = ln; n_pending++; }
    else if (ln)             { assert(n_done < RT_SIMD_WIDTH);    done[n_done].p = sp;       done[n_done].n = ln;       n_done++; }
    if (rn > per_child)      { assert(n_pending < RT_SIMD_WIDTH); pending[n_pending].p = sp + split; pending[n_pending].n = rn; n_pending++; }
    else if (rn)             { assert(n_done < RT_SIMD_WIDTH);    done[n_done].p = sp + split;       done[n_done].n = rn;       n_done++; }
  }

  BVH_Node node;
  memset(&node, 0, sizeof node);
  for (isize i = 0; i < n_done; i++) {
    AABB box = slice_bounds(done[i].p, done[i].n);
    for (int a = 0; a < 3; a++) {
      node.mins[a][i] = box.min.data[a];
      node.maxs[a][i] = box.max.data[a];
    }
    build_node(b, done[i].p, done[i].n, depth - 1, index * RT_SIMD_WIDTH + 1 + i);
  }
  scene->bvh.nodes.data[index] = node;
}

/* scene.c:78-99 + 416-426 */
void oracle_scene_init(Scene *scene, Triangle_Slice src) {
  isize depth = required_depth(src.len);
  if (depth < 1) depth = 1;                           /* Q2 */
  isize n_internal = bvh_n_internal_nodes(depth);
  scene->bvh.depth           = depth;
  scene->bvh.last_row_offset = n_internal;
  scene->bvh.nodes.len       = n_internal;
  scene->bvh.nodes.data      = aligned_alloc(64, (size_t)n_internal * sizeof(BVH_Node));
  memset(scene->bvh.nodes.data, 0, (size_t)n_internal * sizeof(BVH_Node));

  isize slots = bvh_n_leaf_nodes(depth) * RT_SIMD_WIDTH;
  size_t bytes = (TRIANGLES_ALLOCATION_SIZE(slots) + 63) & ~(size_t)63;
  f32 *block = aligned_alloc(64, bytes);
  memset(block, 0, bytes);
  Triangles *t = &scene->triangles;
  t->len = (i32)slots;
  for (int v = 0; v < 3; v++) {
    t->x[v] = block + slots * (0 + v);
    t->y[v] = block + slots * (3 + v);
    t->z[v] = block + slots * (6 + v);
  }
  t->aos = (Triangle_AOS *)(block + slots * 9);

  Build b;
  b.scene   = scene;
  Triangle *work = malloc((size_t)(src.len ? src.len : 1) * sizeof(Triangle));
  b.scratch      = malloc((size_t)(src.len ? src.len : 1) * sizeof(Triangle));
  memcpy(work, src.data, (size_t)src.len * sizeof(Triangle));
  build_node(&b, work, src.len, depth, 0);
  free(work);
  free(b.scratch);
}

void oracle_scene_destroy(Scene *scene) {
  free(scene->bvh.nodes.data);
  free(scene->triangles.x[0]);
  scene->bvh.nodes.data = NULL;
  scene->triangles.x[0] = NULL;
}
