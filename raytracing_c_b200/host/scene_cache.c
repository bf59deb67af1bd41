/* scene_cache.c — the scene cache file (reference scene.c:13-76: scene_save_writer / scene_load_bytes).
 *
 * Same container as the reference's: a 32-byte aligned header {version, n_nodes, n_triangles, bvh_depth, camera},
 * then the BVH_Node array, then the triangle block (nine f32 arrays and the Triangle_AOS records) byte for byte
 * as scene_init laid them out.  One thing differs, and the version field says so: the reference (version 0) writes
 * every Triangle_AOS.shader as the two raw host pointers it holds, so its files are only valid inside the process
 * that wrote them.  Version 1 stores the material INDEX (+1; 0 = none) in `shader.data` and NULL in `shader.proc`;
 * loading re-binds both against the caller's material table and shader callback.  Version 0 files are refused.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rt_host.h"

void rt_host_set_error(char const *msg);

typedef struct __attribute__((aligned(32))) {
  i32    version, n_nodes, n_triangles, bvh_depth;
  Camera camera;
} Scene_File_Header;

enum { SCENE_FILE_VERSION = 1 };

bool scene_save_file(char const *path, Scene const *scene, PBR_Shader_Data const *materials, isize n_materials) {
  isize n_slots = scene->triangles.len, n_nodes = scene->bvh.nodes.len;
  if (!scene->bvh.nodes.data || !scene->triangles.x[0] || !scene->triangles.aos) { rt_host_set_error("scene cache: scene has no buffers"); return false; }
  /* scene_init lays the nine arrays and the records out as ONE block (scene.h:62-63); a scene assembled differently cannot be cached */
  if (scene->triangles.x[1] != scene->triangles.x[0] + n_slots || (f32 *)scene->triangles.aos != scene->triangles.x[0] + n_slots * 9) {
    rt_host_set_error("scene cache: the triangle buffers are not one block");
    return false;
  }
  size_t tri_bytes = TRIANGLES_ALLOCATION_SIZE(n_slots);
  u8 *block = malloc(tri_bytes ? tri_bytes : 1);
  if (!block) { rt_host_set_error("scene cache: out of memory"); return false; }
  memcpy(block, scene->triangles.x[0], tri_bytes);
  Triangle_AOS *aos = (Triangle_AOS *)(block + (size_t)n_slots * 9 * sizeof(f32));
  for (isize i = 0; i < n_slots; i++) {
    isize index = 0;
    if (aos[i].shader.data) {
      PBR_Shader_Data const *m = aos[i].shader.data;
      if (m < materials || m >= materials + n_materials) { free(block); rt_host_set_error("scene cache: a triangle's material is outside the table"); return false; }
      index = (m - materials) + 1;
    }
    aos[i].shader.data = (rawptr)(uintptr_t)index;
    aos[i].shader.proc = NULL;
  }
  Scene_File_Header header;
  memset(&header, 0, sizeof header);
  header.version = SCENE_FILE_VERSION;
  header.n_nodes = (i32)n_nodes;
  header.n_triangles = (i32)n_slots;
  header.bvh_depth = (i32)scene->bvh.depth;
  header.camera = scene->camera;
  FILE *f = fopen(path, "wb");
  bool ok = f && fwrite(&header, sizeof header, 1, f) == 1 &&
            (n_nodes == 0 || fwrite(scene->bvh.nodes.data, sizeof(BVH_Node), (size_t)n_nodes, f) == (size_t)n_nodes) &&
            fwrite(block, 1, tri_bytes, f) == tri_bytes;
  if (f) ok = fclose(f) == 0 && ok;
  free(block);
  if (!ok) rt_host_set_error("scene cache: write failed");
  return ok;
}

bool scene_load_file(char const *path, Scene *scene, PBR_Shader_Data *materials, isize n_materials, Shader_Proc proc) {
  FILE *f = fopen(path, "rb");
  if (!f) { rt_host_set_error("scene cache: cannot open file"); return false; }
  Scene_File_Header header;
  bool ok = fread(&header, sizeof header, 1, f) == 1;
  fseek(f, 0, SEEK_END);
  long file_len = ftell(f);
  if (!ok || header.version != SCENE_FILE_VERSION) {
    fclose(f);
    rt_host_set_error(ok && header.version == 0 ? "scene cache: version 0 files hold raw host pointers and cannot be loaded"
                                                : "scene cache: not a scene file of this version");
    return false;
  }
  /* the three counts must describe one complete 8-ary tree (scene.h:103-119) and account for every byte (scene.c:44-46) */
  isize depth = header.bvh_depth, n_nodes = header.n_nodes, n_slots = header.n_triangles;
  if (depth < 1 || depth > 8 || n_nodes != bvh_n_internal_nodes(depth) || n_slots != bvh_n_leaf_nodes(depth) * RT_SIMD_WIDTH ||
      (size_t)file_len != sizeof header + (size_t)n_nodes * sizeof(BVH_Node) + TRIANGLES_ALLOCATION_SIZE(n_slots)) {
    fclose(f);
    rt_host_set_error("scene cache: header and file size disagree");
    return false;
  }
  size_t node_bytes = (size_t)n_nodes * sizeof(BVH_Node), tri_bytes = TRIANGLES_ALLOCATION_SIZE(n_slots);
  BVH_Node *nodes = rt_host_buffer_alloc(node_bytes);
  f32 *block = rt_host_buffer_alloc((tri_bytes + 63) & ~(size_t)63);
  fseek(f, (long)sizeof header, SEEK_SET);
  ok = nodes && block && fread(nodes, 1, node_bytes, f) == node_bytes && fread(block, 1, tri_bytes, f) == tri_bytes;
  fclose(f);
  Triangle_AOS *aos = ok ? (Triangle_AOS *)(block + (size_t)n_slots * 9) : NULL;
  for (isize i = 0; ok && i < n_slots; i++) {
    uintptr_t index = (uintptr_t)aos[i].shader.data;
    if (index > (uintptr_t)n_materials || aos[i].shader.proc) { ok = false; break; }
    aos[i].shader.data = index ? &materials[index - 1] : NULL;
    aos[i].shader.proc = index ? proc : NULL;
  }
  if (!ok) {
    rt_host_buffer_free(nodes);
    rt_host_buffer_free(block);
    rt_host_set_error("scene cache: short read or a material index outside the table");
    return false;
  }
  scene->camera = header.camera;
  scene->bvh.depth = depth;
  scene->bvh.last_row_offset = n_nodes;
  scene->bvh.nodes.data = nodes;
  scene->bvh.nodes.len = n_nodes;
  for (int v = 0; v < 3; v++) {
    scene->triangles.x[v] = block + n_slots * (0 + v);
    scene->triangles.y[v] = block + n_slots * (3 + v);
    scene->triangles.z[v] = block + n_slots * (6 + v);
  }
  scene->triangles.aos = aos;
  scene->triangles.len = (i32)n_slots;
  return true;
}
