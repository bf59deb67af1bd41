/* ref_glue_rt.c — TEST INFRASTRUCTURE.  Includes the reference's raytracer.c
 * UNMODIFIED (it is one translation unit, as in the reference build) and adds
 * exported doors to its file-local functions so tests can call cast_ray and the
 * traversal directly. */
#include "raytracer.c"

/* raytracer.c:505 */
Color3 ref_cast_ray(Scene const *scene, Ray ray, isize max_bounces) { return cast_ray(scene, ray, max_bounces); }

typedef struct { f32 distance; Vec3 normal, normal_geo, point, tangent, bitangent; Vec2 tex_coords; rawptr shader_data; } Ref_Hit;

/* raytracer.c:497 */
void ref_trace_ray(Scene const *scene, Ray ray, Ref_Hit *out) {
  Hit hit = { .distance = F32_INFINITY, };
  ray_scene_hit(&ray, scene, &hit);
  out->distance = hit.distance; out->normal = hit.normal; out->normal_geo = hit.normal_geo; out->point = hit.point;
  out->tangent = hit.tangent; out->bitangent = hit.bitangent; out->tex_coords = hit.tex_coords; out->shader_data = hit.shader.data;
}

/* raytracer.c:584, lane 0 */
f32 ref_hash12(f32 px, f32 py) {
  Vec2x8 p = { _mm256_set1_ps(px), _mm256_set1_ps(py) };
  return _mm256_cvtss_f32(hash12x8(p));
}

/* raytracer.c's own copy of the generator (common.h:13 is static thread_local in a header): feeds rand_vec3 in lightmap_bake */
u32 *ref_rt_random_state(void) { return &random_state; }
