/* rt_pbr.h — the material record behind Shader.data.
 *
 * In the reference this struct is private to driver.c (driver.c:191-198) and
 * reached through a host function pointer per triangle (scene.h:30-35,50;
 * raytracer.c:535).  A GPU cannot call host pointers, so the record moves to a
 * shared header and the host tells the library which Shader_Proc /
 * Background_Proc values denote "this layout" (rt_gpu_register_pbr_shader,
 * rt_gpu_register_background in rt_gpu.h).  Field names and meaning are the
 * reference's.
 */
#ifndef RT_PBR_H
#define RT_PBR_H

#include "scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  Vec3   base_color, emission;
  f32    roughness, metalness, normal_map_strength, sheen, sheen_tint, anisotropic_strength;
  Image *texture_albedo;
  Image *texture_normal;
  Image *texture_metal_roughness;
  Image *texture_emission;
} PBR_Shader_Data;

#ifdef __cplusplus
}
#endif
#endif
