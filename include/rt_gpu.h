/* rt_gpu.h — additive C ABI of libraytracer_gpu.so (plain C, no Codin and no
 * torch types).  The five reference entry points live in raytracer.h and
 * denoiser.h; everything here is what a GPU back end needs in addition:
 * telling the library which host callbacks it replaces, uploading the scene
 * once, parity hooks, and a device-pointer level for hosts that own the device
 * memory themselves (bench.py wraps torch allocations and NCCL around it).
 *
 * Every call returns 0 on success, non-zero on failure, with the text in
 * rt_gpu_last_error().  There is no CPU fallback: without a CUDA device every
 * compute call fails.
 */
#ifndef RT_GPU_H
#define RT_GPU_H

#include "raytracer.h"
#include "denoiser.h"
#include "rt_pbr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lifecycle ---- */
int         rt_gpu_init(int device);              /* cudaSetDevice + stream; idempotent per device */
void        rt_gpu_shutdown(void);
char const *rt_gpu_last_error(void);
int         rt_gpu_sm_count(void);
/* Measured issue rate of non-fused FP32 multiply/add in lane-operations per second:
 * the roofline denominator for the trace kernel (csrc/rt_peak.cu). */
f64         rt_gpu_measure_fp32_issue(void);

/* ---- callbacks the device code internalises ----
 * Replaces the per-triangle Shader.proc indirect call (reference scene.h:30-35,
 * raytracer.c:535) and Scene.background.proc (scene.h:65-70, raytracer.c:554).
 * A triangle whose proc was not registered makes the upload fail. */
void rt_gpu_register_pbr_shader(Shader_Proc proc);      /* Shader.data is a PBR_Shader_Data*  */
void rt_gpu_register_background(Background_Proc proc);  /* Background.data is an Image* (equirect) */
/* Ready-made identities for hosts that have no CPU shader of their own; calling
 * them on the CPU aborts (there is no CPU path). */
void   rt_gpu_pbr_shader_proc(rawptr data, Shader_Input const *in, Shader_Output *out);
Color3 rt_gpu_background_proc(rawptr image, Vec3 direction);

/* ---- scene residency: upload once, keyed by the Scene pointer ---- */
int  rt_gpu_scene_upload(Scene const *scene);
isize rt_gpu_scene_device_bytes(Scene const *scene);   /* bytes resident for this scene, 0 if absent */
isize rt_gpu_scene_upload_bytes(Scene const *scene);   /* bytes its upload copied host -> device */
void rt_gpu_scene_release(Scene const *scene);

/* ---- options for the raytracer.h entry points ---- */
typedef struct {
  u32   user_seed;        /* rt_seed.h; default 0 */
  i32   sample_begin;     /* render samples [begin, end) of ctx->samples; end 0 = all */
  i32   sample_end;
  i32   slice_samples;    /* samples per progress slice of render_thread_proc (one wavefront chunk
                           * never spans slices); default 64 */
  i32   keep_hit_ids;     /* record the primary-hit slot of sample `sample_begin` */
} RT_GPU_Options;
void rt_gpu_set_options(RT_GPU_Options const *options);
void rt_gpu_get_options(RT_GPU_Options *options);

/* ---- parity hooks: results of the LAST render_thread_proc on this process ---- */
enum {
  RT_GPU_CTR_RAYS = 0, RT_GPU_CTR_NODES, RT_GPU_CTR_LEAVES, RT_GPU_CTR_ACCEPTS,
  RT_GPU_CTR_SHADES, RT_GPU_CTR_MISSES, RT_GPU_CTR_PASSTHROUGH, RT_GPU_CTR_SAMPLES,
};
int rt_gpu_read_accum(f32 *out, isize n_floats);       /* W*H*3 sums of cast_ray, pre-division */
int rt_gpu_read_hit_ids(i32 *out, isize n_pixels);     /* padded slot or -1 */
int rt_gpu_read_counters(u64 out[8]);
int rt_gpu_last_launches(void);                        /* kernels launched since the last entry-point call began */
f64 rt_gpu_last_kernel_ms(void);                       /* CUDA-event time of the last call's trace kernels */

/* ---- per-stage kernel timing (measurement harness): when enabled, every kernel of a render is
 *      bracketed by CUDA events on its launching stream; read = synchronise + sum since enable.
 *      Stages: 0 trace (closest hit), 1 miss (environment), 2 shade (BSDF), 3 accumulate. ---- */
void rt_gpu_stage_profile_enable(i32 on);
int  rt_gpu_stage_profile_read(f64 ms[4], i64 launches[4]);
/* the same split by bounce: slot = stage * 16 + min(bounce, 15) */
int  rt_gpu_stage_profile_read_bounces(f64 ms[64], i64 launches[64]);

/* ---- device-pointer level (all pointers are device memory; stream is a
 *      cudaStream_t passed as void*, NULL = the legacy default stream) ---- */
int rt_gpu_render_accum_device(Scene const *scene, isize width, isize height,
                               isize sample_begin, isize sample_end, isize max_bounces,
                               u32 user_seed, i32 accumulate,
                               f32 *d_accum,          /* W*H*3 */
                               f32 *d_per_sample,     /* optional W*H*(end-begin)*3 */
                               i32 *d_hit_ids,        /* optional W*H */
                               u64 *d_counters,       /* optional 8, atomically added */
                               void *stream);
int rt_gpu_resolve_device(f32 const *d_accum, isize width, isize height, isize samples,
                          u8 *d_pixels, isize stride, i32 components, void *stream);
int rt_gpu_denoise_device(u8 const *d_src, u8 *d_dst, isize width, isize height,
                          isize src_stride, isize dst_stride, i32 components, void *stream);

#ifdef __cplusplus
}
#endif
#endif
