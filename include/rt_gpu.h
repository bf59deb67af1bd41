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

/* ---- lifecycle ----
 * One process drives one device (rt_gpu_init) or several (rt_gpu_init_devices): with N devices the
 * raytracer.h entry points split every frame over them (see RT_GPU_Options.split_mode), the per-device f32
 * accumulators are combined on devices[0] — by one fused reduce+resolve kernel that reads the peers'
 * buffers over NVLink, or by an NCCL reduce (RT_GPU_Options.reduce_mode) — and devices[0] resolves,
 * denoises and returns the image.  The reference has no counterpart: its driver starts N host threads
 * (driver.c:793-803); here `--gpus N` of the host program maps to this call. */
int         rt_gpu_init(int device);              /* = rt_gpu_init_devices(1, &device); idempotent per device */
int         rt_gpu_init_devices(int n_devices, int const *devices);   /* devices NULL = ordinals 0..n-1 */
int         rt_gpu_device_count(void);            /* devices this process drives (0 before init) */
int         rt_gpu_visible_devices(void);         /* CUDA devices visible to the process */
void        rt_gpu_shutdown(void);
char const *rt_gpu_last_error(void);
/* 0 = the most recent raytracer.h / denoiser.h entry-point call on this process succeeded; non-zero = it
 * failed (text in rt_gpu_last_error).  Those entry points return void (reference raytracer.h:51-56), so a
 * host checks this after rendering_context_finish / denoise_image. */
int         rt_gpu_last_status(void);
int         rt_gpu_sm_count(void);
/* Measured issue rate of non-fused FP32 multiply/add in lane-operations per second:
 * the roofline denominator for the trace kernel (csrc/rt_peak.cu). */
f64         rt_gpu_measure_fp32_issue(void);

/* Optional: allocate what a frame of this size needs (accumulators, path-queue workspace on every device) now, e.g.
 * on the thread that created the contexts while the model is still loading. */
int         rt_gpu_prepare_frame(isize width, isize height, isize samples, isize max_bounces);

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

/* ---- scene residency: upload once, keyed by the Scene pointer ----
 * The upload is asynchronous: nodes and triangles go first, textures and the environment follow on a copy
 * stream while the first primary trace already runs; host buffers in pinned memory (rt_gpu_host_alloc) are
 * DMA-read in place, pageable ones pass through a pinned staging ring.  The host buffers must stay alive
 * and unchanged until the next render of the scene has started or rt_gpu_scene_release was called.
 * A host that rebuilds a Scene at the same address (scene_init again) is detected by the buffer pointers
 * and sizes and re-uploaded; one that rewrites buffers in place must call rt_gpu_scene_upload again. */
int  rt_gpu_scene_upload(Scene const *scene);
/* pinned host memory for scene buffers and textures (plug into the host's allocator: H2D then needs no staging) */
void *rt_gpu_host_alloc(size_t bytes);
void  rt_gpu_host_free(void *p);
isize rt_gpu_scene_device_bytes(Scene const *scene);   /* bytes resident for this scene, 0 if absent */
isize rt_gpu_scene_upload_bytes(Scene const *scene);   /* bytes its upload copied host -> device */
void rt_gpu_scene_release(Scene const *scene);

/* ---- options for the raytracer.h entry points ---- */
enum { RT_GPU_SPLIT_AUTO = 0, RT_GPU_SPLIT_SAMPLES = 1, RT_GPU_SPLIT_CHUNKS = 2 };
enum { RT_GPU_REDUCE_P2P = 0, RT_GPU_REDUCE_NCCL = 1 };
typedef struct {
  u32   user_seed;        /* rt_seed.h; default 0 */
  i32   sample_begin;     /* render samples [begin, end) of ctx->samples.  With sample_range_set == 0, */
  i32   sample_end;       /* end 0 = all; with sample_range_set != 0 the range is taken literally (may be empty) */
  i32   slice_samples;    /* samples per progress slice of render_thread_proc (one wavefront chunk never spans
                           * slices); 0 (default) = as many as one chunk of path queues holds */
  i32   keep_hit_ids;     /* record the primary-hit slot of sample `sample_begin` */
  i32   sample_range_set;
  i32   split_mode;       /* N devices: RT_GPU_SPLIT_* — sample ranges in multiples of the 8-sample jitter batch
                           * (raytracer.c:641-697), or the reference's 32x32 chunks dealt round-robin
                           * (raytracer.c:619-637); AUTO = samples when the batches divide evenly, else chunks */
  i32   reduce_mode;      /* N devices: RT_GPU_REDUCE_* */
  i32   pixel_rank;       /* multi-PROCESS hosts: this process renders the 32x32 chunks congruent to */
  i32   pixel_world;      /* pixel_rank modulo pixel_world (<= 1: the whole image) */
  i32   fast_math;        /* 0 (default): every sample bit-identical to the reference arithmetic.  1: the bounce
                           * stages (trace of bounce >= 1, environment, BSDF) run FMA-contracted with hardware
                           * transcendentals (csrc/rt_render_fast.cu); primary-hit ids, ray generation and the film
                           * stay exact.  Faster, statistically equivalent, NOT bit-identical. */
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
/* RGBA8 texels of texture `slot` of a resident scene (first device); out may be NULL to query the size */
int rt_gpu_read_texture(Scene const *scene, i32 slot, u8 *out, isize *width, isize *height);
/* the library's own 16-slot device counter block (first device): slots [0..8) as above, [8] rays whose whole walk
 * was the root-union test, [9] primary rays; pass it as d_counters to the device-level calls to get [8..16) filled */
int rt_gpu_counters_buffer(u64 **d_counters_out);
int rt_gpu_counters_reset(void);                       /* synchronises, then zeroes that block */
int rt_gpu_read_counters_ex(u64 out[16]);              /* all 16 slots, summed over the devices */
int rt_gpu_last_launches(void);                        /* kernels launched since the last entry-point call began */
f64 rt_gpu_last_kernel_ms(void);                       /* CUDA-event time of the last call's trace kernels */
/* wall-clock pieces of the last render_thread_proc: [0] scene upload issued inside the call (ms, 0 if resident),
 * [1] cross-device reduce + resolve (CUDA events), [2] image D2H + host copy, [3] split mode used */
void rt_gpu_last_frame_breakdown(f64 out[4]);

/* ---- lightmap_bake (raytracer.h:56, reference raytracer.c:722-784) with its parity hooks: values_out (optional, W*H*3 f32)
 * = accumulated / samples of every written texel before the u8 store; owner_out (optional, W*H i32) = the triangle slot
 * that owns each texel, -1 = untouched.  Seeds: rt_seed.h, per (texel, sample); RT_GPU_Options.user_seed applies. ---- */
int rt_gpu_lightmap_bake(Image const *lightmap, Scene const *scene, isize samples, f32 *values_out, i32 *owner_out);

/* ---- how a frame is split over `world` devices or processes (pure host arithmetic, no device needed) ----
 * Sample split: batches of 8 samples dealt as evenly as possible; ranks beyond the number of batches get an
 * empty range.  Returns the mode AUTO resolves to for this (samples, world). */
void rt_gpu_shard_samples(i32 rank, i32 world, i32 samples, i32 *begin, i32 *end);
i32  rt_gpu_shard_mode(i32 samples, i32 world, i32 requested_mode);
/* 32x32 chunks rank `rank` owns under the chunk split */
i32  rt_gpu_shard_chunks(i32 rank, i32 world, isize width, isize height);

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
/* The same for one shard of a frame split over `world` processes (one per GPU): the library picks the split
 * (RT_GPU_SPLIT_*; returns the mode used in *mode_used) and renders this rank's share of `samples` into d_accum,
 * which is zero outside the share, so the sum over ranks is the frame. */
int rt_gpu_render_shard_device(Scene const *scene, isize width, isize height, isize samples, isize max_bounces,
                               u32 user_seed, i32 rank, i32 world, i32 split_mode, i32 *mode_used,
                               f32 *d_accum, u64 *d_counters, void *stream);
/* Cross-process film: library-owned accumulator (cudaMalloc, exportable), CUDA-IPC handles, and the fused
 * reduce + resolve over peer-mapped accumulators (parts[0..n) summed in that order; d_sum_out and d_pixels optional). */
int rt_gpu_accum_buffer(isize width, isize height, f32 **d_accum_out);
int rt_gpu_ipc_export(void const *d_ptr, u8 handle_out[64]);
int rt_gpu_ipc_open(u8 const handle[64], void **d_ptr_out);
int rt_gpu_reduce_resolve_device(f32 const *const *d_parts, i32 n_parts, f32 *d_sum_out, isize width, isize height,
                                 isize samples, u8 *d_pixels, isize stride, i32 components, void *stream);
int rt_gpu_resolve_device(f32 const *d_accum, isize width, isize height, isize samples,
                          u8 *d_pixels, isize stride, i32 components, void *stream);
int rt_gpu_denoise_device(u8 const *d_src, u8 *d_dst, isize width, isize height,
                          isize src_stride, isize dst_stride, i32 components, void *stream);

#ifdef __cplusplus
}
#endif
#endif
