// rt_kernels.h — launch wrappers shared between the kernel translation units and
// the C-ABI layer (rt_cabi.cu).  Each returns a cudaError_t value as int.
#pragma once

#include <cuda_runtime.h>
#include "rt_device.cuh"

// wavefront render of samples [p.sample_begin, p.sample_end) into p.accum; `workspace` holds the path queues
// (rt_render_workspace_bytes; a smaller one only means more, smaller chunks).  *n_launches += kernels launched.
size_t rt_render_workspace_bytes(int width, int height, int n_samples, int max_bounces, int slice_samples, int split_world);
int rt_render_chunk_samples(int width, int height, int split_world);
unsigned rt_render_tiles(int width, int height, int split_rank, int split_world);
int rt_launch_render(const RenderParams &p, int sm_count, void *workspace, size_t workspace_bytes,
                     cudaStream_t stream, int *n_launches);
// lightmap_bake (raytracer.c:722-784) on the wavefront: see rt_render.cu
size_t rt_lightmap_workspace_bytes(int width, int height);
int rt_launch_lightmap(const SceneDev &scene, int width, int height, int samples, int max_bounces, uint32_t user_seed,
                       unsigned char *d_pixels, int stride, int components, float *d_values, int *d_owner_out,
                       void *scratch, int sm_count, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                       int *n_launches, unsigned *n_jobs_out);
int rt_launch_resolve(const float *accum, int width, int height, int samples, unsigned char *pixels,
                      int stride, int components, cudaStream_t stream);
// sum of parts.n accumulators (own + peer-mapped) in rank order, optional copy of the sum, optional resolve to u8
int rt_launch_reduce_resolve(const ReduceParts &parts, float *sum_out, int width, int height, int samples,
                             unsigned char *pixels, int stride, int components, cudaStream_t stream);
int rt_launch_denoise(const unsigned char *src, unsigned char *dst, int width, int height,
                      int src_stride, int dst_stride, int components, cudaStream_t stream);
int rt_render_blocks_per_sm(void);
// RGB8 / RGBA8 rows (stride in pixels) -> tightly packed RGBA8 texels
int rt_launch_texel_repack(const unsigned char *src, int width, int height, int stride, int components,
                           uchar4 *dst, cudaStream_t stream);
// RGBA8 texels -> float4 (r, g, b, 0) / 255.999f: the environment's second copy (TextureDev.texels_f32)
int rt_launch_texel_expand(const uchar4 *src, size_t n, float4 *dst, cudaStream_t stream);
// tri_pos / tri_rec from the host's vertex arrays, Triangle_AOS records and per-slot material indices (rt_denoise.cu)
int rt_launch_scene_pack(const float *soa, const float4 *aos, const int *mat_index, int n_slots, float4 *tri_pos, float4 *tri_rec,
                         cudaStream_t stream);
// per-stage CUDA-event timing of rt_launch_render's kernels (off by default)
// slot = stage * RT_STAGE_BOUNCES + min(bounce, RT_STAGE_BOUNCES - 1)
enum { RT_STAGE_TRACE = 0, RT_STAGE_MISS, RT_STAGE_SHADE, RT_STAGE_ACCUMULATE, RT_N_STAGES, RT_STAGE_BOUNCES = 16 };
void rt_stage_profile_enable(int on, int device);
int  rt_stage_profile_read(double ms[RT_N_STAGES * RT_STAGE_BOUNCES], long long launches[RT_N_STAGES * RT_STAGE_BOUNCES]);
