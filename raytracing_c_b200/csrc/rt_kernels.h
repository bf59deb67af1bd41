// rt_kernels.h — launch wrappers shared between the kernel translation units and
// the C-ABI layer (rt_cabi.cu).  Each returns a cudaError_t value as int.
#pragma once

#include <cuda_runtime.h>
#include "rt_device.cuh"

int rt_launch_render(const RenderParams &p, int sm_count, cudaStream_t stream);
int rt_launch_resolve(const float *accum, int width, int height, int samples, unsigned char *pixels,
                      int stride, int components, cudaStream_t stream);
int rt_launch_denoise(const unsigned char *src, unsigned char *dst, int width, int height,
                      int src_stride, int dst_stride, int components, cudaStream_t stream);
int rt_render_blocks_per_sm(void);
