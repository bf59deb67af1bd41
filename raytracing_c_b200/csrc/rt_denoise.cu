// rt_denoise.cu — 3x3 luminance-median denoiser as a shared-memory tiled stencil.
//
// Replaces denoiser_thread_proc (reference denoiser.c:47-127): for each pixel of
// the u8 sRGB image gather the clamp-to-edge 3x3 neighbourhood as u8/255.999,
// luminance on the encoded values, stable ascending insertion sort by luminance
// (strict >), median = element 4, mean of elements 1..7, blend the centre toward
// the median by clamp(|med - centre| - 5*|med - mean|, 0, 0.0125)/0.0125 and store
// (u8)(c*255.999).  Bit-exact against the oracle (no FMA, same summation order).
//
// One block = 32x8 output pixels; the 34x10 halo tile is staged once in shared
// memory as float4 (r, g, b, luminance), so each source byte is converted once
// instead of nine times.  HBM traffic: 3 B in + 3 B out per pixel.
#include <cuda_runtime.h>

#include "rt_kernels.h"
#include "scene.h"

#define TILE_W 32
#define TILE_H 8
#define HALO_W (TILE_W + 2)
#define HALO_H (TILE_H + 2)

__global__ void __launch_bounds__(TILE_W * TILE_H)
rt_denoise_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, int width, int height,
                  int src_stride, int dst_stride, int components) {
  __shared__ float4 tile[HALO_H][HALO_W];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
  const int n_chan = components < 3 ? components : 3;

  for (int i = ty * TILE_W + tx; i < HALO_W * HALO_H; i += TILE_W * TILE_H) {
    int hx = i % HALO_W, hy = i / HALO_W;
    int sx = x0 + hx - 1, sy = y0 + hy - 1;
    sx = sx < 0 ? 0 : (sx >= width  ? width  - 1 : sx);      // denoiser.c:17-20
    sy = sy < 0 ? 0 : (sy >= height ? height - 1 : sy);
    const unsigned char *p = src + (size_t)(sx + sy * src_stride) * (size_t)components;
    float c[3] = { 0, 0, 0 };
    for (int k = 0; k < n_chan; k++) c[k] = (float)p[k] / 255.999f;
    float luma = c[0] * 0.2126f + c[1] * 0.7152f + c[2] * 0.0722f;
    tile[hy][hx] = make_float4(c[0], c[1], c[2], luma);
  }
  __syncthreads();

  const int x = x0 + tx, y = y0 + ty;
  if (x >= width || y >= height) return;

  // insertion in tap order (dy, dx) = (-1,-1) .. (1,1); an element moves past a
  // predecessor only if that predecessor is strictly greater -> stable
  float lum[9];
  int   idx[9];
  #pragma unroll
  for (int k = 0; k < 9; k++) {
    lum[k] = tile[ty + k / 3][tx + k % 3].w;
    idx[k] = k;
    #pragma unroll
    for (int j = k; j > 0; j--) {
      bool swap = lum[j - 1] > lum[j];
      float lo = swap ? lum[j] : lum[j - 1], hi = swap ? lum[j - 1] : lum[j];
      int   li = swap ? idx[j] : idx[j - 1], hi_i = swap ? idx[j - 1] : idx[j];
      lum[j - 1] = lo; lum[j] = hi; idx[j - 1] = li; idx[j] = hi_i;
    }
  }

  float4 centre = tile[ty + 1][tx + 1];
  float4 median = tile[ty + idx[4] / 3][tx + idx[4] % 3];
  float mean = 0;
  #pragma unroll
  for (int k = 1; k < 8; k++) mean += lum[k];
  mean /= 7;
  float noisiness = fabsf(median.w - mean);
  float diff = fabsf(median.w - centre.w) - noisiness * 5;
  diff = clamp1(diff, 0, 0.0125f) / 0.0125f;
  float out[3] = { lerp1(centre.x, median.x, diff), lerp1(centre.y, median.y, diff), lerp1(centre.z, median.z, diff) };
  unsigned char *q = dst + (size_t)(x + y * dst_stride) * (size_t)components;
  for (int k = 0; k < n_chan; k++) q[k] = (unsigned char)(out[k] * 255.999f);
}

int rt_launch_denoise(const unsigned char *src, unsigned char *dst, int width, int height,
                      int src_stride, int dst_stride, int components, cudaStream_t stream) {
  dim3 block(TILE_W, TILE_H);
  dim3 grid((width + TILE_W - 1) / TILE_W, (height + TILE_H - 1) / TILE_H);
  rt_denoise_kernel<<<grid, block, 0, stream>>>(src, dst, width, height, src_stride, dst_stride, components);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- texel repack
// Scene upload helper: the samplers read one 32-bit RGBA8 texel per tap (rt_shade.cuh); host
// images are RGB8 (or RGBA8) rows with a pixel stride.
__global__ void rt_texel_repack_kernel(const unsigned char *__restrict__ src, int width, int height, int stride,
                                       int components, uchar4 *__restrict__ dst) {
  const size_t n = (size_t)width * (size_t)height;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t x = i % (size_t)width, y = i / (size_t)width;
    const unsigned char *p = src + (size_t)components * (x + (size_t)stride * y);
    dst[i] = make_uchar4(p[0], p[1], p[2], 255);
  }
}

int rt_launch_texel_repack(const unsigned char *src, int width, int height, int stride, int components,
                           uchar4 *dst, cudaStream_t stream) {
  const size_t n = (size_t)width * (size_t)height;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  if (grid < 1) grid = 1;
  rt_texel_repack_kernel<<<grid, 256, 0, stream>>>(src, width, height, stride, components, dst);
  return (int)cudaGetLastError();
}

// RGBA8 texels -> (r, g, b) / 255.999f as floats (driver.c:69-88: the division the reference does per tap; IEEE, so the
// same bits as the samplers' 256-entry table holds)
__global__ void rt_texel_expand_kernel(const uchar4 *__restrict__ src, size_t n, float4 *__restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uchar4 t = src[i];
    dst[i] = make_float4((float)t.x / 255.999f, (float)t.y / 255.999f, (float)t.z / 255.999f, 0.0f);
  }
}

int rt_launch_texel_expand(const uchar4 *src, size_t n, float4 *dst, cudaStream_t stream) {
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  if (grid < 1) grid = 1;
  rt_texel_expand_kernel<<<grid, 256, 0, stream>>>(src, n, dst);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ scene pack
// Scene upload helper: the per-slot records the kernels read, built on the device from the host's own buffers.
//   tri_pos[slot] = (p0.xyz, e1.x)(e1.yz, e2.xy)(e2.z, 0, 0, 0) with e1 = p1 - p0, e2 = p2 - p0 — the f32 subtractions
//                   the reference redoes for every ray (raytracer.c:116-122), done once (IEEE, so the same bits as on the host);
//   tri_rec[slot] = the first 96 bytes of Triangle_AOS (normal, normal_a/b/c, tangent, bitangent, three UVs: scene.h:46-51)
//                   as six float4, then the material index in place of the Shader pointer pair.
__global__ void rt_scene_pack_kernel(const float *__restrict__ soa, const float4 *__restrict__ aos, const int *__restrict__ mat_index,
                                     int n_slots, float4 *__restrict__ tri_pos, float4 *__restrict__ tri_rec) {
  const size_t N = (size_t)n_slots;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
    const float p0x = soa[s], p1x = soa[N + s], p2x = soa[2 * N + s];
    const float p0y = soa[3 * N + s], p1y = soa[4 * N + s], p2y = soa[5 * N + s];
    const float p0z = soa[6 * N + s], p1z = soa[7 * N + s], p2z = soa[8 * N + s];
    tri_pos[3 * (size_t)s + 0] = make_float4(p0x, p0y, p0z, __fsub_rn(p1x, p0x));
    tri_pos[3 * (size_t)s + 1] = make_float4(__fsub_rn(p1y, p0y), __fsub_rn(p1z, p0z), __fsub_rn(p2x, p0x), __fsub_rn(p2y, p0y));
    tri_pos[3 * (size_t)s + 2] = make_float4(__fsub_rn(p2z, p0z), 0.0f, 0.0f, 0.0f);
    const float4 *a = aos + 7 * (size_t)s;              // sizeof(Triangle_AOS) == 112 == 7 float4
    #pragma unroll
    for (int k = 0; k < 6; k++) tri_rec[7 * (size_t)s + k] = a[k];
    tri_rec[7 * (size_t)s + 6] = make_float4(__int_as_float(mat_index[s]), 0.0f, 0.0f, 0.0f);
  }
}

int rt_launch_scene_pack(const float *soa, const float4 *aos, const int *mat_index, int n_slots, float4 *tri_pos, float4 *tri_rec,
                         cudaStream_t stream) {
  static_assert(sizeof(Triangle_AOS) == 112, "Triangle_AOS is seven float4 (scene.h)");
  unsigned grid = (unsigned)((n_slots + 255) / 256);
  if (grid > 148u * 8u) grid = 148u * 8u;
  if (grid < 1) grid = 1;
  rt_scene_pack_kernel<<<grid, 256, 0, stream>>>(soa, aos, mat_index, n_slots, tri_pos, tri_rec);
  return (int)cudaGetLastError();
}
