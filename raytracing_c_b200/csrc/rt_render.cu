// rt_render.cu — the path-tracing kernels for sm_100a, wavefront form.
//
// Replaces render_thread_proc's pixel loop and everything under it (reference
// raytracer.c:443-558 traversal + cast_ray, :582-594 jitter hash, :596-720 chunk
// loop, camera, film) — see DESIGN.md for the mapping.
//
// Execution model.  cast_ray's bounce loop (raytracer.c:512-556) is unrolled over
// KERNELS instead of running inside one thread, so every kernel is small enough
// for the instruction cache and runs one kind of work with full warps:
//
//   chunk of S samples x all pixels = N paths, path id = ((tile*S + s)*32 + pixel-in-8x4-tile)
//   trace<primary>   generate camera ray (raytracer.c:644-677), closest hit  -> HIT queue | MISS queue
//   for bounce b = 0 .. max_bounces-1:
//     trace          (b > 0) RAY queue -> closest hit                         -> HIT queue | MISS queue
//     miss           environment * tint + emission (raytracer.c:554)          -> rad[path]
//     shade          back-face pass-through (:516-522) or BSDF (:524-552)     -> RAY queue | rad[path]
//   accumulate       accum[pixel] += rad[path(pixel, s)] in sample order      (raytracer.c:696-700)
//
// Queues are structure-of-float4 arrays in HBM; a record moves WITH its path
// (warp-aggregated atomic append), so every stage reads and writes coalesced
// 16-byte vectors and rays stay compacted.  Queue lengths live on the device: all
// launches of a chunk are enqueued back to back without a host round trip, each
// kernel is a persistent grid that strides over "whatever the previous stage produced".
// Queue ORDER is not deterministic, results are: a path's arithmetic depends only on
// its own record, and the per-pixel sum is taken afterwards in sample order, so the f32
// accumulator equals the sequential CPU loop bit for bit.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdlib>
#include <vector>

#include "rt_stages.cuh"
#include "rt_kernels.h"

// rt_render_fast.cu: the same bounce-stage kernels compiled with FMA contraction and hardware transcendentals
void rt_fast_launch_trace(const StageParams &P, unsigned grid, size_t smem, cudaStream_t stream, bool wide);
void rt_fast_launch_miss(const StageParams &P, unsigned grid, cudaStream_t stream);
void rt_fast_launch_shade(const StageParams &P, unsigned grid, cudaStream_t stream);
int  rt_fast_trace_setup(size_t level_bytes, int *wide_blocks_per_sm);

// ------------------------------------------------------------ camera-relative scene
// All primary rays start at the camera origin o (raytracer.c:612).  Everything in the slab and
// triangle tests that depends on o but not on the direction — (plane - o) per node, and per triangle
// tv = o - p0, qv = tv x e1, e2 . qv (raytracer.c:123-135) — is evaluated here once per render with the
// same f32 expressions the per-ray code uses, so the primary trace reads constants instead.
__global__ void rt_camera_relative_kernel(const SceneDev sc) {
  const float ox = sc.view[0][3], oy = sc.view[1][3], oz = sc.view[2][3];
  const int n_node_floats = sc.n_internal * 48;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_node_floats + sc.n_slots; i += gridDim.x * blockDim.x) {
    if (i < n_node_floats) {
      const int row = (i % 48) / 8;                     // rows: min x y z, max x y z
      const float o = (row % 3 == 0) ? ox : (row % 3 == 1) ? oy : oz;
      sc.nodes_rel[i] = sc.nodes[i] - o;
    } else {
      const int slot = i - n_node_floats;
      const float4 A = sc.tri_pos[3 * slot], B = sc.tri_pos[3 * slot + 1], C = sc.tri_pos[3 * slot + 2];
      const float e1x = A.w, e1y = B.x, e1z = B.y, e2x = B.z, e2y = B.w, e2z = C.x;
      const float tvx = ox - A.x, tvy = oy - A.y, tvz = oz - A.z;
      const float qvx = tvy * e1z - tvz * e1y, qvy = tvz * e1x - tvx * e1z, qvz = tvx * e1y - tvy * e1x;
      sc.tri_rel[4 * slot + 0] = make_float4(tvx, tvy, tvz, e1x);
      sc.tri_rel[4 * slot + 1] = B;
      sc.tri_rel[4 * slot + 2] = make_float4(e2z, qvx, qvy, qvz);
      sc.tri_rel[4 * slot + 3] = make_float4(e2x * qvx + e2y * qvy + e2z * qvz, 0, 0, 0);
    }
  }
}

// ----------------------------------------------------------------- accumulate
// raytracer.c:696-700: color += cast_ray(...) in sample order, one thread per pixel.
__global__ void __launch_bounds__(256)
rt_accumulate_kernel(const __grid_constant__ StageParams P) {
  const unsigned n = P.n_tiles * 32u;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned tile = i >> 5, in_tile = i & 31u;
    int x0, y0;
    tile_origin(P, tile, x0, y0);
    const int px = x0 + (int)(in_tile & 7u), py = y0 + (int)(in_tile >> 3);
    if (px >= P.width || py >= P.height) continue;
    const int pixel = py * P.width + px;
    V3 sum = P.accumulate ? mk3(P.accum[3 * pixel], P.accum[3 * pixel + 1], P.accum[3 * pixel + 2]) : mk3(0, 0, 0);
#if RT_PATH_LAYOUT == 1
    const float4 *r = P.q.rad + (size_t)i * (size_t)P.n_samples;
    const size_t r_step = 1;
#else
    const float4 *r = P.q.rad + ((size_t)tile * (size_t)P.n_samples) * 32 + in_tile;
    const size_t r_step = 32;
#endif
    for (int s = 0; s < P.n_samples; s++) {
      const float4 v = r[(size_t)s * r_step];
      sum = add3(sum, mk3(v.x, v.y, v.z));
      if (P.per_sample) {
        float *dst = P.per_sample + ((size_t)pixel * (size_t)P.per_sample_stride + (size_t)(P.per_sample_offset + s)) * 3;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z;
      }
    }
    P.accum[3 * pixel + 0] = sum.x;
    P.accum[3 * pixel + 1] = sum.y;
    P.accum[3 * pixel + 2] = sum.z;
  }
}

// ----------------------------------------------------------------------- film
// raytracer.c:700-716 + common.h:90-92: /spp, clamp, sRGB OETF, *255.999, truncate.
__device__ __forceinline__ unsigned char film_u8(float sum, float inv_samples) {
  float v = sum * inv_samples;
  if (v != v) v = 0;
  v = clamp1(v, 0, 1);
  v = (v <= 0.0031308f) ? (12.92f * v) : (1.055f * rt_powf(v, 1.0f / 2.4f) - 0.055f);
  v = v * 255.999f;
  return (unsigned char)v;
}

__global__ void rt_resolve_kernel(const float *__restrict__ accum, int width, int height, float inv_samples,
                                  unsigned char *__restrict__ pixels, int stride, int components) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width * height) return;
  int x = i % width, y = i / width;
  unsigned char *dst = pixels + (size_t)components * (size_t)(x + y * stride);
  #pragma unroll
  for (int c = 0; c < 3; c++) dst[c] = film_u8(accum[3 * i + c], inv_samples);
}

// Multi-GPU film: the per-device accumulators are combined and resolved by ONE kernel on the device that owns
// the image.  parts[k] are device pointers — this device's own buffer and its peers' buffers mapped over NVLink
// (cudaDeviceEnablePeerAccess inside one process, cudaIpcOpenMemHandle across processes) — read with 16-byte loads
// and summed in fixed rank order ((p0 + p1) + p2 ...), so the frame does not depend on which GPU finished first.
// Replaces an ncclReduce into a staging buffer followed by rt_resolve_kernel: no round trip through HBM, no second launch.
// One thread = four pixels = three float4 per part.
__global__ void __launch_bounds__(256)
rt_reduce_resolve_kernel(const ReduceParts parts, float *__restrict__ sum_out, int width, int height, float inv_samples,
                         unsigned char *__restrict__ pixels, int stride, int components) {
  const int n_pixels = width * height;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int first = g * 4;
  if (first >= n_pixels) return;
  float v[12];
  const int n_here = n_pixels - first < 4 ? n_pixels - first : 4;
  if (n_here == 4) {
    #pragma unroll
    for (int q = 0; q < 3; q++) {
      // every part's vector is requested before the first add (a peer's accumulator is a ~2 us round trip over NVLink:
      // the loop over parts, left rolled, waited for each in turn); the sum itself stays in rank order
      float4 t[RT_MAX_PARTS];
      #pragma unroll
      for (int k = 0; k < RT_MAX_PARTS; k++)
        if (k < parts.n) t[k] = reinterpret_cast<const float4 *>(parts.part[k])[3 * g + q];
      float4 s = t[0];
      #pragma unroll
      for (int k = 1; k < RT_MAX_PARTS; k++)
        if (k < parts.n) { s.x += t[k].x; s.y += t[k].y; s.z += t[k].z; s.w += t[k].w; }
      v[4 * q] = s.x; v[4 * q + 1] = s.y; v[4 * q + 2] = s.z; v[4 * q + 3] = s.w;
      if (sum_out) reinterpret_cast<float4 *>(sum_out)[3 * g + q] = s;
    }
  } else {
    for (int j = 0; j < 3 * n_here; j++) {
      float s = parts.part[0][3 * first + j];
      for (int k = 1; k < parts.n; k++) s += parts.part[k][3 * first + j];
      v[j] = s;
      if (sum_out) sum_out[3 * first + j] = s;
    }
  }
  if (!pixels) return;
  for (int j = 0; j < n_here; j++) {
    const int i = first + j, x = i % width, y = i / width;
    unsigned char *dst = pixels + (size_t)components * (size_t)(x + y * stride);
    #pragma unroll
    for (int c = 0; c < 3; c++) dst[c] = film_u8(v[3 * j + c], inv_samples);
  }
}

// ------------------------------------------------------------------- launchers
#define RT_PATH_BYTES ((2 + 3 + 1 + 2) * sizeof(float4))       // ray + hit + miss records and the tint / emission(=rad) state of one path
#define RT_CHUNK_PATHS_MAX (512u << 20)

static size_t counts_bytes(int max_bounces) {
  size_t b = (size_t)(max_bounces + 1) * Q_STRIDE * sizeof(unsigned);
  return (b + 255) & ~(size_t)255;
}

// Paths per wavefront chunk.  Every chunk pays a fixed cost: the late bounces hold a few 10^4..10^5 rays and their
// kernels last as long as their slowest warp whatever the chunk's size.  Measured on helmet 1080p x 1024 spp with the
// final kernels of round 2 (128 B per path): 7.00 / 7.16 / 7.25 Gsamples/s at 128 / 256 / 512 Mi paths.  Default
// 512 Mi paths = 69 GB of the 180 GB on the device (a 1080p frame: 256 samples of every pixel per chunk); a device
// that cannot spare it gets smaller chunks (rt_cabi.cu halves the request until it fits).  Path ids are 32-bit: 2^31
// is the ceiling.  RT_GPU_CHUNK_PATHS overrides it (tuning / tests).
static size_t chunk_paths_max() {
  if (const char *e = getenv("RT_GPU_CHUNK_PATHS")) {
    unsigned long long v = strtoull(e, nullptr, 10);
    if (v > (1ull << 31)) v = 1ull << 31;          // path ids and queue lengths are 32-bit
    if (v > 0) return (size_t)v;
  }
  return RT_CHUNK_PATHS_MAX;
}

// samples of every pixel per chunk
static size_t chunk_samples(size_t per_sample, int n_samples, int slice_samples) {
  size_t s = chunk_paths_max() / per_sample;
  if (s < 1) s = 1;
  if (slice_samples > 0 && s > (size_t)slice_samples) s = (size_t)slice_samples;
  if (n_samples > 0 && s > (size_t)n_samples) s = (size_t)n_samples;
  return s;
}

// job tiles (8x4 pixels) one render covers: the whole image, or the 32x32 chunks congruent to split_rank
unsigned rt_render_tiles(int width, int height, int split_rank, int split_world) {
  if (split_world <= 1) return (unsigned)((width + 7) / 8) * (unsigned)((height + 3) / 4);
  const long chunks = (long)((width + 31) / 32) * (long)((height + 31) / 32);
  const long owned = chunks > split_rank ? (chunks - split_rank + split_world - 1) / split_world : 0;
  return (unsigned)owned * 32u;
}

// samples of every pixel one wavefront chunk holds at the preferred queue size
int rt_render_chunk_samples(int width, int height, int split_world) {
  size_t per_sample = (size_t)rt_render_tiles(width, height, 0, split_world) * 32;
  if (per_sample < 32) per_sample = 32;
  return (int)chunk_samples(per_sample, 0, 0);
}

size_t rt_render_workspace_bytes(int width, int height, int n_samples, int max_bounces, int slice_samples, int split_world) {
  size_t per_sample = (size_t)rt_render_tiles(width, height, 0, split_world) * 32;
  if (per_sample < 32) per_sample = 32;
  return counts_bytes(max_bounces > 0 ? max_bounces : 1) +
         chunk_samples(per_sample, n_samples, slice_samples) * per_sample * RT_PATH_BYTES;
}

// occupancy of the trace kernel and its dynamic-shared-memory opt-in, per device (cudaFuncSetAttribute is per device)
static int g_trace_blocks_per_sm[RT_MAX_DEVICES], g_wide_blocks_per_sm[RT_MAX_DEVICES], g_primary_blocks_per_sm[RT_MAX_DEVICES], g_fast_blocks_per_sm[RT_MAX_DEVICES], g_fast_wide_blocks_per_sm[RT_MAX_DEVICES];
static size_t g_level_bytes[RT_MAX_DEVICES];

// ---- optional per-stage timing (bench.py's roofline leg): CUDA events around every launch, on the
// launching stream; read back and summed by rt_stage_profile_read
struct StageEvent { cudaEvent_t a, b; int stage; };
static bool g_profile = false;
static int  g_profile_device = 0;        // launches on other devices of a multi-device frame are not timed
static std::vector<StageEvent> g_stage_events;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t pooled_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

struct StageTimer {
  cudaStream_t st; bool on; StageEvent ev;
  StageTimer(int stage, cudaStream_t s, int device) : st(s), on(g_profile && device == g_profile_device) {
    if (on) { ev.a = pooled_event(); ev.b = pooled_event(); ev.stage = stage; cudaEventRecord(ev.a, st); }
  }
  ~StageTimer() { if (on) { cudaEventRecord(ev.b, st); g_stage_events.push_back(ev); } }
};

void rt_stage_profile_enable(int on, int device) {
  g_profile = on != 0;
  g_profile_device = device;
  for (StageEvent &e : g_stage_events) { g_event_pool.push_back(e.a); g_event_pool.push_back(e.b); }
  g_stage_events.clear();
}

int rt_stage_profile_read(double ms[RT_N_STAGES * RT_STAGE_BOUNCES], long long launches[RT_N_STAGES * RT_STAGE_BOUNCES]) {
  for (int i = 0; i < RT_N_STAGES * RT_STAGE_BOUNCES; i++) { ms[i] = 0; launches[i] = 0; }
  for (StageEvent &e : g_stage_events) {
    cudaError_t err = cudaEventSynchronize(e.b);
    if (err != cudaSuccess) return (int)err;
    float t = 0;
    cudaEventElapsedTime(&t, e.a, e.b);
    ms[e.stage] += t;
    launches[e.stage] += 1;
    g_event_pool.push_back(e.a); g_event_pool.push_back(e.b);
  }
  g_stage_events.clear();
  return 0;
}

static void bind_queues(PathQueues &q, char *w, size_t cb, size_t cap) {
  q.counts = reinterpret_cast<unsigned *>(w); w += cb;
  float4 **arrays[] = { &q.ray_a, &q.ray_b, &q.hit_a, &q.hit_b, &q.hit_h, &q.miss_a, &q.tint, &q.emis };
  for (float4 **a : arrays) { *a = reinterpret_cast<float4 *>(w); w += cap * sizeof(float4); }
  q.rad = q.emis;
}

// Dynamic shared memory opt-in and occupancy of the trace kernels on the current device (once per device and depth).
static int trace_launch_setup(const SceneDev &scene, int *dev_out, size_t *level_bytes_out) {
  // level store: [depth - 1][2][RT_BLOCK] float4 of dynamic shared memory (see rt_trace.cuh)
  // (levels 2 .. depth only: 24 KB per block for the depth-4 trees of the shipped models; the driver's default carve-out is the
  // smallest that holds the resident blocks — forcing a larger one costs up to 6 %, L1 beyond 96 KB gains nothing: measured)
  const size_t level_bytes = (size_t)(scene.depth > 1 ? scene.depth - 1 : 1) * 2 * RT_BLOCK * sizeof(float4);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= RT_MAX_DEVICES) return (int)cudaErrorInvalidDevice;
  if (g_trace_blocks_per_sm[dev] == 0 || level_bytes != g_level_bytes[dev]) {
    int n = 0;
    cudaFuncSetAttribute(rt_trace_kernel<true>,  cudaFuncAttributeMaxDynamicSharedMemorySize, RT_MAX_DEPTH * 2 * RT_BLOCK * 16);
    cudaFuncSetAttribute(rt_trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_MAX_DEPTH * 2 * RT_BLOCK * 16);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rt_trace_kernel<false>, RT_BLOCK, level_bytes) != cudaSuccess || n < 1) n = 1;
    g_trace_blocks_per_sm[dev] = n;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rt_trace_kernel<true>, RT_BLOCK, level_bytes) != cudaSuccess || n < 1) n = 1;
    g_primary_blocks_per_sm[dev] = n;
    cudaFuncSetAttribute(rt_trace_kernel<false, RT_TRACE_MIN_BLOCKS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_MAX_DEPTH * 2 * RT_BLOCK * 16);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rt_trace_kernel<false, RT_TRACE_MIN_BLOCKS_WIDE>, RT_BLOCK, level_bytes) != cudaSuccess || n < 1) n = 1;
    g_wide_blocks_per_sm[dev] = n;
    g_fast_blocks_per_sm[dev] = rt_fast_trace_setup(level_bytes, &g_fast_wide_blocks_per_sm[dev]);
    g_level_bytes[dev] = level_bytes;
  }
  *dev_out = dev;
  *level_bytes_out = level_bytes;
  return 0;
}

int rt_launch_render(const RenderParams &p, int sm_count, void *workspace, size_t workspace_bytes,
                     cudaStream_t stream, int *n_launches) {
  const int n_total = p.sample_end - p.sample_begin;
  if (n_total <= 0 || p.max_bounces < 1) {
    // no samples, or cast_ray's loop body never runs (raytracer.c:512,557): the sum gains nothing
    if (!p.accumulate) cudaMemsetAsync(p.accum, 0, (size_t)p.width * p.height * 3 * sizeof(float), stream);
    return (int)cudaGetLastError();
  }
  size_t level_bytes = 0;
  int dev = 0;
  if (int e = trace_launch_setup(p.scene, &dev, &level_bytes)) return e;

  StageParams P{};
  P.scene = p.scene;
  P.width = p.width; P.height = p.height;
  P.tiles_x = (p.width + 7) / 8; P.tiles_y = (p.height + 3) / 4;
  P.split_rank = p.split_rank; P.split_world = p.split_world > 1 ? p.split_world : 1;
  P.chunks_x = (p.width + 31) / 32;
  P.n_tiles = rt_render_tiles(p.width, p.height, P.split_rank, P.split_world);
  // a path id is at most n_paths (32-bit), so a tile id is below 2^27 and a chunk id below tiles/32 * world + world
  P.div_tiles_x = rt_fastdiv_make((uint32_t)P.tiles_x, 0xffffffffu >> 5);
  P.div_chunks_x = rt_fastdiv_make((uint32_t)P.chunks_x, 0xffffffffu);
  P.max_bounces = p.max_bounces;
  P.user_seed = p.user_seed;
  P.accum = p.accum;
  P.per_sample = p.per_sample;
  P.per_sample_stride = n_total;
  P.counters = p.counters;
  P.counters_ex = p.counters_ex;

  // chunking: as many samples of every pixel per chunk as the workspace holds paths
  const size_t per_sample = (size_t)P.n_tiles * 32;
  const size_t cb = counts_bytes(p.max_bounces);
  if (P.split_world > 1 && !p.accumulate)      // the pixels of other ranks stay zero (the cross-rank sum adds them)
    cudaMemsetAsync(p.accum, 0, (size_t)p.width * p.height * 3 * sizeof(float), stream);
  if (per_sample == 0) return (int)cudaGetLastError();
  if (workspace_bytes < cb + per_sample * RT_PATH_BYTES) return (int)cudaErrorMemoryAllocation;
  size_t cap = (workspace_bytes - cb) / RT_PATH_BYTES;
  size_t fit = chunk_samples(per_sample, n_total, 0);
  if (fit > cap / per_sample) fit = cap / per_sample;
  const int chunk = (int)fit;
  cap = (size_t)chunk * per_sample;
  bind_queues(P.q, static_cast<char *>(workspace), cb, cap);

  rt_camera_relative_kernel<<<(unsigned)sm_count, 256, 0, stream>>>(p.scene);
  const unsigned trace_grid = (unsigned)(sm_count * g_trace_blocks_per_sm[dev]);     // persistent: one wave
  const unsigned wide_grid = (unsigned)(sm_count * g_wide_blocks_per_sm[dev]);
  const unsigned primary_grid = (unsigned)(sm_count * g_primary_blocks_per_sm[dev]);
  const unsigned fast_grid = (unsigned)(sm_count * g_fast_blocks_per_sm[dev]);
  const unsigned fast_wide_grid = (unsigned)(sm_count * g_fast_wide_blocks_per_sm[dev]);
  // grid-stride kernels (miss, shade, accumulate): blocks per SM by queue length.  The items of a queue cost unevenly
  // (texture taps, lobe picks), so the long queues of the first bounces want many short-lived blocks to even out the
  // last wave (bounce-0 miss 2073 / 1950 / 1889 / 1864 / 1856 us at 8 / 16 / 32 / 64 / 128 blocks per SM), the short late
  // queues want few (every block builds its texel table first: bounce-7 shade 23 / 31 / 41 us at 8 / 64 / 128).
  const unsigned flat_grid  = (unsigned)(sm_count * 8);
  auto miss_grid  = [&](int b) { return (unsigned)sm_count * (b == 0 ? 64u : b == 1 ? 32u : 8u); };
  auto shade_grid = [&](int b) { return (unsigned)sm_count * (b == 0 ? 64u : 8u); };
  int launches = 1;
  for (int s0 = p.sample_begin; s0 < p.sample_end; s0 += chunk) {
    const int S = (p.sample_end - s0 < chunk) ? p.sample_end - s0 : chunk;
    P.sample0 = s0; P.n_samples = S;
    P.n_paths = (unsigned)(per_sample * (size_t)S);
    P.div_samples = rt_fastdiv_make((uint32_t)S, P.n_paths);
    P.accumulate = (p.accumulate || s0 > p.sample_begin || P.split_world > 1) ? 1 : 0;
    P.per_sample_offset = s0 - p.sample_begin;
    P.hit_ids = (s0 == p.sample_begin) ? p.hit_ids : nullptr;
    cudaMemsetAsync(P.q.counts, 0, cb, stream);
    P.bounce = 0;
    { StageTimer t(RT_STAGE_TRACE * RT_STAGE_BOUNCES, stream, dev); rt_trace_kernel<true><<<primary_grid, RT_BLOCK, level_bytes, stream>>>(P); }
    launches++;
    // textures and the environment may still be in flight on the upload's copy stream: the primary trace reads
    // nodes and triangles only, everything after it waits here (rt_scene.cu)
    if (s0 == p.sample_begin && p.shading_ready) cudaStreamWaitEvent(stream, p.shading_ready, 0);
    for (int b = 0; b < p.max_bounces; b++) {
      P.bounce = b;
      const int bslot = b < RT_STAGE_BOUNCES ? b : RT_STAGE_BOUNCES - 1;
      if (b > 0) {
        StageTimer t(RT_STAGE_TRACE * RT_STAGE_BOUNCES + bslot, stream, dev);
        if (p.fast) rt_fast_launch_trace(P, b <= RT_TRACE_WIDE_BOUNCES ? fast_wide_grid : fast_grid, level_bytes, stream, b <= RT_TRACE_WIDE_BOUNCES);
        else if (b <= RT_TRACE_WIDE_BOUNCES) rt_trace_kernel<false, RT_TRACE_MIN_BLOCKS_WIDE><<<wide_grid, RT_BLOCK, level_bytes, stream>>>(P);
        else        rt_trace_kernel<false><<<trace_grid, RT_BLOCK, level_bytes, stream>>>(P);
        launches++;
      }
      {
        StageTimer t(RT_STAGE_MISS * RT_STAGE_BOUNCES + bslot, stream, dev);
        if (p.fast) rt_fast_launch_miss(P, miss_grid(b), stream);
        else        rt_miss_kernel<<<miss_grid(b), 256, 0, stream>>>(P);
      }
      {
        StageTimer t(RT_STAGE_SHADE * RT_STAGE_BOUNCES + bslot, stream, dev);
        if (p.fast) rt_fast_launch_shade(P, shade_grid(b), stream);
        else        rt_shade_kernel<<<shade_grid(b), 256, 0, stream>>>(P);
      }
      launches += 2;
    }
    { StageTimer t(RT_STAGE_ACCUMULATE * RT_STAGE_BOUNCES, stream, dev); rt_accumulate_kernel<<<flat_grid * 4, 256, 0, stream>>>(P); }
    launches++;
  }
  if (n_launches) *n_launches += launches;
  return (int)cudaGetLastError();
}

// ===================================================================== lightmap_bake
// Reference raytracer.c:722-784: every triangle, in slot order, rasterises its UV footprint into the lightmap; each
// covered texel gets `samples` cosine-weighted cast_ray estimates from the interpolated surface point; a texel covered by
// several triangles keeps the LAST one's value.  Here: (1) an ownership pass finds that last triangle per texel
// (atomicMax over slots — same coverage arithmetic, f32, no contraction), (2) owned texels are compacted into jobs,
// (3) every (job, sample) is a path of the same wavefront as the renderer — its first ray comes from the texel instead
// of the camera — and (4) the cosine-weighted sum is taken per job in sample order.  Per-(texel, sample) seeds
// (rt_seed.h) replace the reference's two sequential generator streams, as for the renderer.
struct LightmapParams {
  int      width, height, samples;
  int     *owner;          // [W*H] slot of the last covering triangle, -1 = none
  unsigned *jobs;          // [n_jobs] texel index
  unsigned *n_jobs;
  float   *sums;           // [n_jobs][3] running cosine-weighted sums
  float   *cosw;           // [path] cosine of the sample's first direction (0: no direction found)
};

// barycentric weights of texel (x, y) in UV triangle `slot` exactly as raytracer.c:733-745 computes them
__device__ __forceinline__ bool lightmap_weights(const SceneDev &sc, int slot, float W, float H, float px, float py,
                                                 float &w0, float &w1, float &w2) {
  const float4 r4 = __ldg(sc.tri_rec + (size_t)slot * 7 + 4), r5 = __ldg(sc.tri_rec + (size_t)slot * 7 + 5);
  const float p0x = r4.z * W, p0y = r4.w * H, p1x = r5.x * W, p1y = r5.y * H, p2x = r5.z * W, p2y = r5.w * H;
  const float denom = (p1y - p2y) * (p0x - p2x) + (p2x - p1x) * (p0y - p2y);
  w0 = ((p1y - p2y) * (px - p2x) + (p2x - p1x) * (py - p2y)) / denom;
  w1 = ((p2y - p0y) * (px - p2x) + (p0x - p2x) * (py - p2y)) / denom;
  w2 = 1.0f - w0 - w1;
  return w0 >= -RT_EPS && w1 >= -RT_EPS && w2 >= -RT_EPS;
}

// one block per triangle slot: the texels of its UV bounding box (raytracer.c:728-731: float products truncated to i32)
__global__ void __launch_bounds__(128)
rt_lightmap_owner_kernel(const SceneDev sc, const LightmapParams L) {
  const int slot = blockIdx.x;
  const float W = (float)L.width, H = (float)L.height;
  const float4 r4 = __ldg(sc.tri_rec + (size_t)slot * 7 + 4), r5 = __ldg(sc.tri_rec + (size_t)slot * 7 + 5);
  const float ax = r4.z, ay = r4.w, bx = r5.x, by = r5.y, cx = r5.z, cy = r5.w;
  const int min_x = (int)(sel_min(ax, sel_min(bx, cx)) * W), max_x = (int)(sel_max(ax, sel_max(bx, cx)) * W);
  const int min_y = (int)(sel_min(ay, sel_min(by, cy)) * H), max_y = (int)(sel_max(ay, sel_max(by, cy)) * H);
  // texels outside the lightmap are not stored (the reference writes out of bounds): clip the box
  const int x0 = min_x < 0 ? 0 : min_x, x1 = max_x >= L.width ? L.width - 1 : max_x;
  const int y0 = min_y < 0 ? 0 : min_y, y1 = max_y >= L.height ? L.height - 1 : max_y;
  if (x1 < x0 || y1 < y0) return;
  const int bw = x1 - x0 + 1, n = bw * (y1 - y0 + 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int x = x0 + i % bw, y = y0 + i / bw;
    float w0, w1, w2;
    if (lightmap_weights(sc, slot, W, H, (float)x, (float)y, w0, w1, w2)) atomicMax(&L.owner[y * L.width + x], slot);
  }
}

__global__ void __launch_bounds__(256)
rt_lightmap_jobs_kernel(const LightmapParams L) {
  const unsigned n = (unsigned)(L.width * L.height), lane = threadIdx.x & 31u;
  const unsigned n_round = (n + 31u) & ~31u;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool owned = i < n && L.owner[i] >= 0;
    const unsigned pos = warp_append(L.n_jobs, owned, lane);
    if (owned) L.jobs[pos] = i;
  }
}

// common.h:30-42 with the generator state in a register; the final scale is a DOUBLE division narrowed to f32
__device__ __forceinline__ V3 lightmap_rand_vec3(uint32_t &state) {
  for (;;) {
    V3 p;
    p.x = rt_rand_f32(&state) * 2.0f + -1.0f;
    p.y = rt_rand_f32(&state) * 2.0f + -1.0f;
    p.z = rt_rand_f32(&state) * 2.0f + -1.0f;
    const float lensq = dot3(p, p);
    if (RT_EPS < lensq && lensq <= 1) return scale3(p, (float)(1.0 / (double)__fsqrt_rn(lensq)));
  }
}

// first ray of path (job, sample): raytracer.c:747-772
__global__ void __launch_bounds__(256)
rt_lightmap_raygen_kernel(const __grid_constant__ StageParams P, const LightmapParams L, unsigned n_jobs) {
  const SceneDev &sc = P.scene;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned n = n_jobs * (unsigned)P.n_samples, n_round = (n + 31u) & ~31u;
  const float W = (float)L.width, H = (float)L.height;
  for (unsigned path = blockIdx.x * blockDim.x + threadIdx.x; path < n_round; path += gridDim.x * blockDim.x) {
    bool shoot = false;
    V3 o = mk3(0, 0, 0), d = mk3(0, 0, 0);
    uint32_t rng = 0;
    if (path < n) {
      const unsigned job = path / (unsigned)P.n_samples;
      const int s = P.sample0 + (int)(path - job * (unsigned)P.n_samples);
      const unsigned texel = L.jobs[job];
      const int slot = L.owner[texel];
      const int x = (int)(texel % (unsigned)L.width), y = (int)(texel / (unsigned)L.width);
      float w0, w1, w2;
      lightmap_weights(sc, slot, W, H, (float)x, (float)y, w0, w1, w2);
      const float *soa = sc.tri_soa + slot;
      const size_t N = (size_t)sc.n_slots;
      const V3 position = mk3(soa[0] * w0 + soa[N] * w1 + soa[2 * N] * w2,
                              soa[3 * N] * w0 + soa[4 * N] * w1 + soa[5 * N] * w2,
                              soa[6 * N] * w0 + soa[7 * N] * w1 + soa[8 * N] * w2);
      const float4 *rec = sc.tri_rec + (size_t)slot * 7;
      const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
      const V3 normal = mk3(r0.w * w0 + r1.z * w1 + r2.y * w2, r1.x * w0 + r1.w * w1 + r2.z * w2, r1.y * w0 + r2.x * w1 + r2.w * w2);
      o = add3(position, scale3(normal, RT_EPS));
      uint32_t dir_state = rt_path_seed(texel, (uint32_t)s, P.user_seed ^ RT_LIGHTMAP_DIR_SALT);
      rng = rt_path_seed(texel, (uint32_t)s, P.user_seed);
      float cosine = 0;
      for (int tries = 0; tries < RT_LIGHTMAP_MAX_TRIES; tries++) {
        d = lightmap_rand_vec3(dir_state);
        cosine = dot3(d, normal);
        if (cosine > 0) break;
        cosine = 0;
      }
      shoot = cosine > 0;
      L.cosw[path] = cosine;
      if (!shoot) P.q.rad[path] = make_float4(0, 0, 0, 0);
    }
    const unsigned pos = warp_append(&P.q.counts[Q_RAYS], shoot, lane);
    if (shoot) {
      P.q.ray_a[pos] = make_float4(o.x, o.y, o.z, d.x);
      P.q.ray_b[pos] = make_float4(d.y, d.z, __uint_as_float(path), __uint_as_float(rng));
    }
  }
}

// raytracer.c:773: accumulated += cast_ray(...) * cos, in sample order
__global__ void __launch_bounds__(256)
rt_lightmap_accumulate_kernel(const __grid_constant__ StageParams P, const LightmapParams L, unsigned n_jobs, int first_chunk) {
  for (unsigned job = blockIdx.x * blockDim.x + threadIdx.x; job < n_jobs; job += gridDim.x * blockDim.x) {
    V3 sum = first_chunk ? mk3(0, 0, 0) : mk3(L.sums[3 * job], L.sums[3 * job + 1], L.sums[3 * job + 2]);
    const size_t base = (size_t)job * (size_t)P.n_samples;
    for (int s = 0; s < P.n_samples; s++) {
      const float c = L.cosw[base + s];
      if (!(c > 0)) continue;
      const float4 v = P.q.rad[base + s];
      sum = add3(sum, scale3(mk3(v.x, v.y, v.z), c));
    }
    L.sums[3 * job] = sum.x; L.sums[3 * job + 1] = sum.y; L.sums[3 * job + 2] = sum.z;
  }
}

// raytracer.c:777-779: the UN-SCALED f32 goes into the u8 (saturating here; C leaves out-of-range undefined)
__global__ void __launch_bounds__(256)
rt_lightmap_store_kernel(const LightmapParams L, unsigned n_jobs, unsigned char *pixels, int stride, int components, float *values) {
  for (unsigned job = blockIdx.x * blockDim.x + threadIdx.x; job < n_jobs; job += gridDim.x * blockDim.x) {
    const unsigned texel = L.jobs[job];
    const int x = (int)(texel % (unsigned)L.width), y = (int)(texel / (unsigned)L.width);
    unsigned char *dst = pixels + (size_t)components * (size_t)(x + y * stride);
    for (int c = 0; c < 3; c++) {
      const float v = L.sums[3 * job + c] / (float)L.samples;
      if (values) values[3 * (size_t)texel + c] = v;
      dst[c] = !(v == v) ? 0 : (v <= 0 ? 0 : (v >= 255 ? 255 : (unsigned char)v));
    }
  }
}

size_t rt_lightmap_workspace_bytes(int width, int height) {
  const size_t n = (size_t)width * (size_t)height;
  return ((n * (sizeof(int) + sizeof(unsigned) + 3 * sizeof(float)) + 256 + 255) & ~(size_t)255);
}

// `scratch` holds owner / jobs / sums (rt_lightmap_workspace_bytes), `workspace` the path queues plus one f32 per path.
int rt_launch_lightmap(const SceneDev &scene, int width, int height, int samples, int max_bounces, uint32_t user_seed,
                       unsigned char *d_pixels, int stride, int components, float *d_values, int *d_owner_out,
                       void *scratch, int sm_count, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                       int *n_launches, unsigned *n_jobs_out) {
  size_t level_bytes = 0;
  int dev = 0;
  if (int e = trace_launch_setup(scene, &dev, &level_bytes)) return e;
  const size_t n_texels = (size_t)width * (size_t)height;
  LightmapParams L{};
  L.width = width; L.height = height; L.samples = samples;
  char *sp = static_cast<char *>(scratch);
  L.n_jobs = reinterpret_cast<unsigned *>(sp); sp += 256;
  L.owner = reinterpret_cast<int *>(sp);       sp += n_texels * sizeof(int);
  L.jobs = reinterpret_cast<unsigned *>(sp);   sp += n_texels * sizeof(unsigned);
  L.sums = reinterpret_cast<float *>(sp);
  cudaMemsetAsync(L.n_jobs, 0, 256, stream);
  cudaMemsetAsync(L.owner, 0xff, n_texels * sizeof(int), stream);
  const unsigned flat_grid = (unsigned)(sm_count * 8);
  rt_lightmap_owner_kernel<<<(unsigned)scene.n_slots, 128, 0, stream>>>(scene, L);
  rt_lightmap_jobs_kernel<<<flat_grid, 256, 0, stream>>>(L);
  int launches = 2;
  unsigned n_jobs = 0;
  cudaMemcpyAsync(&n_jobs, L.n_jobs, sizeof n_jobs, cudaMemcpyDeviceToHost, stream);
  if (cudaError_t e = cudaStreamSynchronize(stream)) return (int)e;
  if (n_jobs_out) *n_jobs_out = n_jobs;
  if (d_owner_out) cudaMemcpyAsync(d_owner_out, L.owner, n_texels * sizeof(int), cudaMemcpyDeviceToDevice, stream);
  if (n_jobs == 0 || samples < 1 || max_bounces < 1) { if (n_launches) *n_launches += launches; return (int)cudaGetLastError(); }

  StageParams P{};
  P.scene = scene;
  P.width = width; P.height = height;
  P.max_bounces = max_bounces;
  P.user_seed = user_seed;
  const size_t cb = counts_bytes(max_bounces);
  const size_t per_path = RT_PATH_BYTES + sizeof(float);
  if (workspace_bytes < cb + (size_t)n_jobs * per_path) return (int)cudaErrorMemoryAllocation;
  size_t cap = (workspace_bytes - cb) / per_path;
  size_t fit = chunk_samples(n_jobs, samples, 0);
  if (fit > cap / n_jobs) fit = cap / n_jobs;
  const int chunk = (int)fit;
  cap = (size_t)chunk * n_jobs;
  bind_queues(P.q, static_cast<char *>(workspace), cb, cap);
  L.cosw = reinterpret_cast<float *>(static_cast<char *>(workspace) + cb + cap * RT_PATH_BYTES);
  const unsigned trace_grid = (unsigned)(sm_count * g_trace_blocks_per_sm[dev]);
  for (int s0 = 0; s0 < samples; s0 += chunk) {
    const int S = samples - s0 < chunk ? samples - s0 : chunk;
    P.sample0 = s0; P.n_samples = S;
    P.n_paths = n_jobs * (unsigned)S;
    cudaMemsetAsync(P.q.counts, 0, cb, stream);
    rt_lightmap_raygen_kernel<<<flat_grid, 256, 0, stream>>>(P, L, n_jobs);
    launches++;
    for (int b = 0; b < max_bounces; b++) {
      P.bounce = b;
      rt_trace_kernel<false><<<trace_grid, RT_BLOCK, level_bytes, stream>>>(P);
      rt_miss_kernel<<<flat_grid, 256, 0, stream>>>(P);
      rt_shade_kernel<<<flat_grid, 256, 0, stream>>>(P);
      launches += 3;
    }
    rt_lightmap_accumulate_kernel<<<flat_grid, 256, 0, stream>>>(P, L, n_jobs, s0 == 0);
    launches++;
  }
  rt_lightmap_store_kernel<<<flat_grid, 256, 0, stream>>>(L, n_jobs, d_pixels, stride, components, d_values);
  launches++;
  if (n_launches) *n_launches += launches;
  return (int)cudaGetLastError();
}

int rt_launch_resolve(const float *accum, int width, int height, int samples, unsigned char *pixels,
                      int stride, int components, cudaStream_t stream) {
  int n = width * height;
  float inv_samples = 1.0f / (float)samples;
  rt_resolve_kernel<<<(n + 255) / 256, 256, 0, stream>>>(accum, width, height, inv_samples, pixels, stride, components);
  return (int)cudaGetLastError();
}

int rt_launch_reduce_resolve(const ReduceParts &parts, float *sum_out, int width, int height, int samples,
                             unsigned char *pixels, int stride, int components, cudaStream_t stream) {
  const int groups = (width * height + 3) / 4;
  const float inv_samples = 1.0f / (float)samples;
  rt_reduce_resolve_kernel<<<(groups + 255) / 256, 256, 0, stream>>>(parts, sum_out, width, height, inv_samples, pixels, stride, components);
  return (int)cudaGetLastError();
}

int rt_render_blocks_per_sm(void) {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < RT_MAX_DEVICES ? g_trace_blocks_per_sm[dev] : 0;
}
