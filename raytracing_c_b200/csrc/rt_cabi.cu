// rt_cabi.cu — the C ABI of libraytracer_gpu.so.
//
// Exports the reference's own entry points (raytracer.h:51-56, denoiser.h:4) plus the additive rt_gpu_* calls
// declared in include/rt_gpu.h.  Host-side work here:
//   * the reference's threading protocol around render_thread_proc (driver.c:793-818): the thread that claims
//     chunk 0 owns the launch;
//   * one frame over 1..N devices of this process: split (rt_multi.cu), per-device wavefront render
//     (rt_render.cu), fused reduce + resolve on the first device, D2H of the u8 image;
//   * per-call status: the reference's entry points return void, rt_gpu_last_status() says whether the last one failed.
// Scene residency is rt_scene.cu.  There is no CPU fallback: every compute call fails if CUDA does.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <thread>

#include "rt_state.h"
#include "rt_gpu_internal.h"

namespace rt {

State g;
std::mutex g_mutex;

namespace {
std::mutex g_error_mutex;
char g_error_shared[512] = "";
thread_local char g_error_copy[512] = "";
std::atomic<int> g_status{0};
}

int fail(const char *fmt, ...) {
  char text[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(text, sizeof text, fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lock(g_error_mutex);
  memcpy(g_error_shared, text, sizeof text);
  return 1;
}

void clear_error() {
  std::lock_guard<std::mutex> lock(g_error_mutex);
  g_error_shared[0] = 0;
}

}  // namespace rt

using namespace rt;

namespace {

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int init_devices_locked(int n, const int *ids) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail("no CUDA device available (%s); libraytracer_gpu has no CPU fallback", cudaGetErrorString(e));
  if (n < 1 || n > RT_MAX_PARTS) return fail("init: between 1 and %d devices, got %d", RT_MAX_PARTS, n);
  std::vector<int> want;
  for (int k = 0; k < n; k++) {
    const int id = ids ? ids[k] : k;
    if (id < 0 || id >= count || id >= RT_MAX_DEVICES) return fail("device %d out of range (%d visible)", id, count);
    for (int prev : want) if (prev == id) return fail("device %d listed twice", id);
    want.push_back(id);
  }
  if (g.ready) {
    bool same = g.devs.size() == want.size();
    for (size_t k = 0; same && k < want.size(); k++) same = g.devs[k].id == want[k];
    if (same) return 0;
    nccl_shutdown();
    ipc_close_all();
    for (Device &d : g.devs) release_device(d);
    g.devs.clear();
    g.ready = false;
    g.peers_enabled = false;
  }
  g.devs.resize(want.size());
  if (want.size() > 1) {
    // a CUDA context takes ~0.9 s to create; the contexts of different devices can be created side by side
    std::vector<std::thread> makers;
    for (int id : want) makers.emplace_back([id] { if (cudaSetDevice(id) == cudaSuccess) cudaFree(nullptr); });
    for (std::thread &t : makers) t.join();
    cudaGetLastError();
  }
  for (size_t k = 0; k < want.size(); k++) {
    Device &d = g.devs[k];
    CUDA_TRY(cudaSetDevice(want[k]));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, want[k]));
    if (prop.major < 10) return fail("device %d is sm_%d%d; this library is built for sm_100a only", want[k], prop.major, prop.minor);
    d.id = want[k];
    d.sm_count = prop.multiProcessorCount;
    d.l2_persist_max = prop.persistingL2CacheMaxSize > 0 ? (size_t)prop.persistingL2CacheMaxSize : 0;
    d.l2_window_max = prop.accessPolicyMaxWindowSize > 0 ? (size_t)prop.accessPolicyMaxWindowSize : 0;
    CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&d.ev0));
    CUDA_TRY(cudaEventCreate(&d.ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&d.busy, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming));
  }
  CUDA_TRY(cudaSetDevice(want[0]));
  g.ready = true;
  return 0;
}

int ensure_init() {
  if (g.ready) return 0;
  const int zero = 0;
  return init_devices_locked(1, &zero);
}

// Path-queue workspace of one device.  The preferred size (up to 26.6 GB) is a throughput choice, not a requirement:
// on a device that cannot spare it the request is halved down to one sample of every pixel per chunk, and the size
// that was refused is remembered so later calls do not walk the same ladder again.
int ensure_workspace(Device &d, size_t want, size_t floor_bytes, cudaStream_t stream) {
  if (d.workspace_bytes >= want) return 0;
  if (d.workspace_denied && want >= d.workspace_denied && d.workspace_bytes >= floor_bytes) return 0;
  CUDA_TRY(cudaStreamSynchronize(stream));      // an earlier launch may still read the old queues
  CUDA_TRY(cudaStreamSynchronize(d.stream));
  size_t ask = want + want / 50;      // 2 % headroom: a slightly larger request later (another split, another slice) fits
  for (;;) {
    if (d.d_workspace) { cudaFree(d.d_workspace); d.d_workspace = nullptr; d.workspace_bytes = 0; }
    if (cudaMalloc(&d.d_workspace, ask) == cudaSuccess) { d.workspace_bytes = ask; break; }
    cudaGetLastError();                          // clear the allocation failure
    d.d_workspace = nullptr;
    d.workspace_denied = ask;
    if (ask <= floor_bytes) return fail("render: cannot allocate %zu bytes of path queues", ask);
    ask = ask / 2 > floor_bytes ? ask / 2 : floor_bytes;
  }
  return 0;
}

struct Share {            // what one device (or process) renders of a frame
  isize s_begin, s_end;
  int   split_rank, split_world;
};

// One render of `share` on device d, enqueued on `stream`.  Every render of a device uses the device's one workspace
// and rewrites the scene's camera-relative copies, so renders are serialised on the device whatever stream they come
// from: each waits for the event the previous one recorded.
int render_on_device(Device &d, const Scene *scene, isize width, isize height, const Share &share, isize max_bounces,
                     u32 seed, int accumulate, float *d_accum, float *d_per_sample, int *d_hit_ids,
                     unsigned long long *d_counters, cudaStream_t stream, int slice_cap = 0) {
  if (width < 1 || height < 1) return fail("render: empty image");
  CUDA_TRY(cudaSetDevice(d.id));
  auto it = d.scenes.find(scene);
  if (it == d.scenes.end()) return fail("render: scene is not resident on device %d", d.id);
  const DeviceScene &ds = it->second;
  RenderParams p{};
  p.scene = ds.dev;
  const int n_samples = (int)(share.s_end - share.s_begin);
  const size_t want = rt_render_workspace_bytes((int)width, (int)height, n_samples, (int)max_bounces, slice_cap, share.split_world);
  const size_t floor_bytes = rt_render_workspace_bytes((int)width, (int)height, 1, (int)max_bounces, 1, share.split_world);
  if (n_samples > 0 && ensure_workspace(d, want, floor_bytes, stream)) return 1;
  if (d.busy_valid) CUDA_TRY(cudaStreamWaitEvent(stream, d.busy, 0));
  CUDA_TRY(cudaStreamWaitEvent(stream, ds.geom_ready, 0));
  p.shading_ready = ds.tex_ready;
  p.fast = g.options.fast_math != 0;
  set_l2_window(d, ds, stream);
  p.width = (int)width; p.height = (int)height;
  p.split_rank = share.split_rank; p.split_world = share.split_world;
  p.sample_begin = (int)share.s_begin; p.sample_end = (int)share.s_end; p.max_bounces = (int)max_bounces;
  p.user_seed = seed;
  p.accumulate = accumulate;
  p.accum = d_accum; p.per_sample = d_per_sample; p.hit_ids = d_hit_ids;
  p.counters = d_counters;
  p.counters_ex = (d_counters && d_counters == d.d_counters) ? d_counters + 8 : nullptr;   // the library's own buffer has 16 slots
  int e = rt_launch_render(p, d.sm_count, d.d_workspace, d.workspace_bytes, stream, &g.last_launches);
  if (e) return fail("render kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  CUDA_TRY(cudaEventRecord(d.busy, stream));
  d.busy_valid = true;
  return 0;
}

int ensure_frame_buffers(Device &d, size_t n_pixels, bool hit_ids) {
  CUDA_TRY(cudaSetDevice(d.id));
  size_t have = d.accum_floats * sizeof(float);
  if (grow(reinterpret_cast<void **>(&d.d_accum), &have, n_pixels * 3 * sizeof(float))) return 1;
  d.accum_floats = have / sizeof(float);
  if (hit_ids) {
    have = d.hit_pixels * sizeof(int);
    if (grow(reinterpret_cast<void **>(&d.d_hit_ids), &have, n_pixels * sizeof(int))) return 1;
    d.hit_pixels = have / sizeof(int);
  }
  if (!d.d_counters) CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d.d_counters), 16 * sizeof(unsigned long long)));
  return 0;
}

}  // namespace

// ======================================================================= ABI
extern "C" {

int rt_gpu_init(int device) {
  std::lock_guard<std::mutex> lock(g_mutex);
  return init_devices_locked(1, &device);
}

int rt_gpu_init_devices(int n_devices, int const *devices) {
  std::lock_guard<std::mutex> lock(g_mutex);
  return init_devices_locked(n_devices, devices);
}

int rt_gpu_device_count(void) { return g.ready ? (int)g.devs.size() : 0; }

int rt_gpu_visible_devices(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

void rt_gpu_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_mutex);
  nccl_shutdown();
  ipc_close_all();
  jpeg_shutdown();
  for (Device &d : g.devs) release_device(d);
  std::vector<Shader_Proc> pbr = g.pbr_procs;
  std::vector<Background_Proc> bg = g.bg_procs;
  g = State{};
  g.pbr_procs = pbr;
  g.bg_procs = bg;
}

char const *rt_gpu_last_error(void) {
  std::lock_guard<std::mutex> lock(g_error_mutex);
  memcpy(g_error_copy, g_error_shared, sizeof g_error_copy);
  return g_error_copy;
}

int rt_gpu_last_status(void) { return g_status.load(); }

int rt_gpu_sm_count(void) {
  std::lock_guard<std::mutex> lock(g_mutex);
  return ensure_init() ? 0 : g.devs[0].sm_count;
}

f64 rt_gpu_measure_fp32_issue(void) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (ensure_init()) return 0;
  cudaSetDevice(g.devs[0].id);
  return rt_measure_fp32_issue(g.devs[0].sm_count, g.devs[0].stream, nullptr);
}

isize rt_gpu_scene_device_bytes(Scene const *scene) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g.devs.empty()) return 0;
  auto it = g.devs[0].scenes.find(scene);
  return it == g.devs[0].scenes.end() ? 0 : (isize)it->second.bytes;
}

isize rt_gpu_scene_upload_bytes(Scene const *scene) {
  std::lock_guard<std::mutex> lock(g_mutex);
  isize total = 0;
  for (Device &d : g.devs) {
    auto it = d.scenes.find(scene);
    if (it != d.scenes.end()) total += (isize)it->second.h2d_bytes;
  }
  return total;
}

void rt_gpu_register_pbr_shader(Shader_Proc proc) {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (Shader_Proc p : g.pbr_procs) if (p == proc) return;
  g.pbr_procs.push_back(proc);
}

void rt_gpu_register_background(Background_Proc proc) {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (Background_Proc p : g.bg_procs) if (p == proc) return;
  g.bg_procs.push_back(proc);
}

void rt_gpu_pbr_shader_proc(rawptr, Shader_Input const *, Shader_Output *) {
  fprintf(stderr, "rt_gpu_pbr_shader_proc called on the CPU: this build has no CPU shading path\n");
  abort();
}

Color3 rt_gpu_background_proc(rawptr, Vec3) {
  fprintf(stderr, "rt_gpu_background_proc called on the CPU: this build has no CPU shading path\n");
  abort();
}

int rt_gpu_scene_upload(Scene const *scene) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (ensure_init()) return 1;
  if (g.devs.size() > 1 && enable_peers()) return 1;
  return scene_upload_all(scene);
}

void rt_gpu_scene_release(Scene const *scene) {
  std::lock_guard<std::mutex> lock(g_mutex);
  scene_release_all(scene);
}

void *rt_gpu_host_alloc(size_t bytes) {
  void *p = nullptr;
  // portable: pinned for every device this process drives, not only the current one
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    fail("rt_gpu_host_alloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}

void rt_gpu_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// Allocates what a frame of this size will need (accumulators, image, path-queue workspace on every device) ahead of
// the first render, so the allocations (tens of ms for a 69 GB workspace) can overlap the host's model load.
int rt_gpu_prepare_frame(isize width, isize height, isize samples, isize max_bounces) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (ensure_init()) return 1;
  if (width < 1 || height < 1 || samples < 1 || max_bounces < 1) return fail("prepare_frame: empty frame");
  const size_t n_dev = g.devs.size();
  if (n_dev > 1 && enable_peers()) return 1;
  const i32 mode = n_dev > 1 ? rt_gpu_shard_mode((i32)samples, (i32)n_dev, g.options.split_mode) : RT_GPU_SPLIT_SAMPLES;
  const int world = n_dev > 1 && mode == RT_GPU_SPLIT_CHUNKS ? (int)n_dev : 1;
  isize share = samples;
  if (n_dev > 1 && mode == RT_GPU_SPLIT_SAMPLES) share = (samples + (isize)n_dev - 1) / (isize)n_dev;
  isize slice = g.options.slice_samples > 0 ? g.options.slice_samples : rt_render_chunk_samples((int)width, (int)height, world);
  if (slice > share) slice = share;
  for (Device &d : g.devs) {
    if (ensure_frame_buffers(d, (size_t)width * (size_t)height, g.options.keep_hit_ids != 0)) return 1;
    const size_t want = rt_render_workspace_bytes((int)width, (int)height, (int)slice, (int)max_bounces, 0, world);
    const size_t floor_bytes = rt_render_workspace_bytes((int)width, (int)height, 1, (int)max_bounces, 1, world);
    if (ensure_workspace(d, want, floor_bytes, d.stream)) return 1;
  }
  if (n_dev > 1 && g.options.reduce_mode == RT_GPU_REDUCE_NCCL && nccl_prepare()) return 1;
  CUDA_TRY(cudaSetDevice(g.devs[0].id));
  return 0;
}

void rt_gpu_set_options(RT_GPU_Options const *options) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g.options = *options;
  if (g.options.slice_samples < 0) g.options.slice_samples = 0;
  // the communicator costs seconds to create: when the NCCL film is asked for, do it here, not inside the first frame
  if (g.ready && g.devs.size() > 1 && g.options.reduce_mode == RT_GPU_REDUCE_NCCL) nccl_prepare();
}

void rt_gpu_get_options(RT_GPU_Options *options) {
  std::lock_guard<std::mutex> lock(g_mutex);
  *options = g.options;
}

// ------------------------------------------------------------ device-pointer level
int rt_gpu_render_accum_device(Scene const *scene, isize width, isize height, isize sample_begin, isize sample_end,
                               isize max_bounces, u32 user_seed, i32 accumulate, f32 *d_accum, f32 *d_per_sample,
                               i32 *d_hit_ids, u64 *d_counters, void *stream) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (ensure_init()) return 1;
  if (scene_on_devices(scene)) return 1;
  const Share share{sample_begin, sample_end, g.options.pixel_rank, g.options.pixel_world};
  return render_on_device(g.devs[0], scene, width, height, share, max_bounces, user_seed, accumulate, d_accum, d_per_sample,
                          d_hit_ids, reinterpret_cast<unsigned long long *>(d_counters), static_cast<cudaStream_t>(stream));
}

int rt_gpu_render_shard_device(Scene const *scene, isize width, isize height, isize samples, isize max_bounces,
                               u32 user_seed, i32 rank, i32 world, i32 split_mode, i32 *mode_used,
                               f32 *d_accum, u64 *d_counters, void *stream) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (ensure_init()) return 1;
  if (world < 1 || rank < 0 || rank >= world) return fail("render_shard: rank %d of %d", rank, world);
  if (scene_on_devices(scene)) return 1;
  const i32 mode = rt_gpu_shard_mode((i32)samples, world, split_mode);
  if (mode_used) *mode_used = mode;
  Share share{0, samples, 0, 1};
  if (mode == RT_GPU_SPLIT_SAMPLES) {
    i32 a = 0, b = 0;
    rt_gpu_shard_samples(rank, world, (i32)samples, &a, &b);
    share.s_begin = a; share.s_end = b;
  } else {
    share.split_rank = rank; share.split_world = world;
  }
  return render_on_device(g.devs[0], scene, width, height, share, max_bounces, user_seed, 0, d_accum, nullptr, nullptr,
                          reinterpret_cast<unsigned long long *>(d_counters), static_cast<cudaStream_t>(stream));
}

int rt_gpu_accum_buffer(isize width, isize height, f32 **d_accum_out) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (ensure_init()) return 1;
  if (width < 1 || height < 1) return fail("accum_buffer: empty image");
  if (ensure_frame_buffers(g.devs[0], (size_t)width * (size_t)height, false)) return 1;
  *d_accum_out = g.devs[0].d_accum;
  return 0;
}

int rt_gpu_reduce_resolve_device(f32 const *const *d_parts, i32 n_parts, f32 *d_sum_out, isize width, isize height,
                                 isize samples, u8 *d_pixels, isize stride, i32 components, void *stream) {
  if (n_parts < 1 || n_parts > RT_MAX_PARTS) return fail("reduce_resolve: between 1 and %d parts", RT_MAX_PARTS);
  if (samples < 1 || (d_pixels && components < 3)) return fail("reduce_resolve: samples >= 1 and components >= 3 required");
  ReduceParts parts{};
  parts.n = n_parts;
  for (int k = 0; k < n_parts; k++) parts.part[k] = d_parts[k];
  int e = rt_launch_reduce_resolve(parts, d_sum_out, (int)width, (int)height, (int)samples, d_pixels, (int)stride, components,
                                   static_cast<cudaStream_t>(stream));
  if (e) return fail("reduce+resolve kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  g.last_launches++;
  return 0;
}

int rt_gpu_resolve_device(f32 const *d_accum, isize width, isize height, isize samples, u8 *d_pixels, isize stride,
                          i32 components, void *stream) {
  if (samples < 1 || components < 3) return fail("resolve: samples >= 1 and components >= 3 required");
  int e = rt_launch_resolve(d_accum, (int)width, (int)height, (int)samples, d_pixels, (int)stride, components,
                            static_cast<cudaStream_t>(stream));
  if (e) return fail("resolve kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  g.last_launches++;
  return 0;
}

int rt_gpu_denoise_device(u8 const *d_src, u8 *d_dst, isize width, isize height, isize src_stride, isize dst_stride,
                          i32 components, void *stream) {
  if (d_src == d_dst) return fail("denoise: src and dst must differ (reference denoiser.c:130)");
  int e = rt_launch_denoise(d_src, d_dst, (int)width, (int)height, (int)src_stride, (int)dst_stride, components,
                            static_cast<cudaStream_t>(stream));
  if (e) return fail("denoise kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  g.last_launches++;
  return 0;
}

// ----------------------------------------------------------------- parity hooks
int rt_gpu_read_accum(f32 *out, isize n_floats) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g.devs.empty() || !g.devs[0].d_accum || (size_t)n_floats > g.last_pixels * 3) return fail("read_accum: no render of that size has run");
  CUDA_TRY(cudaSetDevice(g.devs[0].id));
  CUDA_TRY(cudaMemcpy(out, g.devs[0].d_accum, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int rt_gpu_read_hit_ids(i32 *out, isize n_pixels) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g.devs.empty() || !g.devs[0].d_hit_ids || !g.last_has_hit_ids || (size_t)n_pixels > g.last_pixels)
    return fail("read_hit_ids: the last render did not keep hit ids (RT_GPU_Options.keep_hit_ids)");
  CUDA_TRY(cudaSetDevice(g.devs[0].id));
  CUDA_TRY(cudaMemcpy(out, g.devs[0].d_hit_ids, (size_t)n_pixels * sizeof(int), cudaMemcpyDeviceToHost));
  if (g.last_split_mode == RT_GPU_SPLIT_CHUNKS && g.devs.size() > 1) {
    // chunk split: every device recorded the pixels of its own chunks, INT_MIN elsewhere
    std::vector<int> other((size_t)n_pixels);
    for (size_t k = 1; k < g.devs.size(); k++) {
      CUDA_TRY(cudaSetDevice(g.devs[k].id));
      CUDA_TRY(cudaMemcpy(other.data(), g.devs[k].d_hit_ids, (size_t)n_pixels * sizeof(int), cudaMemcpyDeviceToHost));
      for (isize i = 0; i < n_pixels; i++) if (other[(size_t)i] > out[i]) out[i] = other[(size_t)i];
    }
    CUDA_TRY(cudaSetDevice(g.devs[0].id));
  }
  return 0;
}

// texels (RGBA8, tightly packed) of texture `slot` of a resident scene on the first device; slot order = first use by
// the materials in triangle-slot order, the environment last unless a material uses it.  *width / *height are set;
// out may be NULL to query the size, slot beyond the last returns non-zero.
int rt_gpu_read_texture(Scene const *scene, i32 slot, u8 *out, isize *width, isize *height) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g.devs.empty()) return fail("read_texture: not initialised");
  Device &d = g.devs[0];
  auto it = d.scenes.find(scene);
  if (it == d.scenes.end()) return fail("read_texture: scene is not resident");
  CUDA_TRY(cudaSetDevice(d.id));
  CUDA_TRY(cudaStreamSynchronize(d.copy));
  TextureDev table[64];
  const char *tab = reinterpret_cast<const char *>(it->second.dev.textures);
  const int n = it->second.n_textures;
  if (slot < 0 || slot >= n || n > 64) return fail("read_texture: slot %d of %d", slot, n);
  CUDA_TRY(cudaMemcpy(table, tab, (size_t)n * sizeof(TextureDev), cudaMemcpyDeviceToHost));
  if (width) *width = table[slot].width;
  if (height) *height = table[slot].height;
  if (out) CUDA_TRY(cudaMemcpy(out, table[slot].texels, (size_t)table[slot].width * (size_t)table[slot].height * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int rt_gpu_read_counters(u64 out[8]) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g.devs.empty() || !g.devs[0].d_counters) return fail("read_counters: no render has run");
  for (int i = 0; i < 8; i++) out[i] = 0;
  for (Device &d : g.devs) {
    if (!d.d_counters) continue;
    u64 part[8];
    CUDA_TRY(cudaSetDevice(d.id));
    CUDA_TRY(cudaMemcpy(part, d.d_counters, sizeof part, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; i++) out[i] += part[i];
  }
  CUDA_TRY(cudaSetDevice(g.devs[0].id));
  return 0;
}

// The library's own 16-slot counter block on the first device: [0..8) as RT_GPU_CTR_*, [8] rays whose whole walk
// was the root-union test (they count as one node visit in [1], as in the reference, but cost 18 instructions),
// [9] primary rays.  Pass it as d_counters to the device-level calls to have the extended slots filled.
int rt_gpu_counters_buffer(u64 **d_counters_out) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (ensure_init()) return 1;
  Device &d = g.devs[0];
  CUDA_TRY(cudaSetDevice(d.id));
  if (!d.d_counters) {
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d.d_counters), 16 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(d.d_counters, 0, 16 * sizeof(unsigned long long)));
  }
  *d_counters_out = reinterpret_cast<u64 *>(d.d_counters);
  return 0;
}

int rt_gpu_counters_reset(void) {
  u64 *p = nullptr;
  if (rt_gpu_counters_buffer(&p)) return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  CUDA_TRY(cudaSetDevice(g.devs[0].id));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemset(p, 0, 16 * sizeof(u64)));
  return 0;
}

int rt_gpu_read_counters_ex(u64 out[16]) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g.devs.empty() || !g.devs[0].d_counters) return fail("read_counters: no render has run");
  for (int i = 0; i < 16; i++) out[i] = 0;
  for (Device &d : g.devs) {
    if (!d.d_counters) continue;
    u64 part[16];
    CUDA_TRY(cudaSetDevice(d.id));
    CUDA_TRY(cudaMemcpy(part, d.d_counters, sizeof part, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; i++) out[i] += part[i];
  }
  CUDA_TRY(cudaSetDevice(g.devs[0].id));
  return 0;
}

int rt_gpu_last_launches(void) { return g.last_launches; }

void rt_gpu_stage_profile_enable(i32 on) {
  std::lock_guard<std::mutex> lock(g_mutex);
  rt_stage_profile_enable(on, g.devs.empty() ? 0 : g.devs[0].id);
}

int rt_gpu_stage_profile_read_bounces(f64 ms[64], i64 launches[64]) {
  std::lock_guard<std::mutex> lock(g_mutex);
  long long n[RT_N_STAGES * RT_STAGE_BOUNCES];
  int e = rt_stage_profile_read(ms, n);
  if (e) return fail("stage profile: %s", cudaGetErrorString((cudaError_t)e));
  for (int i = 0; i < RT_N_STAGES * RT_STAGE_BOUNCES; i++) launches[i] = n[i];
  return 0;
}

int rt_gpu_stage_profile_read(f64 ms[4], i64 launches[4]) {
  f64 all_ms[RT_N_STAGES * RT_STAGE_BOUNCES];
  i64 all_n[RT_N_STAGES * RT_STAGE_BOUNCES];
  if (rt_gpu_stage_profile_read_bounces(all_ms, all_n)) return 1;
  for (int s = 0; s < RT_N_STAGES; s++) {
    ms[s] = 0; launches[s] = 0;
    for (int b = 0; b < RT_STAGE_BOUNCES; b++) { ms[s] += all_ms[s * RT_STAGE_BOUNCES + b]; launches[s] += all_n[s * RT_STAGE_BOUNCES + b]; }
  }
  return 0;
}

f64 rt_gpu_last_kernel_ms(void) { return g.last_kernel_ms; }

void rt_gpu_last_frame_breakdown(f64 out[4]) {
  out[0] = g.last_upload_ms; out[1] = g.last_reduce_ms; out[2] = g.last_d2h_ms; out[3] = (f64)g.last_split_mode;
}

// ------------------------------------------------------- reference entry points
static int render_owner(Rendering_Context *ctx, isize n_chunks) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g.last_launches = 0;
  if (ensure_init()) return 1;
  const Image &im = ctx->image;
  if (im.components < 3 || im.pixel_type != PT_u8 || !im.pixels.data) return fail("render: image must be u8 with >= 3 components");
  if (im.width < 1 || im.height < 1 || im.stride < im.width) return fail("render: empty image or stride < width");
  if (ctx->samples < 1) return fail("render: samples must be >= 1");
  if (!ctx->scene) return fail("render: null scene");
  const size_t n_dev = g.devs.size();
  if (n_dev > 1 && enable_peers()) return 1;
  const size_t n_pixels = (size_t)im.width * (size_t)im.height;
  const size_t image_bytes = (size_t)im.stride * (size_t)im.height * (size_t)im.components;
  const RT_GPU_Options opt = g.options;

  double t0 = now_ms();
  bool was_resident = true;
  for (Device &d : g.devs) was_resident &= d.scenes.find(ctx->scene) != d.scenes.end();
  if (scene_on_devices(ctx->scene)) return 1;
  g.last_upload_ms = was_resident ? 0.0 : now_ms() - t0;

  for (Device &d : g.devs)
    if (ensure_frame_buffers(d, n_pixels, opt.keep_hit_ids != 0)) return 1;
  Device &d0 = g.devs[0];
  CUDA_TRY(cudaSetDevice(d0.id));
  if (grow(reinterpret_cast<void **>(&d0.d_image), &d0.image_bytes, image_bytes)) return 1;
  if (grow_pinned(d0, image_bytes)) return 1;

  // the samples this call renders (a multi-process host narrows them per rank), then their split over the devices
  isize s_begin = opt.sample_begin, s_end = opt.sample_end;
  if (!opt.sample_range_set && s_end <= 0) s_end = ctx->samples;
  if (s_begin < 0) s_begin = 0;
  if (s_end > ctx->samples) s_end = ctx->samples;
  if (s_end < s_begin) s_end = s_begin;
  const i32 mode = n_dev > 1 ? rt_gpu_shard_mode((i32)(s_end - s_begin), (i32)n_dev, opt.split_mode) : RT_GPU_SPLIT_SAMPLES;

  // samples per progress slice: as asked, or as many as one wavefront chunk holds (fewer, larger chunks are faster)
  isize slice = opt.slice_samples;
  if (slice <= 0) {
    const int world = n_dev > 1 && mode == RT_GPU_SPLIT_CHUNKS ? (int)n_dev : (opt.pixel_world > 1 ? opt.pixel_world : 1);
    slice = (isize)rt_render_chunk_samples((int)im.width, (int)im.height, world);
    if (slice < 1) slice = 1;
  }
  g.last_pixels = n_pixels;
  g.last_has_hit_ids = opt.keep_hit_ids != 0;
  g.last_split_mode = mode;
  // rows padded by stride keep whatever the caller had there
  if (im.stride != im.width || im.components != 3) {
    memcpy(d0.h_pinned, im.pixels.data, image_bytes);
    CUDA_TRY(cudaMemcpyAsync(d0.d_image, d0.h_pinned, image_bytes, cudaMemcpyHostToDevice, d0.stream));
  }

  CUDA_TRY(cudaEventRecord(d0.ev0, d0.stream));
  std::vector<Share> shares(n_dev);
  isize max_slices = 0;
  for (size_t k = 0; k < n_dev; k++) {
    Share &sh = shares[k];
    sh = Share{s_begin, s_end, opt.pixel_rank, opt.pixel_world};
    if (n_dev > 1) {
      if (mode == RT_GPU_SPLIT_SAMPLES) {
        i32 a = 0, b = 0;
        rt_gpu_shard_samples((i32)k, (i32)n_dev, (i32)(s_end - s_begin), &a, &b);
        sh.s_begin = s_begin + a; sh.s_end = s_begin + b;
      } else {
        sh.split_rank = (int)k; sh.split_world = (int)n_dev;
      }
    }
    const isize n_slices = (sh.s_end - sh.s_begin + slice - 1) / slice;
    if (n_slices > max_slices) max_slices = n_slices;
  }
  // Path-queue workspaces first, for every device, before anything is launched: with peer access enabled a
  // cudaMalloc maps the new block into every peer and waits for whatever those peers are running.
  for (size_t k = 0; k < n_dev; k++) {
    Device &d = g.devs[k];
    const Share &sh = shares[k];
    const isize n = sh.s_end - sh.s_begin < slice ? sh.s_end - sh.s_begin : slice;
    if (n <= 0) continue;
    CUDA_TRY(cudaSetDevice(d.id));
    const size_t want = rt_render_workspace_bytes((int)im.width, (int)im.height, (int)n, (int)ctx->max_bounces, 0, sh.split_world);
    const size_t floor_bytes = rt_render_workspace_bytes((int)im.width, (int)im.height, 1, (int)ctx->max_bounces, 1, sh.split_world);
    if (ensure_workspace(d, want, floor_bytes, d.stream)) return 1;
  }
  // slices round-robin over the devices, so the host's progress bar (driver.c:810-818) follows all of them
  for (size_t k = 0; k < n_dev; k++) {
    Device &d = g.devs[k];
    CUDA_TRY(cudaSetDevice(d.id));
    CUDA_TRY(cudaMemsetAsync(d.d_counters, 0, 16 * sizeof(unsigned long long), d.stream));
    if (opt.keep_hit_ids && shares[k].split_world > 1)
      CUDA_TRY(cudaMemsetAsync(d.d_hit_ids, 0x80, n_pixels * sizeof(int), d.stream));      // 0x80808080 < -1: "not mine"
    if (shares[k].s_end <= shares[k].s_begin) {
      // an empty share still defines its accumulator: zero
      const Share empty{shares[k].s_begin, shares[k].s_begin, shares[k].split_rank, shares[k].split_world};
      if (render_on_device(d, ctx->scene, im.width, im.height, empty, ctx->max_bounces, opt.user_seed, 0, d.d_accum,
                           nullptr, nullptr, d.d_counters, d.stream))
        return 1;
    }
  }
  for (isize j = 0; j < max_slices; j++) {
    for (size_t k = 0; k < n_dev; k++) {
      const Share &sh = shares[k];
      const isize a = sh.s_begin + j * slice;
      if (a >= sh.s_end) continue;
      const Share part{a, a + slice < sh.s_end ? a + slice : sh.s_end, sh.split_rank, sh.split_world};
      Device &d = g.devs[k];
      const bool first = j == 0;
      if (render_on_device(d, ctx->scene, im.width, im.height, part, ctx->max_bounces, opt.user_seed, first ? 0 : 1, d.d_accum,
                           nullptr, (opt.keep_hit_ids && first && (k == 0 || sh.split_world > 1)) ? d.d_hit_ids : nullptr,
                           d.d_counters, d.stream))
        return 1;
    }
    if (max_slices > 1 && (j % 4 == 3)) {
      // progress for the host's bar; never reaches n_chunks early
      for (Device &d : g.devs) { CUDA_TRY(cudaSetDevice(d.id)); CUDA_TRY(cudaStreamSynchronize(d.stream)); }
      isize progress = n_chunks * (j + 1) / max_slices;
      if (progress >= n_chunks) progress = n_chunks - 1;
      if (progress > ctx->_current_chunk) ctx->_current_chunk = (i32)progress;
    }
  }
  CUDA_TRY(cudaSetDevice(d0.id));
  CUDA_TRY(cudaEventRecord(d0.ev1, d0.stream));

  // ---- film: combine the devices' accumulators on the first one and resolve (raytracer.c:700-716)
  cudaEvent_t r0 = nullptr, r1 = nullptr;
  if (n_dev > 1) {
    for (size_t k = 1; k < n_dev; k++) {
      CUDA_TRY(cudaSetDevice(g.devs[k].id));
      CUDA_TRY(cudaEventRecord(g.devs[k].done, g.devs[k].stream));
    }
    CUDA_TRY(cudaSetDevice(d0.id));
    CUDA_TRY(cudaEventCreate(&r0));
    CUDA_TRY(cudaEventCreate(&r1));
    for (size_t k = 1; k < n_dev; k++) CUDA_TRY(cudaStreamWaitEvent(d0.stream, g.devs[k].done, 0));
    CUDA_TRY(cudaEventRecord(r0, d0.stream));
  }
  if (n_dev > 1 && opt.reduce_mode == RT_GPU_REDUCE_NCCL) {
    if (nccl_reduce_to_first(n_pixels * 3)) return 1;
    CUDA_TRY(cudaSetDevice(d0.id));
    int e = rt_launch_resolve(d0.d_accum, (int)im.width, (int)im.height, (int)ctx->samples, d0.d_image, (int)im.stride,
                              im.components, d0.stream);
    if (e) return fail("resolve kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  } else {
    ReduceParts parts{};
    parts.n = (int)n_dev;
    for (size_t k = 0; k < n_dev; k++) parts.part[k] = g.devs[k].d_accum;
    // one device: the same kernel with a single part (x + nothing) — one film code path for every N
    int e = rt_launch_reduce_resolve(parts, n_dev > 1 ? d0.d_accum : nullptr, (int)im.width, (int)im.height, (int)ctx->samples,
                                     d0.d_image, (int)im.stride, im.components, d0.stream);
    if (e) return fail("reduce+resolve kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  }
  g.last_launches++;
  if (r1) CUDA_TRY(cudaEventRecord(r1, d0.stream));
  CUDA_TRY(cudaMemcpyAsync(d0.h_pinned, d0.d_image, image_bytes, cudaMemcpyDeviceToHost, d0.stream));
  CUDA_TRY(cudaStreamSynchronize(d0.stream));
  t0 = now_ms();
  memcpy(im.pixels.data, d0.h_pinned, image_bytes);
  g.last_d2h_ms = now_ms() - t0;
  // peers keep their accumulators untouched until the reduce has read them: it has, the stream is drained
  for (size_t k = 1; k < n_dev; k++) {
    CUDA_TRY(cudaSetDevice(g.devs[k].id));
    CUDA_TRY(cudaStreamSynchronize(g.devs[k].stream));
  }
  CUDA_TRY(cudaSetDevice(d0.id));
  float ms = 0;
  cudaEventElapsedTime(&ms, d0.ev0, d0.ev1);
  g.last_kernel_ms = ms;
  g.last_reduce_ms = 0;
  if (r0 && r1) {
    cudaEventElapsedTime(&ms, r0, r1);
    g.last_reduce_ms = ms;
    cudaEventDestroy(r0);
    cudaEventDestroy(r1);
  }
  cudaError_t late = cudaGetLastError();
  if (late != cudaSuccess) return fail("render: %s", cudaGetErrorString(late));
  return 0;
}

void render_thread_proc(Rendering_Context *ctx) {
  const isize chunks_x = (ctx->image.width + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  const isize chunks_y = (ctx->image.height + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  const isize n_chunks = chunks_x * chunks_y;
  // raytracer.c:620: claim a chunk; whoever gets chunk 0 launches the whole frame
  i32 c = __atomic_fetch_add(const_cast<i32 *>(&ctx->_current_chunk), 1, __ATOMIC_SEQ_CST);
  if (c == 0) {
    clear_error();
    g_status.store(0);
    if (render_owner(ctx, n_chunks)) {
      g_status.store(1);
      fprintf(stderr, "render_thread_proc: %s\n", rt_gpu_last_error());
    }
    __atomic_store_n(const_cast<i32 *>(&ctx->_current_chunk), (i32)n_chunks + 1, __ATOMIC_SEQ_CST);
  }
  // raytracer.c:622
  __atomic_fetch_add(const_cast<i32 *>(&ctx->n_threads), -1, __ATOMIC_SEQ_CST);
}

bool rendering_context_is_finished(Rendering_Context *ctx) {
  return __atomic_load_n(const_cast<i32 *>(&ctx->n_threads), __ATOMIC_SEQ_CST) == 0;
}

void rendering_context_finish(Rendering_Context *ctx) {
  while (__atomic_load_n(const_cast<i32 *>(&ctx->n_threads), __ATOMIC_SEQ_CST) > 0) sched_yield();
}

// raytracer.h:56 / raytracer.c:722-784 on the GPU (first device): the kernels are in rt_render.cu.  values_out
// (optional, W*H*3 f32) receives accumulated / samples of every written texel before the u8 store, owner_out
// (optional, W*H i32) the slot of the triangle that owns each texel (-1 = untouched; such texels keep their bytes).
int rt_gpu_lightmap_bake(Image const *lightmap, Scene const *scene, isize samples, f32 *values_out, i32 *owner_out) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g.last_launches = 0;
  if (ensure_init()) return 1;
  if (!lightmap || !scene) return fail("lightmap_bake: null argument");
  if (lightmap->pixel_type != PT_u8 || lightmap->components < 3 || !lightmap->pixels.data)      // raytracer.c:723-724
    return fail("lightmap_bake: the lightmap must be u8 with >= 3 components");
  if (lightmap->width < 1 || lightmap->height < 1 || lightmap->stride < lightmap->width) return fail("lightmap_bake: empty image or stride < width");
  if (samples < 1) return fail("lightmap_bake: samples must be >= 1");
  if (scene_on_devices(scene)) return 1;
  Device &d = g.devs[0];
  CUDA_TRY(cudaSetDevice(d.id));
  const DeviceScene &ds = d.scenes.find(scene)->second;
  const int W = (int)lightmap->width, H = (int)lightmap->height;
  const size_t n_texels = (size_t)W * (size_t)H;
  const size_t image_bytes = (size_t)lightmap->stride * (size_t)H * (size_t)lightmap->components;
  const int max_bounces = 8;                                        // raytracer.c:773: cast_ray(scene, r, 8)

  CUDA_TRY(cudaStreamSynchronize(d.stream));
  CUDA_TRY(cudaStreamSynchronize(d.copy));
  if (grow(reinterpret_cast<void **>(&d.d_image), &d.image_bytes, image_bytes)) return 1;
  if (grow_pinned(d, image_bytes)) return 1;
  const size_t scratch_bytes = rt_lightmap_workspace_bytes(W, H) + (values_out ? n_texels * 3 * sizeof(float) : 0) +
                               (owner_out ? n_texels * sizeof(int) : 0);
  if (grow(&d.d_texel_stage, &d.texel_stage_bytes, scratch_bytes)) return 1;       // the upload's raw-texel block doubles as scratch
  char *scratch = static_cast<char *>(d.d_texel_stage);
  float *d_values = values_out ? reinterpret_cast<float *>(scratch + rt_lightmap_workspace_bytes(W, H)) : nullptr;
  int *d_owner = owner_out ? reinterpret_cast<int *>(scratch + rt_lightmap_workspace_bytes(W, H) + (values_out ? n_texels * 3 * sizeof(float) : 0)) : nullptr;
  // path queues: at most one chunk's worth, at least one sample of every texel
  const size_t per_path = rt_render_workspace_bytes(8, 4, 1, max_bounces, 1, 1) / 32 + sizeof(float) + 16;
  size_t paths = n_texels * (size_t)samples;
  const size_t chunk_paths = (size_t)rt_render_chunk_samples(8, 4, 1) * 32;
  if (paths > chunk_paths) paths = chunk_paths;
  if (paths < n_texels) paths = n_texels;
  if (ensure_workspace(d, paths * per_path + 4096, n_texels * per_path + 4096, d.stream)) return 1;

  if (d.busy_valid) CUDA_TRY(cudaStreamWaitEvent(d.stream, d.busy, 0));
  CUDA_TRY(cudaStreamWaitEvent(d.stream, ds.geom_ready, 0));
  CUDA_TRY(cudaStreamWaitEvent(d.stream, ds.tex_ready, 0));
  memcpy(d.h_pinned, lightmap->pixels.data, image_bytes);            // untouched texels keep their bytes
  CUDA_TRY(cudaMemcpyAsync(d.d_image, d.h_pinned, image_bytes, cudaMemcpyHostToDevice, d.stream));
  if (d_values) CUDA_TRY(cudaMemsetAsync(d_values, 0, n_texels * 3 * sizeof(float), d.stream));
  unsigned n_jobs = 0;
  int e = rt_launch_lightmap(ds.dev, W, H, (int)samples, max_bounces, g.options.user_seed, d.d_image, (int)lightmap->stride,
                             lightmap->components, d_values, d_owner, scratch, d.sm_count, d.d_workspace, d.workspace_bytes,
                             d.stream, &g.last_launches, &n_jobs);
  if (e) return fail("lightmap kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  CUDA_TRY(cudaEventRecord(d.busy, d.stream));
  d.busy_valid = true;
  CUDA_TRY(cudaMemcpyAsync(d.h_pinned, d.d_image, image_bytes, cudaMemcpyDeviceToHost, d.stream));
  CUDA_TRY(cudaStreamSynchronize(d.stream));
  memcpy(lightmap->pixels.data, d.h_pinned, image_bytes);
  if (values_out) CUDA_TRY(cudaMemcpy(values_out, d_values, n_texels * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  if (owner_out) CUDA_TRY(cudaMemcpy(owner_out, d_owner, n_texels * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

void lightmap_bake(Image const *lightmap, Scene const *scene, isize samples) {
  clear_error();
  g_status.store(0);
  if (rt_gpu_lightmap_bake(lightmap, scene, samples, nullptr, nullptr)) {
    g_status.store(1);
    fprintf(stderr, "lightmap_bake: %s\n", rt_gpu_last_error());
  }
}

void denoise_image(Image const *src, Image const *dst, isize n_threads) {
  (void)n_threads;
  auto run = [&]() -> int {
    std::lock_guard<std::mutex> lock(g_mutex);
    g.last_launches = 0;
    if (ensure_init()) return 1;
    if (!src->pixels.data || !dst->pixels.data) return fail("denoise: null pixel buffer");
    if (src->pixels.data == dst->pixels.data) return fail("denoise: src and dst must differ (reference denoiser.c:130)");
    if (src->width != dst->width || src->height != dst->height || src->components != dst->components)
      return fail("denoise: src and dst shapes differ (reference denoiser.c:131-132)");
    Device &d = g.devs[0];
    CUDA_TRY(cudaSetDevice(d.id));
    const size_t src_bytes = (size_t)src->stride * (size_t)src->height * (size_t)src->components;
    const size_t dst_bytes = (size_t)dst->stride * (size_t)dst->height * (size_t)dst->components;
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    if (grow(reinterpret_cast<void **>(&d.d_image), &d.image_bytes, src_bytes)) return 1;
    if (grow(reinterpret_cast<void **>(&d.d_image2), &d.image2_bytes, dst_bytes)) return 1;
    if (grow_pinned(d, src_bytes > dst_bytes ? src_bytes : dst_bytes)) return 1;
    memcpy(d.h_pinned, src->pixels.data, src_bytes);
    CUDA_TRY(cudaMemcpyAsync(d.d_image, d.h_pinned, src_bytes, cudaMemcpyHostToDevice, d.stream));
    if (dst->stride != dst->width) CUDA_TRY(cudaMemsetAsync(d.d_image2, 0, dst_bytes, d.stream));
    CUDA_TRY(cudaEventRecord(d.ev0, d.stream));
    int e = rt_launch_denoise(d.d_image, d.d_image2, (int)src->width, (int)src->height, (int)src->stride, (int)dst->stride,
                              src->components, d.stream);
    if (e) return fail("denoise kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
    g.last_launches++;
    CUDA_TRY(cudaEventRecord(d.ev1, d.stream));
    CUDA_TRY(cudaMemcpyAsync(d.h_pinned, d.d_image2, dst_bytes, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    if (dst->stride == dst->width) {
      memcpy(dst->pixels.data, d.h_pinned, dst_bytes);
    } else {
      for (isize y = 0; y < dst->height; y++)
        memcpy(dst->pixels.data + (size_t)y * (size_t)dst->stride * (size_t)dst->components,
               d.h_pinned + (size_t)y * (size_t)dst->stride * (size_t)dst->components,
               (size_t)dst->width * (size_t)dst->components);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, d.ev0, d.ev1);
    g.last_kernel_ms = ms;
    return 0;
  };
  clear_error();
  g_status.store(0);
  if (run()) {
    g_status.store(1);
    fprintf(stderr, "denoise_image: %s\n", rt_gpu_last_error());
  }
}

}  // extern "C"
