// rt_cabi.cu — the C ABI of libraytracer_gpu.so.
//
// Exports the reference's own entry points (raytracer.h:51-56, denoiser.h:4) plus
// the additive rt_gpu_* calls declared in include/rt_gpu.h.  Host-side work here:
//   * scene residency: flatten/regroup the host Scene into the device layout of
//     rt_device.cuh once per Scene pointer (materials de-duplicated by
//     Shader.data, textures by Image*, RGB8 -> RGBA8);
//   * the reference's threading protocol around render_thread_proc
//     (driver.c:793-818): the thread that claims chunk 0 owns the launch;
//   * pinned staging + async copies for the image, CUDA-event timing.
// There is no CPU fallback: every compute call fails if CUDA does.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <sched.h>

#include "rt_gpu.h"
#include "rt_kernels.h"
#include "rt_gpu_internal.h"

namespace {

std::mutex g_mutex;
thread_local char g_error[512] = "";
char g_error_shared[512] = "";

int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
  snprintf(g_error_shared, sizeof g_error_shared, "%s", g_error);
  return 1;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e_));     \
  } while (0)

struct DeviceScene {
  SceneDev dev{};
  std::vector<std::pair<void *, size_t>> allocations;
  size_t bytes = 0;        // resident on the device
  size_t h2d_bytes = 0;    // copied host -> device by the upload
};

struct State {
  bool         ready = false;
  int          device = -1;
  int          sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
  std::vector<Shader_Proc>     pbr_procs;
  std::vector<Background_Proc> bg_procs;
  std::map<const Scene *, DeviceScene> scenes;
  RT_GPU_Options options{0, 0, 0, 64, 0};
  // buffers of the last entry-point render (parity hooks)
  float *d_accum = nullptr;   size_t accum_floats = 0;
  int   *d_hit_ids = nullptr; size_t hit_pixels = 0;
  unsigned long long *d_counters = nullptr;
  void  *d_texel_stage = nullptr; size_t texel_stage_bytes = 0;  // raw texture bytes before the RGBA8 repack
  void  *d_workspace = nullptr; size_t workspace_bytes = 0;     // wavefront path queues (rt_render.cu)
  unsigned char *d_image = nullptr, *d_image2 = nullptr; size_t image_bytes = 0, image2_bytes = 0;
  unsigned char *h_pinned = nullptr; size_t pinned_bytes = 0;
  int    last_launches = 0;
  double last_kernel_ms = 0;
  bool   last_has_hit_ids = false;
  size_t last_pixels = 0;
} g;

int ensure_init() {
  if (g.ready) return 0;
  return rt_gpu_init(g.device < 0 ? 0 : g.device);
}

// Device blocks of released scenes are kept and handed out again by exact size: re-uploading a
// scene (per frame in an interactive host, per step in bench.py's end-to-end leg) then costs no
// cudaMalloc / cudaFree (each a device-wide synchronisation).
std::multimap<size_t, void *> g_block_pool;

int pool_alloc(void **out, size_t bytes) {
  auto it = g_block_pool.find(bytes);
  if (it != g_block_pool.end()) { *out = it->second; g_block_pool.erase(it); return 0; }
  CUDA_TRY(cudaMalloc(out, bytes));
  return 0;
}

int grow_pinned(size_t want);

// Host-side staging copy, split over a few threads above 4 MB: one core copies ~10 GB/s, and the 50 MB of
// the helmet's textures staged per upload would otherwise cost more than their PCIe transfer.
void staged_copy(void *dst, const void *src, size_t n) {
  const size_t kMin = 4u << 20;
  if (n < kMin) { memcpy(dst, src, n); return; }
  const int parts = 4;
  std::thread workers[parts - 1];
  const size_t step = ((n / parts) + 63) & ~(size_t)63;
  for (int i = 1; i < parts; i++) {
    const size_t off = step * i, len = off < n ? (off + step < n && i + 1 < parts ? step : n - off) : 0;
    workers[i - 1] = std::thread([=] { if (len) memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, len); });
  }
  memcpy(dst, src, step < n ? step : n);
  for (auto &t : workers) t.join();
}

// host -> device through the pinned staging buffer, asynchronously on the library's stream;
// staging is reused, so the copy of one buffer is drained before the next one is staged
int upload_bytes(DeviceScene &ds, const void *host, size_t n, const void **out) {
  void *p = nullptr;
  const size_t alloc = n ? n : 16;
  if (pool_alloc(&p, alloc)) return 1;
  ds.allocations.push_back({p, alloc});
  ds.bytes += n;
  if (n) {
    if (grow_pinned(n)) return 1;
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    staged_copy(g.h_pinned, host, n);
    ds.h2d_bytes += n;
    CUDA_TRY(cudaMemcpyAsync(p, g.h_pinned, n, cudaMemcpyHostToDevice, g.stream));
  }
  *out = p;
  return 0;
}

template <typename T>
int upload(DeviceScene &ds, const std::vector<T> &host, const T **out) {
  const void *p = nullptr;
  if (upload_bytes(ds, host.data(), host.size() * sizeof(T), &p)) return 1;
  *out = static_cast<const T *>(p);
  return 0;
}

void release(DeviceScene &ds) {
  for (auto &a : ds.allocations) g_block_pool.emplace(a.second, a.first);
  ds.allocations.clear();
}

void drop_block_pool() {
  for (auto &kv : g_block_pool) cudaFree(kv.second);
  g_block_pool.clear();
}

int grow(void **ptr, size_t *have, size_t want) {
  if (*have >= want && *ptr) return 0;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *have = 0;
  CUDA_TRY(cudaMalloc(ptr, want));
  *have = want;
  return 0;
}

int grow_pinned(size_t want) {
  if (g.pinned_bytes >= want && g.h_pinned) return 0;
  if (g.h_pinned) cudaFreeHost(g.h_pinned);
  g.h_pinned = nullptr;
  g.pinned_bytes = 0;
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&g.h_pinned), want));
  g.pinned_bytes = want;
  return 0;
}

int build_device_scene(const Scene *scene, DeviceScene &ds) {
  const isize depth = scene->bvh.depth;
  if (depth < 1 || depth > RT_MAX_DEPTH) return fail("scene: BVH depth %ld outside [1,%d]", (long)depth, RT_MAX_DEPTH);
  const isize n_nodes = scene->bvh.nodes.len, n_slots = scene->triangles.len;
  if (n_nodes != bvh_n_internal_nodes(depth) || n_slots != bvh_n_leaf_nodes(depth) * RT_SIMD_WIDTH)
    return fail("scene: node/slot counts do not describe a complete 8-ary tree of depth %ld", (long)depth);

  std::vector<float> nodes((size_t)n_nodes * 48);
  memcpy(nodes.data(), scene->bvh.nodes.data, nodes.size() * sizeof(float));

  // per slot: p0.xyz, e1 = p1 - p0, e2 = p2 - p0.  The two edge vectors are the f32
  // subtractions the reference redoes for every ray (raytracer.c:116-122); volatile keeps the
  // host compiler from doing them in any wider type
  std::vector<float4> tri_pos((size_t)n_slots * 3);
  const float *px[3] = { scene->triangles.x[0], scene->triangles.y[0], scene->triangles.z[0] };
  const float *p1[3] = { scene->triangles.x[1], scene->triangles.y[1], scene->triangles.z[1] };
  const float *p2[3] = { scene->triangles.x[2], scene->triangles.y[2], scene->triangles.z[2] };
  for (isize s = 0; s < n_slots; s++) {
    volatile float e1[3], e2[3];
    for (int a = 0; a < 3; a++) { e1[a] = p1[a][s] - px[a][s]; e2[a] = p2[a][s] - px[a][s]; }
    tri_pos[(size_t)s * 3 + 0] = make_float4(px[0][s], px[1][s], px[2][s], e1[0]);
    tri_pos[(size_t)s * 3 + 1] = make_float4(e1[1], e1[2], e2[0], e2[1]);
    tri_pos[(size_t)s * 3 + 2] = make_float4(e2[2], 0.0f, 0.0f, 0.0f);
  }

  // materials / textures, de-duplicated by host pointer
  std::map<const void *, int> material_index, texture_index;
  std::vector<MaterialDev> materials;
  std::vector<const Image *> images;
  auto texture_slot = [&](const Image *im) -> int {
    if (!im) return -1;
    auto it = texture_index.find(im);
    if (it != texture_index.end()) return it->second;
    int slot = (int)images.size();
    images.push_back(im);
    texture_index[im] = slot;
    return slot;
  };

  std::vector<float4> records((size_t)n_slots * 7);
  for (isize s = 0; s < n_slots; s++) {
    const Triangle_AOS &a = scene->triangles.aos[s];
    int mat = 0;
    if (a.shader.proc || a.shader.data) {
      bool known = false;
      for (Shader_Proc p : g.pbr_procs) known |= (p == a.shader.proc);
      if (!known) return fail("scene: triangle slot %ld uses a Shader_Proc that was not registered with rt_gpu_register_pbr_shader", (long)s);
      auto it = material_index.find(a.shader.data);
      if (it == material_index.end()) {
        const PBR_Shader_Data *m = static_cast<const PBR_Shader_Data *>(a.shader.data);
        MaterialDev d{};
        for (int c = 0; c < 3; c++) { d.base[c] = m->base_color.data[c]; d.emission[c] = m->emission.data[c]; }
        d.roughness = m->roughness; d.metalness = m->metalness; d.normal_strength = m->normal_map_strength;
        d.sheen = m->sheen; d.sheen_tint = m->sheen_tint; d.aniso = m->anisotropic_strength;
        d.tex_albedo = texture_slot(m->texture_albedo);
        d.tex_normal = texture_slot(m->texture_normal);
        d.tex_mr = texture_slot(m->texture_metal_roughness);
        d.tex_emission = texture_slot(m->texture_emission);
        mat = (int)materials.size();
        materials.push_back(d);
        material_index[a.shader.data] = mat;
      } else {
        mat = it->second;
      }
    }
    float4 *r = &records[(size_t)s * 7];
    r[0] = make_float4(a.normal.x, a.normal.y, a.normal.z, a.normal_a.x);
    r[1] = make_float4(a.normal_a.y, a.normal_a.z, a.normal_b.x, a.normal_b.y);
    r[2] = make_float4(a.normal_b.z, a.normal_c.x, a.normal_c.y, a.normal_c.z);
    r[3] = make_float4(a.tangent.x, a.tangent.y, a.tangent.z, a.bitangent.x);
    r[4] = make_float4(a.bitangent.y, a.bitangent.z, a.tex_coords_a.x, a.tex_coords_a.y);
    r[5] = make_float4(a.tex_coords_b.x, a.tex_coords_b.y, a.tex_coords_c.x, a.tex_coords_c.y);
    int bits = mat;
    float as_float;
    memcpy(&as_float, &bits, 4);
    r[6] = make_float4(as_float, 0, 0, 0);
  }
  if (materials.empty()) materials.push_back(MaterialDev{});

  // environment
  bool bg_known = false;
  for (Background_Proc p : g.bg_procs) bg_known |= (p == scene->background.proc);
  if (!bg_known || !scene->background.data)
    return fail("scene: Scene.background.proc was not registered with rt_gpu_register_background (or its Image is null)");
  int env_slot = texture_slot(static_cast<const Image *>(scene->background.data));

  std::vector<TextureDev> textures(images.size());
  for (size_t i = 0; i < images.size(); i++) {
    const Image *im = images[i];
    if (im->components < 3 || !im->pixels.data) return fail("scene: texture %zu needs >= 3 u8 components", i);
    // raw texel bytes go up as they are (3 B per texel for RGB8); the RGBA8 repack the samplers read
    // (one 32-bit load per tap) is done by a kernel on the device
    const size_t n_texels = (size_t)im->width * (size_t)im->height;
    const size_t raw_bytes = (size_t)im->stride * (size_t)im->height * (size_t)im->components;
    void *d_texels_raw = nullptr;
    if (pool_alloc(&d_texels_raw, n_texels * sizeof(uchar4))) return 1;
    ds.allocations.push_back({d_texels_raw, n_texels * sizeof(uchar4)});
    ds.bytes += n_texels * sizeof(uchar4);
    size_t have = g.texel_stage_bytes;
    if (grow(&g.d_texel_stage, &have, raw_bytes)) return 1;
    g.texel_stage_bytes = have;
    if (grow_pinned(raw_bytes)) return 1;
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    staged_copy(g.h_pinned, im->pixels.data, raw_bytes);
    ds.h2d_bytes += raw_bytes;
    CUDA_TRY(cudaMemcpyAsync(g.d_texel_stage, g.h_pinned, raw_bytes, cudaMemcpyHostToDevice, g.stream));
    int e = rt_launch_texel_repack(static_cast<const unsigned char *>(g.d_texel_stage), (int)im->width, (int)im->height,
                                   (int)im->stride, im->components, static_cast<uchar4 *>(d_texels_raw), g.stream);
    if (e) return fail("texel repack launch failed: %s", cudaGetErrorString((cudaError_t)e));
    const uchar4 *d_texels = static_cast<const uchar4 *>(d_texels_raw);
    textures[i].texels = d_texels;
    textures[i].width = (int)im->width;
    textures[i].height = (int)im->height;
  }

  SceneDev &dev = ds.dev;
  if (upload(ds, nodes, &dev.nodes)) return 1;
  if (upload(ds, tri_pos, &dev.tri_pos)) return 1;
  if (upload(ds, records, &dev.tri_rec)) return 1;
  {
    // camera-relative copies for the primary trace, filled by rt_camera_relative_kernel at every render
    void *p = nullptr;
    const size_t node_bytes = nodes.size() * sizeof(float), tri_bytes = (size_t)n_slots * 4 * sizeof(float4);
    if (pool_alloc(&p, node_bytes)) return 1;
    ds.allocations.push_back({p, node_bytes}); ds.bytes += node_bytes;
    dev.nodes_rel = static_cast<float *>(p);
    if (pool_alloc(&p, tri_bytes)) return 1;
    ds.allocations.push_back({p, tri_bytes}); ds.bytes += tri_bytes;
    dev.tri_rel = static_cast<float4 *>(p);
  }
  if (upload(ds, materials, &dev.materials)) return 1;
  if (upload(ds, textures, &dev.textures)) return 1;
  dev.env_texture = env_slot;
  dev.depth = (int)depth;
  dev.n_internal = (int)n_nodes;
  dev.n_slots = (int)n_slots;
  // union of the root's child boxes; all-zero padding slots (lo == hi) can never be entered
  // (enter >= leave) and are left out
  for (int a = 0; a < 3; a++) { dev.root_lo[a] = INFINITY; dev.root_hi[a] = -INFINITY; }
  for (int j = 0; j < 8; j++) {
    const float *n0 = nodes.data();
    bool empty = true;
    for (int a = 0; a < 3; a++) empty &= (n0[a * 8 + j] == n0[(3 + a) * 8 + j]);
    if (empty) continue;
    for (int a = 0; a < 3; a++) {
      dev.root_lo[a] = fminf(dev.root_lo[a], n0[a * 8 + j]);
      dev.root_hi[a] = fmaxf(dev.root_hi[a], n0[(3 + a) * 8 + j]);
    }
  }
  // renders may run on another stream (device-pointer level): the scene is complete when this returns
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return 0;
}

void refresh_camera(const Scene *scene, SceneDev &dev) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) dev.view[r][c] = scene->camera.view_matrix.rows[r][c];
  dev.focal_length = scene->camera.focal_length;
}

int scene_on_device(const Scene *scene, SceneDev *out) {
  auto it = g.scenes.find(scene);
  if (it == g.scenes.end()) {
    DeviceScene ds;
    if (build_device_scene(scene, ds)) { release(ds); return 1; }
    it = g.scenes.emplace(scene, std::move(ds)).first;
  }
  refresh_camera(scene, it->second.dev);      // the camera is cheap and may change between frames
  *out = it->second.dev;
  return 0;
}

int render_device_locked(const Scene *scene, isize width, isize height, isize s_begin, isize s_end, isize max_bounces,
                         u32 seed, int accumulate, float *d_accum, float *d_per_sample, int *d_hit_ids,
                         unsigned long long *d_counters, cudaStream_t stream) {
  if (width < 1 || height < 1) return fail("render: empty image");
  RenderParams p{};
  if (scene_on_device(scene, &p.scene)) return 1;
  const size_t want = rt_render_workspace_bytes((int)width, (int)height, (int)(s_end - s_begin), (int)max_bounces,
                                                g.options.slice_samples);
  if (g.workspace_bytes < want) {
    CUDA_TRY(cudaStreamSynchronize(stream));      // an earlier launch may still read the old queues
    // the preferred size (up to 26.6 GB) is a throughput choice, not a requirement: on a device that
    // cannot spare it, halve the request down to one sample of every pixel per chunk
    const size_t floor_bytes = rt_render_workspace_bytes((int)width, (int)height, 1, (int)max_bounces, 1);
    size_t ask = want;
    for (;;) {
      if (g.d_workspace) { cudaFree(g.d_workspace); g.d_workspace = nullptr; g.workspace_bytes = 0; }
      if (cudaMalloc(&g.d_workspace, ask) == cudaSuccess) { g.workspace_bytes = ask; break; }
      cudaGetLastError();                          // clear the allocation failure
      g.d_workspace = nullptr;
      if (ask <= floor_bytes) return fail("render: cannot allocate %zu bytes of path queues", ask);
      ask = ask / 2 > floor_bytes ? ask / 2 : floor_bytes;
    }
  }
  p.width = (int)width; p.height = (int)height;
  p.sample_begin = (int)s_begin; p.sample_end = (int)s_end; p.max_bounces = (int)max_bounces;
  p.user_seed = seed;
  p.accumulate = accumulate;
  p.accum = d_accum; p.per_sample = d_per_sample; p.hit_ids = d_hit_ids;
  p.counters = d_counters;
  int e = rt_launch_render(p, g.sm_count, g.d_workspace, g.workspace_bytes, stream, &g.last_launches);
  if (e) return fail("render kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  return 0;
}

}  // namespace

// ======================================================================= ABI
extern "C" {

int rt_gpu_init(int device) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g.ready && g.device == device) return 0;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail("no CUDA device available (%s); libraytracer_gpu has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail("device %d out of range (%d visible)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  g.device = device;
  g.sm_count = prop.multiProcessorCount;
  if (!g.stream) CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  if (!g.ev0) CUDA_TRY(cudaEventCreate(&g.ev0));
  if (!g.ev1) CUDA_TRY(cudaEventCreate(&g.ev1));
  g.ready = true;
  return 0;
}

void rt_gpu_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (auto &kv : g.scenes) release(kv.second);
  g.scenes.clear();
  cudaFree(g.d_accum); cudaFree(g.d_hit_ids); cudaFree(g.d_counters); cudaFree(g.d_workspace); cudaFree(g.d_texel_stage);
  drop_block_pool();
  cudaFree(g.d_image); cudaFree(g.d_image2);
  if (g.h_pinned) cudaFreeHost(g.h_pinned);
  if (g.ev0) cudaEventDestroy(g.ev0);
  if (g.ev1) cudaEventDestroy(g.ev1);
  if (g.stream) cudaStreamDestroy(g.stream);
  std::vector<Shader_Proc> pbr = g.pbr_procs;
  std::vector<Background_Proc> bg = g.bg_procs;
  g = State{};
  g.pbr_procs = pbr;
  g.bg_procs = bg;
}

char const *rt_gpu_last_error(void) { return g_error[0] ? g_error : g_error_shared; }

int rt_gpu_sm_count(void) { return ensure_init() ? 0 : g.sm_count; }

f64 rt_gpu_measure_fp32_issue(void) {
  if (ensure_init()) return 0;
  std::lock_guard<std::mutex> lock(g_mutex);
  return rt_measure_fp32_issue(g.sm_count, g.stream, nullptr);
}

isize rt_gpu_scene_device_bytes(Scene const *scene) {
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g.scenes.find(scene);
  return it == g.scenes.end() ? 0 : (isize)it->second.bytes;
}

isize rt_gpu_scene_upload_bytes(Scene const *scene) {
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g.scenes.find(scene);
  return it == g.scenes.end() ? 0 : (isize)it->second.h2d_bytes;
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
  if (ensure_init()) return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g.scenes.find(scene);
  if (it != g.scenes.end()) {
    cudaDeviceSynchronize();          // its blocks are reused below: no render may still read them
    release(it->second);
    g.scenes.erase(it);
  }
  SceneDev dev;
  return scene_on_device(scene, &dev);
}

void rt_gpu_scene_release(Scene const *scene) {
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g.scenes.find(scene);
  if (it == g.scenes.end()) return;
  cudaDeviceSynchronize();
  release(it->second);
  g.scenes.erase(it);
}

void rt_gpu_set_options(RT_GPU_Options const *options) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g.options = *options;
  if (g.options.slice_samples < 1) g.options.slice_samples = 64;
}

void rt_gpu_get_options(RT_GPU_Options *options) {
  std::lock_guard<std::mutex> lock(g_mutex);
  *options = g.options;
}

// ------------------------------------------------------------ device-pointer level
int rt_gpu_render_accum_device(Scene const *scene, isize width, isize height, isize sample_begin, isize sample_end,
                               isize max_bounces, u32 user_seed, i32 accumulate, f32 *d_accum, f32 *d_per_sample,
                               i32 *d_hit_ids, u64 *d_counters, void *stream) {
  if (ensure_init()) return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  return render_device_locked(scene, width, height, sample_begin, sample_end, max_bounces, user_seed, accumulate,
                              d_accum, d_per_sample, d_hit_ids, reinterpret_cast<unsigned long long *>(d_counters),
                              static_cast<cudaStream_t>(stream));
}

int rt_gpu_resolve_device(f32 const *d_accum, isize width, isize height, isize samples, u8 *d_pixels, isize stride,
                          i32 components, void *stream) {
  if (ensure_init()) return 1;
  if (samples < 1 || components < 3) return fail("resolve: samples >= 1 and components >= 3 required");
  int e = rt_launch_resolve(d_accum, (int)width, (int)height, (int)samples, d_pixels, (int)stride, components,
                            static_cast<cudaStream_t>(stream));
  if (e) return fail("resolve kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  g.last_launches++;
  return 0;
}

int rt_gpu_denoise_device(u8 const *d_src, u8 *d_dst, isize width, isize height, isize src_stride, isize dst_stride,
                          i32 components, void *stream) {
  if (ensure_init()) return 1;
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
  if (!g.d_accum || (size_t)n_floats > g.last_pixels * 3) return fail("read_accum: no render of that size has run");
  CUDA_TRY(cudaMemcpy(out, g.d_accum, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int rt_gpu_read_hit_ids(i32 *out, isize n_pixels) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!g.d_hit_ids || !g.last_has_hit_ids || (size_t)n_pixels > g.last_pixels)
    return fail("read_hit_ids: the last render did not keep hit ids (RT_GPU_Options.keep_hit_ids)");
  CUDA_TRY(cudaMemcpy(out, g.d_hit_ids, (size_t)n_pixels * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int rt_gpu_read_counters(u64 out[8]) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!g.d_counters) return fail("read_counters: no render has run");
  CUDA_TRY(cudaMemcpy(out, g.d_counters, 8 * sizeof(u64), cudaMemcpyDeviceToHost));
  return 0;
}

int rt_gpu_last_launches(void) { return g.last_launches; }

void rt_gpu_stage_profile_enable(i32 on) {
  std::lock_guard<std::mutex> lock(g_mutex);
  rt_stage_profile_enable(on);
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

// ------------------------------------------------------- reference entry points
static int render_owner(Rendering_Context *ctx, isize n_chunks) {
  if (ensure_init()) return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  const Image &im = ctx->image;
  if (im.components < 3 || im.pixel_type != PT_u8 || !im.pixels.data) return fail("render: image must be u8 with >= 3 components");
  if (ctx->samples < 1) return fail("render: samples must be >= 1");
  CUDA_TRY(cudaSetDevice(g.device));
  const size_t n_pixels = (size_t)im.width * (size_t)im.height;
  const size_t image_bytes = (size_t)im.stride * (size_t)im.height * (size_t)im.components;

  size_t have = g.accum_floats * sizeof(float);
  if (grow(reinterpret_cast<void **>(&g.d_accum), &have, n_pixels * 3 * sizeof(float))) return 1;
  g.accum_floats = have / sizeof(float);
  have = g.hit_pixels * sizeof(int);
  if (grow(reinterpret_cast<void **>(&g.d_hit_ids), &have, n_pixels * sizeof(int))) return 1;
  g.hit_pixels = have / sizeof(int);
  if (!g.d_counters) CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&g.d_counters), 8 * sizeof(unsigned long long)));
  if (grow(reinterpret_cast<void **>(&g.d_image), &g.image_bytes, image_bytes)) return 1;
  if (grow_pinned(image_bytes)) return 1;

  const RT_GPU_Options opt = g.options;
  const isize s_begin = opt.sample_begin;
  const isize s_end = opt.sample_end > 0 ? opt.sample_end : ctx->samples;
  const isize slice = opt.slice_samples > 0 ? opt.slice_samples : 64;

  g.last_launches = 0;
  g.last_pixels = n_pixels;
  g.last_has_hit_ids = opt.keep_hit_ids != 0;
  CUDA_TRY(cudaMemsetAsync(g.d_counters, 0, 8 * sizeof(unsigned long long), g.stream));
  // rows padded by stride keep whatever the caller had there
  if (im.stride != im.width || im.components != 3) {
    memcpy(g.h_pinned, im.pixels.data, image_bytes);
    CUDA_TRY(cudaMemcpyAsync(g.d_image, g.h_pinned, image_bytes, cudaMemcpyHostToDevice, g.stream));
  }

  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  const isize n_slices = (s_end - s_begin + slice - 1) / slice;
  for (isize k = 0; k < n_slices; k++) {
    isize a = s_begin + k * slice, b = a + slice < s_end ? a + slice : s_end;
    if (render_device_locked(ctx->scene, im.width, im.height, a, b, ctx->max_bounces, opt.user_seed, k > 0,
                             g.d_accum, nullptr, (opt.keep_hit_ids && k == 0) ? g.d_hit_ids : nullptr,
                             g.d_counters, g.stream))
      return 1;
    if (n_slices > 1 && (k % 4 == 3)) {
      // progress for the host's bar (driver.c:810-818); never reaches n_chunks early
      CUDA_TRY(cudaStreamSynchronize(g.stream));
      isize progress = n_chunks * (k + 1) / n_slices;
      if (progress >= n_chunks) progress = n_chunks - 1;
      if (progress > ctx->_current_chunk) ctx->_current_chunk = (i32)progress;
    }
  }
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  int e = rt_launch_resolve(g.d_accum, (int)im.width, (int)im.height, (int)ctx->samples, g.d_image, (int)im.stride,
                            im.components, g.stream);
  if (e) return fail("resolve kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
  g.last_launches++;
  CUDA_TRY(cudaMemcpyAsync(g.h_pinned, g.d_image, image_bytes, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  memcpy(im.pixels.data, g.h_pinned, image_bytes);
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev0, g.ev1);
  g.last_kernel_ms = ms;
  return 0;
}

void render_thread_proc(Rendering_Context *ctx) {
  const isize chunks_x = (ctx->image.width + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  const isize chunks_y = (ctx->image.height + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  const isize n_chunks = chunks_x * chunks_y;
  // raytracer.c:620: claim a chunk; whoever gets chunk 0 launches the whole frame
  i32 c = __atomic_fetch_add(const_cast<i32 *>(&ctx->_current_chunk), 1, __ATOMIC_SEQ_CST);
  if (c == 0) {
    if (render_owner(ctx, n_chunks)) fprintf(stderr, "render_thread_proc: %s\n", rt_gpu_last_error());
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

void lightmap_bake(Image const *, Scene const *, isize) {
  // Exported by the reference (raytracer.h:56) but never called;
  // out of scope for the GPU path, reported instead of silently ignored.
  fail("lightmap_bake is not implemented on the GPU path");
  fprintf(stderr, "lightmap_bake: %s\n", rt_gpu_last_error());
}

void denoise_image(Image const *src, Image const *dst, isize n_threads) {
  (void)n_threads;
  auto run = [&]() -> int {
    if (ensure_init()) return 1;
    std::lock_guard<std::mutex> lock(g_mutex);
    if (src->pixels.data == dst->pixels.data) return fail("denoise: src and dst must differ (reference denoiser.c:130)");
    if (src->width != dst->width || src->height != dst->height || src->components != dst->components)
      return fail("denoise: src and dst shapes differ (reference denoiser.c:131-132)");
    CUDA_TRY(cudaSetDevice(g.device));
    const size_t src_bytes = (size_t)src->stride * (size_t)src->height * (size_t)src->components;
    const size_t dst_bytes = (size_t)dst->stride * (size_t)dst->height * (size_t)dst->components;
    if (grow(reinterpret_cast<void **>(&g.d_image), &g.image_bytes, src_bytes)) return 1;
    if (grow(reinterpret_cast<void **>(&g.d_image2), &g.image2_bytes, dst_bytes)) return 1;
    if (grow_pinned(src_bytes > dst_bytes ? src_bytes : dst_bytes)) return 1;
    g.last_launches = 0;
    memcpy(g.h_pinned, src->pixels.data, src_bytes);
    CUDA_TRY(cudaMemcpyAsync(g.d_image, g.h_pinned, src_bytes, cudaMemcpyHostToDevice, g.stream));
    if (dst->stride != dst->width) CUDA_TRY(cudaMemsetAsync(g.d_image2, 0, dst_bytes, g.stream));
    CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
    int e = rt_launch_denoise(g.d_image, g.d_image2, (int)src->width, (int)src->height, (int)src->stride, (int)dst->stride,
                              src->components, g.stream);
    if (e) return fail("denoise kernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
    g.last_launches++;
    CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
    CUDA_TRY(cudaMemcpyAsync(g.h_pinned, g.d_image2, dst_bytes, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    if (dst->stride == dst->width) {
      memcpy(dst->pixels.data, g.h_pinned, dst_bytes);
    } else {
      for (isize y = 0; y < dst->height; y++)
        memcpy(dst->pixels.data + (size_t)y * (size_t)dst->stride * (size_t)dst->components,
               g.h_pinned + (size_t)y * (size_t)dst->stride * (size_t)dst->components,
               (size_t)dst->width * (size_t)dst->components);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    g.last_kernel_ms = ms;
    return 0;
  };
  if (run()) fprintf(stderr, "denoise_image: %s\n", rt_gpu_last_error());
}

}  // extern "C"
