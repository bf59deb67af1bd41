// rt_multi.cu — multi-GPU plumbing of libraytracer_gpu.so.
//
// The reference parallelises one frame over host threads that pull 32x32 chunks from an atomic counter
// (raytracer.c:619-627); its samples are independent (raytracer.c:687-696).  Here a frame is split over GPUs,
// either way the reference's structure offers:
//   * by sample range — every device renders all pixels for a range of sample indices, in multiples of the
//     reference's 8-sample jitter batch (raytracer.c:641-697): perfect load balance, the cross-device sum re-orders
//     the f32 additions of a pixel (ranks are added in fixed order, so the frame is still deterministic);
//   * by chunk — the reference's 32x32 chunks dealt round-robin: every pixel is summed on one device in sample
//     order, the cross-device "sum" only adds zeros, and the frame is bit-identical to the single-GPU one.
// The per-device accumulators are combined on the first device: by rt_reduce_resolve_kernel reading the peers'
// buffers over NVLink (default), or by ncclReduce (RT_GPU_REDUCE_NCCL; libnccl is loaded on demand so that the
// library itself does not depend on it).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>      // types only: the functions are resolved with dlsym when RT_GPU_REDUCE_NCCL is first used

#include <cstdio>
#include <cstring>
#include <string>

#include "rt_state.h"

namespace rt {

int enable_peers() {
  if (g.peers_enabled || g.devs.size() < 2) { g.peers_enabled = true; return 0; }
  const int root = g.devs[0].id;
  for (size_t k = 1; k < g.devs.size(); k++) {
    const int peer = g.devs[k].id;
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, root, peer));
    if (!can) return fail("device %d cannot map device %d's memory (no NVLink/P2P path): use RT_GPU_REDUCE_NCCL", root, peer);
    CUDA_TRY(cudaSetDevice(root));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess(%d): %s", peer, cudaGetErrorString(e));
    cudaGetLastError();
    CUDA_TRY(cudaSetDevice(peer));          // the other direction too: scene fan-out copies read device 0's arena
    e = cudaDeviceEnablePeerAccess(root, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess(%d): %s", root, cudaGetErrorString(e));
    cudaGetLastError();
  }
  CUDA_TRY(cudaSetDevice(root));
  g.peers_enabled = true;
  return 0;
}

// ---------------------------------------------------------------- NCCL, loaded on demand
// Only the handful of entry points the reduce needs; signatures from nccl.h (2.x, stable since 2.0).
namespace {
struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::vector<ncclComm_t> comms;
} nccl;

int nccl_load() {
  if (nccl.lib) return 0;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    nccl.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (nccl.lib) break;
  }
  if (!nccl.lib) return fail("RT_GPU_REDUCE_NCCL: libnccl.so.2 not found (%s)", dlerror());
  auto sym = [&](const char *n) { return dlsym(nccl.lib, n); };
  nccl.CommInitAll = reinterpret_cast<decltype(nccl.CommInitAll)>(sym("ncclCommInitAll"));
  nccl.CommDestroy = reinterpret_cast<decltype(nccl.CommDestroy)>(sym("ncclCommDestroy"));
  nccl.GroupStart = reinterpret_cast<decltype(nccl.GroupStart)>(sym("ncclGroupStart"));
  nccl.GroupEnd = reinterpret_cast<decltype(nccl.GroupEnd)>(sym("ncclGroupEnd"));
  nccl.Reduce = reinterpret_cast<decltype(nccl.Reduce)>(sym("ncclReduce"));
  nccl.GetErrorString = reinterpret_cast<decltype(nccl.GetErrorString)>(sym("ncclGetErrorString"));
  if (!nccl.CommInitAll || !nccl.CommDestroy || !nccl.GroupStart || !nccl.GroupEnd || !nccl.Reduce || !nccl.GetErrorString) {
    dlclose(nccl.lib);
    nccl.lib = nullptr;
    return fail("RT_GPU_REDUCE_NCCL: libnccl lacks an expected symbol");
  }
  return 0;
}
}  // namespace

#define NCCL_TRY(expr)                                                                              \
  do {                                                                                              \
    ncclResult_t r_ = (expr);                                                                       \
    if (r_ != ncclSuccess) return fail("%s failed: %s", #expr, nccl.GetErrorString(r_));                      \
  } while (0)

int nccl_prepare() {
  if (nccl_load()) return 1;
  const int n = (int)g.devs.size();
  if ((int)nccl.comms.size() != n) {
    for (ncclComm_t c : nccl.comms) nccl.CommDestroy(c);
    nccl.comms.assign((size_t)n, nullptr);
    std::vector<int> ids;
    for (Device &d : g.devs) ids.push_back(d.id);
    NCCL_TRY(nccl.CommInitAll(nccl.comms.data(), n, ids.data()));
    cudaSetDevice(g.devs[0].id);
  }
  return 0;
}

int nccl_reduce_to_first(size_t n_floats) {
  if (nccl_prepare()) return 1;
  const int n = (int)g.devs.size();
  NCCL_TRY(nccl.GroupStart());
  for (int k = 0; k < n; k++) {
    Device &d = g.devs[(size_t)k];
    ncclResult_t r = nccl.Reduce(d.d_accum, g.devs[0].d_accum, n_floats, ncclFloat, ncclSum, 0, nccl.comms[(size_t)k], d.stream);
    if (r != ncclSuccess) { nccl.GroupEnd(); return fail("ncclReduce failed: %s", nccl.GetErrorString(r)); }
  }
  NCCL_TRY(nccl.GroupEnd());
  return 0;
}

void nccl_shutdown() {
  for (ncclComm_t c : nccl.comms) if (c) nccl.CommDestroy(c);
  nccl.comms.clear();
}

// ---------------------------------------------------------------- CUDA IPC (one process per GPU)
namespace {
struct IpcMapping { unsigned char handle[64]; void *ptr; };
std::vector<IpcMapping> g_ipc;
}

void ipc_close_all() {
  for (IpcMapping &m : g_ipc) cudaIpcCloseMemHandle(m.ptr);
  g_ipc.clear();
  cudaGetLastError();
}

}  // namespace rt

using namespace rt;

extern "C" {

// Sample split: batches of 8 samples (the reference's jitter batch, raytracer.c:641-697) dealt as evenly as possible.
void rt_gpu_shard_samples(i32 rank, i32 world, i32 samples, i32 *begin, i32 *end) {
  if (world < 1 || rank < 0 || rank >= world || samples < 0) { *begin = *end = 0; return; }
  const i64 batches = ((i64)samples + 7) / 8;
  i64 lo = batches * rank / world * 8, hi = batches * (rank + 1) / world * 8;
  if (lo > samples) lo = samples;
  if (hi > samples) hi = samples;
  *begin = (i32)lo;
  *end = (i32)hi;
}

i32 rt_gpu_shard_mode(i32 samples, i32 world, i32 requested_mode) {
  if (world <= 1) return RT_GPU_SPLIT_SAMPLES;
  if (requested_mode == RT_GPU_SPLIT_SAMPLES || requested_mode == RT_GPU_SPLIT_CHUNKS) return requested_mode;
  const i64 batches = ((i64)samples + 7) / 8;
  // even sample ranges balance perfectly; anything else (spp < 8 x world, or a remainder) is dealt by chunk
  return (samples % 8 == 0 && batches % world == 0) ? RT_GPU_SPLIT_SAMPLES : RT_GPU_SPLIT_CHUNKS;
}

i32 rt_gpu_shard_chunks(i32 rank, i32 world, isize width, isize height) {
  if (world < 1 || rank < 0 || rank >= world || width < 1 || height < 1) return 0;
  const i64 chunks = ((i64)(width + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE) * ((i64)(height + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE);
  return (i32)(chunks > rank ? (chunks - rank + world - 1) / world : 0);
}

int rt_gpu_ipc_export(void const *d_ptr, u8 handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
  memcpy(handle_out, &h, 64);
  return 0;
}

int rt_gpu_ipc_open(u8 const handle[64], void **d_ptr_out) {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (IpcMapping &m : g_ipc)
    if (!memcmp(m.handle, handle, 64)) { *d_ptr_out = m.ptr; return 0; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void *p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  IpcMapping m;
  memcpy(m.handle, handle, 64);
  m.ptr = p;
  g_ipc.push_back(m);
  *d_ptr_out = p;
  return 0;
}

}  // extern "C"
