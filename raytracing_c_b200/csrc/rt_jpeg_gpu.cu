// rt_jpeg_gpu.cu — texture decode on the device.
//
// The reference decodes every glTF image on the host, one after the other (driver.c:620-626, stb_image).  Here a
// host may hand the scene upload COMPRESSED baseline JPEGs (Image.pixel_type == PT_RT_JPEG_BYTES, rt_base.h): nvJPEG
// decodes them on the upload's copy stream into an interleaved RGB block in device memory, and the same repack
// kernel that serves host-decoded textures turns that into the RGBA8 texels the samplers read.  Texels never exist
// in host memory and never cross PCIe decoded (a 2048^2 texture is ~2 MB compressed, 12.6 MB decoded).
// libnvjpeg is resolved with dlopen on first use, so the library has no load-time dependency on it.
//
// PARITY: JPEG decoders are allowed to differ in IDCT rounding and chroma upsampling; nvJPEG's output is NOT
// byte-identical to host/rt_jpeg.c (tests/test_gpu_jpeg.py measures the difference).  The host decoder stays the
// parity path; device decode is an opt-in for time-to-image.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>      // types only

#include <mutex>
#include <vector>

#include "rt_state.h"

namespace rt {

namespace {
struct NvJpeg {
  void *lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t *) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*StateCreate)(nvjpegHandle_t, nvjpegJpegState_t *) = nullptr;
  nvjpegStatus_t (*StateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char *, size_t, int *, nvjpegChromaSubsampling_t *, int *, int *) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char *, size_t, nvjpegOutputFormat_t, nvjpegImage_t *, cudaStream_t) = nullptr;
  nvjpegHandle_t handle = nullptr;
  // one decoder state per image in flight: a state's pinned staging is still being read by the stream after
  // nvjpegDecode returns, so it is not reused before jpeg_begin_batch has drained that stream
  std::vector<nvjpegJpegState_t> states;
  size_t next_state = 0;
} nj;

int nvjpeg_load() {
  if (nj.handle) return 0;
  if (!nj.lib) {
    for (const char *name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      nj.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (nj.lib) break;
    }
    if (!nj.lib) return fail("device JPEG decode: libnvjpeg.so.12 not found (%s)", dlerror());
    auto sym = [&](const char *n) { return dlsym(nj.lib, n); };
    nj.CreateSimple = reinterpret_cast<decltype(nj.CreateSimple)>(sym("nvjpegCreateSimple"));
    nj.Destroy = reinterpret_cast<decltype(nj.Destroy)>(sym("nvjpegDestroy"));
    nj.StateCreate = reinterpret_cast<decltype(nj.StateCreate)>(sym("nvjpegJpegStateCreate"));
    nj.StateDestroy = reinterpret_cast<decltype(nj.StateDestroy)>(sym("nvjpegJpegStateDestroy"));
    nj.GetImageInfo = reinterpret_cast<decltype(nj.GetImageInfo)>(sym("nvjpegGetImageInfo"));
    nj.Decode = reinterpret_cast<decltype(nj.Decode)>(sym("nvjpegDecode"));
    if (!nj.CreateSimple || !nj.Destroy || !nj.StateCreate || !nj.StateDestroy || !nj.GetImageInfo || !nj.Decode)
      return fail("device JPEG decode: libnvjpeg lacks an expected symbol");
  }
  if (nj.CreateSimple(&nj.handle) != NVJPEG_STATUS_SUCCESS) { nj.handle = nullptr; return fail("nvjpegCreateSimple failed"); }
  return 0;
}
}  // namespace

// Call once before the decodes of one upload: waits for the decodes of the previous one, then hands out states from the start.
int jpeg_begin_batch(cudaStream_t stream) {
  if (nj.next_state) CUDA_TRY(cudaStreamSynchronize(stream));
  nj.next_state = 0;
  return 0;
}

// Decodes `bytes` into interleaved RGB8 at d_rgb (pitch 3 * width) on `stream`; width/height must match the header.
int jpeg_decode_device(const unsigned char *bytes, size_t len, int width, int height, unsigned char *d_rgb, cudaStream_t stream) {
  if (nvjpeg_load()) return 1;
  int n_comp = 0, widths[NVJPEG_MAX_COMPONENT] = {0}, heights[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t sub;
  if (nj.GetImageInfo(nj.handle, bytes, len, &n_comp, &sub, widths, heights) != NVJPEG_STATUS_SUCCESS)
    return fail("device JPEG decode: nvjpegGetImageInfo rejected the stream");
  if (widths[0] != width || heights[0] != height) return fail("device JPEG decode: header says %dx%d, the Image %dx%d", widths[0], heights[0], width, height);
  nvjpegImage_t out{};
  out.channel[0] = d_rgb;
  out.pitch[0] = (size_t)width * 3;
  if (nj.next_state == nj.states.size()) {
    nvjpegJpegState_t st = nullptr;
    if (nj.StateCreate(nj.handle, &st) != NVJPEG_STATUS_SUCCESS) return fail("nvjpegJpegStateCreate failed");
    nj.states.push_back(st);
  }
  nvjpegStatus_t st = nj.Decode(nj.handle, nj.states[nj.next_state++], bytes, len, NVJPEG_OUTPUT_RGBI, &out, stream);
  if (st != NVJPEG_STATUS_SUCCESS) return fail("device JPEG decode: nvjpegDecode failed (%d)", (int)st);
  return 0;
}

void jpeg_shutdown() {
  for (nvjpegJpegState_t st : nj.states) nj.StateDestroy(st);
  nj.states.clear();
  nj.next_state = 0;
  if (nj.handle) nj.Destroy(nj.handle);
  nj.handle = nullptr;
}

}  // namespace rt
