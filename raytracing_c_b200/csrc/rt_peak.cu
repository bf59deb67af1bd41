// rt_peak.cu — measures the FP32 issue ceiling the trace kernel is bounded by.
//
// The render kernel is built without FMA contraction (bit-exact parity), so its
// ceiling is the rate at which the SMs issue plain FMUL/FADD lane-operations:
// nominally 148 SMs x 128 lanes x clock.  This kernel runs 8 independent
// multiply-then-add chains per thread (FMUL + FADD, not fused because of
// -fmad=false) and reports lane-operations per second under whatever clock the
// power manager grants, which is the roofline denominator bench.py uses
// (MEASURED_PEAKS.json has no FP32 entry).
#include <cuda_runtime.h>

#include "rt_gpu_internal.h"

__global__ void __launch_bounds__(256) rt_peak_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    #pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
      x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
  }
  float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) out[0] = s;     // keeps the chains alive
}

double rt_measure_fp32_issue(int sm_count, cudaStream_t stream, double *ms_out) {
  float *d = nullptr;
  if (cudaMalloc(&d, 4) != cudaSuccess) return 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096, blocks = sm_count * 8, threads = 256;
  rt_peak_kernel<<<blocks, threads, 0, stream>>>(d, 64, 0.999f, 0.001f);       // warm-up
  cudaEventRecord(e0, stream);
  for (int r = 0; r < 4; r++) rt_peak_kernel<<<blocks, threads, 0, stream>>>(d, iters, 0.999f, 0.001f);
  cudaEventRecord(e1, stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (ms_out) *ms_out = ms;
  double ops = 4.0 * (double)blocks * threads * (double)iters * 8 * 8 * 2;    // FMUL + FADD per chain step
  return ops / (ms * 1e-3);
}
