// rt_gpu_internal.h — helpers shared inside libraytracer_gpu.so only.
#pragma once
#include <cuda_runtime.h>

// lane-operations per second of non-fused FMUL/FADD (see rt_peak.cu)
double rt_measure_fp32_issue(int sm_count, cudaStream_t stream, double *ms_out);
