// nss_microbench.cu -- on-box measurement of the two compute ceilings the EAS kernel is
// bounded by (SURVEY.md 8(d) "Peaks"): the FP32 FMA issue rate and the MUFU (SFU) issue
// rate.  MEASURED_PEAKS.json has HBM and bf16-GEMM numbers only, so bench.py measures
// these here, with CUDA events on the launching stream, and reports roofline fractions
// "of measured".  Not on the product path.
#include <cuda_runtime.h>

#include "../../include/nss_b200.h"

namespace {

constexpr int MB_THREADS = 256;
constexpr int MB_ILP = 8;

// 8 independent FMA chains per thread, 3-register form (the multiplier is loaded at run time).
__global__ void __launch_bounds__(MB_THREADS)
fma_chain_kernel(float a, float b, int iters, float* __restrict__ sink) {
  float x[MB_ILP];
#pragma unroll
  for (int k = 0; k < MB_ILP; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < MB_ILP; ++k) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MB_ILP; ++k) s += x[k];
  if (s == 12345.678f) sink[0] = s;  // never true; keeps the chains alive
}

// 8 independent MUFU.EX2 chains per thread.
__global__ void __launch_bounds__(MB_THREADS)
mufu_chain_kernel(int iters, float* __restrict__ sink) {
  float x[MB_ILP];
#pragma unroll
  for (int k = 0; k < MB_ILP; ++k) x[k] = -(float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < MB_ILP; ++k) {
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[k]));
      x[k] = -y;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MB_ILP; ++k) s += x[k];
  if (s == 12345.678f) sink[0] = s;
}

template <class F>
int time_best(F launch, cudaStream_t s, int reps, float* best_ms) {
  cudaEvent_t a, b;
  cudaError_t e = cudaEventCreate(&a);
  if (e != cudaSuccess) return (int)e;
  e = cudaEventCreate(&b);
  if (e != cudaSuccess) return (int)e;
  float best = 1e30f;
  for (int r = 0; r < reps + 2; ++r) {  // two warm-ups
    cudaEventRecord(a, s);
    launch();
    cudaEventRecord(b, s);
    e = cudaEventSynchronize(b);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (r >= 2 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *best_ms = best;
  return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

}  // namespace

extern "C" int nss_microbench_peaks(double* fp32_tflops, double* mufu_tops, void* stream) {
  if (!fp32_tflops || !mufu_tops) return NSS_E_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* sink = nullptr;
  cudaError_t e = cudaMalloc((void**)&sink, 256);
  if (e != cudaSuccess) return (int)e;
  const int blocks = sms * 8, iters = 1 << 14;
  float ms = 0.f;
  int rc = time_best([&] { fma_chain_kernel<<<blocks, MB_THREADS, 0, s>>>(1.0000001f, 1e-9f, iters, sink); },
                     s, 5, &ms);
  if (rc == 0) {
    *fp32_tflops = 2.0 * (double)blocks * MB_THREADS * MB_ILP * (double)iters / (ms * 1e-3) * 1e-12;
    rc = time_best([&] { mufu_chain_kernel<<<blocks, MB_THREADS, 0, s>>>(iters / 4, sink); }, s, 5, &ms);
    if (rc == 0)
      *mufu_tops = (double)blocks * MB_THREADS * MB_ILP * (double)(iters / 4) / (ms * 1e-3) * 1e-12;
  }
  cudaFree(sink);
  return rc;
}
