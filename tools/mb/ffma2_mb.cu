// Microbenchmark (development aid): issue cost of packed FFMA2 vs scalar FFMA on sm_100a,
// alone and interleaved with ALU (FMNMX) and MUFU work.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float a, float b, int iters, float* out) {
  float2 x[8];
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); m[i] = threadIdx.x * 1e-3f + i; }
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // scalar FFMA x2
        x[i].x = fmaf(x[i].x, a, b);
        x[i].y = fmaf(x[i].y, a, b);
      } else if (MODE == 1) {  // packed
        x[i] = __ffma2_rn(x[i], A, B);
      } else if (MODE == 2) {  // packed + 1 FMNMX
        x[i] = __ffma2_rn(x[i], A, B);
        m[i] = fminf(m[i], x[i].x);
      } else if (MODE == 3) {  // scalar x2 + 1 FMNMX
        x[i].x = fmaf(x[i].x, a, b);
        x[i].y = fmaf(x[i].y, a, b);
        m[i] = fminf(m[i], x[i].x);
      } else if (MODE == 4) {  // packed + 2 FMNMX
        x[i] = __ffma2_rn(x[i], A, B);
        m[i] = fminf(m[i], x[i].x);
        m[i] = fmaxf(m[i], x[i].y);
      } else if (MODE == 5) {  // 2 packed + 1 MUFU
        x[i] = __ffma2_rn(x[i], A, B);
        x[i] = __ffma2_rn(x[i], B, A);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
      } else if (MODE == 6) {  // 4 scalar + 1 MUFU
        x[i].x = fmaf(x[i].x, a, b);
        x[i].y = fmaf(x[i].y, a, b);
        x[i].x = fmaf(x[i].x, b, a);
        x[i].y = fmaf(x[i].y, b, a);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
      } else if (MODE == 7) {  // f16x2 ex2 only
        unsigned h = __float_as_uint(m[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        m[i] = __uint_as_float(h);
      } else if (MODE == 8) {  // f32 ex2 only
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter_thread) {
  int sms = 148, blocks = sms * 8, iters = 4096;
  float* out;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(1.0001f, 0.5f, 16, out);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(1.0001f, 0.5f, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double warp_insts = (double)blocks * 8 * iters * 8 * ops_per_iter_thread;  // warps*iters*8 chains*insts
  double cyc = best * 1e-3 * 1.965e9 * sms * 4;  // SMSP-cycles at max clock
  printf("%-28s %8.3f ms  warp-inst/SMSP-cycle(@1965MHz) %.3f\n", name, best, warp_insts / cyc);
  cudaFree(out);
}

int main() {
  run<0>("2xFFMA", 2);
  run<1>("1xFFMA2", 1);
  run<2>("FFMA2+FMNMX", 2);
  run<3>("2xFFMA+FMNMX", 3);
  run<4>("FFMA2+2xFMNMX", 3);
  run<5>("2xFFMA2+MUFU", 3);
  run<6>("4xFFMA+MUFU", 5);
  run<7>("MUFU.EX2.f16x2", 1);
  run<8>("MUFU.EX2.f32", 1);
  return 0;
}
