// nss_eas_kernel.cu -- sm_100a kernel for the optical-Cherenkov air-shower stage.
// One warp per event (see nss_eas_core.cuh for the per-event algorithm).
//
// Replaces: EAS.__call__ (src/nuspacesim/simulation/eas_optical/eas.py:83-138),
//           CphotAng.__call__/run (cphotang.py:623-724), zsteps.cpp:9-45.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/nss_b200.h"
#include "nss_eas_core.cuh"
#include "nss_eas_tables.h"

namespace {

__constant__ nss::EasTables c_tables;

struct Ops32 {
  static constexpr int W = 32;
  static constexpr unsigned FULL = 0xffffffffu;
  static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
  static __device__ __forceinline__ float scan_f(float v) {
    const int l = lane();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float t = __shfl_up_sync(FULL, v, d);
      if (l >= d) v += t;
    }
    return v;
  }
  static __device__ __forceinline__ double scan_d(double v) {
    const int l = lane();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double t = __shfl_up_sync(FULL, v, d);
      if (l >= d) v += t;
    }
    return v;
  }
  static __device__ __forceinline__ double sum_d(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
  }
  static __device__ __forceinline__ int sum_i(int v) { return __reduce_add_sync(FULL, v); }
  static __device__ __forceinline__ int max_i(int v) { return __reduce_max_sync(FULL, v); }
  static __device__ __forceinline__ int min_i(int v) { return __reduce_min_sync(FULL, v); }
  static __device__ __forceinline__ bool any(bool b) { return __any_sync(FULL, b); }
  static __device__ __forceinline__ double shift_up_d(double v, double carry) {
    double t = __shfl_up_sync(FULL, v, 1);
    return lane() == 0 ? carry : t;
  }
  static __device__ __forceinline__ double last_d(double v) { return __shfl_sync(FULL, v, 31); }
  // payload of the lane holding the largest key; ties -> smallest idx
  static __device__ __forceinline__ void argmax_first(float key, int idx, float payload,
                                                      float& out) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      float k2 = __shfl_xor_sync(FULL, key, d);
      int i2 = __shfl_xor_sync(FULL, idx, d);
      float p2 = __shfl_xor_sync(FULL, payload, d);
      if (k2 > key || (k2 == key && i2 < idx)) { key = k2; idx = i2; payload = p2; }
    }
    out = payload;
  }
};

// mode 0: EAS.__call__ (altDec mask + numPEs/cos tail); mode 1: CphotAng.__call__ only.
template <int MODE>
__global__ void __launch_bounds__(128)
eas_kernel(nss::EasParams P, const double* __restrict__ beta, const double* __restrict__ alt,
           const double* __restrict__ eshow, const uint8_t* __restrict__ mask, long long n,
           double* __restrict__ out0, double* __restrict__ out1, double* __restrict__ dphots,
           double* __restrict__ cang, unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  nss::EasStats st = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = warp; i < n; i += nwarps) {
    const double b = beta[i], a = alt[i], e = eshow[i];
    double d = 0.0, c = (MODE == 0) ? 1.5 : 0.0;  // eas.py:106-107 defaults
    bool run = (mask == nullptr) || (mask[i] != 0);
    if (MODE == 0) run = run && !(a < 0.0 || a > 20.0);  // eas.py:102-103
    if (run) nss::eas_event<Ops32>(c_tables, P, b, a, e, d, c, stats ? &st : nullptr, lane == 0);
    if (lane == 0) {
      if (MODE == 0) {
        double npe, cc;
        nss::eas_tail(P, d, c, npe, cc);
        out0[i] = npe;
        out1[i] = cc;
        if (dphots) dphots[i] = d;
        if (cang) cang[i] = c;
      } else {
        out0[i] = d;
        out1[i] = c;
      }
    }
  }
  if (stats && lane == 0 && st.n_events) {
    atomicAdd(&stats[0], st.n_events);
    atomicAdd(&stats[1], st.n_raw);
    atomicAdd(&stats[2], st.n_val);
    atomicAdd(&stats[3], st.J);
    atomicAdd(&stats[4], st.JNe);
    atomicAdd(&stats[5], st.n_exp);
    atomicAdd(&stats[6], st.n_val_ne);
  }
}

bool g_tables_ready = false;
int g_grid_blocks[2] = {0, 0};

// optional per-launch CUDA-event timing of the EAS kernel (nss_eas_profile_*)
bool g_profile = false;
struct ProfRec { cudaEvent_t a, b; };
ProfRec* g_prof = nullptr;
int g_prof_n = 0, g_prof_cap = 0;

template <int MODE>
int launch(const NssParams* p, const double* beta, const double* alt, const double* eshow,
           const uint8_t* mask, int64_t n, double* o0, double* o1, double* dph, double* cang,
           unsigned long long* stats, void* stream) {
  if (!p) return NSS_E_NULL;
  if (n < 0) return NSS_E_BADARG;
  if (n == 0) return 0;
  if (!beta || !alt || !eshow || !o0 || !o1) return NSS_E_NULL;
  cudaError_t err;
  if (!g_tables_ready) {
    nss::EasTables T;
    nss_eas_tables_init(&T);
    err = cudaMemcpyToSymbol(c_tables, &T, sizeof(T));
    if (err != cudaSuccess) return (int)err;
    g_tables_ready = true;
  }
  if (g_grid_blocks[MODE] == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eas_kernel<MODE>, 128, 0);
    if (err != cudaSuccess) return (int)err;
    g_grid_blocks[MODE] = sms * (per_sm > 0 ? per_sm : 1);
  }
  nss::EasParams P;
  P.detector_altitude = p->detector_altitude;
  P.area = p->telescope_effective_area;
  P.qe = p->quantum_efficiency;
  P.pe_threshold = p->photo_electron_threshold;
  P.has_cloud = p->has_cloud;
  P.cloud_top = (float)p->cloud_top;
  long long need = (n + 3) / 4;  // 4 warps per block
  int blocks = (int)(need < g_grid_blocks[MODE] ? need : g_grid_blocks[MODE]);
  ProfRec* rec = nullptr;
  if (g_profile) {
    if (g_prof_n == g_prof_cap) {
      g_prof_cap = g_prof_cap ? 2 * g_prof_cap : 256;
      g_prof = (ProfRec*)realloc(g_prof, sizeof(ProfRec) * g_prof_cap);
    }
    rec = &g_prof[g_prof_n++];
    cudaEventCreate(&rec->a);
    cudaEventCreate(&rec->b);
    cudaEventRecord(rec->a, (cudaStream_t)stream);
  }
  eas_kernel<MODE><<<blocks, 128, 0, (cudaStream_t)stream>>>(P, beta, alt, eshow, mask, n, o0, o1,
                                                             dph, cang, stats);
  if (rec) cudaEventRecord(rec->b, (cudaStream_t)stream);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int nss_eas_optical(const NssParams* p, const double* beta, const double* altDec,
                               const double* showerEnergy, const uint8_t* mask, int64_t n,
                               double* numPEs, double* costhetaChEff, double* dphots,
                               double* cang, unsigned long long* stats, void* stream) {
  return launch<0>(p, beta, altDec, showerEnergy, mask, n, numPEs, costhetaChEff, dphots, cang,
                   stats, stream);
}

extern "C" int nss_cphotang(const NssParams* p, const double* betaE, const double* alt,
                            const double* Eshow100PeV, int64_t n, double* dphots, double* cang,
                            unsigned long long* stats, void* stream) {
  return launch<1>(p, betaE, alt, Eshow100PeV, nullptr, n, dphots, cang, nullptr, nullptr, stats,
                   stream);
}

extern "C" int nss_eas_profile_enable(int on) {
  g_profile = on != 0;
  return 0;
}

extern "C" int nss_eas_profile_collect(double* total_ms, int64_t* launches) {
  double tot = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(g_prof[i].b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b);
    cudaEventDestroy(g_prof[i].a);
    cudaEventDestroy(g_prof[i].b);
    if (e != cudaSuccess) { g_prof_n = 0; return (int)e; }
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = g_prof_n;
  g_prof_n = 0;
  return 0;
}
