// nss_eas_kernel.cu -- sm_100a kernels for the optical-Cherenkov air-shower stage.
//
//   eas_prep_kernel    one thread per event (FP64): per-event constants + column totals
//   eas_walk_kernel    one warp per event, persistent warps pulling events from a global
//                      counter (showers differ 10x in length); FP32 + MUFU; no shared memory
//   eas_finish_kernel  one thread per event (FP64): angle statistics, area, scaling, tail
//
// See nss_eas_core.cuh for the per-event algorithm.
//
// Replaces: EAS.__call__ (src/nuspacesim/simulation/eas_optical/eas.py:83-138),
//           CphotAng.__call__/run (cphotang.py:623-724), zsteps.cpp:9-45.
#include <cuda_runtime.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/nss_b200.h"
#include "nss_eas_core.cuh"
#include "nss_eas_tables.h"

namespace {

__constant__ nss::EasTables c_tables;

struct Ops32 {
  static constexpr int W = 32;
  static constexpr unsigned FULL = 0xffffffffu;
  static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
  // inclusive Hillis-Steele scan; the shuffle's own in-range predicate guards the add (no
  // lane-index compares: 10 instructions)
  static __device__ __forceinline__ float scan_f(float v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      asm volatile("{\n\t.reg .f32 t;\n\t.reg .pred p;\n\t"
          "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t"
          "@p add.rn.f32 %0, %0, t;\n\t}"
          : "+f"(v)
          : "r"(d));
    }
    return v;
  }
  static __device__ __forceinline__ double sum_d(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
  }
  static __device__ __forceinline__ float sum_f(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
  }
  static __device__ __forceinline__ float max_f(float v) {  // v >= 0: bit pattern is monotone
    return __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(v)));
  }
  static __device__ __forceinline__ int sum_i(int v) { return __reduce_add_sync(FULL, v); }
  static __device__ __forceinline__ int max_i(int v) { return __reduce_max_sync(FULL, v); }
  static __device__ __forceinline__ int min_i(int v) { return __reduce_min_sync(FULL, v); }
  static __device__ __forceinline__ bool any(bool b) { return __any_sync(FULL, b); }
  static __device__ __forceinline__ float lookup11(const float*, float lane_val, int idx) {
    return __shfl_sync(FULL, lane_val, idx);
  }
  static __device__ __forceinline__ float last_f(float v) { return __shfl_sync(FULL, v, 31); }
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
// profile 2: conex = (conex_rows, 6) float64 table, u_profile = one uniform per event picking
// the row (replaces np.random.randint, shower_properties.py:165)
__global__ void __launch_bounds__(256)
eas_prep_kernel(int mode, int profile, const double* __restrict__ beta,
                const double* __restrict__ alt, const double* __restrict__ eshow,
                const uint8_t* __restrict__ mask, const double* __restrict__ conex,
                int conex_rows, const double* __restrict__ u_profile,
                const double* __restrict__ cloud_ev, float cloud_run, long long n,
                nss::EasPrep* __restrict__ prep) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double a = alt[i];
    bool run = (mask == nullptr) || (mask[i] != 0);
    if (mode == 0) run = run && !(a < 0.0 || a > 20.0);  // eas.py:102-103
    const double* row = nullptr;
    if (profile == 2) {
      int r = (int)(u_profile[i] * (double)conex_rows);
      r = r < 0 ? 0 : (r >= conex_rows ? conex_rows - 1 : r);
      row = conex + 6 * (long long)r;
    }
    nss::EasPrep P;
    nss::eas_prep(c_tables, beta[i], a, eshow[i], run, profile, row,
                  cloud_ev ? (float)cloud_ev[i] : cloud_run, P);
    prep[i] = P;
  }
}

// threads per block of the walk: one warp = one event in flight, one warp per block (a block
// ends, and its slot is refilled, with its own event; 4-warp blocks were 4-5 % slower for the
// Gaisser-Hillas instantiation in round 2)
constexpr int walk_block(bool) { return 32; }

// Occupancy (measured, tools/ab_bench.sh, profiles/r02_ab_v20.log).  The instantiations
// without the work counters -- what compute() and the acceptance runner launch -- fit 72
// registers (Greisen: 8 bytes of spill) and run 28 one-warp blocks per SM: 1.183e8 events/s
// against 1.174e8 at 24 blocks (80 registers, no spill), 1.169e8 at 32 (64 registers, 36 bytes)
// and 1.127e8 at 20 (87 registers).  The instantiations WITH the counters (diagnostics: bench
// work counters, ncu captures) need 96 registers and stay at 20 blocks.
#ifndef NSS_LB_G0
#define NSS_LB_G0 28
#endif
#ifndef NSS_LB_G1
#define NSS_LB_G1 28
#endif
constexpr int walk_min_blocks(bool gh, bool stats) {
  return stats ? 20 : (gh ? NSS_LB_G1 : NSS_LB_G0);
}

// STATS: with the algorithmic-work counters (stats != nullptr) or with them compiled out
template <bool GH, bool STATS>
__global__ void __launch_bounds__(walk_block(GH), walk_min_blocks(GH, STATS))
eas_walk_kernel(nss::EasParams P, const nss::EasPrep* __restrict__ prep, long long n,
                nss::EasAcc* __restrict__ acc, unsigned long long* __restrict__ counter,
                unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  nss::EasStats st = {0, 0, 0, 0, 0, 0, 0, 0};
  for (;;) {
    unsigned long long i = 0;
    if (lane == 0) i = atomicAdd(counter, 1ull);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= (unsigned long long)n) break;
    // first 48 bytes: scalars every lane needs; ozA[g] stays in lane g's register
    nss::EasPrep E;
    {
      const float4* src = reinterpret_cast<const float4*>(prep + i);
      float4* dst = reinterpret_cast<float4*>(&E);
      dst[0] = __ldg(src);
      dst[1] = __ldg(src + 1);
      dst[2] = __ldg(src + 2);
    }
    const float ozA_lane =
        __ldg(reinterpret_cast<const float*>(prep + i) + 12 + (lane < 12 ? lane : 0));
    if (GH) {
      const float4* src = reinterpret_cast<const float4*>(&prep[i].gh);
      float4* dst = reinterpret_cast<float4*>(&E.gh);
      dst[0] = __ldg(src);
      dst[1] = __ldg(src + 1);
      dst[2] = __ldg(src + 2);
    }
    nss::EasAcc A;
    const float cloud_top = __shfl_sync(0xffffffffu, ozA_lane, 11);  // EasPrep::cloud
    nss::eas_walk<Ops32, GH, STATS>(c_tables, P, E, ozA_lane, cloud_top, A,
                                    STATS ? &st : nullptr, lane == 0);
    if (lane == 0) acc[i] = A;
  }
  if (STATS && stats && lane == 0 && st.n_events) {
    atomicAdd(&stats[0], st.n_events);
    atomicAdd(&stats[1], st.n_raw);
    atomicAdd(&stats[2], st.n_val);
    atomicAdd(&stats[3], st.J);
    atomicAdd(&stats[4], st.JNe);
    atomicAdd(&stats[5], st.n_exp);
    atomicAdd(&stats[6], st.n_val_ne);
    atomicAdd(&stats[7], st.reserved);
  }
}

__global__ void __launch_bounds__(256)
eas_finish_kernel(int mode, nss::EasParams P, const double* __restrict__ beta,
                  const double* __restrict__ alt, const nss::EasAcc* __restrict__ acc,
                  const nss::EasPrep* __restrict__ prep, long long n,
                  double* __restrict__ out0, double* __restrict__ out1,
                  double* __restrict__ dphots, double* __restrict__ cang) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const nss::EasAcc A = acc[i];
    double d = 0.0, c = 0.0;
    const float cloud_top = P.has_cloud ? __ldg(&prep[i].cloud) : 0.0f;
    nss::eas_finish(P, beta[i], alt[i], A, cloud_top, d, c);
    if (mode == 0) {
      if (A.pad == 0) c = 1.5;  // eas.py:106-107 default for events the mask skipped
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

// Per-device lazily initialised state: the constant tables (one cudaMemcpyToSymbol per device,
// a synchronous copy of 1.5 KB of HOST data, no device synchronisation) and the persistent
// grid size of the two walk instantiations.  Guarded by a mutex so that concurrent first calls
// from several host threads are safe; after that the entry points touch no shared state.
constexpr int MAX_DEVICES = 64;
std::mutex g_init_mutex;
bool g_tables_ready[MAX_DEVICES] = {};
int g_walk_blocks[MAX_DEVICES][2][2] = {};  // [device][profile is Gaisser-Hillas][with counters]
double* g_conex = nullptr;  // device copy of the CONEX Gaisser-Hillas table (registered)
int g_conex_rows = 0;

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

// caller-supplied workspace of one launch: [work counter | EasPrep[n] | EasAcc[n]]
struct Workspace {
  unsigned long long* counter;
  nss::EasPrep* prep;
  nss::EasAcc* acc;
};
inline Workspace carve(void* ws, int64_t n) {
  char* w = (char*)ws;
  Workspace W;
  W.counter = (unsigned long long*)w;
  w += 256;
  W.prep = (nss::EasPrep*)w;
  w += align256((int64_t)sizeof(nss::EasPrep) * n);
  W.acc = (nss::EasAcc*)w;
  return W;
}

// optional per-launch CUDA-event timing of the walk kernel (nss_eas_profile_*)
bool g_profile = false;
struct ProfRec { cudaEvent_t a, b; };
ProfRec* g_prof = nullptr;
int g_prof_n = 0, g_prof_cap = 0;

int launch(int mode, const NssParams* p, const double* beta, const double* alt,
           const double* eshow, const uint8_t* mask, const double* u_profile,
           const double* cloud_ev, int64_t n, double* o0, double* o1, double* dph, double* cang,
           unsigned long long* stats, void* workspace, void* stream) {
  if (!p) return NSS_E_NULL;
  if (n < 0 || p->eas_profile < 0 || p->eas_profile > 2 || p->has_cloud < 0 || p->has_cloud > 2)
    return NSS_E_BADARG;
  if (n == 0) return 0;
  if (!beta || !alt || !eshow || !o0 || !o1 || !workspace) return NSS_E_NULL;
  if (p->eas_profile == 2 && !u_profile) return NSS_E_NULL;
  if (p->has_cloud == 2 && !cloud_ev) return NSS_E_NULL;
  if (p->has_cloud != 2) cloud_ev = nullptr;
  const bool gh = p->eas_profile != 0;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t err;
  int dev = 0;
  if ((err = cudaGetDevice(&dev)) != cudaSuccess) return (int)err;
  if (dev < 0 || dev >= MAX_DEVICES) return NSS_E_BADARG;
  int walk_blocks;
  const double* conex;
  int conex_rows;
  {
    std::lock_guard<std::mutex> lock(g_init_mutex);
    conex = g_conex;  // snapshot of the registered table (nss_eas_set_conex_table)
    conex_rows = g_conex_rows;
    if (p->eas_profile == 2 && !conex) return NSS_E_NULL;
    if (!g_tables_ready[dev]) {
      nss::EasTables T;
      nss_eas_tables_init(&T);
      err = cudaMemcpyToSymbol(c_tables, &T, sizeof(T));
      if (err != cudaSuccess) return (int)err;
      g_tables_ready[dev] = true;
    }
    const bool ws = stats != nullptr;
    if (g_walk_blocks[dev][gh][ws] == 0) {
      int sms = 0, per_sm = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      auto k = gh ? (ws ? eas_walk_kernel<true, true> : eas_walk_kernel<true, false>)
                  : (ws ? eas_walk_kernel<false, true> : eas_walk_kernel<false, false>);
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, walk_block(gh), 0);
      if (err != cudaSuccess) return (int)err;
      g_walk_blocks[dev][gh][ws] = sms * (per_sm > 0 ? per_sm : 1);
    }
    walk_blocks = g_walk_blocks[dev][gh][stats != nullptr];
  }
  const Workspace W = carve(workspace, n);
  nss::EasParams P;
  P.detector_altitude = p->detector_altitude;
  P.area = p->telescope_effective_area;
  P.qe = p->quantum_efficiency;
  P.pe_threshold = p->photo_electron_threshold;
  P.has_cloud = p->has_cloud;
  P.cloud_top = (float)p->cloud_top;
  P.profile = p->eas_profile;
  P.pad = 0;

  const int tb = 256;
  long long g1 = (n + tb - 1) / tb;
  if (g1 > 148LL * 64) g1 = 148LL * 64;
  err = cudaMemsetAsync(W.counter, 0, 8, s);
  if (err != cudaSuccess) return (int)err;
  eas_prep_kernel<<<(int)g1, tb, 0, s>>>(mode, p->eas_profile, beta, alt, eshow, mask, conex,
                                         conex_rows, u_profile, cloud_ev, P.cloud_top, n, W.prep);

  const int wpb = walk_block(gh) / 32;  // warps (= events in flight) per block
  long long need = (n + wpb - 1) / wpb;
  int blocks = (int)(need < walk_blocks ? need : walk_blocks);
  ProfRec* rec = nullptr;
  if (g_profile) {
    if (g_prof_n == g_prof_cap) {
      g_prof_cap = g_prof_cap ? 2 * g_prof_cap : 256;
      g_prof = (ProfRec*)realloc(g_prof, sizeof(ProfRec) * g_prof_cap);
    }
    rec = &g_prof[g_prof_n++];
    cudaEventCreate(&rec->a);
    cudaEventCreate(&rec->b);
    cudaEventRecord(rec->a, s);
  }
  auto walk = gh ? (stats ? eas_walk_kernel<true, true> : eas_walk_kernel<true, false>)
                 : (stats ? eas_walk_kernel<false, true> : eas_walk_kernel<false, false>);
  walk<<<blocks, walk_block(gh), 0, s>>>(P, W.prep, n, W.acc, W.counter, stats);
  if (rec) cudaEventRecord(rec->b, s);
  eas_finish_kernel<<<(int)g1, tb, 0, s>>>(mode, P, beta, alt, W.acc, W.prep, n, o0, o1, dph,
                                           cang);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int64_t nss_eas_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  return 256 + align256((int64_t)sizeof(nss::EasPrep) * n) +
         align256((int64_t)sizeof(nss::EasAcc) * n);
}

extern "C" int nss_eas_optical(const NssParams* p, const double* beta, const double* altDec,
                               const double* showerEnergy, const uint8_t* mask,
                               const double* u_profile, const double* cloud_top_ev, int64_t n,
                               double* numPEs, double* costhetaChEff, double* dphots,
                               double* cang, unsigned long long* stats, void* workspace,
                               void* stream) {
  return launch(0, p, beta, altDec, showerEnergy, mask, u_profile, cloud_top_ev, n, numPEs,
                costhetaChEff, dphots, cang, stats, workspace, stream);
}

extern "C" int nss_cphotang(const NssParams* p, const double* betaE, const double* alt,
                            const double* Eshow100PeV, const double* u_profile,
                            const double* cloud_top_ev, int64_t n, double* dphots, double* cang,
                            unsigned long long* stats, void* workspace, void* stream) {
  return launch(1, p, betaE, alt, Eshow100PeV, nullptr, u_profile, cloud_top_ev, n, dphots,
                cang, nullptr, nullptr, stats, workspace, stream);
}

extern "C" int nss_eas_set_conex_table(const double* rows_host, int nrows) {
  if (!rows_host) return NSS_E_NULL;
  if (nrows <= 0) return NSS_E_BADARG;
  cudaError_t err;
  std::lock_guard<std::mutex> lock(g_init_mutex);
  if (g_conex) {
    cudaDeviceSynchronize();
    cudaFree(g_conex);
    g_conex = nullptr;
    g_conex_rows = 0;
  }
  if ((err = cudaMalloc(&g_conex, sizeof(double) * 6 * (size_t)nrows)) != cudaSuccess)
    return (int)err;
  err = cudaMemcpy(g_conex, rows_host, sizeof(double) * 6 * (size_t)nrows,
                   cudaMemcpyHostToDevice);
  if (err != cudaSuccess) return (int)err;
  g_conex_rows = nrows;
  return 0;
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
