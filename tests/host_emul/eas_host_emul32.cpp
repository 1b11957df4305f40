// TEST BUILD ONLY.  The walk of nss_eas_core.cuh executed as a 32-lane warp ON THE HOST: every
// lane is a fiber (ucontext), every warp collective of the Ops interface is a rendezvous of all
// 32 fibers.  This checks what the W = 1 build (eas_host_emul.cpp) cannot: the chunk-level
// logic -- votes, reductions, the scan carry, the tail skip, per-lane histories -- and it aborts
// when the lanes do not all arrive at the same collective (a collective under a divergent
// branch hangs a real warp).  Never loaded by the product package.
//   g++ -O2 -ffp-contract=off -shared -fPIC -I nuspacesim_b200/csrc tests/host_emul/eas_host_emul32.cpp
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include "nss_eas_core.cuh"
#include "nss_eas_tables.h"

namespace {

constexpr int WARP = 32;
constexpr size_t STACK = 256 * 1024;

struct Warp {
  ucontext_t sched, lane_ctx[WARP];
  char* stacks = nullptr;
  int cur = 0;              // lane being executed
  bool done[WARP];
  // rendezvous state: two alternating exchange buffers (a lane may be one collective ahead)
  double buf[2][WARP];
  int op[2][WARP];
  int phase[WARP];
  int arrived[2];
  bool diverged = false;
};
thread_local Warp* g_w = nullptr;

// every lane deposits (opcode, value) and yields; returns the buffer all 32 values are in
const double* rendezvous(int opcode, double v) {
  Warp& w = *g_w;
  const int l = w.cur, p = w.phase[l] & 1;
  w.buf[p][l] = v;
  w.op[p][l] = opcode;
  w.arrived[p] += 1;
  w.phase[l] += 1;
  swapcontext(&w.lane_ctx[l], &w.sched);  // resumed once all lanes have arrived
  for (int i = 0; i < WARP; ++i)
    if (w.op[p][i] != opcode) w.diverged = true;
  return w.buf[p];
}

enum { OP_SCAN = 1, OP_SUMD, OP_SUMF, OP_MAXF, OP_SUMI, OP_MAXI, OP_MINI, OP_ANY, OP_LOOKUP,
       OP_LAST, OP_ARGMAX_K, OP_ARGMAX_I, OP_ARGMAX_P };

struct Ops32H {
  static constexpr int W = WARP;
  static int lane() { return g_w->cur; }
  static float scan_f(float v) {
    // the GPU's Hillis-Steele order (nss_eas_kernel.cu Ops32::scan_f), so the float32
    // rounding of the prefix sums is the kernel's
    float x = v;
    for (int d = 1; d < WARP; d <<= 1) {
      const double* b = rendezvous(OP_SCAN, (double)x);
      const int l = lane();
      if (l >= d) x += (float)b[l - d];
    }
    return x;
  }
  static double sum_d(double v) {
    for (int d = 16; d > 0; d >>= 1) {  // butterfly, like __shfl_xor
      const double* b = rendezvous(OP_SUMD, v);
      v += b[lane() ^ d];
    }
    return v;
  }
  static float sum_f(float v) {
    for (int d = 16; d > 0; d >>= 1) {
      const double* b = rendezvous(OP_SUMF, (double)v);
      v += (float)b[lane() ^ d];
    }
    return v;
  }
  static float max_f(float v) {
    const double* b = rendezvous(OP_MAXF, (double)v);
    float m = (float)b[0];
    for (int i = 1; i < WARP; ++i) m = (float)b[i] > m ? (float)b[i] : m;
    return m;
  }
  static int sum_i(int v) {
    const double* b = rendezvous(OP_SUMI, (double)v);
    long long s = 0;
    for (int i = 0; i < WARP; ++i) s += (long long)b[i];
    return (int)s;
  }
  static int max_i(int v) {
    const double* b = rendezvous(OP_MAXI, (double)v);
    int m = (int)b[0];
    for (int i = 1; i < WARP; ++i) m = (int)b[i] > m ? (int)b[i] : m;
    return m;
  }
  static int min_i(int v) {
    const double* b = rendezvous(OP_MINI, (double)v);
    int m = (int)b[0];
    for (int i = 1; i < WARP; ++i) m = (int)b[i] < m ? (int)b[i] : m;
    return m;
  }
  static bool any(bool x) {
    const double* b = rendezvous(OP_ANY, x ? 1.0 : 0.0);
    for (int i = 0; i < WARP; ++i)
      if (b[i] != 0.0) return true;
    return false;
  }
  // tab[idx] where lane g < 11 holds tab[g] in lane_value (the index differs per lane)
  static float lookup11(const float*, float lane_value, int idx) {
    const double* b = rendezvous(OP_LOOKUP, (double)lane_value);
    return (float)b[idx & 31];
  }
  static float last_f(float v) {
    const double* b = rendezvous(OP_LAST, (double)v);
    return (float)b[WARP - 1];
  }
  static void argmax_first(float key, int idx, float payload, float& out) {
    double k[WARP], ii[WARP], pp[WARP];
    memcpy(k, rendezvous(OP_ARGMAX_K, (double)key), sizeof(k));
    memcpy(ii, rendezvous(OP_ARGMAX_I, (double)idx), sizeof(ii));
    memcpy(pp, rendezvous(OP_ARGMAX_P, (double)payload), sizeof(pp));
    int best = 0;
    for (int i = 1; i < WARP; ++i)
      if (k[i] > k[best] || (k[i] == k[best] && ii[i] < ii[best])) best = i;
    out = (float)pp[best];
  }
};

struct Job {
  const nss::EasTables* T;
  const nss::EasParams* P;
  const nss::EasPrep* E;
  int profile;
  nss::EasAcc acc[WARP];
  nss::EasStats st;
};
thread_local Job* g_job = nullptr;

void lane_main() {
  Warp& w = *g_w;
  Job& j = *g_job;
  const int l = w.cur;
  const float ozA_lane = (l < 11) ? j.E->ozA[l] : (l == 11 ? j.E->cloud : 0.0f);
  const float cloud_top = Ops32H::lookup11(nullptr, ozA_lane, 11);  // like the kernel's shuffle
  if (j.profile == 0)
    nss::eas_walk<Ops32H, false>(*j.T, *j.P, *j.E, ozA_lane, cloud_top, j.acc[l], &j.st, l == 0);
  else
    nss::eas_walk<Ops32H, true>(*j.T, *j.P, *j.E, ozA_lane, cloud_top, j.acc[l], &j.st, l == 0);
  w.done[l] = true;
  swapcontext(&w.lane_ctx[l], &w.sched);
}

// run one event through the 32-lane walk; returns 0, or 1 if the lanes diverged at a collective
int run_warp(Job& job) {
  Warp w;
  w.stacks = (char*)malloc(STACK * WARP);
  memset(w.done, 0, sizeof(w.done));
  memset(w.phase, 0, sizeof(w.phase));
  w.arrived[0] = w.arrived[1] = 0;
  g_w = &w;
  g_job = &job;
  for (int l = 0; l < WARP; ++l) {
    getcontext(&w.lane_ctx[l]);
    w.lane_ctx[l].uc_stack.ss_sp = w.stacks + STACK * l;
    w.lane_ctx[l].uc_stack.ss_size = STACK;
    w.lane_ctx[l].uc_link = &w.sched;
    makecontext(&w.lane_ctx[l], lane_main, 0);
  }
  int rc = 0;
  for (;;) {
    // run every lane until its next collective (or its end)
    int ndone = 0;
    for (int l = 0; l < WARP; ++l) {
      if (w.done[l]) { ++ndone; continue; }
      w.cur = l;
      swapcontext(&w.sched, &w.lane_ctx[l]);
      if (w.done[l]) ++ndone;
    }
    if (ndone == WARP) break;
    // all live lanes now wait at a collective: every lane must be there, at the same phase
    const int ph = w.phase[0];
    bool ok = (ndone == 0);
    for (int l = 0; l < WARP && ok; ++l) ok = (w.phase[l] == ph);
    if (!ok || w.arrived[(ph - 1) & 1] != WARP) { rc = 1; break; }
    w.arrived[(ph - 1) & 1] = 0;
  }
  if (w.diverged) rc = 1;
  free(w.stacks);
  g_w = nullptr;
  return rc;
}

}  // namespace

// Same arguments as nss_emul_eas_profile (eas_host_emul.cpp).  Returns the number of events
// whose lanes diverged at a collective (must be 0).
extern "C" int nss_emul32_eas_profile(long long n, const double* beta, const double* alt,
                                      const double* eshow, double detector_altitude, double area,
                                      double qe, double thr, int has_cloud, float cloud_top,
                                      double* numPEs, double* cosCh, double* dphots, double* cang,
                                      unsigned long long* stats8, int profile,
                                      const double* gh_rows, const double* cloud_ev) {
  nss::EasTables T;
  nss_eas_tables_init(&T);
  nss::EasParams P;
  P.detector_altitude = detector_altitude;
  P.area = area;
  P.qe = qe;
  P.pe_threshold = thr;
  P.has_cloud = cloud_ev ? 2 : has_cloud;
  P.cloud_top = cloud_top;
  P.profile = profile;
  P.pad = 0;
  int bad = 0;
  nss::EasStats tot = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = 0; i < n; ++i) {
    double d = 0.0, c = 1.5;
    if (!(alt[i] < 0.0 || alt[i] > 20.0)) {
      nss::EasPrep E;
      const float ct = cloud_ev ? (float)cloud_ev[i] : cloud_top;
      nss::eas_prep(T, beta[i], alt[i], eshow[i], true, profile,
                    profile == 2 ? gh_rows + 6 * i : nullptr, ct, E);
      Job job;
      job.T = &T; job.P = &P; job.E = &E; job.profile = profile;
      job.st = nss::EasStats{0, 0, 0, 0, 0, 0, 0, 0};
      bad += run_warp(job);
      nss::eas_finish(P, beta[i], alt[i], job.acc[0], E.cloud, d, c);
      tot.n_events += job.st.n_events; tot.n_raw += job.st.n_raw; tot.n_val += job.st.n_val;
      tot.J += job.st.J; tot.JNe += job.st.JNe; tot.n_exp += job.st.n_exp;
      tot.n_val_ne += job.st.n_val_ne; tot.reserved += job.st.reserved;
    }
    double npe, cc;
    nss::eas_tail(P, d, c, npe, cc);
    numPEs[i] = npe;
    cosCh[i] = cc;
    if (dphots) dphots[i] = d;
    if (cang) cang[i] = c;
  }
  if (stats8) {
    stats8[0] = tot.n_events; stats8[1] = tot.n_raw; stats8[2] = tot.n_val; stats8[3] = tot.J;
    stats8[4] = tot.JNe; stats8[5] = tot.n_exp; stats8[6] = tot.n_val_ne; stats8[7] = tot.reserved;
  }
  return bad;
}
