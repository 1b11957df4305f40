// TEST BUILD ONLY.  Compiles nuspacesim_b200/csrc/nss_eas_core.cuh for the host with a
// lane group of width 1 so the kernel's arithmetic can be compared with the oracle in a
// container that has no GPU.  Never loaded by the product package.
//   g++ -O2 -ffp-contract=off -shared -fPIC -I nuspacesim_b200/csrc tests/host_emul/eas_host_emul.cpp
#include <stdlib.h>

#include "nss_eas_core.cuh"
#include "nss_eas_tables.h"

namespace {
struct Ops1 {
  static constexpr int W = 1;
  static int lane() { return 0; }
  static float scan_f(float v) { return v; }
  static double sum_d(double v) { return v; }
  static float sum_f(float v) { return v; }
  static float max_f(float v) { return v; }
  static int sum_i(int v) { return v; }
  static int max_i(int v) { return v; }
  static int min_i(int v) { return v; }
  static bool any(bool b) { return b; }
  static float lookup11(const float* tab, float, int idx) { return tab[idx]; }
  static float last_f(float v) { return v; }
  static void argmax_first(float, int, float payload, float& out) { out = payload; }
};
}  // namespace

// cloud_ev (optional): per-event cloud tops [km], overriding has_cloud / cloud_top
// profile: 0 Greisen, 1 Gaisser-Hillas Parameterized, 2 Gaisser-Hillas Fluctuated with
// gh_rows[i] = the (Nm, Xm, X0, p1, p2, p3) CONEX row of event i
extern "C" int nss_emul_eas_profile(long long n, const double* beta, const double* alt,
                                    const double* eshow, double detector_altitude, double area,
                                    double qe, double thr, int has_cloud, float cloud_top,
                                    double* numPEs, double* cosCh, double* dphots, double* cang,
                                    unsigned long long* stats6, int profile,
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
  nss::EasStats st = {0, 0, 0, 0, 0, 0, 0, 0};
  // events are independent; tools/emul_eval.py builds this with -fopenmp (the counters are
  // then merged per thread), the test build is serial
#pragma omp parallel
  {
  nss::EasStats stl = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp for schedule(dynamic, 64)
  for (long long i = 0; i < n; ++i) {
    double d = 0.0, c = 1.5;
    if (!(alt[i] < 0.0 || alt[i] > 20.0)) {
      nss::EasPrep E;
      nss::EasAcc A;
      const float ct = cloud_ev ? (float)cloud_ev[i] : cloud_top;
      nss::eas_prep(T, beta[i], alt[i], eshow[i], true, profile,
                    profile == 2 ? gh_rows + 6 * i : nullptr, ct, E);
      if (profile == 0)
        nss::eas_walk<Ops1, false>(T, P, E, 0.0f, E.cloud, A, &stl, true);
      else
        nss::eas_walk<Ops1, true>(T, P, E, 0.0f, E.cloud, A, &stl, true);
      nss::eas_finish(P, beta[i], alt[i], A, E.cloud, d, c);
    }
    double npe, cc;
    nss::eas_tail(P, d, c, npe, cc);
    numPEs[i] = npe;
    cosCh[i] = cc;
    if (dphots) dphots[i] = d;
    if (cang) cang[i] = c;
  }
#pragma omp critical
  {
    st.n_events += stl.n_events; st.n_raw += stl.n_raw; st.n_val += stl.n_val;
    st.J += stl.J; st.JNe += stl.JNe; st.n_exp += stl.n_exp;
    st.reserved += stl.reserved;
  }
  }
  if (stats6) {
    stats6[0] = st.n_events; stats6[1] = st.n_raw; stats6[2] = st.n_val;
    stats6[3] = st.J; stats6[4] = st.JNe; stats6[5] = st.n_exp;
    if (getenv("NSS_EMUL_LIGHT")) stats6[5] = st.reserved;  // steps only counted (tail)
  }
  return 0;
}

extern "C" int nss_emul_eas(long long n, const double* beta, const double* alt,
                            const double* eshow, double detector_altitude, double area,
                            double qe, double thr, int has_cloud, float cloud_top,
                            double* numPEs, double* cosCh, double* dphots, double* cang,
                            unsigned long long* stats6) {
  return nss_emul_eas_profile(n, beta, alt, eshow, detector_altitude, area, qe, thr, has_cloud,
                              cloud_top, numPEs, cosCh, dphots, cang, stats6, 0, nullptr,
                              nullptr);
}

// per-event prep record, for checking the column totals against a brute-force sum
extern "C" int nss_emul_prep(double beta, double alt, double eshow, float* out8, int* out2) {
  nss::EasTables T;
  nss_eas_tables_init(&T);
  nss::EasPrep E;
  nss::eas_prep(T, beta, alt, eshow, true, 0, nullptr, 0.0f, E);
  out8[0] = E.R0; out8[1] = E.q; out8[2] = E.alt; out8[3] = E.Ldet;
  out8[4] = E.Gtot; out8[5] = E.ozA[E.seg0]; out8[6] = E.gb; out8[7] = E.alpha;
  out2[0] = E.n_raw; out2[1] = E.Ieang;
  return 0;
}

// ozone segment index of n altitudes: reference loop and the walk's arithmetic shortcut
extern "C" int nss_emul_ozone_segments(long long n, const float* z, int* slow, int* fast) {
  nss::EasTables T;
  nss_eas_tables_init(&T);
  for (long long i = 0; i < n; ++i) {
    slow[i] = nss::ozone_segment(T, z[i]);
    fast[i] = nss::ozone_segment_fast(T, z[i]);
  }
  return 0;
}
