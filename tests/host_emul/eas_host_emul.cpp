// TEST BUILD ONLY.  Compiles nuspacesim_b200/csrc/nss_eas_core.cuh for the host with a
// lane group of width 1 so the kernel's arithmetic can be compared with the oracle in a
// container that has no GPU.  Never loaded by the product package.
//   g++ -O2 -ffp-contract=off -shared -fPIC -I nuspacesim_b200/csrc tests/host_emul/eas_host_emul.cpp
#include "nss_eas_core.cuh"
#include "nss_eas_tables.h"

namespace {
struct Ops1 {
  static constexpr int W = 1;
  static int lane() { return 0; }
  static float scan_f(float v) { return v; }
  static double scan_d(double v) { return v; }
  static double sum_d(double v) { return v; }
  static int sum_i(int v) { return v; }
  static int max_i(int v) { return v; }
  static int min_i(int v) { return v; }
  static bool any(bool b) { return b; }
  static double shift_up_d(double, double carry) { return carry; }
  static double last_d(double v) { return v; }
  static void argmax_first(float, int, float payload, float& out) { out = payload; }
};
}  // namespace

extern "C" int nss_emul_eas(long long n, const double* beta, const double* alt,
                            const double* eshow, double detector_altitude, double area,
                            double qe, double thr, int has_cloud, float cloud_top,
                            double* numPEs, double* cosCh, double* dphots, double* cang,
                            unsigned long long* stats6) {
  nss::EasTables T;
  nss_eas_tables_init(&T);
  nss::EasParams P;
  P.detector_altitude = detector_altitude;
  P.area = area;
  P.qe = qe;
  P.pe_threshold = thr;
  P.has_cloud = has_cloud;
  P.cloud_top = cloud_top;
  nss::EasStats st = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = 0; i < n; ++i) {
    double d = 0.0, c = 1.5;
    if (!(alt[i] < 0.0 || alt[i] > 20.0))
      nss::eas_event<Ops1>(T, P, beta[i], alt[i], eshow[i], d, c, &st, true);
    double npe, cc;
    nss::eas_tail(P, d, c, npe, cc);
    numPEs[i] = npe;
    cosCh[i] = cc;
    if (dphots) dphots[i] = d;
    if (cang) cang[i] = c;
  }
  if (stats6) {
    stats6[0] = st.n_events; stats6[1] = st.n_raw; stats6[2] = st.n_val;
    stats6[3] = st.J; stats6[4] = st.JNe; stats6[5] = st.n_exp;
  }
  return 0;
}
