"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product path.

Time the UNMODIFIED reference on the host cores: compute()'s optical branch
(src/nuspacesim/compute.py:186-257) stage by stage with the reference's own classes, loaded
through oracle/ref_harness.py from /root/reference (build container) or from the copy
stage_reference() put under oracle/_ref/reference/ (GPU box).  The EAS stage fans out over
`--procs` worker processes exactly where the reference hands its events to dask.bag
(cphotang.py:719-723); everything else is the reference's single-threaded NumPy.

    python -m oracle.ref_bench --events 20000 --procs 16 [--loge 8.0] [--seed 1]

Prints one JSON line.  Run as its own process (it forks workers and installs stub modules).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run(events: int, procs: int, loge: float, seed: int):
    from oracle import ref_harness as H

    warnings.filterwarnings("ignore")
    ref = H.load_reference()
    H.POOL_SIZE = int(procs)
    cfg = H.default_config(ref)
    cfg.simulation.spectrum.log_nu_energy = float(loge)
    cfg.simulation.thrown_events = int(events)
    np.random.seed(seed)
    t = {}
    t0 = time.perf_counter()
    geom = ref.region_geometry.RegionGeom(cfg)
    spec = ref.spectra.Spectra(cfg)
    tau = H.make_taus(ref, cfg)
    eas = ref.eas.EAS(cfg)
    t["construct"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    beta_tr, thetaArr, pathLenArr = geom(events)
    init_lat, init_long = geom.find_lat_long_along_traj(np.zeros_like(beta_tr))
    t["geometry"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    log_e_nu, mc_spec_norm, spec_weights_sum = spec(beta_tr.shape[0])
    tauBeta, tauLorentz, tauEnergy, showerEnergy, tauExitProb = tau(beta_tr, log_e_nu)
    altDec, lenDec = eas.altDec(beta_tr, tauBeta, tauLorentz)
    t["spectra_taus_altdec"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    numPEs, costhetaChEff = eas(beta_tr, altDec, showerEnergy, init_lat, init_long, cloudf=None)
    t["eas"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    mcint, mcintgeo, passEV, mcunc = geom.mcintegral(
        numPEs, costhetaChEff, tauExitProb, cfg.detector.optical.photo_electron_threshold,
        mc_spec_norm, spec_weights_sum, lenDec=lenDec, method="Optical")
    t["mcintegral"] = time.perf_counter() - t0
    total = sum(v for k, v in t.items() if k != "construct")
    return dict(events=int(events), procs=int(procs), seconds=total, events_per_s=events / total,
                eas_seconds=t["eas"], eas_events_per_s=float(beta_tr.shape[0]) / t["eas"],
                stages_s=t, n_valid=int(beta_tr.shape[0]), mcint=float(mcint),
                mcint_unc=float(mcunc), n_pass=int(passEV), log_nu_energy=float(loge),
                source="live" if H.reference_is_live() else "staged")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=float, default=2000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--loge", type=float, default=8.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    print(json.dumps(run(int(a.events), a.procs, a.loge, a.seed)))


if __name__ == "__main__":
    main()
