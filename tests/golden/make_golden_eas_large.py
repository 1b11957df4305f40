"""Generate tests/golden/eas_large.npz: the UNMODIFIED reference's CphotAng.run on ~30 000
realistic events (build container only; needs /root/reference).

    python tests/golden/make_golden_eas_large.py

Inputs (beta_rad, altDec, showerEnergy) are what compute() hands to the EAS stage: they are
produced from PCG64 uniforms by the oracle's float64 geometry / taus / altDec stages (pinned
bit-exactly / to 1e-12 against the reference by tests/test_oracle_golden.py).  Outputs are
CphotAng.run's (photon density, Cherenkov angle) from the reference's own source files run
through oracle/ref_harness.py in a process pool (the reference's own fan-out is a dask bag of
processes, cphotang.py:719-723).

Sets:  c1   sample_input_file physics, log10 E_nu = 8.0        (BASELINE.json configs[0..1])
       e7 / e9 / e10   monospectrum sweep                       (configs[2])
       pl10 power-law spectrum (2, 6, 12), theta_ch_max 10 deg  (configs[3])
"""
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

warnings.filterwarnings("ignore")
SEED = 20241019
_CP = None


def _init(det):
    global _CP
    from oracle import ref_harness as H

    ref = H.load_reference()
    _CP = ref.cphotang.CphotAng(det, "Greisen")


def _run(args):
    b, a, e = args
    with np.errstate(all="ignore"):
        d, c = _CP.run(b, a, e, 0.0, 0.0, None)
    return float(d), float(c)


def main():
    from conftest import load_tables
    from oracle import nss_oracle as O

    tabs = load_tables()
    rng = np.random.Generator(np.random.PCG64(SEED))
    out = {}
    sets = (("c1", dict(log_nu_energy=8.0), 14000), ("e7", dict(log_nu_energy=7.0), 4000),
            ("e9", dict(log_nu_energy=9.0), 4000), ("e10", dict(log_nu_energy=10.0), 4000),
            ("pl10", dict(spectrum="powerspectrum", max_cherenkov_angle=float(np.radians(10.0))),
             6000))
    with mp.Pool(os.cpu_count(), initializer=_init, initargs=(525.0,)) as pool:
        for tag, kw, n in sets:
            p = O.OracleParams(**kw)
            u4, us, ut, ud = rng.random((4, n)), rng.random(n), rng.random(n), rng.random(n)
            g = O.geom_constants(p)
            t = O.geom_throw(g, u4)
            m = t["event_mask"]
            beta = np.radians(t["betaTrSubN"][m])
            le, _, _ = O.spectra(p, int(m.sum()), us[m])
            tb, lor, te, sh, pex = O.taus(p, tabs, beta, le, ut[m])
            alt, _ = O.alt_dec(beta, tb, lor, ud[m])
            ok = ~((alt < 0.0) | (alt > 20.0))  # eas.py:102-103
            beta, alt, sh = beta[ok], alt[ok], sh[ok]
            res = pool.map(_run, zip(beta, alt, sh), chunksize=64)
            d = np.array([r[0] for r in res])
            c = np.array([r[1] for r in res], np.float32)
            out[f"{tag}_beta_rad"], out[f"{tag}_altDec"], out[f"{tag}_showerEnergy"] = beta, alt, sh
            out[f"{tag}_dphots"], out[f"{tag}_cang"] = d, c
            print(tag, len(beta), "events; nonzero", int(np.count_nonzero(d)))
    np.savez_compressed(os.path.join(HERE, "eas_large.npz"), **out)


if __name__ == "__main__":
    main()
