"""Generate tests/golden/eas_gh.npz: the UNMODIFIED reference's CphotAng.run with the two
Gaisser-Hillas longitudinal profiles (build container only; needs /root/reference).

    python tests/golden/make_golden_gh.py

"Gaisser-Hillas Fluctuated" draws a CONEX row per event with np.random.randint
(shower_properties.py:165-166); here np.random.randint is replaced for the duration of each
call by a function returning a recorded index, so the row used is part of the fixture.
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
_CP = {}


def _init():
    from oracle import ref_harness as H

    ref = H.load_reference()
    for prof in ("Gaisser-Hillas Parameterized", "Gaisser-Hillas Fluctuated"):
        _CP[prof] = ref.cphotang.CphotAng(525.0, prof)


def _run(args):
    prof, b, a, e, idx = args
    orig = np.random.randint
    np.random.randint = lambda low=0, high=None, *x, **k: int(idx)
    try:
        with np.errstate(all="ignore"):
            d, c = _CP[prof].run(b, a, e, 0.0, 0.0, None)
    finally:
        np.random.randint = orig
    return float(d), float(c)


def main():
    G = np.load(os.path.join(HERE, "eas_large.npz"))
    rng = np.random.Generator(np.random.PCG64(7))
    out = {}
    with mp.Pool(os.cpu_count(), initializer=_init) as pool:
        for tag, n in (("c1", 3000), ("e10", 1000), ("pl10", 1500)):
            sel = rng.choice(len(G[f"{tag}_beta_rad"]), n, replace=False)
            b, a, e = (G[f"{tag}_{k}"][sel] for k in ("beta_rad", "altDec", "showerEnergy"))
            idx = rng.integers(0, 1000, n)
            for short, prof in (("ghp", "Gaisser-Hillas Parameterized"),
                                ("ghf", "Gaisser-Hillas Fluctuated")):
                res = pool.map(_run, [(prof, *x) for x in zip(b, a, e, idx)], chunksize=32)
                out[f"{tag}_{short}_dphots"] = np.array([r[0] for r in res])
                out[f"{tag}_{short}_cang"] = np.array([r[1] for r in res], np.float32)
            out[f"{tag}_beta_rad"], out[f"{tag}_altDec"], out[f"{tag}_showerEnergy"] = b, a, e
            out[f"{tag}_row"] = idx.astype(np.int32)
            print(tag, n)
    np.savez_compressed(os.path.join(HERE, "eas_gh.npz"), **out)


if __name__ == "__main__":
    main()
