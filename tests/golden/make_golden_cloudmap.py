"""Generate tests/golden/eas_cloudmap.npz: the UNMODIFIED reference's CphotAng.run with the
PressureMapCloud cloud-top function (build container only; needs /root/reference).

    python tests/golden/make_golden_cloudmap.py

cloudf = clouds.altitude_from_pressure_map_v0(map) (atmosphere/clouds.py:85-96) on the
reference's own FITS maps (primary HDU read with numpy: astropy is absent), evaluated at a
latitude per event like compute() does (init_lat in radians, compute.py:214-216).
"""
import importlib
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
MAPS = "/root/reference/src/nuspacesim/data/cloud_maps/nss_map_CloudTopPressure_%02d.v0.fits"
_S = {}


def ref_map(month):
    raw = open(MAPS % month, "rb").read()
    return np.frombuffer(raw[2880:2880 + 361 * 576 * 4], dtype=">f4").reshape(361, 576).astype(
        np.float32)


def _init():
    from oracle import ref_harness as H

    ref = H.load_reference()
    clouds = importlib.import_module("nuspacesim.simulation.atmosphere.clouds")
    _S["cp"] = ref.cphotang.CphotAng(525.0, "Greisen")
    _S["f"] = {m: clouds.altitude_from_pressure_map_v0(ref_map(m)) for m in (1, 7)}


def _run(args):
    month, b, a, e, lat, lon = args
    f = _S["f"][month]
    with np.errstate(all="ignore"):
        d, c = _S["cp"].run(b, a, e, lat, lon, f)
    return float(d), float(c), float(f(lat, lon))


def main():
    G = np.load(os.path.join(HERE, "eas_large.npz"))
    rng = np.random.Generator(np.random.PCG64(11))
    out = {}
    with mp.Pool(os.cpu_count(), initializer=_init) as pool:
        for month, tag, n in ((1, "c1", 1500), (7, "e10", 1000)):
            sel = rng.choice(len(G[f"{tag}_beta_rad"]), n, replace=False)
            b, a, e = (G[f"{tag}_{k}"][sel] for k in ("beta_rad", "altDec", "showerEnergy"))
            lat = rng.uniform(-1.55, 1.55, n)
            lat[: n // 2] = rng.uniform(-0.35, 0.35, n // 2)  # what compute() produces
            lon = rng.uniform(-np.pi, np.pi, n)
            res = pool.map(_run, [(month, *x) for x in zip(b, a, e, lat, lon)], chunksize=32)
            k = f"m{month:02d}"
            out[f"{k}_beta_rad"], out[f"{k}_altDec"], out[f"{k}_showerEnergy"] = b, a, e
            out[f"{k}_lat"], out[f"{k}_lon"] = lat, lon
            out[f"{k}_dphots"] = np.array([r[0] for r in res])
            out[f"{k}_cang"] = np.array([r[1] for r in res], np.float32)
            out[f"{k}_cloud_top"] = np.array([r[2] for r in res])
            print(k, n, "clouded to zero:", int((out[f"{k}_dphots"] == 0).sum()))
    # the lookup itself on a latitude sweep (both months), incl. the exact axis knots
    _init()
    sweep = np.concatenate([np.linspace(-1.6, 1.6, 4001), np.linspace(-90, 90, 361),
                            np.linspace(-180, 180, 576)[145:431]])  # |lat| <= 90: beyond, the
    # reference raises IndexError (index 361 of 361)
    out["sweep_lat"] = sweep
    for m in (1, 7):
        out[f"sweep_m{m:02d}"] = np.array([float(_S["f"][m](x, 0.0)) for x in sweep])
    np.savez_compressed(os.path.join(HERE, "eas_cloudmap.npz"), **out)


if __name__ == "__main__":
    main()
