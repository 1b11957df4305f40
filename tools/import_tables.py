"""Re-encode the nu2tau NssGrid tables (scientific input data, nupyprop /
nuleptonsim outputs) into this package's data directory through this repo's own
HDF5 writer.  Layout per file (reference: src/nuspacesim/utils/grid.py:314-359):
root attrs AXIS0..AXISn, dataset ``__nss_grid_data__``, group ``__nss_grid_axes__``.

    python tools/import_tables.py /root/reference/src/nuspacesim/data
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuspacesim_b200.utils.hdf5_min import H5File, write_hdf5  # noqa: E402


def main(src):
    dst = os.path.join(ROOT, "nuspacesim_b200", "data")
    for d in ("nupyprop_tables", "nupyprop_bdhm_tables", "nuleptonsim_tables"):
        os.makedirs(os.path.join(dst, d), exist_ok=True)
        for fn in sorted(os.listdir(os.path.join(src, d))):
            if not fn.endswith(".h5"):
                continue
            f = H5File(os.path.join(src, d, fn))
            data = np.asarray(f["__nss_grid_data__"][()])
            names = [f.attrs[f"AXIS{i}"] for i in range(data.ndim)]
            axes = {n: np.asarray(f["__nss_grid_axes__"][n][()]) for n in names}
            write_hdf5(
                os.path.join(dst, d, fn),
                {"__nss_grid_data__": data,
                 "__nss_grid_axes__": {"__kind__": "group", "members": axes, "attrs": {}}},
                {f"AXIS{i}": n for i, n in enumerate(names)},
            )
            g = H5File(os.path.join(dst, d, fn))
            assert np.array_equal(g["__nss_grid_data__"][()], data)
            for n in names:
                assert np.array_equal(g["__nss_grid_axes__"][n][()], axes[n])
            print(d, fn, data.shape, names)


def import_conex(src):
    """Columns 4..9 (Nm, Xm, X0, p1, p2, p3) of the CONEX Gaisser-Hillas fit table the
    reference loads for the fluctuated profile (cphotang.py:266-276), as float64 .npy."""
    dst = os.path.join(ROOT, "nuspacesim_b200", "data", "CONEX_table")
    os.makedirs(dst, exist_ok=True)
    fn = "dumpGH_conex_pi_E17_95deg_0km_eposlhc_1394249052_211.dat"
    tab = np.loadtxt(os.path.join(src, "CONEX_table", fn), usecols=(4, 5, 6, 7, 8, 9))
    np.save(os.path.join(dst, "gh_conex_pi_E17_95deg_0km_eposlhc.npy"), tab)
    print("CONEX_table", tab.shape)


def import_cloud_maps(src):
    """The twelve monthly cloud-top pressure maps (primary HDU, 361 x 576 float32, Pa) of
    PressureMapCloud (atmosphere/clouds.py:99-107), gathered into one compressed .npz with keys
    m01..m12 (the FITS files hold ~240 distinct pressure levels each and deflate 3:1)."""
    dst = os.path.join(ROOT, "nuspacesim_b200", "data", "cloud_maps")
    os.makedirs(dst, exist_ok=True)
    out = {}
    for month in range(1, 13):
        raw = open(os.path.join(src, "cloud_maps",
                                f"nss_map_CloudTopPressure_{month:02d}.v0.fits"), "rb").read()
        cards = {raw[i:i + 8].decode().strip(): raw[i + 10:i + 80].decode().split("/")[0].strip()
                 for i in range(0, 2880, 80)}
        assert cards["BITPIX"] == "-32" and cards["NAXIS"] == "2", cards
        n1, n2 = int(cards["NAXIS1"]), int(cards["NAXIS2"])
        out[f"m{month:02d}"] = np.frombuffer(raw[2880:2880 + 4 * n1 * n2], dtype=">f4").reshape(
            n2, n1).astype(np.float32)
    np.savez_compressed(os.path.join(dst, "nss_map_CloudTopPressure.v0.npz"), **out)
    print("cloud_maps", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1])
    import_conex(sys.argv[1])
    import_cloud_maps(sys.argv[1])
