"""Generate tests/golden/taus_raw.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_taus_raw.py

Taus.tau_energy(betas, log_e_nu) called WITHOUT ``u`` -- the path compute() takes -- draws
``np.random.uniform(cdfs[:, 0], 1.0)`` inside grid_cdf_sampler (utils/cdf.py:419-423).  With the
legacy global generator seeded, those draws are ``low + (high - low) * random_sample()``; the
fixture stores the generator's raw (0, 1) outputs next to the reference's results so that the
oracle and the kernels (u_raw) can be checked on the nuleptonsim tables, whose first CDF column
is NOT zero at high energies (up to 0.0134), and on the default nupyprop tables.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402

warnings.filterwarnings("ignore", category=DeprecationWarning)


def main():
    ref = H.load_reference()
    out = {}
    for tag, table, lo, hi in (("nls0", ("nuleptonsim", "0"), 10.0, 12.0),
                               ("np3", ("nupyprop", "3"), 6.0, 12.0)):
        cfg = H.default_config(ref)
        cfg.simulation.tau_shower.id, cfg.simulation.tau_shower.table_version = table
        t = H.make_taus(ref, cfg)
        bax = t.tau_cdf_grid["beta_rad"]
        n = 4096  # one nditer chunk (8192) -> one draw of n uniforms, in event order
        rng = np.random.default_rng(77)
        beta = bax[0] + (bax[-1] - bax[0]) * rng.random(n)  # inside the table: one sampler call
        loge = lo + (hi - lo) * rng.random(n)
        loge[:4] = [hi, hi, 11.5, 11.5]
        seed = 424242
        np.random.seed(seed)
        te = t.tau_energy(beta, loge)
        u_raw = np.random.RandomState(seed).random_sample(n)
        out.update({f"{tag}_beta": beta, f"{tag}_loge": loge, f"{tag}_u_raw": u_raw,
                    f"{tag}_tauEnergy": np.asarray(te, float)})
        c0 = t.tau_cdf_grid.data[..., 0].max()
        print(tag, "max cdf[..., 0] =", c0)
    np.savez_compressed(os.path.join(HERE, "taus_raw.npz"), **out)


if __name__ == "__main__":
    main()
