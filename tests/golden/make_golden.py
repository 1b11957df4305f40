"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's own source files are imported from /root/reference through
oracle/ref_harness.py (astropy/h5py/dask are stubbed, the pybind11 ``zsteps`` is
compiled from the reference's zsteps.cpp).  Every array written here is an
input handed to, or an output returned by, reference code:

  geometry.npz   RegionGeom.throw / accessors / find_lat_long_along_traj / mcintegral
  taus.npz       Taus.tau_exit_prob, grid_cdf_sampler(...), Taus.__call__ arithmetic
  altdec.npz     EAS.altDec
  eas.npz        CphotAng.run per event, EAS.__call__ tail
  pipeline.npz   compute()'s optical branch stage by stage on shared uniforms
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
SEED = 20240601


def cfg_variant(ref, altitude=None, max_ch=None, table=None):
    cfg = H.default_config(ref)
    if altitude is not None:
        cfg.detector.initial_position.altitude = altitude
    if max_ch is not None:
        cfg.simulation.max_cherenkov_angle = float(np.radians(max_ch))
    if table is not None:
        cfg.simulation.tau_shower.id, cfg.simulation.tau_shower.table_version = table
    return cfg


def geom_arrays(geom):
    names = ["thetaTrSubV", "costhetaTrSubV", "phiTrSubV", "phiS", "losPathLen", "thetaS",
             "costhetaNSubV", "costhetaTrSubN", "betaTrSubN", "latS", "longS",
             "elevAngVSubN", "aziAngVSubN", "event_mask"]
    return {n: getattr(geom, n) for n in names}


def main():
    ref = H.load_reference()
    rng = np.random.Generator(np.random.PCG64(SEED))
    out = {}

    # ------------------------------------------------------------ geometry
    geo = {}
    for tag, kw in (("d525", {}), ("d33", dict(altitude=33.0)),
                    ("w10", dict(max_ch=10.0)), ("d1000w20", dict(altitude=1000.0, max_ch=20.0))):
        cfg = cfg_variant(ref, **kw)
        g = ref.region_geometry.RegionGeom(cfg)
        n = 1536
        u = rng.random((4, n))
        # edge cases: u on the closed ends of [0, 1]
        u[:, 0] = [0.0, 0.0, 0.0, 0.0]
        u[:, 1] = [1.0, 1.0, 1.0, 1.0]
        u[:, 2] = [1.0, 0.5, 0.5, 0.0]
        u[:, 3] = [1.0, 0.5, 0.5, 1.0]
        with np.errstate(all="ignore"):
            g.throw(u)
            lat, lon = g.find_lat_long_along_traj(np.zeros_like(g.betas()))
            lat7, lon7 = g.find_lat_long_along_traj(np.full_like(g.betas(), 7.5))
        geo[f"{tag}_u"] = u
        for k, v in geom_arrays(g).items():
            geo[f"{tag}_{k}"] = v
        geo[f"{tag}_init_lat"], geo[f"{tag}_init_lon"] = lat, lon
        geo[f"{tag}_lat7"], geo[f"{tag}_lon7"] = lat7, lon7
        geo[f"{tag}_consts"] = np.array(
            [g.earth_radius, g.core_alt, g.minLOSpathLen, g.maxLOSpathLen,
             g.sinOfMaxThetaTrSubV, g.maxPhiS, g.minPhiS, g.mcnorm])
        geo[f"{tag}_cfg"] = np.array(
            [cfg.detector.initial_position.altitude, cfg.simulation.max_cherenkov_angle,
             cfg.simulation.max_azimuth_angle, cfg.simulation.angle_from_limb])
        # mcintegral on synthetic downstream columns
        nv = int(g.event_mask.sum())
        trig = rng.random(nv) * 40.0
        cth = np.cos(np.radians(rng.random(nv) * 3.0))
        pex = 10 ** (-6 + 5 * rng.random(nv))
        geo[f"{tag}_mc_in"] = np.stack([trig, cth, pex])
        with np.errstate(all="ignore"):
            geo[f"{tag}_mc_out"] = np.array(g.mcintegral(trig, cth, pex, 10.0, 1.0, 1.0), float)
            geo[f"{tag}_mc_out_pl"] = np.array(g.mcintegral(trig, cth, pex, 10.0, 1e-6, 1e6), float)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **geo)

    # ------------------------------------------------------------ taus
    tg = {}
    for tag, table in (("np3", ("nupyprop", "3")), ("np1", ("nupyprop", "1")),
                       ("bdhm1", ("nupyprop_bdhm", "1")), ("nls0", ("nuleptonsim", "0"))):
        cfg = cfg_variant(ref, table=table)
        t = H.make_taus(ref, cfg)
        bax = t.pexit_grid["beta_rad"]
        n = 1536
        bmax_deg = np.degrees(bax[-1])
        beta = np.radians(rng.random(n) * (bmax_deg + 1.5))  # some beyond the table
        beta[:8] = [bax[0] * 0.5, bax[0], bax[-1], bax[-1] * 1.0001, bax[3], bax[10],
                    np.radians(0.05), np.radians(0.95)]
        loge = 6.0 + 6.0 * rng.random(n)
        loge[:6] = [6.0, 12.0, 8.0, 8.25, 11.999, 6.001]
        loge[n // 2:] = 8.0  # monospectrum half
        pex = t.tau_exit_prob(beta.copy(), loge.copy())
        # tau-energy: reference sampler called directly with the beta clamp of
        # taus.py:141-144 applied by hand (its own u plumbing is broken)
        sampler = ref.cdf.grid_cdf_sampler(t.tau_cdf_grid)
        cb = t.tau_cdf_grid["beta_rad"]
        u = rng.random(n)
        u[:4] = [1e-300, 1.0 - 1e-12, 0.5, 1e-9]
        bc = np.clip(beta, cb[0], cb[-1])
        first = ref.cdf.interpn((t.tau_cdf_grid["log_e_nu"], cb), t.tau_cdf_grid.data, (loge, bc))[:, 0]
        u = np.maximum(u, first * (1 + 1e-9) + 1e-300)  # u in (cdf[0], 1)
        z = sampler(loge, bc, u)
        z = np.where(beta > cb[-1], np.finfo(np.float32).eps, z)
        etau = z * 10**loge
        shower = cfg.simulation.tau_shower.etau_frac * etau / 1e8
        lor = etau / ref.taus.massTau
        with np.errstate(invalid="ignore"):
            tb = np.sqrt(1.0 - np.reciprocal(lor**2))
        for k, v in dict(beta=beta, loge=loge, u=u, pexit=pex, z=z, tauEnergy=etau,
                         showerEnergy=shower, tauLorentz=lor, tauBeta=tb).items():
            tg[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "taus.npz"), **tg)

    # ------------------------------------------------------------ altDec
    cfg = cfg_variant(ref)
    e = ref.eas.EAS(cfg)
    n = 1024
    beta = np.radians(0.01 + 41.9 * rng.random(n))
    lor = 10 ** (3 + 6 * rng.random(n))
    tb = np.sqrt(1 - 1 / lor**2)
    u = rng.random(n)
    u[:3] = [1.0, 1e-12, 0.5]
    alt, length = e.altDec(beta, tb, lor, u=u)
    np.savez_compressed(os.path.join(HERE, "altdec.npz"), beta=beta, tauBeta=tb,
                        tauLorentz=lor, u=u, altDec=alt, lenDec=length)

    # ------------------------------------------------------------ pipeline + EAS events
    pipe = {}
    eas_rows = []
    for tag, kw, loge, n in (("c1", {}, 8.0, 1024), ("e9", {}, 9.0, 256),
                             ("e10", {}, 10.0, 256), ("d33", dict(altitude=33.0), 8.0, 384),
                             ("e7", {}, 7.0, 192)):
        cfg = cfg_variant(ref, **kw)
        cfg.simulation.spectrum.log_nu_energy = loge
        g = ref.region_geometry.RegionGeom(cfg)
        t = H.make_taus(ref, cfg)
        e = ref.eas.EAS(cfg)
        u4 = rng.random((4, n))
        u_tau = rng.random(n)
        u_dec = rng.random(n)
        g.throw(u4)
        m = g.event_mask
        beta = g.beta_rad()
        lat, lon = g.find_lat_long_along_traj(np.zeros_like(beta))
        le = np.full(beta.shape, loge)
        pex = t.tau_exit_prob(beta.copy(), le)
        cb = t.tau_cdf_grid["beta_rad"]
        z = ref.cdf.grid_cdf_sampler(t.tau_cdf_grid)(le, np.clip(beta, cb[0], cb[-1]), u_tau[m])
        etau = z * 10**le
        shower = cfg.simulation.tau_shower.etau_frac * etau / 1e8
        lor = etau / ref.taus.massTau
        tb = np.sqrt(1.0 - np.reciprocal(lor**2))
        alt, length = e.altDec(beta, tb, lor, u=u_dec[m])
        # per-event engine outputs (before the EAS.__call__ tail)
        msk = ~((alt < 0.0) | (alt > 20.0))
        dph = np.zeros_like(beta)
        cang = np.full_like(beta, 1.5)
        for i in np.nonzero(msk)[0]:
            d, c = e.CphotAng.run(beta[i], alt[i], shower[i], lat[i], lon[i], None)
            dph[i], cang[i] = d, c
        npe, cth = e(beta, alt, shower, lat, lon, cloudf=None)
        mc = g.mcintegral(npe, cth, pex, cfg.detector.optical.photo_electron_threshold, 1.0, 1.0)
        pipe[f"{tag}_cfg"] = np.array([cfg.detector.initial_position.altitude, loge, n])
        for k, v in dict(u4=u4, u_tau=u_tau, u_dec=u_dec, event_mask=m, beta_rad=beta,
                         tauExitProb=pex, tauEnergy=etau, showerEnergy=shower, tauLorentz=lor,
                         tauBeta=tb, altDec=alt, lenDec=length, dphots=dph, cang=cang,
                         numPEs=npe, costhetaChEff=cth, init_lat=lat, init_lon=lon,
                         mc=np.array(mc, float)).items():
            pipe[f"{tag}_{k}"] = v
        print(tag, "mc =", mc, "valid", int(m.sum()))
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **pipe)

    # engine sweep: the reference's own golden series (beta=42 deg, alt 0..20) plus
    # a grid over beta / altitude / energy / detector altitude / cloud top
    rows = []
    for det in (525.0, 33.0):
        cp = ref.cphotang.CphotAng(det, "Greisen")
        for b_deg in (0.3, 1.0, 2.0, 3.0, 5.0, 10.0, 20.0, 30.0, 42.0):
            for alt in (0.0, 0.5, 1.0, 2.5, 4.0, 6.0, 8.0, 11.0, 14.0, 17.0, 20.0):
                for es in ((1.0, 0.003, 0.05, 30.0) if det == 525.0 else (1.0,)):
                    if det == 33.0 and alt > 19.0:
                        continue
                    d, c = cp.run(np.radians(b_deg), alt, es, 0.0, 0.0, None)
                    rows.append((det, np.radians(b_deg), alt, es, -np.inf, float(d), float(c)))
    cp = ref.cphotang.CphotAng(525.0, "Greisen")
    for top in (2.0, 6.0, 12.0):
        for b_deg in (1.0, 5.0, 20.0):
            for alt in (0.0, 3.0, 8.0, 15.0):
                d, c = cp.run(np.radians(b_deg), alt, 1.0, 0.0, 0.0, lambda *a: np.float32(top))
                rows.append((525.0, np.radians(b_deg), alt, 1.0, top, float(d), float(c)))
    ser = [cp.run(np.radians(42.0), float(a), 1.0, 0.0, 0.0, None) for a in range(21)]
    np.savez_compressed(os.path.join(HERE, "eas.npz"), sweep=np.array(rows),
                        series42=np.array(ser, float))
    print("sweep rows", len(rows))


if __name__ == "__main__":
    main()
