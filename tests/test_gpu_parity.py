"""Parity of the CUDA path (through the C ABI of libnss_b200.so) with the oracle and the
reference-generated golden vectors, plus size-independent properties at large N.

Tolerances (stated):
  * float64 stages (geometry, spectrum, taus, altDec, mcintegral): rtol 1e-12; masks and
    counts exact;
  * float32 EAS engine: photon density / numPEs per event: median < 1e-4, 99th percentile
    < 1e-3, max < 5e-3.  The tail is the reference's own float32 noise: CherArea uses
    DistStep[argmax RN] (cphotang.py:653,617-621) and for showers that develop above ~35 km
    the Greisen profile is flat over ~20 altitude steps at float32 resolution, so WHICH step
    is the maximum is decided by rounding in np.cumsum/np.exp (one step = 1.8e-4 in area);
    Cherenkov angle rtol 3e-5; costhetaChEff rtol 1e-6; trigger mask (numPEs >= threshold)
    exact on every golden set.
"""
import numpy as np
import pytest

from conftest import (CLOUDMAP_TAGS, EAS_LARGE_TAGS, GEOM_TAGS, GH_PROFILES, GH_TAGS, PIPE_TAGS,
                      TAU_TABLES, check_engine_vs_reference, cloud_pressure_map, conex_rows,
                      load_tables, make_config, philox_uniforms, rel_err)
from oracle import nss_oracle as O

pytestmark = pytest.mark.gpu

GEOM_ARRAYS = ("thetaTrSubV", "costhetaTrSubV", "phiTrSubV", "phiS", "losPathLen", "thetaS",
               "costhetaNSubV", "costhetaTrSubN", "betaTrSubN", "latS", "longS", "elevAngVSubN",
               "aziAngVSubN")


@pytest.fixture(scope="module")
def nss():
    import torch

    assert torch.cuda.is_available()
    import nuspacesim_b200

    return nuspacesim_b200


def _geom_cfg(geo, tag):
    alt, mch, maz, lim = geo[tag + "_cfg"]
    return make_config(altitude=alt, max_ch=mch, max_az=maz, limb=lim)


@pytest.mark.parametrize("tag", GEOM_TAGS)
def test_geometry_vs_reference(nss, golden, tag):
    geo = golden["geometry"]
    g = nss.RegionGeom(_geom_cfg(geo, tag))
    np.testing.assert_allclose(
        [g.earth_radius, g.core_alt, g.minLOSpathLen, g.maxLOSpathLen, g.sinOfMaxThetaTrSubV,
         g.maxPhiS, g.minPhiS, g.mcnorm], geo[tag + "_consts"], rtol=1e-14)
    g.throw(geo[tag + "_u"])
    want_mask = geo[tag + "_event_mask"]
    assert np.array_equal(g.event_mask, want_mask)  # bit-exact mask
    for k in GEOM_ARRAYS:
        got, want = getattr(g, k), geo[f"{tag}_{k}"]
        assert np.array_equal(np.isnan(got), np.isnan(want)), k
        ok = ~np.isnan(want)
        # angles near 0 (acos(1-eps)) are compared absolutely
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-11, atol=1e-12, err_msg=k)
    beta, th, pl = g.beta_rad(), g.thetas(), g.pathLens()
    assert beta.size == want_mask.sum()
    np.testing.assert_allclose(beta, np.radians(geo[tag + "_betaTrSubN"][want_mask]), rtol=1e-11,
                               atol=1e-14)
    lat, lon = g.find_lat_long_along_traj(np.zeros_like(beta))
    np.testing.assert_allclose(lat, geo[tag + "_init_lat"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(lon, geo[tag + "_init_lon"], rtol=1e-10, atol=1e-13)
    lat, lon = g.find_lat_long_along_traj(np.full_like(beta, 7.5))
    np.testing.assert_allclose(lat, geo[tag + "_lat7"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(lon, geo[tag + "_lon7"], rtol=1e-10, atol=1e-13)
    trig, cth, pex = geo[tag + "_mc_in"]
    got = g.mcintegral(trig, cth, pex, 10.0, 1.0, 1.0)
    np.testing.assert_allclose(got, geo[tag + "_mc_out"], rtol=1e-11)
    got = g.mcintegral(trig, cth, pex, 10.0, 1e-6, 1e6)
    np.testing.assert_allclose(got, geo[tag + "_mc_out_pl"], rtol=1e-11)
    assert got[2] == geo[tag + "_mc_out_pl"][2]  # pass count exact


def test_geometry_errors_and_edges(nss):
    g = nss.RegionGeom(make_config())
    with pytest.raises(RuntimeError):
        g.throw(None)
    with pytest.raises(RuntimeError):
        g.throw(np.zeros((3, 10)))
    g.throw(np.zeros((4, 0)))  # empty input
    assert g.beta_rad().size == 0
    b, t, p = g(17)
    assert b.size == t.size == p.size <= 17
    # store= hook receives the reference's column names
    seen = {}
    g(33, store=lambda names, cols: seen.update(zip(names, cols)))
    assert list(seen) == ["beta_rad", "theta_rad", "path_len"]
    # Philox throw == throw(u) on the same uniforms
    g2 = nss.RegionGeom(make_config(), seed=99)
    g2.first_event = 12345
    g2.throw(1000)
    u4 = philox_uniforms(99, 12345, 1000)[0]
    g3 = nss.RegionGeom(make_config())
    g3.throw(u4)
    assert np.array_equal(g2.betaTrSubN, g3.betaTrSubN)
    assert np.array_equal(g2.event_mask, g3.event_mask)


def test_no_valid_events_yields_empty_results(nss, capsys):
    """Thrown events that are ALL masked out (compute.py:200-204: 'No valid events thrown'):
    empty columns, no error -- for the accessors, the one-shot and the sliced pipeline."""
    cfg = make_config(max_ch=np.radians(20.0))
    p = O.OracleParams(max_cherenkov_angle=np.radians(20.0))
    rng = np.random.default_rng(11)
    u = rng.random((4, 4000))
    bad = ~O.geom_throw(O.geom_constants(p), u)["event_mask"]
    assert bad.sum() >= 3
    u_bad = np.ascontiguousarray(u[:, bad][:, :3])
    g = nss.RegionGeom(cfg)
    b, t, pl = g(u_bad)
    assert b.size == t.size == pl.size == 0 and not g.event_mask.any()
    assert g.compact(np.arange(3.0), device=False).size == 0
    lat, lon = g.find_lat_long_along_traj(np.zeros(0))
    assert lat.size == lon.size == 0
    U = dict(u4=u_bad, u_tau=np.full(3, 0.5), u_dec=np.full(3, 0.5))
    sim = nss.compute(cfg, uniforms=U)
    assert len(sim) == 0
    assert "No valid events" in capsys.readouterr().out
    # a slice without valid events between slices that have some
    u_mix = np.ascontiguousarray(np.concatenate([u[:, :5], u_bad, u[:, 5:9]], axis=1))
    n = u_mix.shape[1]
    Um = dict(u4=u_mix, u_tau=np.full(n, 0.5), u_dec=np.full(n, 0.5))
    one = nss.compute(cfg, uniforms=Um, slice_events=0)
    sliced = nss.compute(cfg, uniforms=Um, slice_events=1)  # events 5, 6, 7 are empty slices
    assert 0 < len(one) == len(sliced)
    for c in one.colnames:
        assert np.array_equal(np.asarray(one[c]), np.asarray(sliced[c]), equal_nan=True), c


@pytest.mark.parametrize("tag", list(TAU_TABLES))
def test_taus_vs_reference(nss, golden, tag):
    tg = golden["taus"]
    t = nss.Taus(make_config(table=TAU_TABLES[tag]))
    beta, loge, u = tg[tag + "_beta"], tg[tag + "_loge"], tg[tag + "_u"]
    np.testing.assert_allclose(t.tau_exit_prob(beta, loge), tg[tag + "_pexit"], rtol=1e-12)
    np.testing.assert_allclose(t.tau_energy(beta, loge, u), tg[tag + "_tauEnergy"], rtol=1e-12)
    tb, lor, te, sh, pex = t(beta, loge, u=u)
    np.testing.assert_allclose(pex, tg[tag + "_pexit"], rtol=1e-12)
    np.testing.assert_allclose(te, tg[tag + "_tauEnergy"], rtol=1e-12)
    np.testing.assert_allclose(sh, tg[tag + "_showerEnergy"], rtol=1e-12)
    np.testing.assert_allclose(lor, tg[tag + "_tauLorentz"], rtol=1e-12)
    ok = ~np.isnan(tg[tag + "_tauBeta"])
    np.testing.assert_allclose(tb[ok], tg[tag + "_tauBeta"][ok], rtol=1e-12)


def test_taus_errors_and_sampling(nss):
    t = nss.Taus(make_config())
    with pytest.raises(ValueError):  # scipy bounds_error
        t.tau_exit_prob(np.array([0.1, 0.2]), np.array([8.0, 12.5]))
    with pytest.raises(ValueError):
        t.tau_energy(np.array([0.1]), np.array([5.9]), np.array([0.5]))
    with pytest.raises(ValueError):  # u == cdf[0]
        t.tau_energy(np.array([0.1]), np.array([8.0]), np.array([0.0]))
    with pytest.raises(RuntimeError):
        t.tau_energy(np.array([0.1, 0.2]), np.array([8.0]), np.array([0.5, 0.5]))
    assert t.tau_energy(np.array([]), np.array([]), np.array([])).size == 0
    # the reference's own statistical test (test/utils/test_cdf.py:24-41): random sampling
    n = 200000
    rng = np.random.default_rng(5)
    loge = rng.uniform(6, 12, n)
    beta = np.radians(rng.uniform(1, 42, n))
    e = t.tau_energy(beta, loge)
    z = e / 10**loge
    assert not np.isnan(z).any() and not np.all(z == 0) and not np.all(z == 1)
    assert z.min() >= t.tau_cdf_grid["e_tau_frac"][0] and z.max() <= 1.0 + 1e-12


@pytest.mark.parametrize("tag", ["nls0", "np3"])
def test_tau_energy_raw_uniforms(nss, tag):
    """Uniforms the path draws itself (host generator or Philox) are mapped onto (cdf[0], 1) of
    each event's blended CDF row like the reference's np.random.uniform(cdfs[:, 0], 1.0)
    (utils/cdf.py:419-423): bit-level agreement with the reference run on the same generator
    outputs (tests/golden/make_golden_taus_raw.py), also where the nuleptonsim CDF starts at
    0.0134; compute() on that table at high energy no longer leaves the table."""
    import os

    from conftest import GOLDEN

    G = np.load(os.path.join(GOLDEN, "taus_raw.npz"))
    t = nss.Taus(make_config(table=TAU_TABLES[tag]))
    beta, loge, u = G[tag + "_beta"], G[tag + "_loge"], G[tag + "_u_raw"]
    np.testing.assert_allclose(t.tau_energy(beta, loge, u, u_raw=True), G[tag + "_tauEnergy"],
                               rtol=1e-12)
    np.testing.assert_allclose(t(beta, loge, u=u, u_raw=True)[2], G[tag + "_tauEnergy"],
                               rtol=1e-12)
    z = t.tau_energy(beta, loge) / 10**loge  # own draws: never outside the table
    assert np.isfinite(z).all() and z.min() >= t.tau_cdf_grid["e_tau_frac"][0]
    if tag == "nls0":
        with pytest.raises(ValueError):  # a caller's u is used as it is, like the reference's
            t.tau_energy(beta, loge, u)
        cfg = make_config(table=TAU_TABLES[tag], loge=11.5, n=20000)
        sim = nss.compute(cfg, seed=3)  # Philox uniforms, ~1 % of them below cdf[0] = 0.0134
        assert len(sim) > 19000 and np.isfinite(sim["tauEnergy"]).all()
        r = nss.run_acceptance(cfg, 20000, seed=3)
        assert np.isfinite(r[0])


def test_altdec_and_spectrum(nss, golden):
    a = golden["altdec"]
    e = nss.EAS(make_config())
    alt, ln = e.altDec(a["beta"], a["tauBeta"], a["tauLorentz"], a["u"])
    np.testing.assert_allclose(ln, a["lenDec"], rtol=1e-13)
    np.testing.assert_allclose(alt, a["altDec"], rtol=1e-9, atol=1e-12)  # sqrt(...) - R cancels
    from nuspacesim_b200.config import PowerSpectrum

    c = make_config()
    c.simulation.spectrum = PowerSpectrum(index=2.0, lower_bound=6.0, upper_bound=12.0)
    u = np.random.default_rng(2).random(5000)
    le, sn, sw = nss.Spectra(c)(5000, u=u)
    p = O.OracleParams(spectrum="powerspectrum")
    want, wsn, wsw = O.spectra(p, 5000, u)
    np.testing.assert_allclose(le, want, rtol=1e-13)
    assert sn == wsn and sw == wsw
    le0, _, _ = nss.Spectra(make_config(loge=9.5))(7)
    assert np.array_equal(le0, np.full(7, 9.5))


def test_engine_reference_known_answers(nss):
    """The reference's own golden vectors (test/simulation/eas_optical/cphot_ang.py:16-66)."""
    from test_oracle_golden import OLD_CANG, OLD_DPHOTS

    cp = nss.CphotAng(525.0, "Greisen")
    z = np.zeros(21)
    d, c = cp(np.full(21, np.radians(42.0)), np.arange(21.0), np.ones(21), z, z, None)
    assert rel_err(d, OLD_DPHOTS).max() < 5e-4
    assert rel_err(c, OLD_CANG).max() < 5e-6
    assert cp([], [], [], [], [])[0].shape == ()  # empty input convention (:709-716)
    with pytest.raises(ValueError):
        nss.CphotAng(525.0, "no such profile")


def test_engine_sweep_vs_reference_and_oracle(nss, golden):
    sw = golden["eas"]["sweep"]
    for det in (525.0, 33.0):
        cp = nss.CphotAng(det, "Greisen")
        for top in (-np.inf, 2.0, 6.0, 12.0):
            m = (sw[:, 0] == det) & (sw[:, 4] == top)
            if not m.any():
                continue
            z = np.zeros(int(m.sum()))
            cloudf = None if top == -np.inf else (lambda *a, _t=top: np.float32(_t))
            d, c = cp(sw[m, 1], sw[m, 2], sw[m, 3], z, z, cloudf)
            assert np.array_equal(d == 0, sw[m, 5] == 0)  # early-outs identical
            rr = rel_err(d, sw[m, 5])
            assert rr.max() < 5e-3 and np.median(rr) < 1e-4, (det, top, rr.max())
            assert np.percentile(rr, 99) < 1e-3 or rr.size < 200, (det, top)
            assert rel_err(c, sw[m, 6]).max() < 3e-5
            od, oc, _ = O.cphotang(sw[m, 1], sw[m, 2], sw[m, 3], det,
                                   None if top == -np.inf else top)
            ro = rel_err(d, od)
            assert ro.max() < 5e-3 and np.median(ro) < 2e-5


@pytest.mark.parametrize("tag", PIPE_TAGS)
def test_compute_pipeline_vs_reference(nss, golden, tag):
    """compute() on the uniforms the reference consumed: every column + acceptance."""
    pp = golden["pipeline"]
    alt, loge, n = pp[tag + "_cfg"]
    cfg = make_config(altitude=alt, loge=loge, n=int(n))
    U = dict(u4=pp[tag + "_u4"], u_tau=pp[tag + "_u_tau"], u_dec=pp[tag + "_u_dec"])
    sim = nss.compute(cfg, uniforms=U)
    assert sim.colnames == ["beta_rad", "theta_rad", "path_len", "init_lat", "init_lon",
                            "log_e_nu", "tauBeta", "tauLorentz", "tauEnergy", "showerEnergy",
                            "tauExitProb", "altDec", "lenDec", "numPEs", "costhetaChEff"]
    assert len(sim) == int(pp[tag + "_event_mask"].sum())
    for k in ("beta_rad", "tauBeta", "tauLorentz", "tauEnergy", "showerEnergy", "tauExitProb",
              "lenDec", "init_lat", "init_lon"):
        np.testing.assert_allclose(sim[k], pp[f"{tag}_{k}"], rtol=1e-10, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(sim["altDec"], pp[tag + "_altDec"], rtol=1e-8, atol=1e-12)
    rd = rel_err(sim["numPEs"], pp[tag + "_numPEs"])
    assert rd.max() < 5e-3 and np.percentile(rd, 99) < 1e-3 and np.median(rd) < 1e-4
    assert rel_err(sim["costhetaChEff"], pp[tag + "_costhetaChEff"]).max() < 1e-6
    assert np.array_equal(sim["numPEs"] >= 10, pp[tag + "_numPEs"] >= 10)  # trigger mask
    ref = pp[tag + "_mc"]
    assert sim.meta["ONEVPASS"][0] == ref[2]
    np.testing.assert_allclose(sim.meta["OMCINTGO"][0], ref[1], rtol=1e-11)
    np.testing.assert_allclose(sim.meta["OMCINT"][0], ref[0], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(sim.meta["OMCINTUN"][0], ref[3], rtol=1e-9, atol=1e-300)


def test_compute_writes_fits(nss, tmp_path):
    from nuspacesim_b200 import results_table as R

    cfg = make_config(n=300)
    fn = tmp_path / "run.fits"
    sim = nss.compute(cfg, output_file=str(fn), seed=3)
    assert not fn.exists()  # like the reference: only write_stages writes inside compute()
    sim.write(str(fn), format="fits", overwrite=True)  # apps/run.py:209-210
    back = R.read_fits_table(str(fn))
    assert back.colnames == sim.colnames and len(back) == len(sim)
    fn2 = tmp_path / "staged.fits"
    sim2 = nss.compute(cfg, output_file=str(fn2), seed=3, write_stages=True)
    back = R.read_fits_table(str(fn2))
    assert np.array_equal(np.asarray(sim2["numPEs"]), np.asarray(sim["numPEs"]))
    assert back.colnames == sim.colnames and len(back) == len(sim)
    assert np.array_equal(back["numPEs"], sim["numPEs"])
    assert back.meta["ONEVPASS"] == sim.meta["ONEVPASS"][0]


def test_acceptance_mode_equals_oracle_on_same_stream(nss, tables_v3):
    """nss_run_acceptance (in-kernel Philox) vs the oracle fed the identical uniforms."""
    cfg = make_config()
    n, seed, first = 3000, 77, 10**10
    r = nss.AcceptanceRunner(cfg, batch=1024, collect_stats=True)  # 3 batches, ragged tail
    r.enqueue(seed, first, n)
    got = r.finalize(r.sums())
    u4, us, ut, ud, _ = philox_uniforms(seed, first, n)
    mc, cols, t, st = O.run_path(O.OracleParams(), tables_v3, u4, us, ut, ud, nthreads=4)
    s = r.sums()
    assert s[4] == t["event_mask"].sum() and s[5] == n
    assert got[2] == mc[2]  # passing events exact
    np.testing.assert_allclose(got[1], mc[1], rtol=1e-11)
    np.testing.assert_allclose(got[0], mc[0], rtol=1e-11)
    np.testing.assert_allclose(got[3], mc[3], rtol=1e-6)
    ks = r.stats.cpu().numpy()
    assert ks[0] == st["n_raw"] * 0 + np.count_nonzero(~((cols["altDec"] < 0) | (cols["altDec"] > 20)))
    # work counters: floor(DistStep * tan(thetaC)) flips by one on a few steps in 1e5
    assert abs(ks[2] - st["n_val"]) <= 4 and abs(ks[3] - st["J"]) <= 1e-5 * st["J"]


CONFIG_CASES = {
    # BASELINE.json configs[2]: monospectrum energy sweep
    "e7": dict(loge=7.0), "e7.5": dict(loge=7.5), "e8.5": dict(loge=8.5), "e9": dict(loge=9.0),
    "e9.5": dict(loge=9.5), "e10": dict(loge=10.0),
    # configs[3]: power-law spectrum, 525 km, wide theta_ch_max
    "powerlaw_w10": dict(power=(2.0, 6.0, 12.0), max_ch=10.0),
    # balloon altitude, cloud deck, narrow azimuth
    "d33": dict(altitude=33.0), "cloud6": dict(cloud=6.0), "az90_limb4": dict(max_az=90.0, limb=4.0),
}


@pytest.mark.parametrize("tag", sorted(CONFIG_CASES))
def test_acceptance_configs_vs_oracle(nss, tables_v3, tag):
    """In-kernel-RNG acceptance for the BASELINE.json configurations against the oracle fed the
    identical Philox uniforms: counts exact, integrals to FP64 round-off, per-event columns of
    compute() within the stated FP32 tolerance."""
    import math

    from nuspacesim_b200.config import MonoCloud, PowerSpectrum

    kw = dict(CONFIG_CASES[tag])
    power, cloud = kw.pop("power", None), kw.pop("cloud", None)
    cfg = make_config(**{k: (math.radians(v) if k in ("max_ch", "max_az", "limb") else v)
                         for k, v in kw.items()})
    op = O.OracleParams(detector_altitude=cfg.detector.initial_position.altitude,
                        max_cherenkov_angle=cfg.simulation.max_cherenkov_angle,
                        max_azimuth_angle=cfg.simulation.max_azimuth_angle,
                        angle_from_limb=cfg.simulation.angle_from_limb,
                        log_nu_energy=cfg.simulation.spectrum.log_nu_energy)
    if power is not None:
        cfg.simulation.spectrum = PowerSpectrum(index=power[0], lower_bound=power[1],
                                                upper_bound=power[2])
        op.spectrum, op.spec_index, op.spec_lower, op.spec_upper = "powerspectrum", *power
    if cloud is not None:
        cfg.simulation.cloud_model = MonoCloud(altitude=cloud)
        op.cloud_top = cloud
    n, seed, first = 2500, 5, 123456789
    r = nss.AcceptanceRunner(cfg, batch=1 << 12)
    r.enqueue(seed, first, n)
    s = r.sums()
    got = r.finalize(s)
    u4, us, ut, ud, _ = philox_uniforms(seed, first, n)
    mc, cols, t, st = O.run_path(op, tables_v3, u4, us if power is not None else None, ut, ud,
                                 nthreads=4)
    assert s[4] == t["event_mask"].sum() and s[5] == n
    np.testing.assert_allclose(got[1], mc[1], rtol=1e-11)  # geometry-only integral
    # trigger decisions sit downstream of FP32 arithmetic: allow events whose numPEs is within
    # the engine tolerance of the threshold to flip
    thr = cfg.detector.optical.photo_electron_threshold
    near = np.count_nonzero(np.abs(cols["numPEs"] / thr - 1) < 2e-3)
    assert abs(got[2] - mc[2]) <= near
    if near == 0:
        np.testing.assert_allclose(got[0], mc[0], rtol=1e-9)
        np.testing.assert_allclose(got[3], mc[3], rtol=1e-6)
    # the same events through compute(): per-event columns
    U = dict(u4=u4, u_tau=ut, u_dec=ud)
    if power is not None:
        U["u_spec"] = us
    sim = nss.compute(cfg, uniforms=U)
    for k in ("beta_rad", "log_e_nu", "tauExitProb", "tauEnergy", "altDec"):
        np.testing.assert_allclose(sim[k], cols[k], rtol=1e-9, atol=0, err_msg=k)
    live = cols["numPEs"] > 0
    assert np.array_equal(sim["numPEs"] > 0, live)
    re = rel_err(sim["numPEs"][live], cols["numPEs"][live])
    # Conditioning of the reference itself: an event whose first retained step has
    # e2hill = 1150 + 454 ln(s) just above 0 (cphotang.py:561-564) gets an arbitrarily large
    # Hillas weight ~ 1/e2hill (:496-507) and its photon density responds to a 1e-5 relative
    # change of the shower energy by several 1e-3 (normal events: 2e-5).  Such events (about
    # 1 in 1000) are bounded by 5x the oracle's own sensitivity instead of the flat 2.5e-3.
    # The same holds for a step whose quantised solid angle sits on a rounding tie (see
    # O.conditioning).
    sens = O.conditioning(cols["beta_rad"][live], cols["altDec"][live],
                          cols["showerEnergy"][live], op.detector_altitude, op.cloud_top,
                          nthreads=4)
    bound = np.maximum(2.5e-3, 5 * sens)
    assert (re < bound).all() and np.median(re) < 1e-4, (re.max(), np.median(re))
    assert np.count_nonzero(re > 2.5e-3) <= max(1, live.sum() // 500)
    np.testing.assert_allclose(sim["costhetaChEff"], cols["costhetaChEff"], rtol=1e-6)


def test_acceptance_properties_at_scale(nss):
    """Size-independent properties at 1e6 thrown events: determinism, shard invariance
    (any split of the thrown-index range gives the same sums), MC consistency."""
    cfg = make_config()
    n = 1_000_000
    r = nss.AcceptanceRunner(cfg, batch=1 << 19)
    r.enqueue(1234, 0, n)
    a = r.sums()
    r.reset()
    r.enqueue(1234, 0, n)
    assert np.array_equal(a, r.sums())  # bit-reproducible
    r.reset()
    for lo, cnt in ((0, 300_001), (300_001, 199_999), (500_000, 500_000)):
        r.enqueue(1234, lo, cnt)
    b = r.sums()
    assert np.array_equal(a[3:6], b[3:6])  # counts exact under re-sharding
    np.testing.assert_allclose(a[:3], b[:3], rtol=1e-12)
    mcint, geo, npass, unc = r.finalize(a)
    # reference at N = 1e4: 0.1291 +- 0.0040, geo 3792.6
    assert abs(mcint - 0.1291) < 4 * np.hypot(0.0040, unc)
    assert abs(geo - 3792.6) / 3792.6 < 0.05
    assert 0.98 < a[4] / n < 1.0
    r.reset()
    r.enqueue(4321, 0, n)  # another seed: independent estimate agrees within errors
    m2, _, _, u2 = r.finalize(r.sums())
    assert abs(m2 - mcint) < 5 * np.hypot(unc, u2)


def test_acceptance_at_baseline_size(nss):
    """BASELINE.json configs[2] size: 1e8 thrown events of the default physics through the fused
    mode (1.7 s on a B200).  Size-independent properties: re-sharding the thrown-index range
    over 8 "ranks" reproduces the counts exactly and the sums to 1e-12; the integral agrees with
    the 1e6-event estimate of another seed within the combined Monte Carlo errors, and its own
    error scales like 1/sqrt(N)."""
    cfg = make_config()
    n = 100_000_000
    r = nss.AcceptanceRunner(cfg)
    r.enqueue(1234, 0, n)
    a = r.sums()
    mcint, geo, npass, unc = r.finalize(a)
    assert a[5] == n and 0.99 < a[4] / n < 0.995
    r.reset()
    from nuspacesim_b200.compute import shard_range

    for rank in range(8):
        lo, cnt = shard_range(n, rank, 8)
        r.enqueue(1234, lo, cnt)
    b = r.sums()
    assert np.array_equal(a[3:6], b[3:6])
    np.testing.assert_allclose(a[:3], b[:3], rtol=1e-12)
    r.reset()
    r.enqueue(99, 0, 1_000_000)
    m2, _, _, u2 = r.finalize(r.sums())
    assert abs(m2 - mcint) < 5 * np.hypot(unc, u2)
    assert 0.08 < unc / u2 < 0.125  # sqrt(1e6 / 1e8) = 0.1
    assert abs(geo - 3700.0) / 3700.0 < 0.01


def test_mask_scan_and_compaction_ragged(nss):
    import ctypes as C

    import torch

    from nuspacesim_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(0)
    for n in (1, 31, 4096, 4097, 100_003):
        m = (rng.random(n) < 0.37).astype(np.uint8)
        x = rng.random(n)
        md, xd = _lib.to_device(m, "uint8"), _lib.to_device(x)
        pos, nset = _lib.empty(n, "int64"), _lib.zeros(1, "int64")
        ws = _lib.empty(int(L.nss_scan_workspace_bytes(n)), "uint8")
        _lib.check(L.nss_mask_scan(_lib.ptr(md), n, _lib.ptr(pos), _lib.ptr(nset), _lib.ptr(ws),
                                   _lib.stream_ptr()))
        k = int(nset.item())
        assert k == m.sum()
        assert np.array_equal(pos.cpu().numpy(), np.cumsum(m) - m)
        if k == 0:  # nothing to compact: a zero-length buffer has no address
            continue
        out = _lib.empty(k)
        _lib.check(L.nss_compact_f64(_lib.ptr(md), _lib.ptr(pos), _lib.ptr(xd), _lib.ptr(out), n,
                                     0, _lib.stream_ptr()))
        assert np.array_equal(out.cpu().numpy(), x[m.astype(bool)])
        # the multi-column form: one launch, mask and positions read once, per-column scaling
        ncols = 5
        X = rng.random((ncols, n))
        ins = [_lib.to_device(X[c]) for c in range(ncols)]
        outs = [_lib.empty(k) for _ in range(ncols)]
        rad = [0, 1, 0, 1, 0]
        rc = L.nss_compact_columns_f64(
            _lib.ptr(md), _lib.ptr(pos), (C.c_void_p * ncols)(*[t.data_ptr() for t in ins]),
            (C.c_void_p * ncols)(*[t.data_ptr() for t in outs]), (C.c_int * ncols)(*rad), ncols,
            n, _lib.stream_ptr())
        assert rc == 0
        for c in range(ncols):
            want = X[c][m.astype(bool)] * (np.pi / 180.0 if rad[c] else 1.0)
            assert np.array_equal(outs[c].cpu().numpy(), want), c
    assert L.nss_compact_columns_f64(None, None, None, None, None, 17, 5, None) == -1  # > 16 columns
    assert L.nss_compact_columns_f64(None, None, None, None, None, 3, 5, None) == -2


def test_cli_run_writes_results(nss, tmp_path):
    """`python -m nuspacesim_b200 run cfg.toml N --monospectrum E -o out.fits` (apps/run.py:180)."""
    from click.testing import CliRunner

    from nuspacesim_b200 import results_table as R
    from nuspacesim_b200.apps.cli import cli

    r = CliRunner()
    cfg, out = str(tmp_path / "c.toml"), str(tmp_path / "o.fits")
    assert r.invoke(cli, ["create-config", cfg]).exit_code == 0
    res = r.invoke(cli, ["run", cfg, "2e3", "--monospectrum", "9", "--monocloud", "3.0",
                         "--seed", "11", "-o", out])
    assert res.exit_code == 0, res.output
    assert "Monte Carlo Integral [Optical]" in res.output
    t = R.read_fits_table(out)
    assert len(t.colnames) == 15 and 1900 < len(t) <= 2000
    assert np.all(t["log_e_nu"] == 9.0)
    out2 = str(tmp_path / "none.fits")  # -n / --no-result-file (apps/run.py:73-78, 209-210)
    res = r.invoke(cli, ["run", cfg, "500", "--no-result-file", "-o", out2])
    assert res.exit_code == 0 and not (tmp_path / "none.fits").exists()
    acc = r.invoke(cli, ["run", cfg, "2e3", "--monospectrum", "9", "--monocloud", "3.0",
                         "--seed", "11", "--acceptance-only"])
    assert acc.exit_code == 0, acc.output
    assert f"Number of Passing Events [Optical]: {int(t.meta['ONEVPASS'])}" in acc.output


@pytest.mark.parametrize("power", [False, True])
def test_compute_and_acceptance_mode_see_the_same_events(nss, power):
    """compute(seed=s) (columns, stage by stage) and run_acceptance(seed=s) (fused, in-kernel
    RNG) consume the same Philox words per thrown event: identical counts and integrals."""
    from nuspacesim_b200.config import PowerSpectrum

    cfg = make_config(n=20_000)
    if power:
        cfg.simulation.spectrum = PowerSpectrum(index=2.2, lower_bound=7.0, upper_bound=10.5)
    sim = nss.compute(cfg, seed=99)
    mcint, geo, npass, unc = nss.run_acceptance(cfg, 20_000, seed=99)
    assert sim.meta["ONEVPASS"][0] == npass
    np.testing.assert_allclose(sim.meta["OMCINT"][0], mcint, rtol=1e-12)
    np.testing.assert_allclose(sim.meta["OMCINTGO"][0], geo, rtol=1e-12)
    np.testing.assert_allclose(sim.meta["OMCINTUN"][0], unc, rtol=1e-9)
    if power:
        assert sim["log_e_nu"].min() >= 7.0 and sim["log_e_nu"].max() <= 10.5


@pytest.mark.parametrize("tag", EAS_LARGE_TAGS)
def test_engine_large_reference_sample(nss, eas_large, tag):
    """CphotAng.__call__ on the GPU (nss_cphotang through the C ABI) against ~31 000 events of
    the unmodified reference: median < 5e-5, 99 % < 3e-4, 99.9 % < 2e-3, every event below
    max(2.5e-3, 5 x the reference's own conditioning); Cherenkov angle 5e-5."""
    ref = eas_large[tag]
    cp = nss.CphotAng(525.0, "Greisen")
    d, c = cp(ref["beta"], ref["alt"], ref["eshow"], np.zeros_like(ref["beta"]),
              np.zeros_like(ref["beta"]), None)
    check_engine_vs_reference(np.asarray(d), np.asarray(c), ref, f"gpu/{tag}")
    # the EAS.__call__ tail on top: trigger mask against the reference-derived numPEs
    npe_ref = ref["dphots"] * 2.5 * 0.2
    npe = np.asarray(d) * 2.5 * 0.2
    near = np.abs(npe_ref / 10.0 - 1) < 2.5e-3
    assert np.array_equal((npe >= 10.0)[~near], (npe_ref >= 10.0)[~near])


# ------------------------------------------------------------------ Gaisser-Hillas profiles


@pytest.mark.parametrize("short", sorted(GH_PROFILES))
@pytest.mark.parametrize("tag", GH_TAGS)
def test_engine_gaisser_hillas_reference_sample(nss, eas_gh, tag, short):
    """nss_cphotang with the two Gaisser-Hillas longitudinal profiles
    (shower_properties.py:122-197) against the unmodified reference; the fluctuated profile
    on the CONEX rows the reference drew (u_profile = (row + 1/2) / nrows)."""
    ref = eas_gh[tag, short]
    cp = nss.CphotAng(525.0, GH_PROFILES[short])
    z = np.zeros_like(ref["beta"])
    d, c = cp(ref["beta"], ref["alt"], ref["eshow"], z, z, None, u_profile=ref["u_profile"])
    check_engine_vs_reference(np.asarray(d), np.asarray(c), ref, f"gpu/{tag}/{short}",
                              median_tol=7e-5 if short == "ghf" else 5e-5)
    assert np.median(rel_err(d, ref["oracle_dphots"])) < 1e-5


@pytest.mark.parametrize("prof", sorted(GH_PROFILES.values()))
def test_gaisser_hillas_acceptance_and_compute(nss, tables_v3, prof):
    """The fused acceptance mode and compute() with a Gaisser-Hillas profile: same events
    from the same Philox words (the CONEX row uniform is word 3 of stream 1), and both equal
    to the oracle fed the identical uniforms."""
    cfg = make_config(n=4000)
    cfg.simulation.eas_long_profile = prof
    seed, n = 5, 4000
    r = nss.AcceptanceRunner(cfg, batch=1536)
    r.enqueue(seed, 0, n)
    got = r.finalize(r.sums())
    sim = nss.compute(cfg, seed=seed)
    assert sim.meta["ONEVPASS"][0] == got[2]
    np.testing.assert_allclose(sim.meta["OMCINT"][0], got[0], rtol=1e-12)
    u4, us, ut, ud, words = philox_uniforms(seed, 0, n)
    u_prof = (words[1][3].astype(np.float64) + 0.5) / 4294967296.0
    p = O.OracleParams(eas_long_profile=prof)
    mc, cols, t, _ = O.run_path(p, tables_v3, u4, us, ut, ud, nthreads=4, u_profile=u_prof,
                                conex=conex_rows())
    rd = rel_err(sim["numPEs"], cols["numPEs"])
    rows = conex_rows()[np.minimum((u_prof[t["event_mask"]] * 1000).astype(int), 999)]
    sens = O.conditioning(cols["beta_rad"], cols["altDec"], cols["showerEnergy"], nthreads=4,
                          profile=prof, gh_rows=rows)
    assert np.median(rd) < 1e-5 and (rd < np.maximum(2.5e-3, 5 * sens)).all(), rd.max()
    near = np.abs(cols["numPEs"] / 10.0 - 1) < 2.5e-3
    assert np.array_equal((sim["numPEs"] >= 10)[~near], (cols["numPEs"] >= 10)[~near])
    if not near.any():
        assert got[2] == mc[2]
        np.testing.assert_allclose(got[0], mc[0], rtol=1e-11)
    # the profile changes the answer (guards against silently running Greisen)
    g = nss.compute(make_config(n=4000), seed=seed)
    assert not np.allclose(g["numPEs"], sim["numPEs"], rtol=1e-3)


def test_gaisser_hillas_fluctuated_needs_uniforms(nss):
    from nuspacesim_b200 import _lib
    import ctypes

    p = _lib.NssParams()
    p.eas_profile = 2
    x = _lib.to_device(np.ones(4))
    ws = _lib.eas_workspace(4)
    rc = _lib.lib().nss_cphotang(ctypes.byref(p), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None,
                                 None, 4, _lib.ptr(x), _lib.ptr(x), None, _lib.ptr(ws), None)
    assert rc == -2  # null u_profile
    p.eas_profile, p.has_cloud = 0, 2
    rc = _lib.lib().nss_cphotang(ctypes.byref(p), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None,
                                 None, 4, _lib.ptr(x), _lib.ptr(x), None, _lib.ptr(ws), None)
    assert rc == -2  # per-event cloud model without cloud tops
    p.has_cloud = 0
    rc = _lib.lib().nss_cphotang(ctypes.byref(p), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x), None,
                                 None, 4, _lib.ptr(x), _lib.ptr(x), None, None, None)
    assert rc == -2  # no workspace
    with pytest.raises(ValueError):
        nss.CphotAng(525.0, "Linsley")


# ------------------------------------------------------------------ PressureMapCloud


@pytest.mark.parametrize("tag", sorted(CLOUDMAP_TAGS))
def test_pressure_map_cloud_engine(nss, eas_cloudmap, tag):
    """nss_cloud_top_from_map equals the reference's cloudf bit for bit; CphotAng with the
    PressureMapCloud CloudTopHeight clouds out the same events as the reference and agrees on
    the others within the stated tolerance."""
    from nuspacesim_b200.config import PressureMapCloud
    from nuspacesim_b200.simulation.atmosphere.clouds import CloudTopHeight

    cfg = make_config()
    cfg.simulation.cloud_model = PressureMapCloud(month=CLOUDMAP_TAGS[tag])
    ct = CloudTopHeight(cfg)
    tops = ct.device_tops(eas_cloudmap["sweep_lat"]).cpu().numpy()
    assert np.array_equal(tops, eas_cloudmap["sweep_" + tag])
    ref = eas_cloudmap[tag]
    cp = nss.CphotAng(525.0, "Greisen")
    d, c = cp(ref["beta"], ref["alt"], ref["eshow"], ref["lat"], ref["lon"], ct)
    check_engine_vs_reference(np.asarray(d), np.asarray(c), ref, f"gpu/{tag}", cang_tol=1.5e-4)
    # a plain Python callable (evaluated on the host, per event) gives the same answer
    d2, _ = cp(ref["beta"][:64], ref["alt"][:64], ref["eshow"][:64], ref["lat"][:64],
               ref["lon"][:64], lambda la, lo: float(ct(la, lo)))
    assert np.array_equal(np.asarray(d2), np.asarray(d)[:64])


def test_pressure_map_cloud_compute_and_acceptance(nss, tables_v3):
    """compute() and the fused acceptance mode under PressureMapCloud: identical events
    (the cloud top comes from the ground-spot latitude of each trajectory,
    compute.py:214-216), both equal to the oracle fed the identical uniforms."""
    from nuspacesim_b200.config import PressureMapCloud

    seed, n = 21, 6000
    cfg = make_config(n=n, loge=9.0)
    cfg.simulation.cloud_model = PressureMapCloud(month=1)
    r = nss.AcceptanceRunner(cfg, batch=2048)
    r.enqueue(seed, 0, n)
    got = r.finalize(r.sums())
    sim = nss.compute(cfg, seed=seed)
    assert sim.meta["ONEVPASS"][0] == got[2]
    np.testing.assert_allclose(sim.meta["OMCINT"][0], got[0], rtol=1e-12)
    u4, us, ut, ud, _ = philox_uniforms(seed, 0, n)
    p = O.OracleParams(log_nu_energy=9.0, cloud_map=cloud_pressure_map(1))
    mc, cols, t, _ = O.run_path(p, tables_v3, u4, us, ut, ud, nthreads=4)
    assert np.array_equal(sim["numPEs"] == 0, cols["numPEs"] == 0)
    live = cols["numPEs"] > 0
    rd = rel_err(sim["numPEs"][live], cols["numPEs"][live])
    assert np.median(rd) < 1e-5 and rd.max() < 2.5e-3, rd.max()
    # the clouds matter for this configuration
    clear = nss.compute(make_config(n=n, loge=9.0), seed=seed)
    assert np.count_nonzero(clear["numPEs"] > 0) > np.count_nonzero(live)
    near = np.abs(cols["numPEs"] / 10.0 - 1) < 2.5e-3
    if not near.any():
        assert got[2] == mc[2]
        np.testing.assert_allclose(got[0], mc[0], rtol=1e-11)


# ------------------------------------------------------------------ sliced compute() pipeline


@pytest.mark.parametrize("mode", ["philox", "uniforms", "uniforms_numpy", "gh_cloudmap_powerlaw"])
def test_compute_sliced_pipeline_equals_unsliced(nss, mode, tmp_path):
    """compute() pipelined over slices of thrown events returns bit-identical columns (valid
    events keep their thrown order) and the same acceptance as the one-shot path; ragged last
    slice, a slice boundary inside the table, pinned and NumPy uniforms."""
    import torch
    from nuspacesim_b200.config import PowerSpectrum, PressureMapCloud

    n = 50_001
    cfg = make_config(n=n)
    kw = {}
    if mode == "uniforms":
        rng = np.random.default_rng(5)
        u4 = torch.from_numpy(rng.random((4, n))).pin_memory()
        kw["uniforms"] = dict(u4=u4, u_tau=torch.from_numpy(rng.random(n)).pin_memory(),
                              u_dec=rng.random(n))  # one NumPy array on purpose
    if mode == "uniforms_numpy":
        rng = np.random.default_rng(6)
        kw["uniforms"] = dict(u4=rng.random((4, n)), u_tau=rng.random(n), u_dec=rng.random(n))
    if mode == "gh_cloudmap_powerlaw":
        cfg.simulation.eas_long_profile = "Gaisser-Hillas Fluctuated"
        cfg.simulation.cloud_model = PressureMapCloud(month=7)
        cfg.simulation.spectrum = PowerSpectrum(index=2.0, lower_bound=7.0, upper_bound=10.0)
    one = nss.compute(cfg, seed=17, slice_events=0, **kw)
    fn = tmp_path / "sliced.fits"
    many = nss.compute(cfg, seed=17, slice_events=8192, output_file=str(fn), **kw)
    assert many.colnames == one.colnames and len(many) == len(one) > 0.9 * n
    for c in one.colnames:
        assert np.array_equal(one[c], many[c]), c
    assert many.meta["ONEVPASS"][0] == one.meta["ONEVPASS"][0]
    for k in ("OMCINT", "OMCINTGO"):
        np.testing.assert_allclose(many.meta[k][0], one.meta[k][0], rtol=1e-13)
    np.testing.assert_allclose(many.meta["OMCINTUN"][0], one.meta["OMCINTUN"][0], rtol=1e-10)
    from nuspacesim_b200 import results_table as R

    assert not fn.exists()  # compute() leaves the final write to its caller (apps/run.py:209-210)
    many.write(str(fn), format="fits", overwrite=True)
    back = R.read_fits_table(str(fn))
    assert np.array_equal(back["numPEs"], one["numPEs"])


# ------------------------------------------------------------------ sharded column mode


def _shard_worker(rank, world, port, n, seed, q):
    import os
    import sys

    import torch
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)  # both ranks share the one GPU of the test box
    import nuspacesim_b200 as nss
    from conftest import make_config

    sim = nss.compute(make_config(n=n), seed=seed, shard=True, slice_events=4096)
    q.put((rank, {c: np.asarray(sim[c]).copy() for c in sim.colnames},
           {k: sim.meta[k][0] for k in ("OMCINT", "OMCINTGO", "ONEVPASS", "OMCINTUN")}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_compute_sharded_over_ranks(nss):
    """compute(shard=True) with two ranks (gloo; both on this box's GPU): each returns its own
    block of rows, the blocks concatenate to the one-process table bit for bit, and every rank
    holds the global acceptance."""
    import socket

    import torch.multiprocessing as mp

    n, seed, world = 30_001, 23, 2
    one = nss.compute(make_config(n=n), seed=seed, slice_events=0)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, n, seed, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(len(r[1]["numPEs"]) > 0.4 * len(one) for r in res)
    for c in one.colnames:
        assert np.array_equal(np.concatenate([r[1][c] for r in res]), one[c]), c
    for r in res:
        assert r[2]["ONEVPASS"] == one.meta["ONEVPASS"][0]
        np.testing.assert_allclose(r[2]["OMCINT"], one.meta["OMCINT"][0], rtol=1e-13)
        np.testing.assert_allclose(r[2]["OMCINTGO"], one.meta["OMCINTGO"][0], rtol=1e-13)
        np.testing.assert_allclose(r[2]["OMCINTUN"], one.meta["OMCINTUN"][0], rtol=1e-10)


def test_eas_entry_points_are_reentrant(nss, eas_large):
    """Two nss_cphotang launches in flight at once on different streams, each with its own
    caller-supplied workspace, return bit for bit what the same launches return one after the
    other: the entry points keep no per-call state of their own (include/nss_b200.h)."""
    import ctypes

    import torch

    from nuspacesim_b200 import _lib

    L = _lib.lib()
    p = nss.CphotAng(525.0, "Greisen").params
    sets = []
    for tag in ("c1", "e10"):
        ref = eas_large[tag]
        sets.append([_lib.to_device(ref[k]) for k in ("beta", "alt", "eshow")])

    def launch(x, stream, ws):
        n = x[0].numel()
        d, c = _lib.empty(n), _lib.empty(n)
        rc = L.nss_cphotang(ctypes.byref(p), _lib.ptr(x[0]), _lib.ptr(x[1]), _lib.ptr(x[2]), None,
                            None, n, _lib.ptr(d), _lib.ptr(c), None, _lib.ptr(ws),
                            ctypes.c_void_p(stream.cuda_stream))
        assert rc == 0
        return d, c

    ws = [_lib.eas_workspace(x[0].numel()) for x in sets]
    main = torch.cuda.current_stream()
    serial = [launch(x, main, w) for x, w in zip(sets, ws)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(3):  # interleave the two event sets on two streams, several times over
        conc = [launch(x, s, w) for x, s, w in zip(sets, streams, ws)]
        torch.cuda.synchronize()
        for (d0, c0), (d1, c1) in zip(serial, conc):
            assert torch.equal(d0, d1) and torch.equal(c0, c1)


@pytest.mark.parametrize("profile", [0, 1, 2])
def test_walk_with_and_without_work_counters_is_identical(nss, eas_large, profile):
    """The walk kernel has two instantiations per profile: without the algorithmic-work
    counters (72 registers, what compute() and the acceptance runner launch) and with them
    (stats != NULL: bench work counters, ncu captures).  Same source; ptxas fuses multiplies
    and adds per instantiation, so a last-bit difference can move a tail decision (one chunk
    evaluated instead of continued: < 1e-6 of an event).  Almost every event must agree bit
    for bit, all of them to 5e-6, and the counters must be filled."""
    import ctypes

    from nuspacesim_b200 import _lib

    L = _lib.lib()
    p = nss.CphotAng(525.0, ["Greisen", "Gaisser-Hillas Parameterized",
                             "Gaisser-Hillas Fluctuated"][profile]).params
    ref = eas_large["c1"]
    x = [_lib.to_device(ref[k][:6000]) for k in ("beta", "alt", "eshow")]
    n = x[0].numel()
    u = _lib.to_device(np.random.default_rng(3).random(n)) if profile == 2 else None
    ws = _lib.eas_workspace(n)

    def run(stats):
        d, c = _lib.empty(n), _lib.empty(n)
        rc = L.nss_cphotang(ctypes.byref(p), _lib.ptr(x[0]), _lib.ptr(x[1]), _lib.ptr(x[2]),
                            _lib.ptr(u), None, n, _lib.ptr(d), _lib.ptr(c), _lib.ptr(stats),
                            _lib.ptr(ws), _lib.stream_ptr())
        assert rc == 0
        return d.cpu().numpy(), c.cpu().numpy()

    st = _lib.zeros(8, "int64")
    d0, c0 = run(None)
    d1, c1 = run(st)
    assert np.mean(d0 == d1) > 0.9 and np.mean(c0 == c1) > 0.9, (np.mean(d0 == d1), np.mean(c0 == c1))
    np.testing.assert_allclose(d0, d1, rtol=5e-6)
    np.testing.assert_allclose(c0, c1, rtol=5e-6)
    k = st.cpu().numpy()
    assert k[0] == n and k[2] > 100 * n and k[3] > k[2]  # events, retained steps, radial bins
    assert (d0 > 0).mean() > 0.9
