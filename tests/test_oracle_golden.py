"""The oracle (oracle/nss_oracle.py + oracle/cphot_oracle.c) against outputs of the
UNMODIFIED reference (tests/golden/*.npz, made by tests/golden/make_golden.py) and against
the reference's own known-answer vectors.  CPU only."""
import numpy as np
import pytest

from conftest import (CLOUDMAP_TAGS, EAS_LARGE_TAGS, GEOM_TAGS, GH_PROFILES, GH_TAGS, PIPE_TAGS,
                      TAU_TABLES, check_engine_vs_reference, cloud_pressure_map, conex_rows,
                      load_tables, rel_err)
from oracle import nss_oracle as O

# Known-answer vectors copied from the reference's test data
# (test/simulation/eas_optical/cphot_ang.py:16-66: beta = 42 deg, alt = 0..20 km,
#  Eshow = 100 PeV, 525 km orbit, no cloud).  The reference's current code reproduces them
#  to <= 3.0e-4, hence the tolerance.
OLD_DPHOTS = np.array([
    1.6210271e03, 1.8776989e03, 2.1580850e03, 2.4297334e03, 2.6037090e03, 2.5893950e03,
    2.6665376e03, 1.9930570e03, 1.2905229e03, 7.2669135e02, 3.5919400e02, 1.5781609e02,
    6.3526283e01, 2.4607666e01, 9.4170065e00, 3.5231948e00, 1.3178004e00, 5.0189447e-01,
    1.9809218e-01, 8.2580306e-02, 3.6842916e-02])
OLD_CANG = np.array([
    1.1297, 1.0476834, 0.9589647, 0.86524945, 0.7722871, 0.68627226, 0.6104225, 0.54487425,
    0.48861432, 0.4402858, 0.39882657, 0.3633839, 0.33331406, 0.30719605, 0.28355855,
    0.2634559, 0.24582747, 0.22999273, 0.21532202, 0.20063952, 0.18477382])


def _params(tagcfg):
    alt, mch, maz, lim = tagcfg
    return O.OracleParams(detector_altitude=alt, max_cherenkov_angle=mch, max_azimuth_angle=maz,
                          angle_from_limb=lim)


@pytest.mark.parametrize("tag", GEOM_TAGS)
def test_geometry_bit_exact(golden, tag):
    geo = golden["geometry"]
    g = O.geom_constants(_params(geo[tag + "_cfg"]))
    c = geo[tag + "_consts"]
    mine = [g["earth_radius"], g["core_alt"], g["minLOSpathLen"], g["maxLOSpathLen"],
            g["sinOfMaxThetaTrSubV"], g["maxPhiS"], g["minPhiS"], g["mcnorm"]]
    assert np.array_equal(np.array(mine), c)
    with np.errstate(all="ignore"):
        t = O.geom_throw(g, geo[tag + "_u"])
    for k, v in t.items():
        assert np.array_equal(v, geo[f"{tag}_{k}"], equal_nan=True), k
    for d, names in ((0.0, ("init_lat", "init_lon")), (7.5, ("lat7", "lon7"))):
        la, lo = O.geom_lat_long(g, t, d)
        assert np.array_equal(la, geo[f"{tag}_{names[0]}"])
        assert np.array_equal(lo, geo[f"{tag}_{names[1]}"])
    trig, cth, pex = geo[tag + "_mc_in"]
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(O.mcintegral(g, t, trig, cth, pex, 10.0),
                                   geo[tag + "_mc_out"], rtol=1e-13)
        np.testing.assert_allclose(O.mcintegral(g, t, trig, cth, pex, 10.0, 1e-6, 1e6),
                                   geo[tag + "_mc_out_pl"], rtol=1e-13)


def test_geometry_shape_errors():
    g = O.geom_constants(O.OracleParams())
    with pytest.raises(RuntimeError):
        O.geom_throw(g, np.zeros((3, 5)))


@pytest.mark.parametrize("tag", list(TAU_TABLES))
def test_taus(golden, tag):
    tg = golden["taus"]
    T = load_tables(*TAU_TABLES[tag])
    tb, lor, te, sh, pex = O.taus(O.OracleParams(), T, tg[tag + "_beta"], tg[tag + "_loge"],
                                  tg[tag + "_u"])
    assert rel_err(pex, tg[tag + "_pexit"]).max() < 1e-13
    assert rel_err(te, tg[tag + "_tauEnergy"]).max() < 1e-13
    assert rel_err(sh, tg[tag + "_showerEnergy"]).max() < 1e-13
    assert rel_err(lor, tg[tag + "_tauLorentz"]).max() < 1e-13
    assert rel_err(tb, tg[tag + "_tauBeta"]).max() < 1e-13


def test_taus_errors(tables_v3):
    T = tables_v3
    with pytest.raises(ValueError):
        O.tau_exit_prob(T["pexit"], T["pexit_log_e_nu"], T["pexit_beta_rad"], np.array([0.1]),
                        np.array([12.5]))
    with pytest.raises(ValueError):  # u == cdf[0]: no crossing
        O.tau_energy(T["cdf"], T["cdf_log_e_nu"], T["cdf_beta_rad"], T["cdf_e_tau_frac"],
                     np.array([0.1]), np.array([8.0]), np.array([0.0]))
    assert O.cdf_sample(T["cdf"], T["cdf_log_e_nu"], T["cdf_beta_rad"], T["cdf_e_tau_frac"],
                        np.array([]), np.array([]), np.array([])).size == 0


def test_survey_known_answers(tables_v3):
    """Known answers recorded from the reference in SURVEY.md 8(c)."""
    T = tables_v3
    p = O.OracleParams()
    g = O.geom_constants(p)
    u = np.array([[.1, .5, .9, .25], [.2, .4, .6, .8], [.3, .3, .3, .3], [.05, .35, .65, .95]])
    t = O.geom_throw(g, u)
    np.testing.assert_allclose(t["losPathLen"], [2354.714068140746, 1858.414446554829,
                                                 1550.94979827645, 1298.220903303413], rtol=1e-12)
    np.testing.assert_allclose(t["betaTrSubN"], [2.431471593603, 10.250049946441,
                                                 15.643323527213, 18.149591448875], rtol=1e-11)
    pex = O.tau_exit_prob(T["pexit"], T["pexit_log_e_nu"], T["pexit_beta_rad"],
                          np.radians([0.05, 0.5, 1, 5.5, 20, 41.9, 43]),
                          np.array([8, 8.1, 6, 9.33, 12, 10, 8.0]))
    np.testing.assert_allclose(pex, [9.295e-04, 1.186998561145e-03, 1.3e-06, 1.126520859035e-03,
                                     2.249e-05, 1.905916404001e-06, 1.192091745974e-07],
                               rtol=2e-10)
    z = O.cdf_sample(T["cdf"], T["cdf_log_e_nu"], T["cdf_beta_rad"], T["cdf_e_tau_frac"],
                     np.array([8.1, 6, 9.33, 12, 10]), np.radians([0.5, 1, 5.5, 20, 41.9]),
                     np.array([.1, .5, .9, .999, .3]))
    np.testing.assert_allclose(z, [4.129418955339e-01, 8.369675031351e-01, 6.194613349434e-01,
                                   3.000980635461e-04, 1.723158132956e-04], rtol=1e-11)
    alt, ln = O.alt_dec(np.radians([1, 10, 20.0]), np.array([0.999999, 1 - 1e-10, 1 - 1e-14]),
                        np.array([1e3, 1e6, 1e8]), 0.5)
    np.testing.assert_allclose(ln, [6.032436589927e-05, 6.032442621766e-02, 6.032442622370],
                               rtol=1e-11)
    np.testing.assert_allclose(alt[1:], [1.047550335534e-02, 2.065735124765], rtol=1e-9)


def test_altdec_bit_exact(golden):
    a = golden["altdec"]
    alt, ln = O.alt_dec(a["beta"], a["tauBeta"], a["tauLorentz"], a["u"])
    assert np.array_equal(alt, a["altDec"]) and np.array_equal(ln, a["lenDec"])


def test_engine_reference_known_answers():
    """cphot_oracle.c against the reference's own golden vectors (cphot_ang.py:16-66)."""
    d, c, _ = O.cphotang(np.full(21, np.radians(42.0)), np.arange(21.0), np.ones(21))
    assert rel_err(d, OLD_DPHOTS).max() < 4e-4
    assert rel_err(c, OLD_CANG).max() < 3e-6


def test_engine_sweep(golden):
    """cphot_oracle.c against CphotAng.run of today's reference on a parameter sweep."""
    e = golden["eas"]
    d, c, _ = O.cphotang(np.full(21, np.radians(42.0)), np.arange(21.0), np.ones(21))
    assert rel_err(d, e["series42"][:, 0]).max() < 4e-4
    assert rel_err(c, e["series42"][:, 1]).max() < 3e-6
    sw = e["sweep"]
    for det in (525.0, 33.0):
        for top in (-np.inf, 2.0, 6.0, 12.0):
            m = (sw[:, 0] == det) & (sw[:, 4] == top)
            if not m.any():
                continue
            d, c, st = O.cphotang(sw[m, 1], sw[m, 2], sw[m, 3], det,
                                  None if top == -np.inf else top)
            assert np.array_equal(d == 0, sw[m, 5] == 0)  # cloud / empty early-outs
            rd = rel_err(d, sw[m, 5])
            assert rd.max() < 1e-3 and np.median(rd) < 1e-4
            assert rel_err(c, sw[m, 6]).max() < 2e-6


@pytest.mark.parametrize("tag", EAS_LARGE_TAGS)
def test_engine_large_reference_sample(eas_large, tag):
    """cphot_oracle.c against ~31 000 events of the unmodified reference
    (BASELINE.json configs 1-4 inputs)."""
    ref = eas_large[tag]
    check_engine_vs_reference(ref["oracle_dphots"], ref["oracle_cang"], ref, f"oracle/{tag}")


@pytest.mark.parametrize("short", sorted(GH_PROFILES))
@pytest.mark.parametrize("tag", GH_TAGS)
def test_engine_gaisser_hillas_reference_sample(eas_gh, tag, short):
    """cphot_oracle.c with the Gaisser-Hillas profiles (shower_properties.py:122-197) against
    the unmodified reference, the fluctuated one on the recorded CONEX rows."""
    ref = eas_gh[tag, short]
    check_engine_vs_reference(ref["oracle_dphots"], ref["oracle_cang"], ref, f"oracle/{tag}/{short}",
                              median_tol=7e-5 if short == "ghf" else 5e-5)


@pytest.mark.parametrize("tag", ("c1", "e10", "pl10"))
def test_reference_float32_accumulation_is_the_residual(tag):
    """Where the 3e-5 between the reference and any exact implementation comes from:
    np.einsum("zje,zw->", ..., dtype=float32) (cphotang.py:521) adds its ~1e5 float32 terms one
    after the other into a float32.  With that accumulation replayed (accum="reference") the
    oracle reproduces the unmodified reference to a median of 4e-7 (libm-level); summed
    exactly (accum="exact", what the CUDA kernel is checked against) the same terms give
    3.0e-5 more.  Same for Gaisser-Hillas Fluctuated, whose longer profiles lose 5e-5."""
    import os

    from conftest import GOLDEN

    nt = os.cpu_count() or 1
    G = np.load(os.path.join(GOLDEN, "eas_large.npz"))
    b, a, e = (G[f"{tag}_{k}"][:3000] for k in ("beta_rad", "altDec", "showerEnergy"))
    ref = G[f"{tag}_dphots"][:3000]
    d_ref, _, _ = O.cphotang(b, a, e, 525.0, None, nthreads=nt, accum="reference")
    d_ex, _, _ = O.cphotang(b, a, e, 525.0, None, nthreads=nt, accum="exact")
    r = rel_err(d_ref, ref)
    assert np.median(r) < 1e-6 and np.percentile(r, 90) < 1e-5, (np.median(r), np.percentile(r, 90))
    deficit = (d_ex - d_ref) / d_ex
    assert 2.0e-5 < np.median(deficit) < 4.5e-5 and np.percentile(deficit, 2) > 0, np.median(deficit)
    assert abs(np.median((d_ex - ref) / ref) - np.median(deficit)) < 3e-6
    if tag == "c1":
        H = np.load(os.path.join(GOLDEN, "eas_gh.npz"))
        b, a, e = (H[f"c1_{k}"] for k in ("beta_rad", "altDec", "showerEnergy"))
        rows = conex_rows()[H["c1_row"]]
        d, _, _ = O.cphotang(b, a, e, 525.0, None, nthreads=nt, accum="reference",
                             profile="Gaisser-Hillas Fluctuated", gh_rows=rows)
        assert np.median(rel_err(d, H["c1_ghf_dphots"])) < 1e-6


def test_einsum_float32_order_replayed_bit_for_bit():
    """The accumulation order assumed by accum="reference", checked against NumPy itself:
    for float32 operands einsum("zje,zw->") equals the sequential float32 running sum over
    (z, j, e) in C order of svtrm[z,j,e] * sum_w sigval[z,w]."""
    rng = np.random.default_rng(2)
    sv = (rng.random((150, 40, 8)) ** 6 * 1e3).astype(np.float32)
    sv[:, 25:, :] *= (rng.random((150, 15, 8)) < 0.5)  # masked-out bins are zeros
    sg = (rng.random((150, 28)) ** 3).astype(np.float32)
    want = np.einsum("zje,zw->", sv, sg, dtype=np.float32)
    sgs = sg.sum(axis=1, dtype=np.float32)
    got = np.cumsum((sv * sgs[:, None, None]).ravel(), dtype=np.float32)[-1]
    exact = np.einsum("zje,zw->", sv.astype(np.float64), sg.astype(np.float64))
    assert got == want
    assert abs(float(want) - exact) / exact > 1e-6  # and it is measurably not the exact sum


def test_conex_table_is_the_reference_table():
    """The shipped .npy holds columns 4..9 of the reference's CONEX file (cphotang.py:266-276)."""
    import os

    T = conex_rows()
    assert T.shape == (1000, 6) and T.dtype == np.float64 and np.isfinite(T).all()
    src = ("/root/reference/src/nuspacesim/data/CONEX_table/"
           "dumpGH_conex_pi_E17_95deg_0km_eposlhc_1394249052_211.dat")
    if os.path.exists(src):
        assert np.array_equal(T, np.loadtxt(src, usecols=(4, 5, 6, 7, 8, 9)))


@pytest.mark.parametrize("tag", sorted(CLOUDMAP_TAGS))
def test_pressure_map_cloud(eas_cloudmap, tag):
    """PressureMapCloud (clouds.py:85-107): the lookup + standard-atmosphere conversion is
    bit-identical to the reference's cloudf on a latitude sweep and on the event sample; the
    engine under per-event cloud tops agrees with the reference within the stated tolerance
    and clouds out exactly the same events."""
    month = CLOUDMAP_TAGS[tag]
    m = cloud_pressure_map(month)
    assert np.array_equal(O.cloud_top_pressure_map(m, eas_cloudmap["sweep_lat"]),
                          eas_cloudmap["sweep_" + tag])
    ref = eas_cloudmap[tag]
    assert np.array_equal(ref["oracle_tops"], ref["cloud_top"])
    assert 0 < np.count_nonzero(ref["dphots"] == 0) < len(ref["dphots"])
    check_engine_vs_reference(ref["oracle_dphots"], ref["oracle_cang"], ref, f"oracle/{tag}")


def test_cloud_maps_are_the_reference_maps():
    import os

    for month in (1, 6, 12):
        m = cloud_pressure_map(month)
        assert m.shape == (361, 576) and m.dtype == np.float32
        src = f"/root/reference/src/nuspacesim/data/cloud_maps/nss_map_CloudTopPressure_{month:02d}.v0.fits"
        if os.path.exists(src):
            raw = open(src, "rb").read()
            want = np.frombuffer(raw[2880:2880 + 361 * 576 * 4], dtype=">f4").reshape(361, 576)
            assert np.array_equal(m, want)


@pytest.mark.parametrize("tag", PIPE_TAGS)
def test_pipeline(golden, tables_v3, tag):
    """compute()'s optical branch stage by stage on shared uniforms."""
    pp = golden["pipeline"]
    alt, loge, n = pp[tag + "_cfg"]
    p = O.OracleParams(detector_altitude=alt, log_nu_energy=loge)
    mc, cols, t, st = O.run_path(p, tables_v3, pp[tag + "_u4"], None, pp[tag + "_u_tau"],
                                 pp[tag + "_u_dec"], nthreads=4)
    assert np.array_equal(t["event_mask"], pp[tag + "_event_mask"])
    for k in ("beta_rad", "tauBeta", "tauLorentz", "tauEnergy", "showerEnergy", "tauExitProb",
              "altDec", "lenDec"):
        assert rel_err(cols[k], pp[f"{tag}_{k}"]).max() < 1e-13, k
    rd = rel_err(cols["numPEs"], pp[tag + "_numPEs"])
    assert rd.max() < 2.5e-3 and np.median(rd) < 1e-4
    assert rel_err(cols["costhetaChEff"], pp[tag + "_costhetaChEff"]).max() < 1e-6
    # trigger mask and pass count are exact; integrals agree to the float32 tolerance
    assert np.array_equal(cols["numPEs"] >= 10, pp[tag + "_numPEs"] >= 10)
    ref = pp[tag + "_mc"]
    assert mc[2] == ref[2]
    np.testing.assert_allclose(mc[1], ref[1], rtol=1e-13)
    if ref[0] > 0:
        np.testing.assert_allclose(mc[0], ref[0], rtol=1e-13)
        np.testing.assert_allclose(mc[3], ref[3], rtol=1e-12)


def test_oracle_against_live_reference():
    """When /root/reference is mounted (build container), run the reference itself."""
    from oracle import ref_harness as H

    if not H.reference_available():
        pytest.skip("reference tree not mounted")
    import warnings

    warnings.filterwarnings("ignore", category=DeprecationWarning)
    ref = H.load_reference()
    cp = ref.cphotang.CphotAng(525.0, "Greisen")
    rng = np.random.default_rng(3)
    b, a, e = np.radians(rng.uniform(0.5, 40, 12)), rng.uniform(0, 12, 12), 10 ** rng.uniform(-2, 1, 12)
    want = np.array([cp.run(b[i], a[i], e[i], 0.0, 0.0, None) for i in range(12)], float)
    d, c, _ = O.cphotang(b, a, e)
    assert rel_err(d, want[:, 0]).max() < 1e-3
    assert rel_err(c, want[:, 1]).max() < 2e-6


@pytest.mark.parametrize("tag", ["nls0", "np3"])
def test_tau_energy_raw_uniforms_vs_reference(tag):
    """Taus.tau_energy WITHOUT u, as compute() calls it: the reference draws
    np.random.uniform(cdfs[:, 0], 1.0) (utils/cdf.py:419-423).  The fixture
    (tests/golden/make_golden_taus_raw.py) holds the legacy generator's raw outputs and the
    reference's results; the nuleptonsim CDF starts at up to 0.0134, so using the raw draws as
    they are would leave the table."""
    import os

    from conftest import GOLDEN

    G = np.load(os.path.join(GOLDEN, "taus_raw.npz"))
    T = load_tables(*TAU_TABLES[tag])
    args = (T["cdf"], T["cdf_log_e_nu"], T["cdf_beta_rad"], T["cdf_e_tau_frac"],
            G[tag + "_beta"], G[tag + "_loge"], G[tag + "_u_raw"])
    te = O.tau_energy(*args, u_raw=True)
    np.testing.assert_allclose(te, G[tag + "_tauEnergy"], rtol=1e-13)
    if tag == "nls0":
        with pytest.raises(ValueError):
            O.tau_energy(*args)
    else:  # first CDF column identically 0: the mapping is the identity
        assert np.array_equal(O.tau_energy(*args), te)
