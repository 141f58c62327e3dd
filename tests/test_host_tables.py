"""CPU: the host side of the product -- loaders, table builder, parameter handling -- checked
against golden outputs of the unmodified reference, with the packed tables evaluated by the
numpy walk-through in oracle/table_emul.py (the same algebra the kernels run).

No kernel is launched here; what is tested is that the tables the product uploads, evaluated
with the kernels' formulas in float64, reproduce the reference within the parity tolerances.
"""
import copy

import numpy as np
import pytest

from oracle import table_emul as E

RTOL, ATOL, C2_ATOL = 1e-9, 1e-13, 1e-6


@pytest.fixture(scope="module")
def fit(boss_blocks):
    from victor_b200 import CCFFit
    model, data = boss_blocks
    return CCFFit(copy.deepcopy(model), copy.deepcopy(data))


@pytest.fixture(scope="module")
def packed(fit):
    from victor_b200 import tables as T
    mt = T.build_model_tables(fit, fit.model)
    like = fit.fit_options["likelihood"]
    ft = T.build_fit_tables(fit, like)
    mu, W = T.mu_projection_weights(fit.poles_s)
    return mt, ft, mu, W


def test_loaders_match_reference_state(fit, golden):
    t = golden("boss_tables")
    assert abs(fit.iaH - float(t["iaH"])) < 1e-16
    for key in ("r", "s", "beta", "sv_rmu", "r_for_sv", "mu_for_sv"):
        np.testing.assert_allclose(getattr(fit, key), t[key], rtol=1e-14, atol=0)
    r31 = np.append([0.01], fit.r)
    np.testing.assert_allclose(fit.delta(r31), t["delta_r31"], rtol=1e-13)
    np.testing.assert_allclose(fit.integrated_delta(r31), t["Delta_r31"], rtol=1e-13)
    np.testing.assert_allclose(fit.get_interpolated_real_multipoles(0.37), t["xi_r_beta037"], rtol=1e-14)
    np.testing.assert_allclose(fit.icov[0], t["icov_first"], rtol=1e-9, atol=1e-6)
    a = golden("boss_notebook_anchors")
    np.testing.assert_allclose(fit.multipole_datavector(0.37), a["data_vector"], rtol=1e-14)
    assert abs(np.linalg.slogdet(fit.get_interpolated_covariance(0.37))[1] - float(a["slogdet_cov"])) < 1e-9


def test_covariance_bracket_uses_last_index(fit):
    """The reference's bracket: lower neighbour and the LAST grid index (ccf_fit.py:225-227)."""
    g = fit.beta_covmat
    lo, hi, t = fit._bracket(0.37)
    assert (lo, hi) == (12, 30) and abs(t - 0.045370370370370484) < 1e-15
    assert fit._bracket(g[0] - 0.01) == (0, 0, 0.0)
    assert fit._bracket(g[-1] + 0.01) == (30, 30, 0.0)
    assert fit._bracket(g[5]) == (5, 5, 0.0)
    cov = fit.get_interpolated_covariance(0.37)
    np.testing.assert_allclose(cov, (1 - t) * fit.covmat[12] + t * fit.covmat[30], rtol=1e-15)


def test_weight_vectors(packed):
    from victor_b200 import tables as T
    mt, ft, mu, W = packed
    assert mu.shape == (100,) and W.shape == (2, 100)
    assert abs(W[0].sum() - 1.0) < 1e-13          # monopole of a constant is the constant
    assert abs(W[1].sum()) < 1e-4                 # quadrature of L_2: small, not exactly zero
    x, w = T.velocity_nodes(50)
    assert abs(w.sum() - 12.0) < 1e-12            # integral of 1 over [-6, 6]
    assert abs((w * x).sum()) < 0.02 and w[-1] != w[0]   # scipy's even-N end correction is one-sided
    # the rule integrates a unit Gaussian over +-6 sigma to ~1e-7 (what the 50 nodes can resolve)
    assert abs((w * np.exp(-0.5 * x * x)).sum() / np.sqrt(2 * np.pi) - 1) < 1e-6
    _, W3 = T.mu_projection_weights([0, 2, 4])
    f = 0.3 + 0.5 * (3 * mu ** 2 - 1) / 2 - 0.2 * (35 * mu ** 4 - 30 * mu ** 2 + 3) / 8
    np.testing.assert_allclose(W3 @ f, [0.3, 0.5, -0.2], atol=2e-4)
    mu_odd, W_odd = T.mu_projection_weights([0, 1])
    assert mu_odd[0] == -1.0 and abs(W_odd[0].sum() - 1.0) < 1e-12


def test_cells_reproduce_scipy_splines(fit, packed):
    """Cell cubics + bucket search == FITPACK ext=3 splines, including the clamped ranges."""
    from scipy.interpolate import InterpolatedUnivariateSpline, PchipInterpolator
    mt = packed[0]
    rng = np.random.default_rng(5)
    u = np.concatenate([rng.uniform(0, 170, 20000), mt.knots, mt.knots - 1e-12, mt.knots + 1e-12,
                        [0.0, 1e-300, 0.005, 0.01, 1.999999, 2.0, 118.0, 147.0, 1e4]])
    cell, t = E._cells(mt, u)
    assert np.all(cell >= 1) and np.all(cell <= mt.ncell - 1) and np.all(t >= 0)
    inside = (u >= mt.knots[0])
    assert np.all(mt.origin[cell][inside] <= u[inside]) and np.all(u < mt.upper[cell])
    sv = InterpolatedUnivariateSpline(fit.r_for_sv, fit.sv_rmu[0], ext=3)
    np.testing.assert_allclose(E._horner(mt.sv, cell, t), sv(u), rtol=1e-13, atol=1e-15)
    r31 = np.append([0.01], fit.r)
    v0 = InterpolatedUnivariateSpline(r31, r31 * fit.integrated_delta(r31), ext=3)
    np.testing.assert_allclose(E._horner(mt.v0, cell, t), v0(u), rtol=1e-12, atol=1e-13)
    for beta in (0.37, 0.12, 0.70, float(fit.beta[7])):
        xi0 = PchipInterpolator(fit.beta, fit.real_multipoles["0"], axis=0)(beta)
        want = InterpolatedUnivariateSpline(fit.r, xi0, ext=3)(u)
        xc = E.xi_cells(mt, np.array([beta]))[0, 0]
        np.testing.assert_allclose(E._horner(xc, cell, t), want, rtol=1e-11, atol=1e-14)


def test_bucket_map_nonuniform_knots():
    from victor_b200 import tables as T
    rng = np.random.default_rng(11)
    knots = np.unique(np.concatenate([[0.01], np.cumsum(rng.uniform(0.11, 0.13, 25)), [0.5, 1.7]]))
    inv_h, entry, maxscan = T.bucket_map(knots)
    assert maxscan >= 1 and np.any(entry < 0)

    class M:
        pass
    m = M()
    m.inv_h, m.bucket_base, m.maxscan = inv_h, entry, maxscan
    m.upper = np.concatenate([knots, [np.inf]])
    m.origin = np.concatenate([[knots[0]], knots])
    u = np.concatenate([rng.uniform(0, knots[-1] * 1.3, 50000), knots, np.nextafter(knots, 0), [0.0]])
    cell, _ = E._cells(m, u)
    want = np.maximum(np.searchsorted(knots, u, side="right"), 1)
    assert np.array_equal(cell, want)


def test_tables_reproduce_reference_streaming(packed, fit, golden):
    """All 80 golden rows (64 seeded + 16 edge rows): multipoles, chi2 and lnL from the packed
    tables, against the unmodified reference."""
    from victor_b200.model import params_to_rows
    mt, ft, mu, W = packed
    g = golden("boss_streaming_points")
    rows = params_to_rows(g["params"])
    mult, _ = E.theory_multipoles(mt, rows, np.asarray(fit.s, float), mu, W)
    theory = mult.reshape(len(rows), -1)
    np.testing.assert_allclose(theory, g["theory"], rtol=RTOL, atol=ATOL)
    for a in (0, 30):
        scale = np.abs(g["theory"][:, a:a + 30]).max(axis=1)
        err = np.abs(theory[:, a:a + 30] - g["theory"][:, a:a + 30]).max(axis=1)
        assert np.all(err <= 1e-11 * scale)         # two orders inside the 1e-9 contract
    chi2, lnl = E.chi2_lnl(ft, rows[:, 1], theory)
    np.testing.assert_allclose(chi2, g["chi2"], rtol=0, atol=C2_ATOL)
    np.testing.assert_allclose(lnl, g["lnl"], rtol=0, atol=C2_ATOL)
    assert np.abs(chi2 - g["chi2"]).max() < 1e-8    # measured headroom: ~1e-11


def test_likelihood_forms_from_tables(packed, fit, golden):
    from victor_b200 import tables as T
    from victor_b200.model import params_to_rows
    mt, _, mu, W = packed
    g = golden("boss_forms")
    rows = params_to_rows(g["params"])
    mult, _ = E.theory_multipoles(mt, rows, np.asarray(fit.s, float), mu, W)
    theory = mult.reshape(len(rows), -1)
    for form in ("gaussian", "hartlap", "percival", "sellentin"):
        ft = T.build_fit_tables(fit, {"form": form, "nmocks": 1000, "nparams": 4})
        chi2, lnl = E.chi2_lnl(ft, rows[:, 1], theory)
        np.testing.assert_allclose(chi2, g[f"{form}_chi2"], rtol=0, atol=C2_ATOL)
        np.testing.assert_allclose(lnl, g[f"{form}_lnl"], rtol=0, atol=C2_ATOL)
    with pytest.raises(Exception):
        T.likelihood_constants({"form": "nonsense"}, 60)


def test_example_config_tables(example_block, golden):
    """Non-uniform knots, fixed real-space input, caller-supplied s grid, poles 0/2/4."""
    from victor_b200 import CCFModel, tables as T
    from victor_b200.model import params_to_rows
    g = golden("example_points")
    m = CCFModel(copy.deepcopy(example_block))
    assert abs(m.iaH - float(g["iaH"])) < 1e-16
    np.testing.assert_allclose(m.sv_rmu, g["sv_rmu"], rtol=1e-13)
    mt = T.build_model_tables(m, m.model)
    mu, W = T.mu_projection_weights([0, 2, 4])
    P = {"fsigma8": g["params"][:, 0], "sigma_v": g["params"][:, 1], "epsilon": g["params"][:, 2]}
    mult, _ = E.theory_multipoles(mt, params_to_rows(P), g["s"], mu, W)
    np.testing.assert_allclose(mult.reshape(3, -1), g["streaming_theory"], rtol=RTOL, atol=ATOL)


def test_params_to_rows_conventions():
    from victor_b200 import InputError
    from victor_b200.model import params_to_rows
    rows = params_to_rows({"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 0.95, "alpha": 1.01})
    apar = 1.01 * 0.95 ** (-2 / 3)
    assert rows.shape == (1, 10) and rows[0, 4] == apar and rows[0, 3] == 0.95 * apar   # ccf_model.py:589-592
    rows = params_to_rows({"fsigma8": [0.1, 0.2, 0.3], "beta": 0.4, "sigma_v": 300.0, "aperp": 1.0, "apar": 1.0,
                           "b": 1.9, "Av": 0, "chi2_unused": 7})                       # cobaya passes extras
    assert rows.shape == (3, 10) and np.all(rows[:, 1] == 0.4) and np.all(rows[:, 5] == 1.0)
    rows = params_to_rows(np.array([[0.47, 0.37, 380.0]]))
    assert rows.shape == (1, 10) and rows[0, 3] == 1.0 and rows[0, 4] == 1.0
    with pytest.raises(InputError):
        params_to_rows(np.zeros((2, 2)))


def test_input_errors(boss_blocks):
    from victor_b200 import CCFFit, CCFModel, InputError
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    bad = copy.deepcopy(model)
    bad["input_model_data_file"] = "nope.hdf5"
    with pytest.raises(InputError):
        CCFModel(bad)
    bad = copy.deepcopy(model)
    bad["realspace_ccf"]["ccf_keys"] = ["r", "missing"]
    with pytest.raises(InputError):
        CCFModel(bad)
    bad = copy.deepcopy(model)
    bad["matter_ccf"]["template_sigma8"] = None
    with pytest.raises(InputError):
        CCFModel(bad)
    bad = copy.deepcopy(model)
    bad["velocity_pdf"]["dispersion"] = {"model": "constant"}
    with pytest.raises(InputError):
        CCFModel(bad)
    badd = copy.deepcopy(data)
    badd["covariance_matrix"]["cov_key"] = "missing"
    with pytest.raises(InputError):
        CCFFit(copy.deepcopy(model), badd)
    badd = copy.deepcopy(data)
    badd["covariance_matrix"] = {"data_file": "data/boss_dr12_cmass/cmass_fixed_D_covariance.npz",
                                 "cov_key": "covmat", "fixed_beta": False, "beta_key": "beta"}
    with pytest.raises(InputError):      # 60x60 matrix where a beta stack is announced
        CCFFit(copy.deepcopy(model), badd)


def test_no_cpu_fallback_without_a_gpu(fit):
    """On a box without a CUDA device the product must fail loudly, not compute on the CPU."""
    from victor_b200 import _lib
    if _lib.load().vb200_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fit.log_likelihood({"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0})


VARIANTS = [("dispersion", {"rsd_model": "dispersion"}), ("kaiser", {"rsd_model": "kaiser"}),
            ("anisotropic", {"assume_isotropic": False})]
MORE = [("euclid", {"rsd_model": "euclid_special"}),
        ("kaiser_noshift", {"rsd_model": "kaiser", "kaiser_coord_shift": False}),
        ("kaiser_approx", {"rsd_model": "kaiser", "kaiser_approximation": True}),
        ("kaiser_mq", {"rsd_model": "kaiser"}),
        ("aniso_dispersion", {"rsd_model": "dispersion", "assume_isotropic": False}),
        ("aniso_kaiser", {"rsd_model": "kaiser", "assume_isotropic": False})]


def _tables_vs_golden(fit, kw, rows, want_theory, want_chi2, want_lnl):
    from victor_b200 import tables as T
    opts = fit._merged_options(kw)
    mt = T.build_model_tables(fit, opts)
    ft = T.build_fit_tables(fit, fit.fit_options["likelihood"])
    mu, W = T.mu_projection_weights(fit.poles_s)
    mult, _ = E.theory_multipoles(mt, rows, np.asarray(fit.s, float), mu, W)
    theory = mult.reshape(len(rows), -1)
    np.testing.assert_allclose(theory, want_theory, rtol=RTOL, atol=ATOL)
    chi2, lnl = E.chi2_lnl(ft, rows[:, 1], theory)
    np.testing.assert_allclose(chi2, want_chi2, rtol=0, atol=C2_ATOL)
    np.testing.assert_allclose(lnl, want_lnl, rtol=0, atol=C2_ATOL)


@pytest.mark.parametrize("name,kw", VARIANTS)
def test_tables_reproduce_reference_variants(fit, golden, name, kw):
    """dispersion / kaiser / anisotropic input from the packed tables."""
    from victor_b200.model import params_to_rows
    g = golden("boss_variant_points")
    _tables_vs_golden(fit, kw, params_to_rows(g["params"]), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])


@pytest.mark.parametrize("name,kw", MORE)
def test_tables_reproduce_reference_more_variants(fit, golden, name, kw):
    from victor_b200.model import params_to_rows
    g = golden("boss_more_variants")
    rows = params_to_rows(g["params"])
    rows[:, 6:8] = g["MQ"]
    _tables_vs_golden(fit, kw, rows, g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])


@pytest.mark.parametrize("aniso", [False, True])
def test_tables_measured_model_from_data(boss_blocks, golden, aniso):
    """realspace_ccf from_data coordinates with the measured model file and the MD covariances."""
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    model["realspace_ccf"]["assume_isotropic"] = not aniso
    kind = "anisotropic" if aniso else "isotropic"
    data["covariance_matrix"]["data_file"] = f"data/boss_dr12_cmass/cmass_variable_{kind}_MD_covariance.npz"
    fm = CCFFit(model, data)
    if aniso:
        g = golden("boss_more_variants")
        rows = params_to_rows(g["measured_params"])
        _tables_vs_golden(fm, {}, rows, g["measured_aniso_theory"], g["measured_aniso_chi2"], g["measured_aniso_lnl"])
        _tables_vs_golden(fm, {"rsd_model": "dispersion"}, rows, g["measured_aniso_dispersion_theory"],
                          g["measured_aniso_dispersion_chi2"], g["measured_aniso_dispersion_lnl"])
    else:
        g = golden("boss_measured_model")
        _tables_vs_golden(fm, {}, params_to_rows(g["params"]), g["theory"], g["chi2"], g["lnl"])


def test_example_config_variants(example_block, golden):
    from victor_b200 import CCFModel, tables as T
    from victor_b200.model import params_to_rows
    g = golden("example_points")
    m = CCFModel(copy.deepcopy(example_block))
    mu, W = T.mu_projection_weights([0, 2, 4])
    P = {"fsigma8": g["params"][:, 0], "sigma_v": g["params"][:, 1], "epsilon": g["params"][:, 2]}
    for name in ("dispersion", "kaiser"):
        mt = T.build_model_tables(m, m._merged_options({"rsd_model": name}))
        mult, _ = E.theory_multipoles(mt, params_to_rows(P), g["s"], mu, W)
        np.testing.assert_allclose(mult.reshape(3, -1), g[f"{name}_theory"], rtol=RTOL, atol=ATOL)


def sv2d_blocks(boss_blocks):
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_sv2d_inputs.npz"
    model["velocity_pdf"]["dispersion"]["template_keys"] = ["rsv", "musv", "sigmav2d"]
    return model, data


def test_bicubic_dispersion_patches_match_scipy(boss_blocks):
    """sigma_v(r, mu) template: patches == RectBivariateSpline.ev with bispeu's argument clamping."""
    from scipy.interpolate import RectBivariateSpline
    from victor_b200 import CCFFit, tables as T
    fm = CCFFit(*sv2d_blocks(boss_blocks))
    assert not fm.sv_isotropic and fm.sv_rmu.shape == (10, 25)
    mt = T.build_model_tables(fm, fm.model)
    spl = RectBivariateSpline(fm.r_for_sv, fm.mu_for_sv, fm.sv_rmu.T)
    rng = np.random.default_rng(2)
    u = np.concatenate([rng.uniform(0, 170, 20000), mt.knots])
    m = np.concatenate([rng.uniform(-1, 1, 20000), np.linspace(-1, 1, len(mt.knots))])
    cell, t = E._cells(mt, u)
    np.testing.assert_allclose(E._sv(mt, cell, t, m), spl.ev(u, m), rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                                     ("aniso_streaming", {"assume_isotropic": False})])
def test_tables_sigma_v_r_mu_template(boss_blocks, golden, name, kw):
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    fm = CCFFit(*sv2d_blocks(boss_blocks))
    g = golden("boss_sv2d")
    np.testing.assert_allclose(fm.sv_rmu, g["sv_rmu"], rtol=1e-14)
    _tables_vs_golden(fm, kw, params_to_rows(g["params"]), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])


LB_CASES = [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}), ("kaiser", {"rsd_model": "kaiser"}),
            ("bias25", {"bias": 2.5})]


@pytest.mark.parametrize("name,kw", LB_CASES)
def test_tables_linear_bias_matter_model(boss_blocks, golden, name, kw):
    """matter_ccf model 'linear_bias' (ccf_model.py:358-370): V0 / D0 become beta power tables."""
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    fm = CCFFit(model, data)
    g = golden("boss_linear_bias")
    _tables_vs_golden(fm, kw, params_to_rows(g["params"]), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])


def test_tables_linear_bias_from_data(boss_blocks, golden):
    """linear_bias with the measured model: growth term beta * bias (ccf_model.py:429-430)."""
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    data["covariance_matrix"]["data_file"] = "data/boss_dr12_cmass/cmass_variable_isotropic_MD_covariance.npz"
    fm = CCFFit(model, data)
    g = golden("boss_linear_bias")
    _tables_vs_golden(fm, {}, params_to_rows(g["measured_params"]), g["measured_theory"], g["measured_chi2"],
                      g["measured_lnl"])


def test_direct_model_calls_from_tables(fit, golden):
    """Odd multipoles (mu grid [-1, 1]), a bare integer pole, a fine caller-supplied s grid, xi(s, mu)
    at negative mu -- the notebook-style calls of SURVEY.md 3.4 -- from the packed tables."""
    from victor_b200 import tables as T
    from victor_b200.model import params_to_rows
    g = golden("boss_misc_calls")
    mt = T.build_model_tables(fit, fit.model)
    prm = {"p0": {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0},
           "p1": {"fsigma8": 0.8, "beta": 0.45, "sigma_v": 250, "aperp": 1.03, "apar": 0.96}}
    s = np.asarray(fit.s, float)
    for tag, p in prm.items():
        rows = params_to_rows(dict(p))
        for poles, key, grid in (([0, 1, 2], "odd_012", s), ([1], "pole1", s), ([0, 2, 4], "fine_024", g["s_fine"]),
                                 ([2], "fine_bare2", g["s_fine"])):
            mu, W = T.mu_projection_weights(poles)
            assert (mu[0] == -1.0) == any(ell % 2 for ell in poles)
            mult, _ = E.theory_multipoles(mt, rows, grid, mu, W)
            np.testing.assert_allclose(mult.reshape(-1), g[f"{tag}_{key}"], rtol=RTOL, atol=ATOL)
    rows = params_to_rows(dict(prm["p1"]))
    xi = E.theory_xi(mt, rows, s, np.linspace(-1, 1, 11))[0]
    np.testing.assert_allclose(xi, g["xi_negmu"], rtol=RTOL, atol=ATOL)
    xi = E.theory_xi(mt, rows, np.sort(g["xi_unsorted_s"]), np.sort(g["xi_unsorted_mu"]))[0]
    np.testing.assert_allclose(xi, g["xi_unsorted"], rtol=RTOL, atol=ATOL)


def fixed_blocks(boss_blocks):
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/fixed_inputs_model.npz"
    model["realspace_ccf"]["reconstruction"] = False
    data["redshift_space_ccf"].update(reconstruction=False, data_file="tests/golden/fixed_inputs_data.npz")
    data["covariance_matrix"] = {"data_file": "tests/golden/fixed_inputs_cov.npz", "cov_key": "covmat"}
    return model, data


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                                     ("gaussian", {"likelihood": {"form": "gaussian"}})])
def test_tables_no_reconstruction_anywhere(boss_blocks, golden, name, kw):
    """1-D real-space and data multipoles, one covariance matrix: no beta dependence, beta not even
    required in the parameters (ccf_model.py:105-111, 585; ccf_fit.py:60-64, 130-133, 210-211)."""
    from victor_b200 import CCFFit, tables as T
    from victor_b200.model import params_to_rows
    fm = CCFFit(*fixed_blocks(boss_blocks))
    assert fm.fixed_real_input and fm.fixed_data and fm.fixed_covmat and fm.covmat.shape == (60, 60)
    g = golden("boss_fixed_everything")
    P = {"fsigma8": g["params"][:, 0], "sigma_v": g["params"][:, 2], "aperp": g["params"][:, 3], "apar": g["params"][:, 4]}
    rows = params_to_rows(P)
    assert np.all(np.isnan(rows[:, 1]))                 # beta absent -> NaN column, must not matter
    opts = fm._merged_options({k: v for k, v in kw.items() if k != "likelihood"})
    mt = T.build_model_tables(fm, opts)
    ft = T.build_fit_tables(fm, kw.get("likelihood", fm.fit_options["likelihood"]))
    mu, W = T.mu_projection_weights(fm.poles_s)
    mult, _ = E.theory_multipoles(mt, rows, np.asarray(fm.s, float), mu, W)
    theory = mult.reshape(len(rows), -1)
    np.testing.assert_allclose(theory, g[f"{name}_theory"], rtol=RTOL, atol=ATOL)
    chi2, lnl = E.chi2_lnl(ft, rows[:, 1], theory)
    np.testing.assert_allclose(chi2, g[f"{name}_chi2"], rtol=0, atol=C2_ATOL)
    np.testing.assert_allclose(lnl, g[f"{name}_lnl"], rtol=0, atol=C2_ATOL)


# ---------------------------------------------------------------------------------------------
# real-space input as xi(r, mu), the empirical velocity correction, a bias given with the
# parameters, the velocity-template mean model
# ---------------------------------------------------------------------------------------------
def rmu_blocks(boss_blocks, fixed=False):
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_rmu_inputs.npz"
    model["realspace_ccf"].update(format="rmu", ccf_keys=["r", "mu_rmu", "xi_rmu_fixed" if fixed else "xi_rmu"])
    if fixed:
        model["realspace_ccf"]["reconstruction"] = False
    return model, data


def test_rmu_input_is_converted_like_the_reference(boss_blocks, golden):
    """'rmu' format (ccf_model.py:154-181): multipoles 0, 2, 4 from a linear interp2d + 200-point trapezoid."""
    from victor_b200 import CCFFit, CCFModel, InputError
    g = golden("boss_rmu")
    fm = CCFFit(*rmu_blocks(boss_blocks))
    assert list(fm.poles_r) == [0, 2, 4]
    for ell in (0, 2, 4):
        assert fm.real_multipoles[f"{ell}"].shape == (31, 30)
        np.testing.assert_allclose(fm.real_multipoles[f"{ell}"], g[f"real_multipole_{ell}"], rtol=1e-13, atol=1e-16)
    cm = CCFModel(rmu_blocks(boss_blocks, fixed=True)[0])
    for ell in (0, 2, 4):
        np.testing.assert_allclose(cm.real_multipoles[f"{ell}"], g[f"fixed_real_multipole_{ell}"], rtol=1e-13,
                                   atol=1e-16)
    bad = rmu_blocks(boss_blocks)[0]
    bad["realspace_ccf"]["ccf_keys"] = ["r", "xi_rmu"]
    with pytest.raises(InputError):
        CCFModel(bad)
    bad = rmu_blocks(boss_blocks)[0]
    bad["realspace_ccf"]["ccf_keys"] = ["r", "mu_rmu", "xi_rmu_fixed"]      # shape lacks the beta axis
    with pytest.raises(InputError):
        CCFModel(bad)


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("aniso_streaming", {"assume_isotropic": False}),
                                     ("aniso_dispersion", {"assume_isotropic": False, "rsd_model": "dispersion"})])
def test_tables_rmu_input(boss_blocks, golden, name, kw):
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    fm = CCFFit(*rmu_blocks(boss_blocks))
    g = golden("boss_rmu")
    _tables_vs_golden(fm, kw, params_to_rows(g["params"]), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])


def test_tables_rmu_input_without_reconstruction(boss_blocks, golden):
    from victor_b200 import CCFModel, tables as T
    from victor_b200.model import params_to_rows
    g = golden("boss_rmu")
    cm = CCFModel(rmu_blocks(boss_blocks, fixed=True)[0])
    mt = T.build_model_tables(cm, cm._merged_options({"assume_isotropic": False}))
    assert mt.n_ell == 3 and not mt.beta_dependent
    mu, W = T.mu_projection_weights([0, 2, 4])
    s = np.load("tests/golden/fixed_inputs_data.npz")["s"]
    mult, _ = E.theory_multipoles(mt, params_to_rows(g["params"][:3]), s, mu, W)
    np.testing.assert_allclose(mult.reshape(3, -1), g["fixed_aniso_theory"], rtol=RTOL, atol=ATOL)


def velocity_rows(g, av=False, bias=False):
    from victor_b200.model import params_to_rows
    rows = params_to_rows(g["params"])
    if av:
        rows[:, 8] = g["Av"]
    if bias:
        rows[:, 9] = g["bias"]
    return rows


EMP_CASES = [("emp_streaming", {}), ("emp_dispersion", {"rsd_model": "dispersion"}), ("emp_kaiser", {"rsd_model": "kaiser"})]


@pytest.mark.parametrize("name,kw", EMP_CASES)
def test_tables_empirical_velocity_correction(boss_blocks, golden, name, kw):
    """velocity_pdf.mean.empirical_corr (ccf_model.py:451-459): v_r gains (1 + Av delta(r)), its slope comes
    from a finite difference on a 100-point grid; both are linear in Av -> V0 = v0 + Av v0b, D0 = d0 + Av d0b."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["velocity_pdf"]["mean"]["empirical_corr"] = True
    fm = CCFFit(model, data)
    g = golden("boss_velocity_options")
    _tables_vs_golden(fm, kw, velocity_rows(g, av=True), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])
    if name == "emp_dispersion":     # no Av among the parameters: 0, but still the finite-difference slope
        _tables_vs_golden(fm, kw, velocity_rows(g)[:3], g["emp_noAv_dispersion_theory"],
                          g["emp_noAv_dispersion_chi2"], g["emp_noAv_dispersion_lnl"])


@pytest.mark.parametrize("name,kw", [("rowbias_streaming", {}), ("rowbias_dispersion", {"rsd_model": "dispersion"})])
def test_tables_bias_given_with_the_parameters(boss_blocks, golden, name, kw):
    """params.get('bias', model['bias']) (ccf_model.py:359): delta and Delta scale as 1 / bias."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    fm = CCFFit(model, data)
    g = golden("boss_velocity_options")
    _tables_vs_golden(fm, kw, velocity_rows(g, bias=True), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])


def test_tables_empirical_correction_with_linear_bias(boss_blocks, golden):
    from victor_b200 import CCFModel, tables as T
    model = copy.deepcopy(boss_blocks[0])
    model["input_model_data_file"] = "tests/golden/fixed_inputs_model.npz"
    model["realspace_ccf"]["reconstruction"] = False
    model["matter_ccf"]["model"] = "linear_bias"
    model["velocity_pdf"]["mean"]["empirical_corr"] = True
    cm = CCFModel(model)
    g = golden("boss_velocity_options")
    mt = T.build_model_tables(cm, cm._merged_options({"rsd_model": "dispersion"}))
    assert mt.v0b is not None and mt.linear_bias and not mt.vd_beta_dependent
    mu, W = T.mu_projection_weights([0, 2])
    s = np.load("tests/golden/fixed_inputs_data.npz")["s"]
    mult, _ = E.theory_multipoles(mt, velocity_rows(g, av=True, bias=True), s, mu, W)
    np.testing.assert_allclose(mult.reshape(len(g["params"]), -1), g["emp_linbias_fixed_dispersion_theory"],
                               rtol=RTOL, atol=ATOL)
    # with a reconstruction-dependent monopole delta * Delta is no cubic in beta: refused, not approximated
    model = copy.deepcopy(boss_blocks[0])
    model["matter_ccf"]["model"] = "linear_bias"
    model["velocity_pdf"]["mean"]["empirical_corr"] = True
    cm = CCFModel(model)
    with pytest.raises(NotImplementedError):
        T.build_model_tables(cm, cm.model)


def vtemplate_blocks(boss_blocks):
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_vtemplate_inputs.npz"
    model["velocity_pdf"]["mean"].update(model="template", template_fsigma8=0.45, z_sim=0.5,
                                         template_hubble_ratio=1.02, template_keys=["rvel", "vr_template"])
    return model, data


@pytest.mark.parametrize("name,kw", [("vtemplate_streaming", {}), ("vtemplate_dispersion", {"rsd_model": "dispersion"}),
                                     ("vtemplate_kaiser", {"rsd_model": "kaiser"})])
def test_tables_velocity_template_mean_model(boss_blocks, golden, name, kw):
    """velocity_pdf.mean.model 'template' (ccf_model.py:227-246, 439-443, 483-488)."""
    from victor_b200 import CCFFit, InputError, tables as T
    fm = CCFFit(*vtemplate_blocks(boss_blocks))
    assert fm.has_velocity_template and fm.template_fsigma8 == 0.45
    g = golden("boss_velocity_options")
    _tables_vs_golden(fm, kw, velocity_rows(g), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])
    mt = T.build_model_tables(fm, fm._merged_options(kw))
    assert mt.growth_mode == T.GROWTH_VELOCITY_TEMPLATE and abs(mt.growth_scale - 1.02 * 1.5 / 1.57) < 1e-15
    bad = vtemplate_blocks(boss_blocks)[0]
    del bad["velocity_pdf"]["mean"]["template_fsigma8"]
    with pytest.raises(InputError):
        CCFFit(bad, vtemplate_blocks(boss_blocks)[1])


def test_profile_helpers_match_the_reference(fit, golden):
    """delta_profiles / velocity_terms (ccf_model.py:328-492): host helpers the notebooks plot."""
    g = golden("boss_helpers")
    p1 = {"fsigma8": 0.8, "beta": 0.45, "sigma_v": 250, "aperp": 1.03, "apar": 0.96}
    r = np.asarray(fit.r, float)
    for tag, kw in (("template", {}), ("linear_bias", {"matter_model": "linear_bias"})):
        d, D = fit.delta_profiles(r, dict(p1), **kw)
        np.testing.assert_allclose(d, g[f"delta_{tag}"], rtol=1e-13, atol=1e-16)
        np.testing.assert_allclose(D, g[f"Delta_{tag}"], rtol=1e-13, atol=1e-16)
    for tag, kw, extra in (("linear", {}, {}), ("empirical", {"empirical_corr": True}, {"Av": 0.7}),
                           ("linear_bias", {"matter_model": "linear_bias"}, {"bias": 2.2})):
        prm = dict(p1)
        prm.update(extra)
        vr, dvr = fit.velocity_terms(r, prm, **kw)
        np.testing.assert_allclose(vr, g[f"vr_{tag}"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(dvr, g[f"dvr_{tag}"], rtol=1e-12, atol=1e-13)
    with pytest.raises(NotImplementedError):
        fit.delta_profiles(r, dict(p1), matter_model="excursion_set")
    # host accessors of CCFFit (ccf_fit.py:166-323)
    np.testing.assert_allclose(fit.correlation_matrix(0.37), g["corrmat_037"], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(fit.diagonal_errors(0.37), g["errors_037"], rtol=1e-14)
    np.testing.assert_allclose(fit.get_interpolated_redshift_multipoles(0.41), g["data_multipoles_041"], rtol=1e-14)
    np.testing.assert_array_equal(fit.get_interpolated_precision(0.10), g["precision_below_grid"])
    np.testing.assert_array_equal(fit.get_interpolated_covariance(float(fit.beta_covmat[9])), g["covariance_on_node"])


def test_grid_interpolator_follows_legacy_interp2d():
    """Shapes and edge handling of the interp2d stand-in (what theory_xi_2D returns)."""
    from victor_b200.utils import GridInterpolator2D, fn_from_multipoles, multipoles_from_fn
    x, y = np.linspace(0, 4, 9), np.linspace(-1, 1, 5)
    z = y[:, None] * 2 + x[None, :] * 3                      # bilinear: reproduced exactly
    f = GridInterpolator2D(x, y, z)
    assert f(1.3, 0.2).shape == (1,) and abs(f(1.3, 0.2)[0] - (0.4 + 3.9)) < 1e-14
    assert f([3.0, 1.0], 0.0).shape == (2,) and np.allclose(f([3.0, 1.0], 0.0), [3.0, 9.0])    # sorted
    assert f(1.0, [0.5, -0.5]).shape == (2, 1)
    assert np.allclose(f([-5.0, 9.0], [0.0]), [0.0, 12.0])   # moved to the edges
    with pytest.raises(ValueError):
        GridInterpolator2D(x, y, z.T)
    r = np.linspace(1, 50, 20)
    mult = np.array([np.sin(r / 9), 0.3 * np.cos(r / 7)])
    back = multipoles_from_fn(fn_from_multipoles(r, [0, 2], mult), r, ell=[0, 2])
    np.testing.assert_allclose(back["0"], mult[0], atol=2e-4)
    np.testing.assert_allclose(back["2"], mult[1], atol=2e-4)


def test_pairwise_points_equal_the_grid_diagonal(fit, packed):
    """theory_xi_2D evaluates separate (s, mu) points; on the CPU the same points are the diagonal of
    the outer-product grid of the table walk-through (the GPU test holds the kernel to the reference)."""
    from victor_b200.model import params_to_rows
    mt = packed[0]
    sperp, spar, s, mu = fit._sky_grid(85)
    assert s.shape == (50, 50) and mu.min() < -0.99 and abs(s[0, 0] - np.hypot(0.01, 85)) < 1e-12
    pick = np.array([0, 777, 1249, 2499])
    rows = params_to_rows({"fsigma8": 0.8, "beta": 0.45, "sigma_v": 250, "aperp": 1.03, "apar": 0.96})
    sp, mp = s.ravel()[pick], mu.ravel()[pick]
    grid = E.theory_xi(mt, rows, sp, mp)[0]                   # [nmu][ns]
    assert np.all(np.isfinite(np.diag(grid)))


def test_cell_tables_on_random_knot_sets():
    """Property test (hypothesis): for arbitrary strictly increasing knot sets -- clustered, on a lattice,
    spanning decades -- the bucket table finds the searchsorted cell and the cell cubics equal the
    FITPACK ext=3 spline everywhere, below, between and above the knots."""
    from hypothesis import given, settings, strategies as st
    from scipy.interpolate import InterpolatedUnivariateSpline
    from victor_b200 import tables as T

    @settings(max_examples=60, deadline=None, derandomize=True, database=None)
    @given(st.integers(5, 40), st.floats(1e-3, 50.0), st.sampled_from(["uniform", "jitter", "cluster", "lattice"]),
           st.integers(0, 2 ** 31 - 1))
    def check(nk, scale, kind, seed):
        rng = np.random.default_rng(seed)
        if kind == "uniform":
            x = scale * (0.5 + np.arange(nk))
        elif kind == "lattice":
            x = scale * np.sort(rng.choice(np.arange(1, 4 * nk), nk, replace=False)).astype(float)
        elif kind == "jitter":
            x = scale * np.cumsum(rng.uniform(0.8, 1.2, nk))
        else:
            x = scale * np.cumsum(np.where(rng.uniform(size=nk) < 0.3, rng.uniform(0.01, 0.05, nk), rng.uniform(0.5, 2, nk)))
        y = rng.standard_normal(nk)
        extra = np.sort(rng.uniform(x[0], x[-1], 3))                     # knots of another spline in the union
        knots = T.union_knots(x, extra)
        inv_h, entry, maxscan = T.bucket_map(knots)

        class M:
            pass
        m = M()
        m.inv_h, m.bucket_base, m.maxscan = inv_h, entry, maxscan
        m.upper = np.concatenate([knots, [np.inf]])
        m.origin = np.concatenate([[knots[0]], knots])
        u = np.concatenate([rng.uniform(0, 1.5 * knots[-1], 3000), knots, np.nextafter(knots, 0),
                            np.nextafter(knots, np.inf), [0.0, 10 * knots[-1]]])
        cell, t = E._cells(m, u)
        want_cell = np.maximum(np.searchsorted(knots, u, side="right"), 1)
        # a u within an ulp or two of a knot may land in the cell on either side when the bucket spacing is not a
        # lattice of the knots (floor(u inv_h) against bucket edges rounded the other way): the splines are
        # continuous there, so both cells give the same value -- which is what the comparison below holds them to
        off = cell != want_cell
        gap = np.abs(u[off][:, None] - knots[None, :]).min(axis=1) if off.any() else np.zeros(0)
        assert np.all(np.abs(cell - want_cell)[off] == 1) and np.all(gap <= 4 * np.spacing(u[off]))
        coef = T.spline_cells(x, y, knots)
        want = InterpolatedUnivariateSpline(x, y, ext=3)(u)
        np.testing.assert_allclose(E._horner(coef, cell, t), want, rtol=1e-9, atol=1e-9 * np.abs(y).max())

    check()


def loader_option_blocks(boss_blocks):
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/loader_inputs_model.npz"
    model["cosmology"] = {"Omega_m": 0.29, "Omega_K": 0.01}
    model["realspace_ccf"].update(ccf_keys=["r", "monopole_sims", "quadrupole_sims"], simulation_number=2)
    model["matter_ccf"].update(template_keys=["rDelta", "Delta"], integrated=True)
    model["velocity_pdf"]["dispersion"]["filter"] = False
    data["redshift_space_ccf"].update(data_file="tests/golden/loader_inputs_data.npz",
                                      ccf_keys=["s", "monopole_sims", "quadrupole_sims"], simulation_number=1)
    return model, data


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"})])
def test_tables_loader_options(boss_blocks, golden, name, kw):
    """simulation_number, an integrated matter template, an unfiltered dispersion template and a non-default
    cosmology (ccf_model.py:99-297, ccf_fit.py:44-114): loader state and results against the reference."""
    from victor_b200 import CCFFit, InputError
    from victor_b200.model import params_to_rows
    fm = CCFFit(*loader_option_blocks(boss_blocks))
    g = golden("boss_loader_options")
    assert abs(fm.iaH - float(g["iaH"])) < 1e-16
    r31 = np.append([0.01], fm.r)
    np.testing.assert_allclose(fm.sv_rmu, g["sv_rmu"], rtol=1e-14)
    np.testing.assert_allclose(fm.delta(r31), g["delta_r31"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(fm.integrated_delta(r31), g["Delta_r31"], rtol=1e-13)
    np.testing.assert_array_equal(fm.real_multipoles["0"], g["real_mono"])
    np.testing.assert_array_equal(fm.redshift_multipoles["0"], g["data_mono"])
    _tables_vs_golden(fm, kw, params_to_rows(g["params"]), g[f"{name}_theory"], g[f"{name}_chi2"], g[f"{name}_lnl"])
    bad = loader_option_blocks(boss_blocks)
    bad[0]["realspace_ccf"]["simulation_number"] = 1.0
    with pytest.raises(InputError):
        CCFFit(*bad)


def test_one_bad_grid_covariance_fails_only_its_own_points(boss_blocks):
    """The reference checks slogdet per evaluation and returns (-inf, inf) only for points whose blended matrix
    fails it (ccf_fit.py:445-450): a non-positive-definite grid matrix must not stop the fit from loading."""
    import copy as _copy
    from victor_b200 import CCFFit, tables as T
    model, data = boss_blocks
    f = CCFFit(_copy.deepcopy(model), _copy.deepcopy(data))
    f.covmat = np.array(f.covmat, copy=True)
    f.covmat[4] = -f.covmat[4]
    ft = T.build_fit_tables(f, f.fit_options["likelihood"])
    assert np.isnan(ft.logdet[4]) and np.all(np.isnan(ft.lam[4]))
    ok = np.ones(len(ft.logdet), dtype=bool)
    ok[4] = False
    assert np.all(np.isfinite(ft.logdet[ok])) and np.all(np.isfinite(ft.lam[ok]))
    # walk-through with the kernels' algebra: beta in interval 4 fails, the others evaluate
    g = np.asarray(f.beta_covmat)
    betas = np.array([0.5 * (g[4] + g[5]), 0.5 * (g[7] + g[8]), g[4], 0.37])
    theory = np.tile(f.multipole_datavector(0.37), (len(betas), 1)) * 1.01
    chi2, lnl = E.chi2_lnl(ft, betas, theory)
    assert lnl[0] == -np.inf and chi2[0] == np.inf and lnl[2] == -np.inf
    assert np.isfinite(lnl[1]) and np.isfinite(lnl[3]) and np.isfinite(chi2[1])


def test_engine_key_follows_option_changes(fit, monkeypatch):
    """fit.model / fit.fit_options may be edited between calls (the reference re-reads both every call,
    ccf_fit.py:379-381, ccf_model.py:565-567): the engine is resolved from the current options each time,
    and `niter` (ccf_model.py:661) is part of the key."""
    built = []

    def fake_engine(self, opts, need_fit=False):
        built.append((opts["rsd_model"], int(opts.get("niter", 5)), self._fit_key(opts) if need_fit else None))
        return object()

    from victor_b200 import CCFFit
    monkeypatch.setattr(CCFFit, "_engine", fake_engine)
    fit._fit_engine({})
    old_model, old_like = fit.model["rsd_model"], fit.fit_options["likelihood"]
    try:
        fit.model["rsd_model"] = "dispersion"
        fit._fit_engine({})
        fit.fit_options["likelihood"] = {"form": "gaussian"}
        fit._fit_engine({"niter": 3})
    finally:
        fit.model["rsd_model"], fit.fit_options["likelihood"] = old_model, old_like
    assert [b[0] for b in built] == [old_model, "dispersion", "dispersion"]
    assert built[2][1] == 3 and built[2][2][0] == "gaussian" and built[0][2][0] == "sellentin"
    from victor_b200 import tables as T
    assert T.build_model_tables(fit, fit._merged_options({"rsd_model": "dispersion", "niter": 2})).niter == 2
