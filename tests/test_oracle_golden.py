"""CPU: the oracle (oracle/ccf_oracle.py) against outputs of the UNMODIFIED reference.

The golden files under tests/golden/ were written by oracle/make_golden.py, which imports the
reference from /root/reference through oracle/refshim.py (dev container only).  The reference
ships no tests of its own; its only published numbers are the chi2 / lnL pairs printed in
notebooks/victor_usage_demo.ipynb cell 22, checked here to the printed digits.

The oracle calls the same scipy routines in the same order as the reference, so agreement is
expected at the level of a few ulp; the bounds below (1e-12 relative on multipoles, 1e-9 on
chi2 / lnL) leave room for BLAS / scipy build differences between boxes.
"""
import copy

import numpy as np
import pytest

from oracle.ccf_oracle import OracleFit, OracleModel

COLS = ("fsigma8", "beta", "sigma_v", "aperp", "apar")
TH_RTOL, TH_ATOL, C2_ATOL = 1e-12, 1e-15, 1e-9


def as_params(row):
    return dict(zip(COLS, map(float, row)))


@pytest.fixture(scope="module")
def orc(boss_blocks):
    model, data = boss_blocks
    return OracleFit(copy.deepcopy(model), copy.deepcopy(data))


def check_rows(orc, P, theory, chi2, lnl, **kw):
    for i, row in enumerate(P):
        prm = as_params(row)
        th = orc.theory_multipole_vector(orc.s, dict(prm), orc.poles_s, **kw)
        np.testing.assert_allclose(th, theory[i], rtol=TH_RTOL, atol=TH_ATOL)
        l, c = orc.log_likelihood(dict(prm), **kw)
        if np.isfinite(chi2[i]):
            assert abs(c - chi2[i]) < C2_ATOL and abs(l - lnl[i]) < C2_ATOL
        else:
            assert c == chi2[i] and l == lnl[i]


def test_notebook_cell22_numbers(orc, golden):
    """The five printed pairs of the usage-demo notebook (2 d.p.); the anisotropic chi2 differs by
    0.01 from the print because scipy's Simpson end rule changed since the notebook was run
     -- the unmodified reference gives 64.40 on this scipy as well."""
    a = golden("boss_notebook_anchors")
    p0 = {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0}
    printed = {"streaming": (65.01, 284.76, {}), "dispersion": (65.03, 284.76, {"rsd_model": "dispersion"}),
               "kaiser": (103.90, 266.81, {"rsd_model": "kaiser"}),
               "anisotropic": (64.39, 285.06, {"assume_isotropic": False}),
               "likelihood_interp": (64.80, 285.30, {"beta_interpolation": "likelihood"})}
    for name, (c2p, lnlp, kw) in printed.items():
        lnl, c2 = orc.log_likelihood(dict(p0), **kw)
        tol = 0.015 if name == "anisotropic" else 0.0051
        assert abs(c2 - c2p) < tol and abs(lnl - lnlp) < tol, (name, c2, lnl)
        assert abs(c2 - float(a[f"{name}_chi2"])) < C2_ATOL
        assert abs(lnl - float(a[f"{name}_lnl"])) < C2_ATOL
        if f"{name}_theory" in a.files:
            th = orc.theory_multipole_vector(orc.s, dict(p0), orc.poles_s, **kw)
            np.testing.assert_allclose(th, a[f"{name}_theory"], rtol=TH_RTOL, atol=TH_ATOL)
    assert abs(np.linalg.slogdet(orc.covariance_at(0.37))[1] - float(a["slogdet_cov"])) < 1e-9
    np.testing.assert_allclose(orc.data_vector(0.37), a["data_vector"], rtol=1e-14)


def test_host_state_matches_reference_loaders(orc, golden):
    t = golden("boss_tables")
    assert abs(orc.iaH - float(t["iaH"])) < 1e-16
    for key in ("r", "s", "beta", "sv_rmu", "r_for_sv", "mu_for_sv"):
        np.testing.assert_allclose(getattr(orc, key), t[key], rtol=1e-14, atol=0)
    r31 = np.append([0.01], orc.r)
    np.testing.assert_allclose(orc.delta(r31), t["delta_r31"], rtol=1e-13)
    np.testing.assert_allclose(orc.integrated_delta(r31), t["Delta_r31"], rtol=1e-13)
    np.testing.assert_allclose(orc.icov[0], t["icov_first"], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(orc.icov[-1], t["icov_last"], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(orc.real_multipoles_at(0.37), t["xi_r_beta037"], rtol=1e-14)


def test_streaming_rows_seeded_and_edges(orc, golden):
    """A slice of the seeded bench rows plus every hand-picked edge row (beta on / off / outside
    the grid, prior corners, narrow and wide pdf, zero outflow)."""
    g = golden("boss_streaming_points")
    idx = np.r_[0:6, 64:80]
    check_rows(orc, g["params"][idx], g["theory"][idx], g["chi2"][idx], g["lnl"][idx])


@pytest.mark.parametrize("name,kw", [("dispersion", {"rsd_model": "dispersion"}),
                                     ("kaiser", {"rsd_model": "kaiser"}),
                                     ("anisotropic", {"assume_isotropic": False})])
def test_variant_rows(orc, golden, name, kw):
    g = golden("boss_variant_points")
    idx = np.array([0, 3, 6, 7, 9, 10])
    check_rows(orc, g["params"][idx], g[f"{name}_theory"][idx], g[f"{name}_chi2"][idx],
               g[f"{name}_lnl"][idx], **kw)


def test_likelihood_interpolation_rows(orc, golden):
    g = golden("boss_variant_points")
    for i, row in enumerate(g["params"]):
        if not np.isfinite(g["likelihood_interp_chi2"][i]):
            continue
        l, c = orc.log_likelihood(as_params(row), beta_interpolation="likelihood")
        assert abs(c - g["likelihood_interp_chi2"][i]) < C2_ATOL
        assert abs(l - g["likelihood_interp_lnl"][i]) < C2_ATOL


def test_likelihood_forms_and_fixed_covariance(orc, golden, boss_blocks):
    g = golden("boss_forms")
    for form in ("gaussian", "hartlap", "percival", "sellentin"):
        like = {"form": form, "nmocks": 1000, "nparams": 4}
        for i, row in enumerate(g["params"][:2]):
            l, c = orc.log_likelihood(as_params(row), likelihood=like)
            assert abs(c - g[f"{form}_chi2"][i]) < C2_ATOL and abs(l - g[f"{form}_lnl"][i]) < C2_ATOL
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    data["covariance_matrix"] = {"data_file": "data/boss_dr12_cmass/cmass_fixed_D_covariance.npz",
                                 "cov_key": "covmat", "fixed_beta": True}
    ofix = OracleFit(model, data)
    for i, row in enumerate(g["params"][:2]):
        l, c = ofix.log_likelihood(as_params(row))
        assert abs(c - g["fixedcov_chi2"][i]) < C2_ATOL and abs(l - g["fixedcov_lnl"][i]) < C2_ATOL


def test_cobaya_block_astar(golden, repo_root):
    import yaml
    g = golden("boss_cobaya_block")
    with open(f"{repo_root}/config/boss_cobaya_config.yaml") as fh:
        blk = yaml.full_load(fh)["likelihood"]["CCFLikelihood"]
    blk["model"]["dir"] = blk["data"]["dir"] = repo_root
    oc = OracleFit(blk["model"], blk["data"])
    for i, row in enumerate(g["params"]):
        pr = dict(fsigma8=row[0], beta=row[1], sigma_v=row[2], epsilon=row[3], alpha=1, astar=row[4],
                  b=1.9, Av=0, M=1, Q=1)
        th = oc.theory_multipole_vector(oc.s, dict(pr), oc.poles_s)
        np.testing.assert_allclose(th, g["theory"][i], rtol=TH_RTOL, atol=TH_ATOL)
        l, c = oc.log_likelihood(dict(pr))
        assert abs(c - g["chi2"][i]) < C2_ATOL and abs(l - g["lnl"][i]) < C2_ATOL


def test_measured_model_from_data_coordinates(golden, boss_blocks):
    """*_measured_model + isotropic MD covariance: realspace_ccf from_data coordinates (15-point
    beta grid), SURVEY.md 8(f) rank 2."""
    g = golden("boss_measured_model")
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    data["covariance_matrix"]["data_file"] = "data/boss_dr12_cmass/cmass_variable_isotropic_MD_covariance.npz"
    om = OracleFit(model, data)
    np.testing.assert_allclose(om.beta_covmat, g["beta_covmat"], rtol=1e-14)
    check_rows(om, g["params"][:2], g["theory"][:2], g["chi2"][:2], g["lnl"][:2])


def test_example_config(golden, example_block):
    """Non-uniform r grid, no beta dependence, poles 0/2/4, all three rsd models."""
    g = golden("example_points")
    om = OracleModel(copy.deepcopy(example_block))
    assert abs(om.iaH - float(g["iaH"])) < 1e-16
    np.testing.assert_allclose(om.r, g["r"], rtol=1e-15)
    np.testing.assert_allclose(om.sv_rmu, g["sv_rmu"], rtol=1e-13)
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                     ("kaiser", {"rsd_model": "kaiser"})):
        row = g["params"][1]
        pr = dict(fsigma8=row[0], sigma_v=row[1], epsilon=row[2])
        th = om.theory_multipole_vector(g["s"], pr, [0, 2, 4], **kw)
        np.testing.assert_allclose(th, g[f"{name}_theory"][1], rtol=1e-11, atol=1e-14)


def test_more_variants(orc, golden):
    """euclid_special, kaiser options with M and Q, anisotropic input under dispersion / kaiser."""
    g = golden("boss_more_variants")
    cases = {"euclid": {"rsd_model": "euclid_special"},
             "kaiser_noshift": {"rsd_model": "kaiser", "kaiser_coord_shift": False},
             "kaiser_approx": {"rsd_model": "kaiser", "kaiser_approximation": True},
             "kaiser_mq": {"rsd_model": "kaiser"},
             "aniso_dispersion": {"rsd_model": "dispersion", "assume_isotropic": False},
             "aniso_kaiser": {"rsd_model": "kaiser", "assume_isotropic": False}}
    for name, kw in cases.items():
        for i in (1, 5):
            prm = as_params(g["params"][i])
            prm.update(M=float(g["MQ"][i, 0]), Q=float(g["MQ"][i, 1]))
            th = orc.theory_multipole_vector(orc.s, dict(prm), orc.poles_s, **kw)
            np.testing.assert_allclose(th, g[f"{name}_theory"][i], rtol=TH_RTOL, atol=TH_ATOL)
            l, c = orc.log_likelihood(dict(prm), **kw)
            assert abs(c - g[f"{name}_chi2"][i]) < C2_ATOL and abs(l - g[f"{name}_lnl"][i]) < C2_ATOL


def test_sigma_v_r_mu_template(golden, boss_blocks):
    g = golden("boss_sv2d")
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_sv2d_inputs.npz"
    model["velocity_pdf"]["dispersion"]["template_keys"] = ["rsv", "musv", "sigmav2d"]
    om = OracleFit(model, data)
    np.testing.assert_allclose(om.sv_rmu, g["sv_rmu"], rtol=1e-14)
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"})):
        check_rows(om, g["params"][[1, 6]], g[f"{name}_theory"][[1, 6]], g[f"{name}_chi2"][[1, 6]],
                   g[f"{name}_lnl"][[1, 6]], **kw)


def test_linear_bias_matter_model(golden, boss_blocks):
    g = golden("boss_linear_bias")
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    om = OracleFit(model, data)
    for name, kw in (("streaming", {}), ("kaiser", {"rsd_model": "kaiser"}), ("bias25", {"bias": 2.5})):
        check_rows(om, g["params"][[2, 7]], g[f"{name}_theory"][[2, 7]], g[f"{name}_chi2"][[2, 7]],
                   g[f"{name}_lnl"][[2, 7]], **kw)


def test_rmu_format_input(golden, boss_blocks):
    """Real-space ccf given as xi(r, mu) (ccf_model.py:154-181)."""
    g = golden("boss_rmu")
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_rmu_inputs.npz"
    model["realspace_ccf"].update(format="rmu", ccf_keys=["r", "mu_rmu", "xi_rmu"])
    om = OracleFit(model, data)
    for ell in (0, 2, 4):
        np.testing.assert_allclose(om.real_multipoles[f"{ell}"], g[f"real_multipole_{ell}"], rtol=1e-13, atol=1e-16)
    check_rows(om, g["params"][[1, 5]], g["aniso_streaming_theory"][[1, 5]], g["aniso_streaming_chi2"][[1, 5]],
               g["aniso_streaming_lnl"][[1, 5]], assume_isotropic=False)


def test_velocity_options(golden, boss_blocks):
    """Empirical correction with Av, bias among the parameters, velocity-template mean model
    (ccf_model.py:227-246, 359, 439-459, 483-488)."""
    g = golden("boss_velocity_options")
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["velocity_pdf"]["mean"]["empirical_corr"] = True
    om = OracleFit(model, data)
    for i in (1, 3):
        prm = as_params(g["params"][i])
        prm["Av"] = float(g["Av"][i])
        th = om.theory_multipole_vector(om.s, dict(prm), om.poles_s, rsd_model="dispersion")
        np.testing.assert_allclose(th, g["emp_dispersion_theory"][i], rtol=TH_RTOL, atol=TH_ATOL)
        l, c = om.log_likelihood(dict(prm))
        assert abs(c - g["emp_streaming_chi2"][i]) < C2_ATOL and abs(l - g["emp_streaming_lnl"][i]) < C2_ATOL

    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    om = OracleFit(model, data)
    prm = as_params(g["params"][4])
    prm["bias"] = float(g["bias"][4])
    l, c = om.log_likelihood(dict(prm))
    assert abs(c - g["rowbias_streaming_chi2"][4]) < C2_ATOL and abs(l - g["rowbias_streaming_lnl"][4]) < C2_ATOL

    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_vtemplate_inputs.npz"
    model["velocity_pdf"]["mean"].update(model="template", template_fsigma8=0.45, z_sim=0.5,
                                         template_hubble_ratio=1.02, template_keys=["rvel", "vr_template"])
    om = OracleFit(model, data)
    check_rows(om, g["params"][[2]], g["vtemplate_streaming_theory"][[2]], g["vtemplate_streaming_chi2"][[2]],
               g["vtemplate_streaming_lnl"][[2]])
    check_rows(om, g["params"][[5]], g["vtemplate_dispersion_theory"][[5]], g["vtemplate_dispersion_chi2"][[5]],
               g["vtemplate_dispersion_lnl"][[5]], rsd_model="dispersion")


def test_loader_options(golden, boss_blocks):
    """simulation_number, integrated matter template, unfiltered dispersion template, non-default cosmology."""
    g = golden("boss_loader_options")
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/loader_inputs_model.npz"
    model["cosmology"] = {"Omega_m": 0.29, "Omega_K": 0.01}
    model["realspace_ccf"].update(ccf_keys=["r", "monopole_sims", "quadrupole_sims"], simulation_number=2)
    model["matter_ccf"].update(template_keys=["rDelta", "Delta"], integrated=True)
    model["velocity_pdf"]["dispersion"]["filter"] = False
    data["redshift_space_ccf"].update(data_file="tests/golden/loader_inputs_data.npz",
                                      ccf_keys=["s", "monopole_sims", "quadrupole_sims"], simulation_number=1)
    om = OracleFit(model, data)
    assert abs(om.iaH - float(g["iaH"])) < 1e-16
    check_rows(om, g["params"][[1, 5]], g["dispersion_theory"][[1, 5]], g["dispersion_chi2"][[1, 5]],
               g["dispersion_lnl"][[1, 5]], rsd_model="dispersion")
