"""The plain C table walk (oracle/table_walk.c, test infrastructure): held to golden outputs of the unmodified
reference on the CPU, and used on the GPU box to recompute whole bench batches row by row."""
import copy

import numpy as np
import pytest

from oracle import table_walk as TW

RTOL, ATOL, C2_ATOL = 1e-9, 1e-13, 1e-6
GPU_ATOL = 1e-12    # CUDA path (MUFU seeds + one Newton step by default): see tests/test_gpu_parity.py


def measured_blocks(boss_blocks):
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    model["realspace_ccf"]["assume_isotropic"] = False
    data["covariance_matrix"]["data_file"] = "data/boss_dr12_cmass/cmass_variable_anisotropic_MD_covariance.npz"
    return model, data


def sv2d_blocks(boss_blocks):
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_sv2d_inputs.npz"
    model["velocity_pdf"]["dispersion"]["template_keys"] = ["rsv", "musv", "sigmav2d"]
    return model, data


@pytest.fixture(scope="module")
def fit(boss_blocks):
    from victor_b200 import CCFFit
    TW.build()
    return CCFFit(copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1]))


def test_streaming_rows_match_the_reference(fit, golden):
    from victor_b200.model import params_to_rows
    g = golden("boss_streaming_points")
    tw = TW.TableWalk(fit)
    th, c2, ll = tw.likelihood(params_to_rows(g["params"]), want_theory=True)
    np.testing.assert_allclose(th, g["theory"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(c2, g["chi2"], rtol=0, atol=C2_ATOL)
    np.testing.assert_allclose(ll, g["lnl"], rtol=0, atol=C2_ATOL)
    assert np.abs(th - g["theory"]).max() < 1e-13 and np.nanmax(np.abs(c2 - g["chi2"])) < 1e-9   # measured: 1e-15, 6e-12


def test_xi_and_options_match_the_reference(fit, golden, boss_blocks):
    from victor_b200 import CCFFit, tables as T
    from victor_b200.model import params_to_rows
    e = golden("boss_epsilon_points")
    tw = TW.TableWalk(fit)
    rows = params_to_rows({"fsigma8": e["params"][:, 0], "beta": e["params"][:, 1], "sigma_v": e["params"][:, 2],
                           "epsilon": e["params"][:, 3], "alpha": e["params"][:, 4]})
    xi, _ = tw.theory(rows, np.asarray(fit.s, float), e["mu"])
    np.testing.assert_allclose(xi, e["xi_smu"], rtol=RTOL, atol=ATOL)
    # anisotropic real-space input (three-term Legendre sum), likelihood forms
    v = golden("boss_variant_points")
    th, c2, ll = TW.TableWalk(fit, options={"assume_isotropic": False}).likelihood(params_to_rows(v["params"]), True)
    np.testing.assert_allclose(th, v["anisotropic_theory"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(c2, v["anisotropic_chi2"], rtol=0, atol=C2_ATOL)
    f = golden("boss_forms")
    for form in ("gaussian", "hartlap", "percival"):
        like = {"form": form, "nmocks": 1000, "nparams": 4}
        _, c2, ll = TW.TableWalk(fit, likelihood=like).likelihood(params_to_rows(f["params"]))
        np.testing.assert_allclose(ll, f[f"{form}_lnl"], rtol=0, atol=C2_ATOL)
    # empirical velocity correction, velocity template, linear_bias with a bias among the parameters
    o = golden("boss_velocity_options")
    rows = params_to_rows(o["params"])
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["velocity_pdf"]["mean"]["empirical_corr"] = True
    r2 = rows.copy()
    r2[:, 8] = o["Av"]
    _, c2, _ = TW.TableWalk(CCFFit(model, data)).likelihood(r2)
    np.testing.assert_allclose(c2, o["emp_streaming_chi2"], rtol=0, atol=C2_ATOL)
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    r2 = rows.copy()
    r2[:, 9] = o["bias"]
    _, c2, _ = TW.TableWalk(CCFFit(model, data)).likelihood(r2)
    np.testing.assert_allclose(c2, o["rowbias_streaming_chi2"], rtol=0, atol=C2_ATOL)
    assert T.NPAR == 10


def test_other_rsd_models_match_the_reference(fit, golden, boss_blocks):
    """dispersion, kaiser (+ M, Q, no coordinate shift, linear approximation), euclid_special, with isotropic and
    anisotropic real-space input, from-data coordinates, a sigma_v(r, mu) template."""
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    v = golden("boss_variant_points")
    rows = params_to_rows(v["params"])
    for name, kw in (("dispersion", {"rsd_model": "dispersion"}), ("kaiser", {"rsd_model": "kaiser"})):
        th, c2, ll = TW.TableWalk(fit, options=kw).likelihood(rows, want_theory=True)
        np.testing.assert_allclose(th, v[f"{name}_theory"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(c2, v[f"{name}_chi2"], rtol=0, atol=C2_ATOL)
        np.testing.assert_allclose(ll, v[f"{name}_lnl"], rtol=0, atol=C2_ATOL)
    mv = golden("boss_more_variants")
    rows = params_to_rows(mv["params"])
    rows[:, 6:8] = mv["MQ"]
    for name, kw in (("euclid", {"rsd_model": "euclid_special"}),
                     ("kaiser_noshift", {"rsd_model": "kaiser", "kaiser_coord_shift": False}),
                     ("kaiser_approx", {"rsd_model": "kaiser", "kaiser_approximation": True}),
                     ("kaiser_mq", {"rsd_model": "kaiser"}),
                     ("aniso_dispersion", {"rsd_model": "dispersion", "assume_isotropic": False})):
        th, c2, _ = TW.TableWalk(fit, options=kw).likelihood(rows, want_theory=True)
        np.testing.assert_allclose(th, mv[f"{name}_theory"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(c2, mv[f"{name}_chi2"], rtol=0, atol=C2_ATOL)
    # from-data coordinates (measured model + MD covariance) and a sigma_v(r, mu) template
    fm = CCFFit(*measured_blocks(boss_blocks))
    rows = params_to_rows(mv["measured_params"])
    for name, kw in (("measured_aniso", {}), ("measured_aniso_dispersion", {"rsd_model": "dispersion"})):
        th, c2, _ = TW.TableWalk(fm, options=kw).likelihood(rows, want_theory=True)
        np.testing.assert_allclose(th, mv[f"{name}_theory"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(c2, mv[f"{name}_chi2"], rtol=0, atol=C2_ATOL)
    sv = golden("boss_sv2d")
    fs = CCFFit(*sv2d_blocks(boss_blocks))
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"})):
        th, c2, _ = TW.TableWalk(fs, options=kw).likelihood(params_to_rows(sv["params"]), want_theory=True)
        np.testing.assert_allclose(th, sv[f"{name}_theory"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(c2, sv[f"{name}_chi2"], rtol=0, atol=C2_ATOL)


@pytest.mark.gpu
def test_whole_bench_batch_against_the_c_table_walk(fit):
    """The WHOLE BASELINE batch (all 65 536 seeded rows of bench.py), every row recomputed on the CPU by the C table
    walk (~15 s on 16 host cores): multipoles to 1e-9 relative, chi-square and lnL to 1e-6 absolute, row by row."""
    from bench import synthetic_batch
    from victor_b200.model import params_to_rows
    rows = params_to_rows(synthetic_batch(65536))
    lnl, chi2, th = fit.log_likelihood_batch(rows, return_theory=True)
    wth, wc2, wll = TW.TableWalk(fit).likelihood(rows, want_theory=True)
    np.testing.assert_allclose(th, wth, rtol=RTOL, atol=GPU_ATOL)
    np.testing.assert_allclose(chi2, wc2, rtol=0, atol=C2_ATOL)
    np.testing.assert_allclose(lnl, wll, rtol=0, atol=C2_ATOL)
    scale = np.abs(wth).reshape(len(rows), 2, -1).max(axis=2)
    err = np.abs(th - wth).reshape(len(rows), 2, -1).max(axis=2)
    assert np.all(err <= RTOL * scale)


@pytest.mark.gpu
@pytest.mark.parametrize("kw,n", [({"rsd_model": "dispersion"}, 4096), ({"assume_isotropic": False}, 8192),
                                  ({"rsd_model": "kaiser"}, 16384), ({"rsd_model": "euclid_special"}, 16384)])
def test_general_kernel_batches_against_the_c_table_walk(fit, kw, n):
    """The general kernel, thousands of seeded rows per model, every row recomputed by the C table walk."""
    from bench import synthetic_batch
    from victor_b200.model import params_to_rows
    rows = params_to_rows(synthetic_batch(65536)[:n])
    rng = np.random.default_rng(8)
    rows[:, 6] = rng.uniform(0.8, 1.2, n)          # M, Q (kaiser forms)
    rows[:, 7] = rng.uniform(0.8, 1.2, n)
    lnl, chi2, th = fit.log_likelihood_batch(rows, return_theory=True, **kw)
    wth, wc2, wll = TW.TableWalk(fit, options=kw).likelihood(rows, want_theory=True)
    np.testing.assert_allclose(th, wth, rtol=RTOL, atol=GPU_ATOL)
    np.testing.assert_allclose(chi2, wc2, rtol=0, atol=C2_ATOL)
    np.testing.assert_allclose(lnl, wll, rtol=0, atol=C2_ATOL)


@pytest.mark.gpu
@pytest.mark.parametrize("which,kw,n", [("measured", {}, 4096), ("measured", {"rsd_model": "dispersion"}, 2048),
                                        ("sv2d", {}, 4096), ("sv2d", {"rsd_model": "dispersion"}, 2048)])
def test_from_data_and_sv2d_batches_against_the_c_table_walk(boss_blocks, which, kw, n):
    """From-data coordinates (measured model, anisotropic, MD covariance) and a sigma_v(r, mu) template: thousands
    of seeded rows, every row recomputed by the C table walk."""
    from bench import synthetic_batch
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    fm = CCFFit(*(measured_blocks(boss_blocks) if which == "measured" else sv2d_blocks(boss_blocks)))
    rows = params_to_rows(synthetic_batch(65536)[:n])
    if which == "measured":
        rows[:, 1] = np.clip(rows[:, 1], 0.25, 0.55)     # the measured model's beta grid is narrower
    lnl, chi2, th = fm.log_likelihood_batch(rows, return_theory=True, **kw)
    wth, wc2, wll = TW.TableWalk(fm, options=kw).likelihood(rows, want_theory=True)
    np.testing.assert_allclose(th, wth, rtol=RTOL, atol=GPU_ATOL)
    np.testing.assert_allclose(chi2, wc2, rtol=0, atol=C2_ATOL)
    np.testing.assert_allclose(lnl, wll, rtol=0, atol=C2_ATOL)
    fm.close()
