"""GPU parity: the CUDA path, called through the C ABI (ctypes), against golden outputs of the
unmodified reference and against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): multipoles relative 1e-9 -- applied as rtol = 1e-9 with
atol = 1e-12 elementwise (|xi_2| crosses zero, where a pure relative test is ill-defined; 1e-12 is 1e-3 of the
relative tolerance at the multipoles' own scale, |xi_0| ~ 1, |xi_2| ~ 0.1) AND as an inf-norm-relative bound per
multipole; chi-square and lnL absolute 1e-6.  Measured margins of the default kernels over 16,384 rows of the
bench batch (profiles/r02c_parity_report.jsonl): 5.5e-12 inf-norm-relative, 2.8e-13 absolute, 7.5e-10 in chi-square.
"""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL, CHI2_ATOL = 1e-9, 1e-12, 1e-6


def assert_theory(got, want, ns=None):
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    got2 = np.atleast_2d(got)
    want2 = np.atleast_2d(want)
    ns = ns or want2.shape[1]
    for a in range(0, want2.shape[1], ns):
        scale = np.abs(want2[:, a:a + ns]).max(axis=1)
        err = np.abs(got2[:, a:a + ns] - want2[:, a:a + ns]).max(axis=1)
        assert np.all(err <= RTOL * scale + ATOL)


@pytest.fixture(scope="module")
def fit(boss_blocks):
    from victor_b200 import CCFFit
    model, data = boss_blocks
    f = CCFFit(copy.deepcopy(model), copy.deepcopy(data))
    yield f
    f.close()


def test_native_library_is_the_path():
    from victor_b200 import _lib
    lib = _lib.load()
    assert lib.vb200_device_count() >= 1


def test_fast_math_primitives():
    """Hand-rolled exp / rsqrt / rcp (libvictor_b200_probes.so runs the kernels' own device functions) against
    numpy: a few ulp for the default forms, and the stated bounds for the cheaper variants."""
    from victor_b200 import _probes
    lib = _probes.load()
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(1e-6, 400.0, 200000), 10.0 ** rng.uniform(-8, 8, 50000),
                        np.array([0.0, 1e-300, 1.0, 2.0, 1300.0])])
    x = np.ascontiguousarray(x)
    n = len(x)
    out = np.empty(_probes.SELFTEST_OUTPUTS * n)
    rc = lib.vb200p_math_selftest(0, x.ctypes.data, n, out.ctypes.data)
    assert rc == 0, _probes.last_error()
    g, rs, rc_, g5, g52, g53, rs_n, rc_n, gb, gbm = (out[i * n:(i + 1) * n] for i in range(10))
    want = np.exp(-0.5 * x)
    # |z| <= 10: the range that carries the integral; a few ulp.  Beyond it the single-constant
    # range reduction loses ~1e-17 * z^2 relative, irrelevant where exp(-z^2/2) < 2e-22.
    core = x <= 100.0
    tail = (x > 100.0) & (want > 1e-280)
    for got in (g, g5):
        assert np.max(np.abs(got[core] / want[core] - 1)) < 4e-15
        assert np.max(np.abs(got[tail] / want[tail] - 1)) < 2e-13
    # the scaled-argument forms take zs = sqrt(x) * scale, rounded twice before it is squared: compare with the
    # exponential of the argument they were actually given, in extended precision
    def scaled_want(scale, table):
        zs = (np.sqrt(x) * scale).astype(np.longdouble)
        return np.exp(-(zs * zs) * (np.log(np.longdouble(2)) / table)).astype(np.float64)
    w32, w1024 = scaled_want(4.804489635145799, 32), scaled_want(27.178297609216609367, 1024)
    # FP32 tails of the remainder polynomial: 1.2e-14 / 3.6e-12 by construction (common.cuh)
    assert np.max(np.abs(g52[core] / w32[core] - 1)) < 5e-14
    assert np.max(np.abs(g53[core] / w32[core] - 1)) < 8e-12
    # 1024-entry table + degree-3 remainder: as exact as the degree-5 form; the conversion-unit range reduction
    # rounds zs^2 once (relative 1.1e-16 of the exponent: x / 2 * 1.1e-16 of the result, as libm's exp(-0.5 * x))
    assert np.max(np.abs(gbm[core] / w1024[core] - 1)) < 4e-15
    assert np.max(np.abs(gb[core] / w1024[core] - 1) - 0.5 * x[core] * 1.2e-16) < 4e-15
    # ... and saturates: exp(-x/2) of a huge argument is 0, not a wrapped exponent
    assert gb[-1] < 1e-280 and np.all(np.isfinite(gb))
    pos = (x > 1e-290)
    assert np.max(np.abs(rs[pos] * np.sqrt(x[pos]) - 1)) < 1e-15
    assert np.max(np.abs(rc_[pos] * x[pos] - 1)) < 1e-15
    # one Newton step on the 2^-20 MUFU seeds: 3/8 e^2 and e^2
    assert np.max(np.abs(rs_n[pos] * np.sqrt(x[pos]) - 1)) < 2e-12
    assert np.max(np.abs(rc_n[pos] * x[pos] - 1)) < 2e-12


@pytest.mark.parametrize("fast", [1, 0])
def test_boss_streaming_golden(fit, golden, fast):
    g = golden("boss_streaming_points")
    eng, _ = fit._fit_engine({})
    eng.set_option("fast_math", fast)
    try:
        lnl, chi2, theory = fit.log_likelihood_batch(g["params"], return_theory=True)
    finally:
        eng.set_option("fast_math", 1)
    assert_theory(theory, g["theory"], ns=len(fit.s))
    np.testing.assert_allclose(chi2, g["chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g["lnl"], rtol=0, atol=CHI2_ATOL)


@pytest.mark.parametrize("opts", [{"ilp": 1}, {"ilp": 4}, {"exp_degree": 5, "newton": 3}, {"exp_degree": 5, "newton": 2},
                                  {"exp_degree": 3, "newton": 2}, {"threads": 256}, {"threads": 64}])
def test_kernel_variants_hold_parity(fit, golden, opts):
    """Every tuning variant of K1 must meet the same bar as the default."""
    g = golden("boss_streaming_points")
    eng, _ = fit._fit_engine({})
    defaults = {"ilp": 0, "exp_degree": 0, "threads": 0, "newton": 0}        # 0 = the library's default
    try:
        for k, v in opts.items():
            eng.set_option(k, v)
        lnl, chi2, theory = fit.log_likelihood_batch(g["params"], return_theory=True)
    finally:
        for k, v in defaults.items():
            eng.set_option(k, v)
    assert_theory(theory, g["theory"], ns=len(fit.s))
    np.testing.assert_allclose(chi2, g["chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g["lnl"], rtol=0, atol=CHI2_ATOL)


def test_notebook_anchor_single_point(fit, golden):
    a = golden("boss_notebook_anchors")
    params = {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0}
    lnl, chi2 = fit.log_likelihood(dict(params))
    assert abs(chi2 - 65.01) < 0.005 and abs(lnl - 284.76) < 0.005      # notebook cell 22
    assert abs(chi2 - float(a["streaming_chi2"])) < CHI2_ATOL
    assert abs(lnl - float(a["streaming_lnl"])) < CHI2_ATOL
    th = fit.theory_multipole_vector(fit.s, dict(params), fit.poles_s)
    assert_theory(th, a["streaming_theory"], ns=len(fit.s))
    c2, cov = fit.chi_squared(dict(params))
    assert abs(c2 - chi2) < 1e-9
    assert abs(np.linalg.slogdet(cov)[1] - float(a["slogdet_cov"])) < 1e-9
    mp = fit.theory_multipoles(fit.s, dict(params), poles=[0, 2])
    assert set(mp) == {"0", "2"} and mp["0"].shape == (30,)


def test_epsilon_rows_and_theory_xi(fit, golden):
    g = golden("boss_epsilon_points")
    P = {"fsigma8": g["params"][:, 0], "beta": g["params"][:, 1], "sigma_v": g["params"][:, 2],
         "epsilon": g["params"][:, 3], "alpha": g["params"][:, 4]}
    lnl, chi2, theory = fit.log_likelihood_batch(P, return_theory=True)
    assert_theory(theory, g["theory"], ns=len(fit.s))
    np.testing.assert_allclose(chi2, g["chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g["lnl"], rtol=0, atol=CHI2_ATOL)
    xi = fit.theory_xi_batch(fit.s, g["mu"], P)
    np.testing.assert_allclose(xi, g["xi_smu"], rtol=RTOL, atol=ATOL)
    one = fit.theory_xi(*np.meshgrid(fit.s, g["mu"]),
                        {k: float(v[1]) for k, v in P.items()})
    np.testing.assert_allclose(one, g["xi_smu"][1], rtol=RTOL, atol=ATOL)


def test_likelihood_forms_and_fixed_covariance(fit, golden, boss_blocks):
    from victor_b200 import CCFFit
    g = golden("boss_forms")
    for form in ("gaussian", "hartlap", "percival", "sellentin"):
        like = {"form": form, "nmocks": 1000, "nparams": 4}
        lnl, chi2 = fit.log_likelihood_batch(g["params"], likelihood=like)
        np.testing.assert_allclose(chi2, g[f"{form}_chi2"], rtol=0, atol=CHI2_ATOL)
        np.testing.assert_allclose(lnl, g[f"{form}_lnl"], rtol=0, atol=CHI2_ATOL)
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    data["covariance_matrix"] = {"data_file": "data/boss_dr12_cmass/cmass_fixed_D_covariance.npz",
                                 "cov_key": "covmat", "fixed_beta": True}
    ffix = CCFFit(model, data)
    lnl, chi2 = ffix.log_likelihood_batch(g["params"])
    np.testing.assert_allclose(chi2, g["fixedcov_chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g["fixedcov_lnl"], rtol=0, atol=CHI2_ATOL)
    ffix.close()


def test_likelihood_interpolation_mode(fit, golden):
    g = golden("boss_variant_points")
    ok = np.isfinite(g["likelihood_interp_chi2"])
    lnl, chi2 = fit.log_likelihood_batch(g["params"][ok], beta_interpolation="likelihood")
    np.testing.assert_allclose(chi2, g["likelihood_interp_chi2"][ok], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g["likelihood_interp_lnl"][ok], rtol=0, atol=CHI2_ATOL)


def test_cobaya_block_astar(golden, repo_root):
    """boss_cobaya_config-style block: rescale_templates_independent_of_AP absent -> astar rescaling."""
    import yaml
    from victor_b200 import CCFFit
    g = golden("boss_cobaya_block")
    with open(f"{repo_root}/config/boss_cobaya_config.yaml") as fh:
        blk = yaml.full_load(fh)["likelihood"]["CCFLikelihood"]
    blk["model"]["dir"] = blk["data"]["dir"] = repo_root
    cc = CCFFit(blk["model"], blk["data"])
    P = {"fsigma8": g["params"][:, 0], "beta": g["params"][:, 1], "sigma_v": g["params"][:, 2],
         "epsilon": g["params"][:, 3], "alpha": 1.0, "astar": g["params"][:, 4]}
    lnl, chi2, theory = cc.log_likelihood_batch(P, return_theory=True)
    assert_theory(theory, g["theory"], ns=len(cc.s))
    np.testing.assert_allclose(chi2, g["chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g["lnl"], rtol=0, atol=CHI2_ATOL)
    cc.close()


def test_example_config_multipoles(example_block, golden):
    """Non-uniform knots, no beta dependence, poles 0/2/4 on a caller-supplied s grid."""
    from victor_b200 import CCFModel
    g = golden("example_points")
    m = CCFModel(copy.deepcopy(example_block))
    assert abs(m.iaH - float(g["iaH"])) < 1e-15
    P = {"fsigma8": g["params"][:, 0], "sigma_v": g["params"][:, 1], "epsilon": g["params"][:, 2]}
    th = m.theory_multipole_vector_batch(g["s"], P, poles=[0, 2, 4])
    assert_theory(th, g["streaming_theory"], ns=len(g["s"]))
    th2 = m.theory_multipole_vector_batch(m.r, P, poles=[0, 2])
    assert_theory(th2, g["streaming_theory_rgrid"], ns=len(m.r))
    m.close()


def test_batch_invariances(fit, golden):
    """Row order, batch split, block split and n=1 vs n=many give bit-identical rows."""
    g = golden("boss_streaming_points")
    P = g["params"]
    lnl, chi2, th = fit.log_likelihood_batch(P, return_theory=True)
    perm = np.random.default_rng(3).permutation(len(P))
    lnl_p, chi2_p, th_p = fit.log_likelihood_batch(P[perm], return_theory=True)
    assert np.array_equal(th_p, th[perm]) and np.array_equal(chi2_p, chi2[perm]) and np.array_equal(lnl_p, lnl[perm])
    eng, _ = fit._fit_engine({})
    eng.set_option("tiny", 0)          # (calls of <= 2 rows otherwise run k_small: same values, another summation order)
    l1, c1, t1 = fit.log_likelihood_batch(P[5:6], return_theory=True)
    eng.set_option("tiny", 1)
    assert np.array_equal(t1[0], th[5]) and c1[0] == chi2[5] and l1[0] == lnl[5]
    l1, c1, t1 = fit.log_likelihood_batch(P[5:6], return_theory=True)
    assert np.abs(t1[0] - th[5]).max() < 1e-14 and abs(c1[0] - chi2[5]) < 1e-9 and abs(l1[0] - lnl[5]) < 1e-9
    for nsplit in (1, 3, 30):
        eng.set_option("nsplit", nsplit)
        l2, c2, t2 = fit.log_likelihood_batch(P, return_theory=True)
        assert np.array_equal(t2, th) and np.array_equal(c2, chi2)
    eng.set_option("nsplit", 0)


def test_fused_likelihood_epilogue_is_bit_identical(fit):
    """Batch mode: chi2 / lnL from the epilogue of K1 (one block per row) equal the separate K2 launch bit
    for bit, for the general kernel (default) and the tuned one (option fuse = 2)."""
    rng = np.random.default_rng(11)
    n = 2048                                                   # >= 6 blocks per SM: one block per row
    P = np.column_stack([rng.uniform(0.05, 1.5, n), rng.uniform(0.1, 0.7, n), rng.uniform(100, 500, n),
                         rng.uniform(0.9, 1.1, n), rng.uniform(0.9, 1.1, n)])
    P[7, 1] = np.nan                                           # a failing row goes through the same guard
    for kw in ({}, {"rsd_model": "dispersion"}, {"rsd_model": "kaiser", "assume_isotropic": False}):
        eng, _ = fit._fit_engine(kw)
        out = {}
        # (the dispersion model runs on the tuned kernel, which has no fused form -- chi2 is < 1 % of its step:
        # the general kernel's epilogue is what is checked for it)
        eng.set_option("tuned", 0 if kw.get("rsd_model") == "dispersion" else 1)
        for fuse in (0, 2):
            eng.set_option("fuse", fuse)
            before = eng.launch_count()
            out[fuse] = fit.log_likelihood_batch(P, **kw)
            launches = eng.launch_count() - before
            assert launches == (1 if fuse else 2)
        eng.set_option("fuse", 1)
        assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
        assert out[0][0][7] == -np.inf and out[0][1][7] == np.inf
        lnl_t, chi2_t, _ = fit.log_likelihood_batch(P, return_theory=True, **kw)
        assert np.array_equal(lnl_t, out[0][0]) and np.array_equal(chi2_t, out[0][1])
        eng.set_option("tuned", 1)


def test_bucketed_chi2_is_bit_identical(fit, boss_blocks):
    """K2 of large batches groups the rows by covariance bracket (count, scatter, k_chi2_bucketed: both precision
    matrices of a group in shared memory, two rows per fetched element).  chi2 and lnL equal those of the row-by-row
    kernel ("bucket" 0) bit for bit -- on rows of every kind (beta on a grid value, outside the grid at both ends,
    NaN), for sizes around the tile and threshold edges, for every likelihood form, and with ONE fixed covariance."""
    from victor_b200 import CCFFit
    rng = np.random.default_rng(23)

    def table(n, grid):
        P = np.column_stack([rng.uniform(0.05, 1.5, n), rng.uniform(0.1, 0.7, n), rng.uniform(100, 500, n),
                             rng.uniform(0.9, 1.1, n), rng.uniform(0.9, 1.1, n)])
        P[3, 1] = np.nan
        P[4:36, 1] = grid[:32] if len(grid) >= 32 else grid[0]            # exactly on grid values
        P[40, 1], P[41, 1] = grid[0] - 0.01, grid[-1] + 0.01              # beyond both ends
        P[50:80, 1] = 0.3011                                              # one crowded bracket next to empty ones
        return P

    grid = np.asarray(fit.beta_covmat if hasattr(fit, "beta_covmat") else fit.beta_ccf, dtype=float)
    eng, _ = fit._fit_engine({})
    for n in (4096, 4097, 4127, 20000):
        P = table(n, grid)
        out = {}
        for bucket in (0, 1):
            eng.set_option("bucket", bucket)
            before = eng.launch_count()
            out[bucket] = fit.log_likelihood_batch(P)
            assert eng.launch_count() - before == (4 if bucket else 2)         # K1 + (count, scatter,) chi2
        eng.set_option("bucket", 1)
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]), n
        assert out[1][0][3] == -np.inf and out[1][1][3] == np.inf
        again = fit.log_likelihood_batch(P)                                    # whatever order the scatter produced
        assert np.array_equal(again[0], out[1][0]) and np.array_equal(again[1], out[1][1])
    P = table(6000, grid)
    for form in ("gaussian", "hartlap", "percival", "sellentin"):
        like = {"form": form, "nmocks": 1000, "nparams": 4}
        e2, _ = fit._fit_engine({"likelihood": like})
        res = {}
        for bucket in (0, 1):
            e2.set_option("bucket", bucket)
            res[bucket] = fit.log_likelihood_batch(P, likelihood=like)
        e2.set_option("bucket", 1)
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]), form
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    data["covariance_matrix"] = {"data_file": "data/boss_dr12_cmass/cmass_fixed_D_covariance.npz",
                                 "cov_key": "covmat", "fixed_beta": True}
    ffix = CCFFit(model, data)
    e3, _ = ffix._fit_engine({})
    res = {}
    for bucket in (0, 1):
        e3.set_option("bucket", bucket)
        res[bucket] = ffix.log_likelihood_batch(P)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    ffix.close()


def test_small_calls_replayed_as_graph(fit, golden):
    """MCMC-sized calls (host rows, n <= 256) go through a captured CUDA graph; changing n or an option
    rebuilds it; results equal the plain submissions bit for bit and the golden values.  (Calls of up to two
    rows run the one-launch k_small kernel, whose summation order differs in the last bits: "tiny" 0 switches it
    off here so that every call size goes through the batch kernels; test_small_row_kernel covers k_small.)"""
    g = golden("boss_streaming_points")
    P = g["params"]
    eng, _ = fit._fit_engine({})
    eng.set_option("tiny", 0)
    try:
        ref_l, ref_c = fit.log_likelihood_batch(P)                 # 80 rows: still the small path
        np.testing.assert_allclose(ref_c, g["chi2"], rtol=0, atol=CHI2_ATOL)
        for graph in (1, 0, 1):
            eng.set_option("graph", graph)
            for sl in (slice(0, 1), slice(3, 6), slice(0, 1), slice(0, 1), slice(0, 80), slice(79, 80)):
                l, c = fit.log_likelihood_batch(P[sl])
                assert np.array_equal(l, ref_l[sl]) and np.array_equal(c, ref_c[sl])
            before = eng.launch_count()
            for i in range(5):
                lnl, chi2 = fit.log_likelihood({"fsigma8": float(P[i, 0]), "beta": float(P[i, 1]), "sigma_v": float(P[i, 2]),
                                                "aperp": float(P[i, 3]), "apar": float(P[i, 4])})
                assert chi2 == ref_c[i] and lnl == ref_l[i]
            assert eng.launch_count() - before == 10               # K1 + K2 per call, graph or not
        eng.set_option("fast_math", 0)                             # a different kernel variant: the graph is rebuilt
        l0, c0 = fit.log_likelihood_batch(P[:2])
        eng.set_option("fast_math", 1)
        l1, c1 = fit.log_likelihood_batch(P[:2])
        assert np.array_equal(c1, ref_c[:2]) and np.allclose(c0, c1, rtol=0, atol=1e-8) and not np.array_equal(c0, c1)
    finally:
        eng.set_option("tiny", 1)


@pytest.mark.parametrize("kw,gname,prefix", [({}, "boss_streaming_points", ""),
                                             ({"rsd_model": "dispersion"}, "boss_variant_points", "dispersion_"),
                                             ({"assume_isotropic": False}, "boss_variant_points", "anisotropic_")])
def test_small_row_kernel(fit, golden, kw, gname, prefix):
    """Calls of one or two rows (an MCMC step) run k_small (k1_small.cuh): one launch, the velocity nodes of a
    (s, mu) pair over 16 lanes + a shuffle butterfly, chi-square by the last block to retire.  Its summation order is
    per-lane runs of consecutive nodes, then the butterfly -- not the batch kernel's single chain -- so it is held
    (a) to the goldens of the unmodified reference at the usual tolerances, (b) to the batch kernels within 1e-13 of
    the multipoles' scale / 1e-9 in chi-square, and (c) to itself bit for bit on repetition (tickets recycle)."""
    import torch
    from victor_b200.model import params_to_rows
    g = golden(gname)
    N = min(12, len(g["params"]))
    P = g["params"][:N]
    want_th, want_c, want_l = g[f"{prefix}theory"][:N], g[f"{prefix}chi2"][:N], g[f"{prefix}lnl"][:N]
    eng, _ = fit._fit_engine(kw)
    eng.set_option("tiny", 0)
    bl, bc, bt = fit.log_likelihood_batch(P, return_theory=True, **kw)              # batch kernels
    single = [fit.log_likelihood_batch(P[i:i + 1], **kw) for i in range(3)]
    eng.set_option("tiny", 1)
    try:
        for n in (1, 2):
            for lo in range(0, N - n + 1, n):
                before = eng.launch_count()
                l, c = fit.log_likelihood_batch(P[lo:lo + n], **kw)                  # staged host path (graph replay)
                assert eng.launch_count() - before == 1
                np.testing.assert_allclose(c, want_c[lo:lo + n], rtol=0, atol=CHI2_ATOL)
                np.testing.assert_allclose(l, want_l[lo:lo + n], rtol=0, atol=CHI2_ATOL)
                np.testing.assert_allclose(c, bc[lo:lo + n], rtol=0, atol=1e-9)
                l2, c2, t2 = fit.log_likelihood_batch(P[lo:lo + n], return_theory=True, **kw)   # general path, theory out
                assert np.array_equal(c2, c) and np.array_equal(l2, l)
                assert_theory(t2, want_th[lo:lo + n], ns=len(fit.s))
                scale = np.abs(bt[lo:lo + n]).reshape(n, -1, len(fit.s)).max(axis=2, keepdims=True)
                assert (np.abs(t2 - bt[lo:lo + n]).reshape(n, -1, len(fit.s)) / scale).max() < 1e-13
        for i in range(3):                                                           # same row, same bits, every time
            for _ in range(20):
                l, c = fit.log_likelihood_batch(P[i:i + 1], **kw)
                first = first if _ else (l, c)
                assert np.array_equal(l, first[0]) and np.array_equal(c, first[1])
            assert abs(c[0] - single[i][1][0]) < 1e-9
        # device-resident rows and results
        rows = torch.from_numpy(params_to_rows(P[:2])).cuda()
        ld, cd = fit.log_likelihood_device(rows, **kw)
        l4, c4 = fit.log_likelihood_batch(P[:2], **kw)
        assert np.array_equal(ld.cpu().numpy(), l4) and np.array_equal(cd.cpu().numpy(), c4)
        # a failing row still ends in the NaN guard
        bad = np.array(P[:2], copy=True)
        bad[1, 1] = np.nan
        l, c = fit.log_likelihood_batch(bad, **kw)
        assert l[1] == -np.inf and c[1] == np.inf and np.isfinite(l[0])
    finally:
        eng.set_option("tiny", 1)


def test_chunked_host_outputs_are_bit_identical(fit):
    """Bulk outputs bound for host memory go in row chunks with the device-to-host copies overlapping later chunks;
    any chunking gives the rows the single launch gives."""
    rng = np.random.default_rng(21)
    n = 5003
    P = np.column_stack([rng.uniform(0.05, 1.5, n), rng.uniform(0.2, 0.6, n), rng.uniform(100, 500, n),
                         rng.uniform(0.9, 1.1, n), rng.uniform(0.9, 1.1, n)])
    for kw in ({}, {"rsd_model": "dispersion"}):
        eng, _ = fit._fit_engine(kw)
        out = {}
        for chunks in (1, 5, 16):
            eng.set_option("chunks", chunks)
            out[chunks] = fit.log_likelihood_batch(P, return_theory=True, **kw)
        eng.set_option("chunks", 0)
        for chunks in (5, 16):
            for a, b in zip(out[1], out[chunks]):
                assert np.array_equal(a, b)
    s = np.linspace(5.0, 100.0, 24)
    mu = np.linspace(0, 1, 64)
    meng = fit._engine(fit._merged_options({}))
    ref = None
    for chunks in (1, 7):
        meng.set_option("chunks", chunks)
        xi = fit.theory_xi_batch(s, mu, P[:700])
        mult = fit.theory_multipoles_batch(s, P[:700], poles=[0, 2, 4], mu_nodes=64)
        if ref is None:
            ref = (xi, mult)
        else:
            assert np.array_equal(xi, ref[0]) and np.array_equal(mult, ref[1])
    meng.set_option("chunks", 0)


def test_against_oracle_fresh_points(fit, boss_blocks):
    """Seeded rows that are not in the golden files, checked against the CPU oracle."""
    from oracle.ccf_oracle import OracleFit
    model, data = boss_blocks
    orc = OracleFit(copy.deepcopy(model), copy.deepcopy(data))
    rng = np.random.default_rng(777)
    n = 12
    P = np.column_stack([rng.uniform(0.05, 1.5, n), rng.uniform(0.2, 0.6, n), rng.uniform(100, 500, n),
                         rng.uniform(0.9, 1.1, n), rng.uniform(0.9, 1.1, n)])
    lnl, chi2, th = fit.log_likelihood_batch(P, return_theory=True)
    for i in range(n):
        prm = dict(zip(("fsigma8", "beta", "sigma_v", "aperp", "apar"), map(float, P[i])))
        want = orc.theory_multipole_vector(orc.s, prm, orc.poles_s)
        assert_theory(th[i], want, ns=len(orc.s))
        wl, wc = orc.log_likelihood(prm)
        assert abs(chi2[i] - wc) < CHI2_ATOL and abs(lnl[i] - wl) < CHI2_ATOL


def test_nan_rows_follow_reference_convention(fit):
    P = np.array([[0.47, 0.37, 380.0, 1.0, 1.0], [np.nan, 0.37, 380.0, 1.0, 1.0],
                  [0.47, np.nan, 380.0, 1.0, 1.0]])
    lnl, chi2 = fit.log_likelihood_batch(P)
    assert np.isfinite(lnl[0]) and np.isfinite(chi2[0])
    assert lnl[1] == -np.inf and chi2[1] == np.inf
    assert lnl[2] == -np.inf and chi2[2] == np.inf


def test_failed_point_is_logged(fit, caplog):
    """A failing single-point call returns (-inf, inf) and reports it through logging, where the reference
    prints (ccf_fit.py:477-481)."""
    import logging
    with caplog.at_level(logging.WARNING, logger="victor_b200"):
        lnl, chi2 = fit.log_likelihood({"fsigma8": float("nan"), "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0})
    assert lnl == -np.inf and chi2 == np.inf
    assert any("Likelihood evaluation failed" in r.message for r in caplog.records)


def test_full_batch_properties(fit):
    """BASELINE size (65,536 rows): size-independent properties instead of a CPU re-computation."""
    from bench import synthetic_batch
    P = synthetic_batch(65536)
    lnl, chi2, th = fit.log_likelihood_batch(P, return_theory=True)
    assert np.all(np.isfinite(lnl)) and np.all(chi2 > 0)
    # first rows are the golden rows
    g = np.load(__file__.replace("test_gpu_parity.py", "golden/boss_streaming_points.npz"))
    assert_theory(th[:64], g["theory"][:64], ns=len(fit.s))
    np.testing.assert_allclose(chi2[:64], g["chi2"][:64], rtol=0, atol=CHI2_ATOL)
    # a strided re-evaluation in a different batch shape reproduces the same bits
    idx = np.arange(0, 65536, 257)
    l2, c2, t2 = fit.log_likelihood_batch(P[idx], return_theory=True)
    assert np.array_equal(t2, th[idx]) and np.array_equal(c2, chi2[idx]) and np.array_equal(l2, lnl[idx])
    # sellentin lnL is a monotone function of chi2 at fixed covariance normalisation: check the
    # closed form against the returned chi2 through the host-side covariance blend
    for i in idx[:8]:
        cov = fit.get_interpolated_covariance(P[i, 1])
        want = -1000 * np.log(1 + chi2[i] / 999) / 2 - 0.5 * np.linalg.slogdet(cov)[1]
        assert abs(want - lnl[i]) < 1e-8


# ---------------------------------------------------------------------------------------------
# general kernel: dispersion, kaiser, euclid_special, anisotropic input, from-data coordinates
# 
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kw", [("dispersion", {"rsd_model": "dispersion"}), ("kaiser", {"rsd_model": "kaiser"}),
                                     ("anisotropic", {"assume_isotropic": False})])
def test_variant_models_golden(fit, golden, name, kw):
    g = golden("boss_variant_points")
    eng, _ = fit._fit_engine(kw)
    for fast in (0, 1):          # libm arithmetic, then the hand-rolled rsqrt / reciprocal / exp (default)
        eng.set_option("fast_math", fast)
        lnl, chi2, theory = fit.log_likelihood_batch(g["params"], return_theory=True, **kw)
        assert_theory(theory, g[f"{name}_theory"], ns=len(fit.s))
        np.testing.assert_allclose(chi2, g[f"{name}_chi2"], rtol=0, atol=CHI2_ATOL)
        np.testing.assert_allclose(lnl, g[f"{name}_lnl"], rtol=0, atol=CHI2_ATOL)
    a = golden("boss_notebook_anchors")
    l0, c0 = fit.log_likelihood({"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0}, **kw)
    assert abs(c0 - float(a[f"{name}_chi2"])) < CHI2_ATOL and abs(l0 - float(a[f"{name}_lnl"])) < CHI2_ATOL


@pytest.mark.parametrize("name,kw", [
    ("euclid", {"rsd_model": "euclid_special"}),
    ("kaiser_noshift", {"rsd_model": "kaiser", "kaiser_coord_shift": False}),
    ("kaiser_approx", {"rsd_model": "kaiser", "kaiser_approximation": True}),
    ("kaiser_mq", {"rsd_model": "kaiser"}),
    ("aniso_dispersion", {"rsd_model": "dispersion", "assume_isotropic": False}),
    ("aniso_kaiser", {"rsd_model": "kaiser", "assume_isotropic": False})])
def test_more_variant_models_golden(fit, golden, name, kw):
    g = golden("boss_more_variants")
    P = dict(zip(("fsigma8", "beta", "sigma_v", "aperp", "apar"), g["params"].T))
    P.update(M=g["MQ"][:, 0], Q=g["MQ"][:, 1])
    lnl, chi2, theory = fit.log_likelihood_batch(P, return_theory=True, **kw)
    assert_theory(theory, g[f"{name}_theory"], ns=len(fit.s))
    np.testing.assert_allclose(chi2, g[f"{name}_chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g[f"{name}_lnl"], rtol=0, atol=CHI2_ATOL)


@pytest.mark.parametrize("aniso", [False, True])
def test_measured_model_from_data(boss_blocks, golden, aniso):
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    model["realspace_ccf"]["assume_isotropic"] = not aniso
    kind = "anisotropic" if aniso else "isotropic"
    data["covariance_matrix"]["data_file"] = f"data/boss_dr12_cmass/cmass_variable_{kind}_MD_covariance.npz"
    fm = CCFFit(model, data)
    if aniso:
        g = golden("boss_more_variants")
        cases = [("measured_aniso", {}), ("measured_aniso_dispersion", {"rsd_model": "dispersion"})]
        P = g["measured_params"]
    else:
        g = golden("boss_measured_model")
        cases = [(None, {})]
        P = g["params"]
    for name, kw in cases:
        pre = f"{name}_" if name else ""
        lnl, chi2, theory = fm.log_likelihood_batch(P, return_theory=True, **kw)
        assert_theory(theory, g[f"{pre}theory"], ns=len(fm.s))
        np.testing.assert_allclose(chi2, g[f"{pre}chi2"], rtol=0, atol=CHI2_ATOL)
        np.testing.assert_allclose(lnl, g[f"{pre}lnl"], rtol=0, atol=CHI2_ATOL)
    fm.close()


def test_example_config_other_models(example_block, golden):
    from victor_b200 import CCFModel
    g = golden("example_points")
    m = CCFModel(copy.deepcopy(example_block))
    P = {"fsigma8": g["params"][:, 0], "sigma_v": g["params"][:, 1], "epsilon": g["params"][:, 2]}
    for name in ("dispersion", "kaiser"):
        th = m.theory_multipole_vector_batch(g["s"], P, poles=[0, 2, 4], rsd_model=name)
        assert_theory(th, g[f"{name}_theory"], ns=len(g["s"]))
    m.close()


def test_general_kernel_agrees_with_tuned_on_streaming(fit, golden):
    """The anisotropic streaming kernels (general kernel, and the tuned kernel's n_ell = 2 variant) against the
    isotropic tuned kernel on the same rows: an anisotropic table with a zero quadrupole is the isotropic model."""
    g = golden("boss_streaming_points")
    eng_iso, _ = fit._fit_engine({})
    mt = eng_iso.model_tables
    import dataclasses
    from victor_b200.engine import Engine
    xi2 = np.zeros((2,) + mt.xi_tab.shape[1:])
    xi2[0] = mt.xi_tab[0]
    mt2 = dataclasses.replace(mt, n_ell=2, ells=np.array([0, 2], dtype=np.int32), xi_tab=xi2)
    eng = Engine(mt2, fit._fit_tables(fit._merged_options({})), device=None)
    from victor_b200.model import params_to_rows
    for tuned in (1, 0):
        eng.set_option("tuned", tuned)
        th2, c2, l2 = eng.likelihood(params_to_rows(g["params"]), want_theory=True)
        assert_theory(th2, g["theory"], ns=len(fit.s))
        np.testing.assert_allclose(c2, g["chi2"], rtol=0, atol=CHI2_ATOL)
    eng.close()


@pytest.mark.parametrize("kw", [{}, {"rsd_model": "dispersion"}])
def test_tuned_from_data_kernels_agree_with_general(boss_blocks, kw):
    """Real-space ccf measured from data (anisotropic measured model + MD covariance, ccf_model.py:675-679): xi is
    looked up at fiducial coordinates -- K1Cfg::kFromData of the tuned kernel against the general kernel, 2048 rows."""
    from bench import synthetic_batch
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    model["realspace_ccf"]["assume_isotropic"] = False
    data["covariance_matrix"]["data_file"] = "data/boss_dr12_cmass/cmass_variable_anisotropic_MD_covariance.npz"
    fm = CCFFit(model, data)
    rows = params_to_rows(synthetic_batch(65536)[20000:22048])
    eng, _ = fm._fit_engine(kw)
    out = {}
    for tuned in (1, 0):
        eng.set_option("tuned", tuned)
        before = eng.launch_count()
        out[tuned] = eng.likelihood(rows, want_theory=True)
    eng.set_option("tuned", 1)
    (th1, c1, l1), (th0, c0, l0) = out[1], out[0]
    for i in (3, 700):                                        # single rows: the one-launch kernel, same model
        before = eng.launch_count()
        ls, cs, ts = fm.log_likelihood_batch(rows[i:i + 1], return_theory=True, **kw)
        assert eng.launch_count() - before == 1
        assert abs(cs[0] - c1[i]) < 1e-8 and np.abs(ts[0] - th1[i]).max() < 1e-12
    ok = np.isfinite(l0) & np.isfinite(l1)
    assert ok.sum() >= len(rows) - 2
    scale = np.abs(th0[ok]).reshape(ok.sum(), 2, -1).max(axis=2, keepdims=True)
    err = (np.abs(th1[ok] - th0[ok]).reshape(ok.sum(), 2, -1) / scale).max(axis=(1, 2))
    if kw:   # the dispersion iteration amplifies last-bit differences in a few rows
        assert np.quantile(err, 0.99) < 1e-11 and err.max() < 1e-7, (np.quantile(err, 0.99), err.max())
    else:
        assert err.max() < 1e-11 and np.max(np.abs(c1[ok] - c0[ok])) < 1e-7
    fm.close()


@pytest.mark.parametrize("kw", [{"rsd_model": "dispersion"}, {"assume_isotropic": False},
                                {"rsd_model": "dispersion", "assume_isotropic": False}])
def test_tuned_wide_kernels_agree_with_general(fit, kw):
    """Dispersion / anisotropic streaming run on the tuned kernel (k1_streaming.cuh: kModel, kNEll); the general
    kernel computes the same model with another schedule (and the iteration written with 1/u).  2048 rows of
    the bench batch, both against each other far inside the parity tolerances, for every block split."""
    from bench import synthetic_batch
    from victor_b200.model import params_to_rows
    rows = params_to_rows(synthetic_batch(65536)[1000:3048])
    eng, _ = fit._fit_engine(kw)
    out = {}
    try:
        for tuned in (1, 0):
            eng.set_option("tuned", tuned)
            out[tuned] = eng.likelihood(rows, want_theory=True)
        eng.set_option("tuned", 1)
        th, c2, _ = eng.likelihood(rows[:300], want_theory=True)   # 300 rows: the s range is split over blocks
        np.testing.assert_array_equal(th, out[1][0][:300])
        np.testing.assert_array_equal(c2, out[1][1][:300])
    finally:
        eng.set_option("tuned", 1)
    th1, c1, l1 = out[1]
    th0, c0, l0 = out[0]
    scale = np.abs(th0).reshape(len(rows), 2, -1).max(axis=2, keepdims=True)
    err = np.abs(th1 - th0).reshape(len(rows), 2, -1) / scale
    assert err.max() < 1e-11, err.max()
    assert np.max(np.abs(c1 - c0)) < 1e-7 and np.max(np.abs(l1 - l0)) < 1e-7


def test_dense_grid_against_restatement(fit):
    """BASELINE configs[3]: 200 mu x 100 velocity nodes, poles 0/2/4.  The reference hard-codes its
    grids, so the CPU reference here is the numpy table walk-through (oracle/table_emul.py), itself
    pinned to the unmodified reference at the default sizes (tests/test_host_tables.py)."""
    from oracle import table_emul as E
    from victor_b200 import tables as T
    from victor_b200.model import params_to_rows
    from bench import synthetic_batch
    P = synthetic_batch(65536)[:6]
    kw = {"velocity_nodes": 100, "mu_nodes": 200}
    got = fit.theory_multipole_vector_batch(fit.s, P, poles=[0, 2, 4], **kw)
    mt = T.build_model_tables(fit, fit._merged_options(kw), nx=100)
    mu, W = T.mu_projection_weights([0, 2, 4], nmu=200)
    want, _ = E.theory_multipoles(mt, params_to_rows(P), np.asarray(fit.s, float), mu, W)
    assert_theory(got, want.reshape(len(P), -1), ns=len(fit.s))
    # and the denser quadrature moves the answer only at the level the default grid resolves
    base = fit.theory_multipole_vector_batch(fit.s, P, poles=[0, 2, 4])
    assert 1e-9 < np.abs(base - got).max() < 1e-4


def test_multi_device_fit_single_gpu(boss_blocks, golden):
    """MultiDeviceFit with every visible GPU: concatenated shard outputs == single-context output."""
    import torch
    from victor_b200 import CCFFit
    from victor_b200.batch import MultiDeviceFit
    model, data = boss_blocks
    devices = list(range(torch.cuda.device_count()))
    mf = MultiDeviceFit(lambda d: CCFFit(copy.deepcopy(model), copy.deepcopy(data), device=d), devices)
    g = golden("boss_streaming_points")
    lnl, chi2 = mf.log_likelihood_batch(g["params"])
    np.testing.assert_allclose(chi2, g["chi2"], rtol=0, atol=CHI2_ATOL)
    one = CCFFit(copy.deepcopy(model), copy.deepcopy(data), device=0)
    l1, c1 = one.log_likelihood_batch(g["params"])
    assert np.array_equal(c1, chi2) and np.array_equal(l1, lnl)     # bit-identical wherever a row lands
    mf.close()
    one.close()


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                                     ("aniso_streaming", {"assume_isotropic": False})])
def test_sigma_v_r_mu_template(boss_blocks, golden, name, kw):
    """3-key dispersion template: bicubic sigma_v(r, mu) with mu clamped below 0."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_sv2d_inputs.npz"
    model["velocity_pdf"]["dispersion"]["template_keys"] = ["rsv", "musv", "sigmav2d"]
    fm = CCFFit(model, data)
    g = golden("boss_sv2d")
    lnl, chi2, theory = fm.log_likelihood_batch(g["params"], return_theory=True, **kw)
    assert_theory(theory, g[f"{name}_theory"], ns=len(fm.s))
    np.testing.assert_allclose(chi2, g[f"{name}_chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g[f"{name}_lnl"], rtol=0, atol=CHI2_ATOL)
    fm.close()


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                                     ("kaiser", {"rsd_model": "kaiser"}), ("bias25", {"bias": 2.5})])
def test_linear_bias_matter_model(boss_blocks, golden, name, kw):
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    fm = CCFFit(model, data)
    g = golden("boss_linear_bias")
    lnl, chi2, theory = fm.log_likelihood_batch(g["params"], return_theory=True, **kw)
    assert_theory(theory, g[f"{name}_theory"], ns=len(fm.s))
    np.testing.assert_allclose(chi2, g[f"{name}_chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g[f"{name}_lnl"], rtol=0, atol=CHI2_ATOL)
    fm.close()


def test_linear_bias_from_data(boss_blocks, golden):
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    data["covariance_matrix"]["data_file"] = "data/boss_dr12_cmass/cmass_variable_isotropic_MD_covariance.npz"
    fm = CCFFit(model, data)
    g = golden("boss_linear_bias")
    lnl, chi2, theory = fm.log_likelihood_batch(g["measured_params"], return_theory=True)
    assert_theory(theory, g["measured_theory"], ns=len(fm.s))
    np.testing.assert_allclose(chi2, g["measured_chi2"], rtol=0, atol=CHI2_ATOL)
    fm.close()


def _check_fit(fm, rows, g, name, **kw):
    lnl, chi2, theory = fm.log_likelihood_batch(rows, return_theory=True, **kw)
    assert_theory(theory, g[f"{name}_theory"], ns=len(fm.s))
    np.testing.assert_allclose(chi2, g[f"{name}_chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g[f"{name}_lnl"], rtol=0, atol=CHI2_ATOL)


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("aniso_streaming", {"assume_isotropic": False}),
                                     ("aniso_dispersion", {"assume_isotropic": False, "rsd_model": "dispersion"})])
def test_rmu_format_input(boss_blocks, golden, name, kw):
    """Real-space ccf given as xi(r, mu) (ccf_model.py:154-181): converted at load, three real-space poles."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_rmu_inputs.npz"
    model["realspace_ccf"].update(format="rmu", ccf_keys=["r", "mu_rmu", "xi_rmu"])
    fm = CCFFit(model, data)
    g = golden("boss_rmu")
    _check_fit(fm, g["params"], g, name, **kw)
    fm.close()


def test_rmu_format_input_without_reconstruction(boss_blocks, golden):
    from victor_b200 import CCFModel
    model = copy.deepcopy(boss_blocks[0])
    model["input_model_data_file"] = "tests/golden/model_rmu_inputs.npz"
    model["realspace_ccf"].update(format="rmu", ccf_keys=["r", "mu_rmu", "xi_rmu_fixed"], reconstruction=False)
    cm = CCFModel(model)
    g = golden("boss_rmu")
    s = np.load("tests/golden/fixed_inputs_data.npz")["s"]
    P = g["params"][:3]
    prm = {"fsigma8": P[:, 0], "beta": P[:, 1], "sigma_v": P[:, 2], "aperp": P[:, 3], "apar": P[:, 4]}
    th = cm.theory_multipole_vector_batch(s, prm, [0, 2, 4], assume_isotropic=False)
    assert_theory(th, g["fixed_aniso_theory"], ns=len(s))
    cm.close()


def _velocity_params(g, av=False, bias=False, n=None):
    P = g["params"][:n]
    prm = {"fsigma8": P[:, 0], "beta": P[:, 1], "sigma_v": P[:, 2], "aperp": P[:, 3], "apar": P[:, 4]}
    if av:
        prm["Av"] = g["Av"][:n]
    if bias:
        prm["bias"] = g["bias"][:n]
    return prm


@pytest.mark.parametrize("name,kw", [("emp_streaming", {}), ("emp_dispersion", {"rsd_model": "dispersion"}),
                                     ("emp_kaiser", {"rsd_model": "kaiser"})])
def test_empirical_velocity_correction(boss_blocks, golden, name, kw):
    """velocity_pdf.mean.empirical_corr with Av among the parameters (ccf_model.py:451-459)."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["velocity_pdf"]["mean"]["empirical_corr"] = True
    fm = CCFFit(model, data)
    g = golden("boss_velocity_options")
    _check_fit(fm, _velocity_params(g, av=True), g, name, **kw)
    if name == "emp_dispersion":
        _check_fit(fm, _velocity_params(g, n=3), g, "emp_noAv_dispersion", **kw)
        # single-point call, reference style
        i = 3
        prm = {k: float(v[i]) for k, v in _velocity_params(g, av=True).items()}
        lnl, chi2 = fm.log_likelihood(prm, **kw)
        assert abs(chi2 - g["emp_dispersion_chi2"][i]) < CHI2_ATOL and abs(lnl - g["emp_dispersion_lnl"][i]) < CHI2_ATOL
    fm.close()


@pytest.mark.parametrize("name,kw", [("rowbias_streaming", {}), ("rowbias_dispersion", {"rsd_model": "dispersion"})])
def test_bias_given_with_the_parameters(boss_blocks, golden, name, kw):
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["matter_ccf"]["model"] = "linear_bias"
    fm = CCFFit(model, data)
    g = golden("boss_velocity_options")
    _check_fit(fm, _velocity_params(g, bias=True), g, name, **kw)
    fm.close()


def test_empirical_correction_with_linear_bias(boss_blocks, golden):
    from victor_b200 import CCFModel
    model = copy.deepcopy(boss_blocks[0])
    model["input_model_data_file"] = "tests/golden/fixed_inputs_model.npz"
    model["realspace_ccf"]["reconstruction"] = False
    model["matter_ccf"]["model"] = "linear_bias"
    model["velocity_pdf"]["mean"]["empirical_corr"] = True
    cm = CCFModel(model)
    g = golden("boss_velocity_options")
    s = np.load("tests/golden/fixed_inputs_data.npz")["s"]
    th = cm.theory_multipole_vector_batch(s, _velocity_params(g, av=True, bias=True), [0, 2], rsd_model="dispersion")
    assert_theory(th, g["emp_linbias_fixed_dispersion_theory"], ns=len(s))
    cm.close()


@pytest.mark.parametrize("name,kw", [("vtemplate_streaming", {}), ("vtemplate_dispersion", {"rsd_model": "dispersion"}),
                                     ("vtemplate_kaiser", {"rsd_model": "kaiser"})])
def test_velocity_template_mean_model(boss_blocks, golden, name, kw):
    """velocity_pdf.mean.model 'template' (ccf_model.py:227-246, 439-443, 483-488); the streaming case runs
    on the tuned kernel (only the per-row amplitude differs)."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/model_vtemplate_inputs.npz"
    model["velocity_pdf"]["mean"].update(model="template", template_fsigma8=0.45, z_sim=0.5,
                                         template_hubble_ratio=1.02, template_keys=["rvel", "vr_template"])
    fm = CCFFit(model, data)
    g = golden("boss_velocity_options")
    _check_fit(fm, _velocity_params(g), g, name, **kw)
    fm.close()


def test_two_dimensional_helpers(fit, golden):
    """theory_xi_2D (2500 scalar theory_xi calls in the reference, one pairwise launch here) and
    xi_2D_from_multipoles (ccf_model.py:862-934), compared through the returned interpolators."""
    g = golden("boss_helpers")
    p1 = {"fsigma8": 0.8, "beta": 0.45, "sigma_v": 250, "aperp": 1.03, "apar": 0.96}
    gx, gy = np.linspace(0.01, 85), np.linspace(-85, 85)
    f1 = fit.theory_xi_2D(dict(p1), rmax=85)
    np.testing.assert_allclose(f1(gx, gy), g["xi2d_grid"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(f1(g["qx"], g["qy"]), g["xi2d_q"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(f1(12.5, -40.0), g["xi2d_scalar"], rtol=RTOL, atol=ATOL)
    assert_theory(fit.theory_multipole_vector(fit.s, dict(p1), [0, 1, 2, 3, 4]), g["five_poles"], ns=30)
    assert_theory(fit.theory_multipole_vector(fit.s, dict(p1), [0, 2, 4, 6], rsd_model="dispersion"), g["even_four"], ns=30)
    f2 = fit.xi_2D_from_multipoles(dict(p1), rmax=85)
    np.testing.assert_allclose(f2(gx, gy), g["from_multipoles_grid"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(f2(g["qx"], g["qy"]), g["from_multipoles_q"], rtol=RTOL, atol=1e-12)
    f3 = fit.xi_2D_from_multipoles(dict(p1), rmax=60, rsd_model="dispersion")
    np.testing.assert_allclose(f3(g["qx"], g["qy"]), g["from_multipoles_disp60_q"], rtol=RTOL, atol=1e-12)


def test_pairwise_points_match_the_grid_kernel(fit):
    """vb200_theory_pairs against vb200_theory on the same points, batched, both K1 kernels."""
    rng = np.random.default_rng(5)
    P = {"fsigma8": rng.uniform(0.2, 1.2, 7), "beta": rng.uniform(0.2, 0.6, 7), "sigma_v": rng.uniform(150, 450, 7),
         "aperp": rng.uniform(0.95, 1.05, 7), "apar": rng.uniform(0.95, 1.05, 7)}
    s = np.sort(rng.uniform(1.0, 110.0, 300))
    mu = rng.uniform(-1, 1, 300)
    for kw in ({}, {"rsd_model": "dispersion"}, {"assume_isotropic": False}):
        pairs = fit.theory_xi_pairs_batch(s, mu, P, **kw)
        assert pairs.shape == (7, 300)
        order = np.argsort(mu)
        grid = fit.theory_xi_batch(s, mu[order], P, **kw)           # [n][nmu][ns]
        want = grid[:, np.argsort(order), np.arange(300)]
        np.testing.assert_array_equal(pairs, want)                  # same arithmetic per point: bit-identical


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"})])
def test_loader_options(boss_blocks, golden, name, kw):
    """simulation_number, integrated matter template, unfiltered dispersion template, non-default cosmology."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/loader_inputs_model.npz"
    model["cosmology"] = {"Omega_m": 0.29, "Omega_K": 0.01}
    model["realspace_ccf"].update(ccf_keys=["r", "monopole_sims", "quadrupole_sims"], simulation_number=2)
    model["matter_ccf"].update(template_keys=["rDelta", "Delta"], integrated=True)
    model["velocity_pdf"]["dispersion"]["filter"] = False
    data["redshift_space_ccf"].update(data_file="tests/golden/loader_inputs_data.npz",
                                      ccf_keys=["s", "monopole_sims", "quadrupole_sims"], simulation_number=1)
    fm = CCFFit(model, data)
    g = golden("boss_loader_options")
    _check_fit(fm, g["params"], g, name, **kw)
    fm.close()


def test_direct_model_calls(fit, golden):
    """Notebook-style calls: odd poles, bare-integer poles, fine s grid, theory_xi on
    unsorted meshgrid input (sorted / uniqued like the reference) and at negative mu."""
    g = golden("boss_misc_calls")
    prm = {"p0": {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0},
           "p1": {"fsigma8": 0.8, "beta": 0.45, "sigma_v": 250, "aperp": 1.03, "apar": 0.96}}
    for tag, p in prm.items():
        assert_theory(fit.theory_multipole_vector(fit.s, dict(p), [0, 1, 2]), g[f"{tag}_odd_012"], ns=30)
        assert_theory(fit.theory_multipole_vector(fit.s, dict(p), 1), g[f"{tag}_pole1"], ns=30)
        assert_theory(fit.theory_multipole_vector(g["s_fine"], dict(p), [0, 2, 4]), g[f"{tag}_fine_024"], ns=120)
        mp = fit.theory_multipoles(g["s_fine"], dict(p), poles=2)
        assert list(mp) == ["2"]
        assert_theory(mp["2"], g[f"{tag}_fine_bare2"], ns=120)
    S, M = np.meshgrid(g["xi_unsorted_s"], g["xi_unsorted_mu"])
    np.testing.assert_allclose(fit.theory_xi(S, M, dict(prm["p1"])), g["xi_unsorted"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(fit.theory_xi(fit.s, np.linspace(-1, 1, 11), dict(prm["p1"])), g["xi_negmu"],
                               rtol=RTOL, atol=ATOL)


def test_empty_and_large_batches(fit):
    """n = 0 returns empty arrays; a 300k-row table runs as one launch (bench batches are 65,536)."""
    lnl, chi2 = fit.log_likelihood_batch(np.empty((0, 5)))
    assert lnl.shape == (0,) and chi2.shape == (0,)
    from bench import synthetic_batch
    P = synthetic_batch(300000, seed=5)
    lnl, chi2 = fit.log_likelihood_batch(P)
    assert np.all(np.isfinite(lnl)) and chi2.min() > 0
    l2, c2 = fit.log_likelihood_batch(P[123456:123461])
    assert np.array_equal(c2, chi2[123456:123461]) and np.array_equal(l2, lnl[123456:123461])
    l3, c3 = fit.log_likelihood_batch(P[123456:123458])          # two rows: the one-launch k_small kernel
    assert np.abs(c3 - chi2[123456:123458]).max() < 1e-9 and np.abs(l3 - lnl[123456:123458]).max() < 1e-9


def test_hostile_rows_do_not_break_the_context(fit, golden):
    """Rows far outside any prior (zeros, negatives, huge values, inf, NaN in every column): every
    table index must stay in range -- a stray access would surface as a CUDA error here or corrupt the
    golden check that follows (compute-sanitizer is closed on this GPU pool, so this is the bounds test)."""
    rng = np.random.default_rng(99)
    base = np.array([0.47, 0.37, 380.0, 1.0, 1.0])
    rows = [base.copy()]
    for col in range(5):
        for val in (0.0, -1.0, 1e-300, 1e-8, 1e8, 1e300, np.inf, -np.inf, np.nan):
            r = base.copy()
            r[col] = val
            rows.append(r)
    rows += list(np.abs(rng.standard_cauchy((200, 5))) * np.array([1.0, 0.5, 400.0, 1.0, 1.0]))
    P = np.array(rows)
    for kw in ({}, {"rsd_model": "dispersion"}, {"rsd_model": "kaiser"}, {"assume_isotropic": False}):
        lnl, chi2 = fit.log_likelihood_batch(P, **kw)
        assert lnl.shape == (len(P),)
        bad = ~np.isfinite(lnl)
        assert np.all(chi2[bad] == np.inf) and np.all(lnl[bad] == -np.inf)      # ccf_fit.py:477-481
    g = golden("boss_streaming_points")
    lnl, chi2 = fit.log_likelihood_batch(g["params"])
    np.testing.assert_allclose(chi2, g["chi2"], rtol=0, atol=CHI2_ATOL)


def test_long_s_grids_split_over_blocks(fit):
    """A 1500-point s grid for many rows: the library splits the s range over blocks by itself and
    the rows agree with the same s values evaluated in short grids."""
    s = np.linspace(1.0, 150.0, 1500)
    from bench import synthetic_batch
    P = synthetic_batch(65536)[:1000]
    big = fit.theory_multipole_vector_batch(s, P, poles=[0, 2])
    assert big.shape == (1000, 3000) and np.all(np.isfinite(big))
    part = fit.theory_multipole_vector_batch(s[400:420], P[:7], poles=[0, 2])
    np.testing.assert_array_equal(part[:, :20], big[:7, 400:420])
    np.testing.assert_array_equal(part[:, 20:], big[:7, 1900:1920])


@pytest.mark.parametrize("name,kw", [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                                     ("gaussian", {"likelihood": {"form": "gaussian"}})])
def test_no_reconstruction_anywhere(boss_blocks, golden, name, kw):
    """Fixed real-space input, fixed data vector, one covariance matrix; beta is not a parameter."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    model["input_model_data_file"] = "tests/golden/fixed_inputs_model.npz"
    model["realspace_ccf"]["reconstruction"] = False
    data["redshift_space_ccf"].update(reconstruction=False, data_file="tests/golden/fixed_inputs_data.npz")
    data["covariance_matrix"] = {"data_file": "tests/golden/fixed_inputs_cov.npz", "cov_key": "covmat"}
    fm = CCFFit(model, data)
    g = golden("boss_fixed_everything")
    P = {"fsigma8": g["params"][:, 0], "sigma_v": g["params"][:, 2], "aperp": g["params"][:, 3], "apar": g["params"][:, 4]}
    lnl, chi2, theory = fm.log_likelihood_batch(P, return_theory=True, **kw)
    assert_theory(theory, g[f"{name}_theory"], ns=len(fm.s))
    np.testing.assert_allclose(chi2, g[f"{name}_chi2"], rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, g[f"{name}_lnl"], rtol=0, atol=CHI2_ATOL)
    one = {k: float(v[0]) for k, v in P.items()}
    l1, c1 = fm.log_likelihood(one, **kw)                 # single point without a 'beta' key
    assert abs(c1 - g[f"{name}_chi2"][0]) < CHI2_ATOL and abs(l1 - g[f"{name}_lnl"][0]) < CHI2_ATOL
    fm.close()


# ---- random rows of the WHOLE bench batch against the scipy oracle (not only its first rows or hand-picked edges)
_POOL_ORACLE = None


def _pool_init(root, kw):
    global _POOL_ORACLE
    import os
    import sys
    import warnings
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"
    warnings.filterwarnings("ignore")
    if root not in sys.path:
        sys.path.insert(0, root)
    import yaml
    from oracle.ccf_oracle import OracleFit
    with open(os.path.join(root, "config", "boss_config.yaml")) as fh:
        info = yaml.full_load(fh)
    info["model"]["dir"] = info["data"]["dir"] = root
    _POOL_ORACLE = (OracleFit(info["model"], info["data"]), kw)


def _pool_eval(row):
    orc, kw = _POOL_ORACLE
    prm = dict(zip(("fsigma8", "beta", "sigma_v", "aperp", "apar"), map(float, row)))
    th = orc.theory_multipole_vector(orc.s, dict(prm), orc.poles_s, **kw)
    lnl, chi2 = orc.log_likelihood(dict(prm), **kw)
    return th, chi2, lnl


@pytest.mark.parametrize("kw,n", [({}, 256), ({"rsd_model": "dispersion"}, 128), ({"assume_isotropic": False}, 128),
                                  ({"rsd_model": "kaiser"}, 64)])
def test_random_rows_of_the_whole_batch_against_the_scipy_oracle(fit, repo_root, kw, n):
    """`n` rows drawn at random from all 65,536 rows of the bench batch, each evaluated by oracle/ccf_oracle.py
    (the scipy restatement pinned to the unmodified reference) in a process pool, against the CUDA path.  This
    comparison does not share the product's packed tables, unlike the C table walk."""
    import multiprocessing as mp
    import os
    from bench import synthetic_batch
    batch = synthetic_batch(65536)
    pick = np.sort(np.random.default_rng(2026).choice(len(batch), size=n, replace=False))
    assert pick[-1] > 60000 and pick[0] < 5000                       # spread over the whole batch
    rows = batch[pick]
    with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, 32), initializer=_pool_init,
                                      initargs=(repo_root, kw)) as pool:
        res = pool.map(_pool_eval, list(rows), chunksize=2)
    want_th = np.array([r[0] for r in res])
    want_c2 = np.array([r[1] for r in res])
    want_ll = np.array([r[2] for r in res])
    lnl, chi2, th = fit.log_likelihood_batch(rows, return_theory=True, **kw)
    assert_theory(th, want_th, ns=len(fit.s))
    np.testing.assert_allclose(chi2, want_c2, rtol=0, atol=CHI2_ATOL)
    np.testing.assert_allclose(lnl, want_ll, rtol=0, atol=CHI2_ATOL)


def test_device_resident_likelihood(fit, golden):
    """CCFFit.log_likelihood_device: rows and results stay on the GPU (torch tensors as buffers only)."""
    import torch
    from victor_b200.model import params_to_rows
    g = golden("boss_streaming_points")
    lnl_h, chi2_h = fit.log_likelihood_batch(g["params"])
    lnl_d, chi2_d = fit.log_likelihood_device(g["params"])
    assert lnl_d.is_cuda and chi2_d.dtype == torch.float64
    assert np.array_equal(lnl_d.cpu().numpy(), lnl_h) and np.array_equal(chi2_d.cpu().numpy(), chi2_h)
    rows = torch.from_numpy(params_to_rows(g["params"])).to(lnl_d.device)
    slot = torch.full((2, len(rows) + 5), float("nan"), dtype=torch.float64, device=rows.device)
    a, b = fit.log_likelihood_device(rows, out=slot, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(a.cpu().numpy(), lnl_h) and np.array_equal(slot[1, :len(rows)].cpu().numpy(), chi2_h)
    assert torch.isnan(slot[:, len(rows):]).all()
    with pytest.raises(ValueError):
        fit.log_likelihood_device(rows.float())
    from victor_b200.batch import likelihood_sharded
    l2, c2, (lo, hi) = likelihood_sharded(fit, params_to_rows(g["params"]))      # no process group: one slice
    assert (lo, hi) == (0, len(lnl_h)) and np.array_equal(l2, lnl_h) and np.array_equal(c2, chi2_h)


def test_small_row_kernel_general_epilogue(boss_blocks, tmp_path):
    """k_small with a data vector longer than 64 (three multipoles, p = 90): the last block falls back from the
    staged-matrix epilogue to block_chi2.  Builder-made inputs (the BOSS data with the quadrupole repeated as a
    hexadecapole and a block covariance); checked against the batch kernels and the C table walk."""
    from oracle.table_walk import TableWalk
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    src = np.load(f"{data['dir']}/{data['redshift_space_ccf']['data_file']}")
    keys = data["redshift_space_ccf"]["ccf_keys"]
    arrays = {k: src[k] for k in src.files}
    arrays["hexadecapole_copy"] = 0.5 * src[keys[2]]
    np.savez(tmp_path / "data3.npz", **arrays)
    cov = np.load(f"{data['dir']}/{data['covariance_matrix']['data_file']}")
    ckey = data["covariance_matrix"]["cov_key"]
    c60 = cov[ckey]
    c90 = np.zeros((c60.shape[0], 90, 90))
    c90[:, :60, :60] = c60
    c90[:, 60:, 60:] = c60[:, 30:, 30:] * 0.5
    carr = {k: cov[k] for k in cov.files}
    carr[ckey] = c90
    np.savez(tmp_path / "cov3.npz", **carr)
    data["dir"] = str(tmp_path)
    data["redshift_space_ccf"]["data_file"] = "data3.npz"
    data["redshift_space_ccf"]["ccf_keys"] = list(keys) + ["hexadecapole_copy"]
    data["covariance_matrix"]["data_file"] = "cov3.npz"
    fit3 = CCFFit(model, data)
    assert len(fit3.poles_s) == 3
    P = np.array([[0.47, 0.37, 380.0, 1.0, 1.0], [0.9, 0.52, 210.0, 1.05, 0.95], [0.2, 0.25, 450.0, 0.93, 1.08]])
    eng, _ = fit3._fit_engine({})
    eng.set_option("tiny", 0)
    bl, bc = fit3.log_likelihood_batch(P)
    eng.set_option("tiny", 1)
    wt, wc, wl = TableWalk(fit3).likelihood(params_to_rows(P), want_theory=True)
    for i in range(len(P)):
        before = eng.launch_count()
        l, c, t = fit3.log_likelihood_batch(P[i:i + 1], return_theory=True)
        assert eng.launch_count() - before == 1
        assert abs(c[0] - bc[i]) < 1e-9 and abs(l[0] - bl[i]) < 1e-9
        assert abs(c[0] - wc[i]) < CHI2_ATOL and abs(l[0] - wl[i]) < CHI2_ATOL
        assert_theory(t, wt[i:i + 1], ns=len(fit3.s))
    fit3.close()


def test_hostile_single_rows_through_k_small(fit, golden):
    """The one-launch kernel with rows far outside any prior, one call each: every block must still draw its ticket
    (a hang here would mean a block left before the counter), results follow the NaN convention, and the context
    keeps working -- the golden check at the end would show a corrupted ticket counter or scratch row."""
    base = np.array([0.47, 0.37, 380.0, 1.0, 1.0])
    rows = []
    for col in range(5):
        for val in (0.0, -1.0, 1e-300, 1e8, 1e300, np.inf, -np.inf, np.nan):
            r = base.copy()
            r[col] = val
            rows.append(r)
    for kw in ({}, {"rsd_model": "dispersion"}, {"assume_isotropic": False}):
        eng, _ = fit._fit_engine(kw)
        for r in rows:
            before = eng.launch_count()
            lnl, chi2 = fit.log_likelihood_batch(r[None, :], **kw)
            assert eng.launch_count() - before == 1
            if not np.isfinite(lnl[0]):
                assert lnl[0] == -np.inf and chi2[0] == np.inf                   # ccf_fit.py:477-481
        two = fit.log_likelihood_batch(np.array([rows[5], base]), **kw)          # a failing and a good row together
        good = fit.log_likelihood_batch(base[None, :], **kw)
        assert two[0][1] == good[0][0] and two[1][1] == good[1][0]
    g = golden("boss_streaming_points")
    for i in (0, 17, 63):
        lnl, chi2 = fit.log_likelihood_batch(g["params"][i:i + 1])
        assert abs(chi2[0] - g["chi2"][i]) < CHI2_ATOL and abs(lnl[0] - g["lnl"][i]) < CHI2_ATOL
