#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz from the UNMODIFIED reference.

Runs in the dev container only (needs /root/reference).  Imports the reference through
``oracle/refshim.py`` and records, at full float64 precision, the outputs of its public
API on seeded inputs.  The committed .npz files pin the oracle (``oracle/ccf_oracle.py``)
and, through it and directly, the CUDA path.

    python oracle/make_golden.py            # rewrites every tests/golden/*.npz

Inputs are the reference's own HDF5 files and YAML configs, read where they lie.
"""
import copy
import os
import sys

import numpy as np
import scipy
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import refshim  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
SEED = 20251018  # same stream as bench.py / SURVEY 8(d)
PARAM_COLS = ("fsigma8", "beta", "sigma_v", "aperp", "apar")


def synthetic_batch(n, seed=SEED):
    """Prior-box parameter rows, columns PARAM_COLS."""
    rng = np.random.default_rng(seed)
    P = np.empty((n, 5))
    P[:, 0] = rng.uniform(0.05, 1.5, n)
    P[:, 1] = rng.uniform(0.2, 0.6, n)
    P[:, 2] = rng.uniform(100.0, 500.0, n)
    P[:, 3] = rng.uniform(0.9, 1.1, n)
    P[:, 4] = rng.uniform(0.9, 1.1, n)
    return P


def row_to_params(row):
    return {k: float(v) for k, v in zip(PARAM_COLS, row)}


def meta():
    return dict(scipy_version=scipy.__version__, numpy_version=np.__version__,
                reference="seshnadathur/victor 0.1.4 (unmodified, via oracle/refshim.py)")


def boss_blocks(config="config/boss_config.yaml"):
    with open(os.path.join(REF, config)) as fh:
        info = yaml.full_load(fh)
    if "likelihood" in info:  # cobaya-style file
        blk = info["likelihood"]["CCFLikelihood"]
        model, data = blk["model"], blk["data"]
    else:
        model, data = info["model"], info["data"]
    model["dir"] = data["dir"] = REF
    return model, data


def edge_rows(beta_grid):
    """Hand-picked rows: grid nodes, outside-grid beta, prior corners (SURVEY 4 tier 2)."""
    rows = [
        [0.47, 0.37, 380.0, 1.0, 1.0],
        [0.47, float(beta_grid[5]), 380.0, 1.0, 1.0],      # exactly on a grid node
        [0.47, float(beta_grid[0]), 380.0, 1.0, 1.0],      # first node
        [0.47, float(beta_grid[-1]), 380.0, 1.0, 1.0],     # last node
        [0.47, 0.12, 380.0, 1.0, 1.0],                     # below the grid (PCHIP extrapolates)
        [0.47, 0.70, 380.0, 1.0, 1.0],                     # above the grid
        [0.47, float(beta_grid[7]) + 1e-9, 380.0, 1.0, 1.0],
        [0.47, float(beta_grid[7]) - 1e-9, 380.0, 1.0, 1.0],
        [0.05, 0.2, 100.0, 0.8, 1.2],
        [1.5, 0.6, 500.0, 1.2, 0.8],
        [1.5, 0.2, 100.0, 1.2, 1.2],
        [0.05, 0.6, 500.0, 0.8, 0.8],
        [0.47, 0.37, 60.0, 1.0, 1.0],                      # narrow pdf: under-resolved integral
        [0.47, 0.37, 800.0, 1.0, 1.0],                     # wide pdf: r -> 0 crossings
        [0.0, 0.37, 380.0, 1.0, 1.0],                      # no coherent outflow
        [0.47, 0.37, 380.0, 1.02, 0.97],
    ]
    return np.array(rows)


def run_points(ccf, P, **kw):
    n = len(P)
    p = len(ccf.s) * len(ccf.poles_s)
    theory = np.empty((n, p))
    chi2 = np.empty(n)
    lnl = np.empty(n)
    for i, row in enumerate(P):
        params = row_to_params(row)
        theory[i] = ccf.theory_multipole_vector(ccf.s, params, ccf.poles_s, **kw)
        lnl[i], chi2[i] = ccf.log_likelihood(params, **kw)
    return theory, chi2, lnl


def golden_boss(victor):
    model, data = boss_blocks()
    ccf = victor.CCFFit(copy.deepcopy(model), copy.deepcopy(data))

    # (1) notebook cell 22 anchors: epsilon-style parameters and the option variants
    p0 = {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0}
    variants = [("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                ("kaiser", {"rsd_model": "kaiser"}), ("anisotropic", {"assume_isotropic": False}),
                ("likelihood_interp", {"beta_interpolation": "likelihood"})]
    anchors = {}
    for name, kw in variants:
        lnl, c2 = ccf.log_likelihood(dict(p0), **kw)
        anchors[f"{name}_lnl"], anchors[f"{name}_chi2"] = lnl, c2
        if "beta_interpolation" not in kw:
            anchors[f"{name}_theory"] = ccf.theory_multipole_vector(ccf.s, dict(p0), ccf.poles_s, **kw)
    cov = ccf.get_interpolated_covariance(0.37)
    anchors["slogdet_cov"] = np.linalg.slogdet(cov)[1]
    anchors["data_vector"] = ccf.multipole_datavector(0.37)
    np.savez(os.path.join(OUT, "boss_notebook_anchors.npz"), **anchors, **meta())

    # (2) host-table level quantities
    r31 = np.append([0.01], ccf.r)
    tables = dict(iaH=ccf.iaH, r=ccf.r, s=ccf.s, beta=ccf.beta, sv_rmu=ccf.sv_rmu,
                  r_for_sv=ccf.r_for_sv, mu_for_sv=ccf.mu_for_sv,
                  delta_r31=ccf.delta(r31), Delta_r31=ccf.integrated_delta(r31),
                  icov_first=ccf.icov[0], icov_last=ccf.icov[-1],
                  xi_r_beta037=ccf.get_interpolated_real_multipoles(0.37))
    np.savez(os.path.join(OUT, "boss_tables.npz"), **tables, **meta())

    # (3) seeded batch rows (the first rows of the bench batch) + edge rows, streaming
    P = np.vstack([synthetic_batch(65536)[:64], edge_rows(ccf.beta)])
    theory, chi2, lnl = run_points(ccf, P)
    np.savez(os.path.join(OUT, "boss_streaming_points.npz"), params=P, theory=theory, chi2=chi2,
             lnl=lnl, param_cols=np.array(PARAM_COLS), **meta())

    # (4) epsilon / alpha parameterisation and theory_xi on the model grid
    eps_rows = np.array([[0.47, 0.37, 380.0, 1.0, 1.0], [0.6, 0.45, 300.0, 0.95, 1.01],
                         [0.3, 0.25, 450.0, 1.08, 0.98]])  # fsigma8, beta, sigma_v, epsilon, alpha
    eth, ec2, elnl, exi = [], [], [], []
    mu = np.linspace(0, 1, 100)
    for row in eps_rows:
        pr = dict(fsigma8=row[0], beta=row[1], sigma_v=row[2], epsilon=row[3], alpha=row[4])
        eth.append(ccf.theory_multipole_vector(ccf.s, dict(pr), ccf.poles_s))
        a, b = ccf.log_likelihood(dict(pr))
        elnl.append(a)
        ec2.append(b)
        exi.append(ccf.theory_xi(*np.meshgrid(ccf.s, mu), dict(pr)))
    np.savez(os.path.join(OUT, "boss_epsilon_points.npz"), params=eps_rows, theory=np.array(eth),
             chi2=np.array(ec2), lnl=np.array(elnl), xi_smu=np.array(exi), mu=mu, **meta())

    # (5) next-row variants on a few seeded rows
    Pv = np.vstack([P[:6], edge_rows(ccf.beta)[[0, 4, 5, 8, 9]]])
    out = dict(params=Pv)
    for name, kw in variants[1:4]:
        th, c2, ll = run_points(ccf, Pv, **kw)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = th, c2, ll
    c2l, lll = [], []
    for row in Pv:
        if not (ccf.beta_ccf[0] < row[1] <= ccf.beta_ccf[-1]):
            c2l.append(np.nan)
            lll.append(np.nan)
            continue
        a, b = ccf.log_likelihood(row_to_params(row), beta_interpolation="likelihood")
        lll.append(a)
        c2l.append(b)
    out["likelihood_interp_chi2"], out["likelihood_interp_lnl"] = np.array(c2l), np.array(lll)
    np.savez(os.path.join(OUT, "boss_variant_points.npz"), **out, **meta())

    # (6) likelihood forms and the fixed-covariance path
    forms = {}
    for form in ("gaussian", "hartlap", "percival", "sellentin"):
        like = {"form": form, "nmocks": 1000, "nparams": 4}
        vals = [ccf.log_likelihood(row_to_params(r), likelihood=like) for r in P[:4]]
        forms[f"{form}_lnl"] = np.array([v[0] for v in vals])
        forms[f"{form}_chi2"] = np.array([v[1] for v in vals])
    dfix = copy.deepcopy(data)
    dfix["covariance_matrix"] = {
        "data_file": "data/BOSS_DR12_CMASS_data/CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_fixed_D_covariance.hdf5",
        "cov_key": "covmat", "fixed_beta": True}
    cfix = victor.CCFFit(copy.deepcopy(model), dfix)
    vals = [cfix.log_likelihood(row_to_params(r)) for r in P[:4]]
    forms["fixedcov_lnl"] = np.array([v[0] for v in vals])
    forms["fixedcov_chi2"] = np.array([v[1] for v in vals])
    forms["params"] = P[:4]
    np.savez(os.path.join(OUT, "boss_forms.npz"), **forms, **meta())

    # (7) cobaya-config likelihood block (velocity_independent_of_AP defaults True, astar)
    cm, cd = boss_blocks("config/boss_cobaya_config.yaml")
    cc = victor.CCFFit(cm, cd)
    rows = np.array([[0.47, 0.37, 380.0, 1.0, 1.0], [0.55, 0.41, 350.0, 1.03, 1.0],
                     [0.40, 0.33, 420.0, 0.97, 1.02]])  # fsigma8, beta, sigma_v, epsilon, astar
    th, c2, ll = [], [], []
    for row in rows:
        pr = dict(fsigma8=row[0], beta=row[1], sigma_v=row[2], epsilon=row[3], alpha=1, astar=row[4],
                  b=1.9, Av=0, M=1, Q=1)
        th.append(cc.theory_multipole_vector(cc.s, dict(pr), cc.poles_s))
        a, b = cc.log_likelihood(dict(pr))
        ll.append(a)
        c2.append(b)
    np.savez(os.path.join(OUT, "boss_cobaya_block.npz"), params=rows, theory=np.array(th),
             chi2=np.array(c2), lnl=np.array(ll), **meta())

    # (8) measured real-space model + MD covariance ("from_data" coordinates), anisotropic
    mm = copy.deepcopy(model)
    mm["input_model_data_file"] = ("data/BOSS_DR12_CMASS_data/"
                                   "CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_measured_model.hdf5")
    mm["realspace_ccf"]["from_data"] = True
    dm = copy.deepcopy(data)
    dm["covariance_matrix"]["data_file"] = (
        "data/BOSS_DR12_CMASS_data/"
        "CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_variable_isotropic_MD_covariance.hdf5")
    cmd = victor.CCFFit(mm, dm)
    Pm = P[:4].copy()
    Pm[:, 1] = np.clip(Pm[:, 1], 0.25, 0.55)
    th, c2, ll = run_points(cmd, Pm)
    np.savez(os.path.join(OUT, "boss_measured_model.npz"), params=Pm, theory=th, chi2=c2, lnl=ll,
             beta_covmat=cmd.beta_covmat, **meta())


def golden_more(victor):
    """Option combinations beyond notebook cell 22: euclid_special, kaiser without the coordinate
    shift / in the linear approximation (with M, Q), anisotropic input under the dispersion and
    kaiser models, anisotropic input with from-data coordinates (measured model file)."""
    model, data = boss_blocks()
    ccf = victor.CCFFit(copy.deepcopy(model), copy.deepcopy(data))
    P = np.vstack([synthetic_batch(65536)[:4], edge_rows(ccf.beta)[[0, 9]]])
    MQ = np.array([[1.0, 1.0], [0.9, 1.1], [1.05, 0.8], [1.2, 1.0], [1.0, 1.3], [0.95, 0.9]])
    out = dict(params=P, MQ=MQ)
    cases = [("euclid", {"rsd_model": "euclid_special"}),
             ("kaiser_noshift", {"rsd_model": "kaiser", "kaiser_coord_shift": False}),
             ("kaiser_approx", {"rsd_model": "kaiser", "kaiser_approximation": True}),
             ("kaiser_mq", {"rsd_model": "kaiser"}),
             ("aniso_dispersion", {"rsd_model": "dispersion", "assume_isotropic": False}),
             ("aniso_kaiser", {"rsd_model": "kaiser", "assume_isotropic": False})]
    for name, kw in cases:
        th, c2, ll = [], [], []
        for row, mq in zip(P, MQ):
            prm = row_to_params(row)
            prm.update(M=float(mq[0]), Q=float(mq[1]))
            th.append(ccf.theory_multipole_vector(ccf.s, dict(prm), ccf.poles_s, **kw))
            a, b = ccf.log_likelihood(dict(prm), **kw)
            ll.append(a)
            c2.append(b)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = np.array(th), np.array(c2), np.array(ll)
    mm = copy.deepcopy(model)
    mm["input_model_data_file"] = ("data/BOSS_DR12_CMASS_data/"
                                   "CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_measured_model.hdf5")
    mm["realspace_ccf"]["from_data"] = True
    mm["realspace_ccf"]["assume_isotropic"] = False
    dm = copy.deepcopy(data)
    dm["covariance_matrix"]["data_file"] = (
        "data/BOSS_DR12_CMASS_data/"
        "CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_variable_anisotropic_MD_covariance.hdf5")
    cmd = victor.CCFFit(mm, dm)
    Pm = P[:4].copy()
    Pm[:, 1] = np.clip(Pm[:, 1], 0.25, 0.55)
    for name, kw in (("measured_aniso", {}), ("measured_aniso_dispersion", {"rsd_model": "dispersion"})):
        th, c2, ll = run_points(cmd, Pm, **kw)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = th, c2, ll
    out["measured_params"] = Pm
    np.savez(os.path.join(OUT, "boss_more_variants.npz"), **out, **meta())


def golden_sv2d(victor):
    """sigma_v(r, mu) dispersion template (3 template keys): no shipped file has one, so the BOSS
    model file is extended with a synthetic anisotropic template.  The inputs are saved next to
    the outputs (tests/golden/model_sv2d_inputs.npz) so the product reads the very same arrays."""
    import tempfile
    model, data = boss_blocks()
    from victor_b200.io_hdf5 import read_hdf5
    src = read_hdf5(os.path.join(REF, model["input_model_data_file"]))
    rng = np.random.default_rng(SEED + 7)
    musv = np.concatenate([[0.0], np.sort(rng.uniform(0.05, 0.95, 8)), [1.0]])       # non-uniform, 10 knots
    base = src["sigmav"]
    sv2d = base[:, None] * (1 + 0.15 * musv[None, :] ** 2 - 0.05 * musv[None, :]) \
        * (1 + 0.01 * rng.standard_normal((len(base), len(musv))))
    inputs = dict(src)
    inputs["musv"] = musv
    inputs["sigmav2d"] = sv2d
    np.savez(os.path.join(OUT, "model_sv2d_inputs.npz"), **inputs)
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "model_sv2d.npy"), inputs, allow_pickle=True)
    mm = copy.deepcopy(model)
    mm["dir"] = tmp
    mm["input_model_data_file"] = "model_sv2d.npy"
    mm["velocity_pdf"]["dispersion"]["template_keys"] = ["rsv", "musv", "sigmav2d"]
    ccf = victor.CCFFit(mm, copy.deepcopy(data))
    P = np.vstack([synthetic_batch(65536)[:5], edge_rows(ccf.beta)[[0, 8, 13]]])
    out = dict(params=P, sv_rmu=ccf.sv_rmu, mu_for_sv=ccf.mu_for_sv)
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                     ("aniso_streaming", {"assume_isotropic": False})):
        th, c2, ll = run_points(ccf, P, **kw)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = th, c2, ll
    np.savez(os.path.join(OUT, "boss_sv2d.npz"), **out, **meta())


def golden_linear_bias(victor):
    """matter_ccf model 'linear_bias' (ccf_model.py:358-370): delta and Delta from the real-space
    monopole itself; growth term fsigma8 / sigma8_template, or beta * bias with from-data input."""
    model, data = boss_blocks()
    mm = copy.deepcopy(model)
    mm["matter_ccf"]["model"] = "linear_bias"
    ccf = victor.CCFFit(mm, copy.deepcopy(data))
    P = np.vstack([synthetic_batch(65536)[:5], edge_rows(ccf.beta)[[0, 4, 9]]])
    out = dict(params=P)
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}), ("kaiser", {"rsd_model": "kaiser"}),
                     ("bias25", {"bias": 2.5})):
        th, c2, ll = run_points(ccf, P, **kw)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = th, c2, ll
    md = copy.deepcopy(mm)
    md["input_model_data_file"] = ("data/BOSS_DR12_CMASS_data/"
                                   "CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_measured_model.hdf5")
    md["realspace_ccf"]["from_data"] = True
    dm = copy.deepcopy(data)
    dm["covariance_matrix"]["data_file"] = (
        "data/BOSS_DR12_CMASS_data/"
        "CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_variable_isotropic_MD_covariance.hdf5")
    cmd = victor.CCFFit(md, dm)
    Pm = P[:4].copy()
    Pm[:, 1] = np.clip(Pm[:, 1], 0.25, 0.55)
    th, c2, ll = run_points(cmd, Pm)
    out["measured_params"] = Pm
    out["measured_theory"], out["measured_chi2"], out["measured_lnl"] = th, c2, ll
    np.savez(os.path.join(OUT, "boss_linear_bias.npz"), **out, **meta())


def golden_misc(victor):
    """Direct model calls the notebooks make: odd multipoles (mu grid [-1, 1]), a bare
    integer ``poles``, arbitrary s grids, theory_xi on unsorted / meshgrid inputs."""
    model, data = boss_blocks()
    ccf = victor.CCFFit(copy.deepcopy(model), copy.deepcopy(data))
    p0 = {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0}
    p1 = {"fsigma8": 0.8, "beta": 0.45, "sigma_v": 250, "aperp": 1.03, "apar": 0.96}
    s_fine = np.linspace(0.01, 120, 120)       # usage-demo cell 12
    out = dict(s_fine=s_fine)
    for tag, prm in (("p0", p0), ("p1", p1)):
        out[f"{tag}_odd_012"] = ccf.theory_multipole_vector(ccf.s, dict(prm), [0, 1, 2])
        out[f"{tag}_pole1"] = ccf.theory_multipole_vector(ccf.s, dict(prm), 1)
        out[f"{tag}_fine_024"] = ccf.theory_multipole_vector(s_fine, dict(prm), [0, 2, 4])
        mp = ccf.theory_multipoles(s_fine, dict(prm), poles=2)
        out[f"{tag}_fine_bare2"] = mp["2"]
    mu_u = np.array([0.9, 0.1, 0.5, 0.3, 0.7])
    s_u = np.array([40.0, 10.0, 25.0, 80.0])
    S, M = np.meshgrid(s_u, mu_u)
    out["xi_unsorted"] = ccf.theory_xi(S, M, dict(p1))
    out["xi_unsorted_s"], out["xi_unsorted_mu"] = s_u, mu_u
    out["xi_negmu"] = ccf.theory_xi(ccf.s, np.linspace(-1, 1, 11), dict(p1))
    np.savez(os.path.join(OUT, "boss_misc_calls.npz"), **out, **meta())


def golden_fixed(victor):
    """No reconstruction anywhere: 1-D real-space multipoles, 1-D data multipoles, one covariance
    matrix (ccf_model.py:105-111, ccf_fit.py:60-64, 130-133).  No shipped file is laid out like that, so
    row 12 of the BOSS tables is written out as 1-D arrays; inputs saved next to the outputs."""
    import tempfile
    from victor_b200.io_hdf5 import read_hdf5
    model, data = boss_blocks()
    msrc = read_hdf5(os.path.join(REF, model["input_model_data_file"]))
    dsrc = read_hdf5(os.path.join(REF, data["redshift_space_ccf"]["data_file"]))
    csrc = read_hdf5(os.path.join(REF, "data/BOSS_DR12_CMASS_data/"
                                       "CMASS_zobovVoids_reconRs10_0.43z0.7_medianRvcut_fixed_D_covariance.hdf5"))
    minp = {k: v for k, v in msrc.items() if k != "beta"}
    minp["monopole"], minp["quadrupole"] = msrc["monopole"][12], msrc["quadrupole"][12]
    dinp = {"s": dsrc["s"], "monopole": dsrc["monopole"][12], "quadrupole": dsrc["quadrupole"][12]}
    np.savez(os.path.join(OUT, "fixed_inputs_model.npz"), **minp)
    np.savez(os.path.join(OUT, "fixed_inputs_data.npz"), **dinp)
    np.savez(os.path.join(OUT, "fixed_inputs_cov.npz"), covmat=csrc["covmat"])
    tmp = tempfile.mkdtemp()
    for name, arrs in (("m.npy", minp), ("d.npy", dinp), ("c.npy", {"covmat": csrc["covmat"]})):
        np.save(os.path.join(tmp, name), arrs, allow_pickle=True)
    mm, dd = copy.deepcopy(model), copy.deepcopy(data)
    mm["dir"] = dd["dir"] = tmp
    mm["input_model_data_file"] = "m.npy"
    mm["realspace_ccf"]["reconstruction"] = False
    dd["redshift_space_ccf"].update(reconstruction=False, data_file="d.npy")
    dd["covariance_matrix"] = {"data_file": "c.npy", "cov_key": "covmat"}
    ccf = victor.CCFFit(mm, dd)
    P = np.vstack([synthetic_batch(65536)[:5], edge_rows(ccf.r * 0 + 0.4)[[0, 8, 9]]])
    out = dict(params=P)
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                     ("gaussian", {"likelihood": {"form": "gaussian"}})):
        th, c2, ll = [], [], []
        for row in P:
            prm = row_to_params(row)
            del prm["beta"]                       # no beta anywhere: the reference must not need it
            th.append(ccf.theory_multipole_vector(ccf.s, dict(prm), ccf.poles_s, **kw))
            a, b = ccf.log_likelihood(dict(prm), **kw)
            ll.append(a)
            c2.append(b)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = np.array(th), np.array(c2), np.array(ll)
    np.savez(os.path.join(OUT, "boss_fixed_everything.npz"), **out, **meta())


def golden_rmu(victor):
    """Real-space ccf given as xi(r, mu) ('rmu' format, ccf_model.py:154-181): converted to multipoles
    0, 2, 4 at load through a linear interp2d.  No shipped file has that layout, so xi_0 + xi_2 L_2 of
    the BOSS model (plus a small mu^4 term) is tabulated on a non-uniform mu grid, with and without
    the reconstruction axis; inputs saved next to the outputs (tests/golden/model_rmu_inputs.npz)."""
    import tempfile
    from scipy.special import legendre
    from victor_b200.io_hdf5 import read_hdf5
    model, data = boss_blocks()
    src = read_hdf5(os.path.join(REF, model["input_model_data_file"]))
    rng = np.random.default_rng(SEED + 11)
    mu = np.concatenate([[0.0], np.sort(rng.uniform(0.03, 0.97, 18)), [1.0]])
    xi = (src["monopole"][:, :, None] + src["quadrupole"][:, :, None] * legendre(2)(mu)[None, None, :]
          + 0.02 * src["monopole"][:, :, None] * legendre(4)(mu)[None, None, :])
    inputs = dict(src)
    inputs["mu_rmu"] = mu
    inputs["xi_rmu"] = xi                     # (nbeta, nr, nmu)
    inputs["xi_rmu_fixed"] = xi[12]           # (nr, nmu)
    np.savez(os.path.join(OUT, "model_rmu_inputs.npz"), **inputs)
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "model_rmu.npy"), inputs, allow_pickle=True)
    mm = copy.deepcopy(model)
    mm["dir"] = tmp
    mm["input_model_data_file"] = "model_rmu.npy"
    mm["realspace_ccf"].update(format="rmu", ccf_keys=["r", "mu_rmu", "xi_rmu"])
    ccf = victor.CCFFit(mm, copy.deepcopy(data))
    P = np.vstack([synthetic_batch(65536)[:4], edge_rows(ccf.beta)[[0, 9]]])
    out = dict(params=P, poles_r=ccf.poles_r)
    for ell in ccf.poles_r:
        out[f"real_multipole_{ell}"] = ccf.real_multipoles[f"{ell}"]
    for name, kw in (("streaming", {}), ("aniso_streaming", {"assume_isotropic": False}),
                     ("aniso_dispersion", {"assume_isotropic": False, "rsd_model": "dispersion"})):
        th, c2, ll = run_points(ccf, P, **kw)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = th, c2, ll
    # without the reconstruction axis (CCFModel only: the data still depend on beta)
    mf = copy.deepcopy(mm)
    mf["realspace_ccf"].update(reconstruction=False, ccf_keys=["r", "mu_rmu", "xi_rmu_fixed"])
    cm = victor.CCFModel(mf)
    for ell in cm.poles_r:
        out[f"fixed_real_multipole_{ell}"] = cm.real_multipoles[f"{ell}"]
    out["fixed_aniso_theory"] = np.array([cm.theory_multipole_vector(ccf.s, row_to_params(row), [0, 2, 4],
                                                                     assume_isotropic=False) for row in P[:3]])
    np.savez(os.path.join(OUT, "boss_rmu.npz"), **out, **meta())


def golden_velocity(victor):
    """Mean-velocity options of the velocity pdf (ccf_model.py:385-492): the empirical correction
    (1 + Av delta(r)) with Av a parameter, a bias given with the parameters (linear_bias matter model),
    and the 'template' mean model (a velocity profile read from file; none is shipped, so one is derived
    from the BOSS matter template and saved in tests/golden/model_vtemplate_inputs.npz)."""
    import tempfile
    from victor_b200.io_hdf5 import read_hdf5
    model, data = boss_blocks()
    P = np.vstack([synthetic_batch(65536)[:4], edge_rows(np.linspace(0.16, 0.65, 31))[[0, 9]]])
    Av = np.array([0.0, 0.8, -0.5, 1.5, 0.3, -1.0])
    out = dict(params=P, Av=Av)

    # (a) empirical correction, matter template
    me = copy.deepcopy(model)
    me["velocity_pdf"]["mean"]["empirical_corr"] = True
    ccf = victor.CCFFit(me, copy.deepcopy(data))
    for name, kw in (("emp_streaming", {}), ("emp_dispersion", {"rsd_model": "dispersion"}),
                     ("emp_kaiser", {"rsd_model": "kaiser"})):
        th, c2, ll = [], [], []
        for row, av in zip(P, Av):
            prm = row_to_params(row)
            prm["Av"] = float(av)
            th.append(ccf.theory_multipole_vector(ccf.s, dict(prm), ccf.poles_s, **kw))
            a, b = ccf.log_likelihood(dict(prm), **kw)
            ll.append(a)
            c2.append(b)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = np.array(th), np.array(c2), np.array(ll)
    # Av absent from the parameters: defaults to 0, but the slope still comes from the finite-difference branch
    th, c2, ll = run_points(ccf, P[:3], rsd_model="dispersion")
    out["emp_noAv_dispersion_theory"], out["emp_noAv_dispersion_chi2"], out["emp_noAv_dispersion_lnl"] = th, c2, ll

    # (b) bias given with the parameters, linear_bias matter model (beta-dependent monopole)
    mb = copy.deepcopy(model)
    mb["matter_ccf"]["model"] = "linear_bias"
    cb = victor.CCFFit(mb, copy.deepcopy(data))
    bias = np.array([1.9, 2.3, 1.6, 2.0, 2.8, 1.2])
    out["bias"] = bias
    for name, kw in (("rowbias_streaming", {}), ("rowbias_dispersion", {"rsd_model": "dispersion"})):
        th, c2, ll = [], [], []
        for row, b_ in zip(P, bias):
            prm = row_to_params(row)
            prm["bias"] = float(b_)
            th.append(cb.theory_multipole_vector(cb.s, dict(prm), cb.poles_s, **kw))
            a, b = cb.log_likelihood(dict(prm), **kw)
            ll.append(a)
            c2.append(b)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = np.array(th), np.array(c2), np.array(ll)

    # (c) empirical correction + linear_bias without reconstruction (fixed real-space input), with row bias
    fx = np.load(os.path.join(OUT, "fixed_inputs_model.npz"))
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "m.npy"), {k: fx[k] for k in fx.files}, allow_pickle=True)
    mf = copy.deepcopy(model)
    mf["dir"] = tmp
    mf["input_model_data_file"] = "m.npy"
    mf["realspace_ccf"]["reconstruction"] = False
    mf["matter_ccf"]["model"] = "linear_bias"
    mf["velocity_pdf"]["mean"]["empirical_corr"] = True
    cm = victor.CCFModel(mf)
    s = read_hdf5(os.path.join(REF, data["redshift_space_ccf"]["data_file"]))["s"]
    th = []
    for row, av, b_ in zip(P, Av, bias):
        prm = row_to_params(row)
        prm.update(Av=float(av), bias=float(b_))
        th.append(cm.theory_multipole_vector(s, dict(prm), [0, 2], rsd_model="dispersion"))
    out["emp_linbias_fixed_dispersion_theory"] = np.array(th)

    # (d) velocity template
    src = read_hdf5(os.path.join(REF, model["input_model_data_file"]))
    rv = np.linspace(0.5, 130.0, 40)
    base = victor.CCFModel(copy.deepcopy(model))
    vr_t = -0.52 * rv * base.integrated_delta(rv) / (3 * base.iaH) * (1 + 0.1 * np.sin(rv / 15.0))
    inputs = dict(src)
    inputs["rvel"], inputs["vr_template"] = rv, vr_t
    np.savez(os.path.join(OUT, "model_vtemplate_inputs.npz"), **inputs)
    np.save(os.path.join(tmp, "model_vt.npy"), inputs, allow_pickle=True)
    mv = copy.deepcopy(model)
    mv["dir"] = tmp
    mv["input_model_data_file"] = "model_vt.npy"
    mv["velocity_pdf"]["mean"].update(model="template", template_fsigma8=0.45, z_sim=0.5,
                                      template_hubble_ratio=1.02, template_keys=["rvel", "vr_template"])
    cv = victor.CCFFit(mv, copy.deepcopy(data))
    for name, kw in (("vtemplate_streaming", {}), ("vtemplate_dispersion", {"rsd_model": "dispersion"}),
                     ("vtemplate_kaiser", {"rsd_model": "kaiser"})):
        th, c2, ll = run_points(cv, P, **kw)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = th, c2, ll
    np.savez(os.path.join(OUT, "boss_velocity_options.npz"), **out, **meta())


def golden_helpers(victor):
    """Callers either side of the path that the notebooks use: delta_profiles / velocity_terms (host
    helpers), theory_xi_2D (2500 scalar theory_xi calls in the reference: ~6 minutes here) and
    xi_2D_from_multipoles.  The interp2d objects are recorded through their grid values and a few
    off-grid evaluations."""
    model, data = boss_blocks()
    ccf = victor.CCFFit(copy.deepcopy(model), copy.deepcopy(data))
    p1 = {"fsigma8": 0.8, "beta": 0.45, "sigma_v": 250, "aperp": 1.03, "apar": 0.96}
    out = {}
    r = np.asarray(ccf.r, float)
    for tag, kw in (("template", {}), ("linear_bias", {"matter_model": "linear_bias"})):
        d, D = ccf.delta_profiles(r, dict(p1), **kw)
        out[f"delta_{tag}"], out[f"Delta_{tag}"] = d, D
    for tag, kw, extra in (("linear", {}, {}), ("empirical", {"empirical_corr": True}, {"Av": 0.7}),
                           ("linear_bias", {"matter_model": "linear_bias"}, {"bias": 2.2})):
        prm = dict(p1)
        prm.update(extra)
        vr, dvr = ccf.velocity_terms(r, prm, **kw)
        out[f"vr_{tag}"], out[f"dvr_{tag}"] = vr, dvr
    qx = np.array([0.5, 7.3, 22.0, 41.7, 84.0])
    qy = np.array([-80.0, -33.3, -2.0, 0.0, 11.1, 60.5])
    out["qx"], out["qy"] = qx, qy
    f2 = ccf.xi_2D_from_multipoles(dict(p1), rmax=85)
    out["from_multipoles_grid"] = f2(np.linspace(0.01, 85), np.linspace(-85, 85))
    out["from_multipoles_q"] = f2(qx, qy)
    f2 = ccf.xi_2D_from_multipoles(dict(p1), rmax=60, rsd_model="dispersion")
    out["from_multipoles_disp60_q"] = f2(qx, qy)
    out["corrmat_037"] = ccf.correlation_matrix(0.37)
    out["errors_037"] = ccf.diagonal_errors(0.37)
    out["data_multipoles_041"] = ccf.get_interpolated_redshift_multipoles(0.41)
    out["precision_below_grid"] = ccf.get_interpolated_precision(0.10)
    out["covariance_on_node"] = ccf.get_interpolated_covariance(float(ccf.beta_covmat[9]))
    out["five_poles"] = ccf.theory_multipole_vector(ccf.s, dict(p1), [0, 1, 2, 3, 4])        # more than three at once
    out["even_four"] = ccf.theory_multipole_vector(ccf.s, dict(p1), [0, 2, 4, 6], rsd_model="dispersion")
    f1 = ccf.theory_xi_2D(dict(p1), rmax=85)
    out["xi2d_grid"] = f1(np.linspace(0.01, 85), np.linspace(-85, 85))
    out["xi2d_q"] = f1(qx, qy)
    out["xi2d_scalar"] = f1(12.5, -40.0)
    np.savez(os.path.join(OUT, "boss_helpers.npz"), **out, **meta())


def golden_loader_options(victor):
    """Input-side options of the loaders that no shipped configuration uses (ccf_model.py:99-297, ccf_fit.py:44-164):
    `simulation_number` (files holding several realisations), an *integrated* matter template, an unfiltered
    dispersion template, a non-default cosmology.  Inputs derived from the BOSS files and saved next to the outputs."""
    import tempfile
    from victor_b200.io_hdf5 import read_hdf5
    model, data = boss_blocks()
    msrc = read_hdf5(os.path.join(REF, model["input_model_data_file"]))
    dsrc = read_hdf5(os.path.join(REF, data["redshift_space_ccf"]["data_file"]))
    base = victor.CCFModel(copy.deepcopy(model))
    rng = np.random.default_rng(SEED + 23)
    wob = 1 + 0.03 * rng.standard_normal((3, 1, 1))
    minp = dict(msrc)
    minp["monopole_sims"] = msrc["monopole"][None] * wob
    minp["quadrupole_sims"] = msrc["quadrupole"][None] * wob[::-1]
    minp["rDelta"] = np.linspace(1.5, 140.0, 45)
    minp["Delta"] = base.integrated_delta(minp["rDelta"])
    dinp = dict(dsrc)
    dinp["monopole_sims"] = dsrc["monopole"][None] * (1 + 0.01 * rng.standard_normal((3, 1, 1)))
    dinp["quadrupole_sims"] = dsrc["quadrupole"][None] * (1 + 0.01 * rng.standard_normal((3, 1, 1)))
    np.savez(os.path.join(OUT, "loader_inputs_model.npz"), **minp)
    np.savez(os.path.join(OUT, "loader_inputs_data.npz"), **dinp)
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "m.npy"), minp, allow_pickle=True)
    np.save(os.path.join(tmp, "d.npy"), dinp, allow_pickle=True)
    mm, dd = copy.deepcopy(model), copy.deepcopy(data)
    mm["dir"] = tmp
    mm["input_model_data_file"] = "m.npy"
    mm["cosmology"] = {"Omega_m": 0.29, "Omega_K": 0.01}
    mm["realspace_ccf"].update(ccf_keys=["r", "monopole_sims", "quadrupole_sims"], simulation_number=2)
    mm["matter_ccf"].update(template_keys=["rDelta", "Delta"], integrated=True)
    mm["velocity_pdf"]["dispersion"]["filter"] = False
    dd["redshift_space_ccf"].update(data_file=os.path.join(tmp, "d.npy"), ccf_keys=["s", "monopole_sims", "quadrupole_sims"],
                                    simulation_number=1)
    ccf = victor.CCFFit(mm, dd)
    P = np.vstack([synthetic_batch(65536)[:4], edge_rows(ccf.beta)[[0, 9]]])
    r31 = np.append([0.01], ccf.r)
    out = dict(params=P, iaH=ccf.iaH, sv_rmu=ccf.sv_rmu, delta_r31=ccf.delta(r31), Delta_r31=ccf.integrated_delta(r31),
               real_mono=ccf.real_multipoles["0"], data_mono=ccf.redshift_multipoles["0"])
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"})):
        th, c2, ll = run_points(ccf, P, **kw)
        out[f"{name}_theory"], out[f"{name}_chi2"], out[f"{name}_lnl"] = th, c2, ll
    np.savez(os.path.join(OUT, "boss_loader_options.npz"), **out, **meta())


def golden_example(victor):
    with open(os.path.join(REF, "config/example_model_input.yaml")) as fh:
        model = yaml.full_load(fh)["model"]
    model["dir"] = REF
    ccf = victor.CCFModel(model)
    s = np.linspace(0.01, 3, 100)
    rows = np.array([[0.47, 7.0, 1.0], [0.3, 4.0, 0.97], [0.7, 10.0, 1.04]])  # fsigma8, sigma_v, epsilon
    out = dict(params=rows, s=s, iaH=ccf.iaH, r=ccf.r, sv_rmu=ccf.sv_rmu, r_for_sv=ccf.r_for_sv)
    for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"}),
                     ("kaiser", {"rsd_model": "kaiser"})):
        th = []
        for row in rows:
            pr = dict(fsigma8=row[0], sigma_v=row[1], epsilon=row[2])
            th.append(ccf.theory_multipole_vector(s, pr, [0, 2, 4], **kw))
        out[f"{name}_theory"] = np.array(th)
    # model-grid evaluation as well (s = r grid, poles 0,2)
    th = []
    for row in rows:
        pr = dict(fsigma8=row[0], sigma_v=row[1], epsilon=row[2])
        th.append(ccf.theory_multipole_vector(ccf.r, pr, [0, 2]))
    out["streaming_theory_rgrid"] = np.array(th)
    np.savez(os.path.join(OUT, "example_points.npz"), **out, **meta())


def golden_example_fit(victor):
    """BASELINE.json configs[0], fit half: chi-square and lnL of the example void model against a data vector
    and covariance.  The reference ships NO data or covariance file for its example configuration
    (config/example_model_input.yaml has a model block only), so the INPUTS here are builder-made and flagged
    ``non_reference_inputs`` in the fixture: a seeded synthetic data vector (the reference's own theory at a
    fiducial point plus Gaussian noise) and a synthetic positive-definite covariance, written as .npy dict files
    -- a format the unmodified reference reads itself.  The OUTPUTS are the unmodified reference's
    CCFFit.chi_squared / log_likelihood on those inputs (ccf_fit.py:325-354, 356-483: fixed data vector, fixed
    covariance, so no log-det term)."""
    import tempfile
    with open(os.path.join(REF, "config/example_model_input.yaml")) as fh:
        model = yaml.full_load(fh)["model"]
    model["dir"] = REF
    base = victor.CCFModel(copy.deepcopy(model))
    rng = np.random.default_rng(SEED + 7)
    s = np.linspace(0.2, 2.6, 20)
    fid = dict(fsigma8=0.47, sigma_v=7.0, epsilon=1.0)
    truth = base.theory_multipole_vector(s, dict(fid), [0, 2])
    p = len(truth)
    sig = 0.02 * (1 + np.abs(truth))
    idx = np.arange(p)
    corr = 0.4 ** np.abs(idx[:, None] - idx[None, :])
    corr[:20, 20:] *= 0.5
    corr[20:, :20] *= 0.5
    cov = corr * np.outer(sig, sig)
    dvec = truth + np.linalg.cholesky(cov) @ rng.standard_normal(p)
    rows = np.array([[0.47, 7.0, 1.0], [0.3, 4.0, 0.97], [0.7, 10.0, 1.04], [0.55, 6.0, 1.0], [0.2, 9.0, 1.1]])
    out = dict(params=rows, s=s, xi0=dvec[:20], xi2=dvec[20:], covmat=cov, non_reference_inputs=True)
    with tempfile.TemporaryDirectory() as tmp:
        np.save(os.path.join(tmp, "example_data.npy"), {"s": s, "xi0": dvec[:20], "xi2": dvec[20:]}, allow_pickle=True)
        np.save(os.path.join(tmp, "example_cov.npy"), {"covmat": cov}, allow_pickle=True)
        data = {"dir": tmp,
                "redshift_space_ccf": {"data_file": "example_data.npy", "reconstruction": False, "format": "multipoles",
                                       "ccf_keys": ["s", "xi0", "xi2"]},
                "covariance_matrix": {"data_file": "example_cov.npy", "cov_key": "covmat"}}
        for form, like in (("gaussian", {"form": "Gaussian"}), ("sellentin", {"form": "Sellentin", "nmocks": 500}),
                           ("hartlap", {"form": "Hartlap", "nmocks": 500})):
            d = copy.deepcopy(data)
            d["likelihood"] = like
            fit = victor.CCFFit(copy.deepcopy(model), d)
            for name, kw in (("streaming", {}), ("dispersion", {"rsd_model": "dispersion"})):
                if form != "gaussian" and name != "streaming":
                    continue
                th, c2, ll = [], [], []
                for row in rows:
                    pr = dict(fsigma8=row[0], sigma_v=row[1], epsilon=row[2])
                    th.append(fit.theory_multipole_vector(fit.s, dict(pr), fit.poles_s, **kw))
                    l, c = fit.log_likelihood(dict(pr), **kw)
                    c_only, covm = fit.chi_squared(dict(pr), **kw)
                    assert c_only == c and covm.shape == (p, p)
                    c2.append(c)
                    ll.append(l)
                out[f"{form}_{name}_theory"], out[f"{form}_{name}_chi2"], out[f"{form}_{name}_lnl"] = (
                    np.array(th), np.array(c2), np.array(ll))
    np.savez(os.path.join(OUT, "example_fit.npz"), **out, **meta())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    v = refshim.install(REF)
    which = sys.argv[1:] or ["boss", "more", "sv2d", "linear_bias", "misc", "fixed", "rmu", "velocity", "helpers", "loader", "example",
                             "example_fit"]
    if "boss" in which:
        golden_boss(v)
    if "more" in which:
        golden_more(v)
    if "sv2d" in which:
        golden_sv2d(v)
    if "linear_bias" in which:
        golden_linear_bias(v)
    if "misc" in which:
        golden_misc(v)
    if "fixed" in which:
        golden_fixed(v)
    if "rmu" in which:
        golden_rmu(v)
    if "velocity" in which:
        golden_velocity(v)
    if "helpers" in which:
        golden_helpers(v)
    if "loader" in which:
        golden_loader_options(v)
    if "example" in which:
        golden_example(v)
    if "example_fit" in which:
        golden_example_fit(v)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
