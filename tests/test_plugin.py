"""The cobaya plugin surface (victor/likelihoods/CCFLikelihood.py:6-42): construction from the
model / data blocks or from a config file, class defaults from the yaml next to the class,
``calculate`` filling ``state``.  cobaya itself is not in this image; the stand-in base class
gives the plugin the part of the protocol it uses."""
import copy
import os

import numpy as np
import pytest
import yaml


def cobaya_block(repo_root):
    with open(os.path.join(repo_root, "config", "boss_cobaya_config.yaml")) as fh:
        info = yaml.full_load(fh)
    blk = info["likelihood"]["CCFLikelihood"]
    blk["model"]["dir"] = blk["data"]["dir"] = repo_root
    return info, blk


def test_defaults_and_initialize_from_blocks(repo_root):
    from victor_b200.likelihoods import CCFLikelihood
    info, blk = cobaya_block(repo_root)
    like = CCFLikelihood({"model": copy.deepcopy(blk["model"]), "data": copy.deepcopy(blk["data"])})
    assert like.get_can_provide_params() == ["fsigma8"]
    # class defaults of the reference yaml (CCFLikelihood.yaml:9-41)
    assert like.params["sigma_v"] == 380 and like.params["astar"] == 1 and like.params["b"] == 1.9
    assert like.params["chi2_ccf_correct"]["derived"] is True
    assert like.ccf.poles_s.tolist() == [0, 2] and like.ccf.covmat.shape == (31, 60, 60)
    # rescale_templates_independent_of_AP absent in this block -> the reference default True
    assert like.ccf.model["velocity_independent_of_AP"] is True
    # the sampler block of the cobaya file names parameters the plugin understands
    sampled = set(info["params"])
    assert {"fsigma8", "beta", "sigma_v", "epsilon"} <= sampled


def test_initialize_from_config_file(repo_root, monkeypatch):
    from victor_b200.likelihoods import CCFLikelihood
    monkeypatch.chdir(repo_root)                      # config paths are relative, as in the reference
    like = CCFLikelihood({"config_file": "config/boss_config.yaml"})
    assert like.ccf.model["velocity_independent_of_AP"] is False
    assert like.ccf.fit_options["likelihood"]["form"] == "sellentin"
    with pytest.raises(KeyError):
        CCFLikelihood({"config_file": "config/does_not_exist.yaml"})


@pytest.mark.gpu
def test_calculate_fills_state(repo_root, golden):
    from victor_b200.likelihoods import CCFLikelihood
    _, blk = cobaya_block(repo_root)
    like = CCFLikelihood({"model": blk["model"], "data": blk["data"]})
    g = golden("boss_cobaya_block")
    for i, row in enumerate(g["params"]):
        eps = float(row[3])
        values = dict(fsigma8=float(row[0]), beta=float(row[1]), sigma_v=float(row[2]), epsilon=eps, alpha=1,
                      aperp=eps ** (1 / 3), apar=eps ** (-2 / 3), astar=float(row[4]), b=1.9, Av=0, M=1, Q=1)
        state = {}
        like.calculate(state, want_derived=True, **values)
        assert abs(state["logp"] - g["lnl"][i]) < 1e-6
        assert abs(state["derived"]["chi2_ccf_correct"] - g["chi2"][i]) < 1e-6
    like.ccf.close()


@pytest.mark.gpu
def test_metropolis_chain_reproducible(repo_root):
    """Config 5 in miniature: a short Metropolis chain driven through ``calculate`` -- every step is
    one n = 1 call on the persistent context -- is reproducible and stays finite."""
    from victor_b200.likelihoods import CCFLikelihood
    _, blk = cobaya_block(repo_root)
    like = CCFLikelihood({"model": blk["model"], "data": blk["data"]})

    def chain(seed, steps=40):
        rng = np.random.default_rng(seed)
        x = np.array([0.47, 0.37, 380.0, 1.0])
        step = np.array([0.02, 0.005, 10.0, 0.005])

        def logp(v):
            st = {}
            like.calculate(st, fsigma8=v[0], beta=v[1], sigma_v=v[2], epsilon=v[3], alpha=1, astar=1)
            return st["logp"]

        cur = logp(x)
        trace = [cur]
        for _ in range(steps):
            y = x + step * rng.standard_normal(4)
            new = logp(y)
            if np.log(rng.uniform()) < new - cur:
                x, cur = y, new
            trace.append(cur)
        return np.array(trace)

    a, b = chain(1), chain(1)
    assert np.array_equal(a, b) and np.all(np.isfinite(a)) and a.max() >= a[0]
    like.ccf.close()
