"""CPU: the N > 1 path.  Rows shard contiguously over ranks with no data-path collective; the
optional gather of (lnL, chi2) is exercised with world_size 2 over gloo.  The per-slice
evaluation is stood in for by the numpy table walk-through (oracle/table_emul.py) -- on the GPU
box the same code runs with CCFFit.log_likelihood_batch over NCCL (tests/test_gpu_parity.py)."""
import copy
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    from victor_b200.batch import shard_bounds
    assert shard_bounds(10, 1) == [(0, 10)]
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    assert shard_bounds(65536, 8)[-1] == (57344, 65536)
    for n, w in ((1, 8), (7, 2), (1048576, 8), (100, 3)):
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        shard_bounds(4, 0)


def test_unsharded_call_is_a_passthrough():
    from victor_b200.batch import evaluate_sharded
    rows = np.arange(30.0).reshape(3, 10)
    lnl, chi2, (lo, hi) = evaluate_sharded(lambda r: (-r[:, 0], r[:, 1] ** 2), rows)
    assert (lo, hi) == (0, 3) and np.array_equal(lnl, -rows[:, 0]) and np.array_equal(chi2, rows[:, 1] ** 2)


def _worker(rank, world, port, n_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import yaml
    from oracle import table_emul as E
    from victor_b200 import CCFFit, tables as T
    from victor_b200.batch import evaluate_sharded, likelihood_sharded
    from victor_b200.model import params_to_rows
    from bench import synthetic_batch

    dist.init_process_group("gloo", rank=rank, world_size=world)
    with open(os.path.join(ROOT, "config", "boss_config.yaml")) as fh:
        info = yaml.full_load(fh)
    info["model"]["dir"] = info["data"]["dir"] = ROOT
    fit = CCFFit(info["model"], info["data"])
    mt = T.build_model_tables(fit, fit.model)
    ft = T.build_fit_tables(fit, fit.fit_options["likelihood"])
    mu, W = T.mu_projection_weights(fit.poles_s)
    calls = []

    def evaluate(rows):
        calls.append(len(rows))
        mult, _ = E.theory_multipoles(mt, rows, np.asarray(fit.s, float), mu, W)
        chi2, lnl = E.chi2_lnl(ft, rows[:, 1], mult.reshape(len(rows), -1))
        return lnl, chi2

    rows = params_to_rows(synthetic_batch(65536)[:n_rows])
    lnl, chi2, (lo, hi) = evaluate_sharded(evaluate, rows, gather=True)
    lnl_s, chi2_s, _ = evaluate_sharded(evaluate, rows, gather=False)
    ncalls = len(calls)
    # likelihood_sharded is the device-to-device form for NCCL; on any other backend it must route through
    # evaluate_sharded with the fit's own host-array evaluator
    fit.log_likelihood_batch = lambda part, **kw: evaluate(part)
    lnl_l, chi2_l, (lo_l, hi_l) = likelihood_sharded(fit, rows, gather=True)
    assert (lo_l, hi_l) == (lo, hi) and np.array_equal(lnl_l, lnl) and np.array_equal(chi2_l, chi2)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lnl=lnl, chi2=chi2, lo=lo, hi=hi, calls=np.array(calls[:ncalls]),
             lnl_slice=lnl_s, chi2_slice=chi2_s)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_two_gloo(tmp_path, boss_blocks, golden):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_rows = 7                                     # odd: slices of 4 and 3 rows
    mp.spawn(_worker, args=(2, port, n_rows, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in (0, 1))
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 4, 4, 7)
    assert list(r0["calls"]) == [4, 4] and list(r1["calls"]) == [3, 3]     # each rank touched only its slice
    # the gathered vectors are identical on both ranks and equal to slice-concatenation
    assert np.array_equal(r0["lnl"], r1["lnl"]) and np.array_equal(r0["chi2"], r1["chi2"])
    assert np.array_equal(r0["chi2"], np.concatenate([r0["chi2_slice"], r1["chi2_slice"]]))
    # and they are the reference's numbers: the first rows of the seeded batch are golden rows
    g = golden("boss_streaming_points")
    np.testing.assert_allclose(r0["chi2"], g["chi2"][:n_rows], rtol=0, atol=1e-6)
    np.testing.assert_allclose(r0["lnl"], g["lnl"][:n_rows], rtol=0, atol=1e-6)
