"""GPU, world_size 2 over NCCL (runs only where two GPUs are visible: `gpurun --gpus 2`): rows
sharded over ranks, per-row results gathered, bit-identical to the single-GPU evaluation."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, copy
import numpy as np, torch, torch.distributed as dist, yaml
sys.path.insert(0, os.environ["VB200_ROOT"])
from victor_b200 import CCFFit
from victor_b200.batch import evaluate_sharded, likelihood_sharded
from victor_b200.model import params_to_rows
from bench import synthetic_batch
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
root = os.environ["VB200_ROOT"]
info = yaml.full_load(open(os.path.join(root, "config", "boss_config.yaml")))
info["model"]["dir"] = info["data"]["dir"] = root
fit = CCFFit(info["model"], info["data"], device=local)
rows = params_to_rows(synthetic_batch(65536)[:1001])
lnl, chi2, (lo, hi) = evaluate_sharded(fit.log_likelihood_batch, rows, gather=True)
# the device-to-device form: slices evaluated into device buffers, all-gather over NCCL, one D2H
lnl_d, chi2_d, (lo_d, hi_d) = likelihood_sharded(fit, rows, gather=True)
assert (lo_d, hi_d) == (lo, hi) and np.array_equal(lnl_d, lnl) and np.array_equal(chi2_d, chi2)
for m in (1000, 1001, 37, 1001):          # the gather buffers are kept between calls: other sizes, then the first again
    l_e, c_e, _ = evaluate_sharded(fit.log_likelihood_batch, rows[:m], gather=True)
    l_d, c_d, _ = likelihood_sharded(fit, rows[:m], gather=True)
    assert np.array_equal(l_d, l_e) and np.array_equal(c_d, c_e), m
lnl_s, chi2_s, _ = likelihood_sharded(fit, rows, gather=False)
assert np.array_equal(lnl_s, lnl[lo:hi]) and np.array_equal(chi2_s, chi2[lo:hi])
np.savez(os.path.join(os.environ["VB200_OUT"], f"rank{rank}.npz"), lnl=lnl, chi2=chi2, lo=lo, hi=hi)
if rank == 0:
    l1, c1 = fit.log_likelihood_batch(rows)
    np.savez(os.path.join(os.environ["VB200_OUT"], "single.npz"), lnl=l1, chi2=c1)
dist.destroy_process_group()
fit.close()
'''


def test_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, VB200_ROOT=ROOT, VB200_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    r0, r1, one = (np.load(tmp_path / f) for f in ("rank0.npz", "rank1.npz", "single.npz"))
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 501, 501, 1001)
    assert np.array_equal(r0["chi2"], r1["chi2"]) and np.array_equal(r0["lnl"], r1["lnl"])
    assert np.array_equal(r0["chi2"], one["chi2"]) and np.array_equal(r0["lnl"], one["lnl"])
