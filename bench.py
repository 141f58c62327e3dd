#!/usr/bin/env python
"""Benchmark of the B200 likelihood hot path (BASELINE.json: likelihood evals/sec, batch 64K).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--no-cpu] [--no-extra]

One "step" = one pass of the hot path (theory multipoles + chi-square + lnL) over one batch of
65,536 synthetic parameter rows at the BOSS DR12 CMASS configuration (config/boss_config.yaml; BASELINE.json
configs[2]).  N > 1 is launched by torchrun, one rank per GPU; every rank owns its own 65,536-row batch (weak
scaling, no data-path collective; NCCL is only used for the barrier and the max-over-ranks time).

Printed JSON line (rank 0): the bench contract of the task description, plus
  roofline      FP64 CUDA-core roofline of the dominant kernel (k_multipoles): algorithmic flops per launch /
                CUDA-event duration against the NOMINAL FP64 peak, 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s
                (MEASURED_PEAKS.json and the profiling guide carry no FP64 figure); the DFMA-chain rate probed on
                this GPU in this run is a side key (fp64_probe_tflops).
  cpu_baseline  the reference's CPU path timed on all host cores on a bounded sample of the same batch: the
                UNMODIFIED reference (baseline/_ref, imported behind oracle/refshim.py) when it is importable
                (kind "reference"), else the oracle port (kind "port").
  e2e           same metric through CCFFit.log_likelihood_batch with HOST arrays (pinned), H2D of the parameter
                rows and D2H of chi2 / lnL inside the timed region.
  extra         the other BASELINE.json configurations and the north star's second model, timed in the same run
                (each max over ranks, CUDA events or wall clock around a synchronised region):
                  dispersion   65,536 rows per GPU with rsd_model 'dispersion' (own roofline block)
                  dense_sweep  configs[3]: one 1,048,576-row sweep sharded over the ranks (strong scaling),
                               200 mu x 100 velocity nodes, l = 0, 2, 4
                  mcmc         configs[4]: one chain per GPU, n = 1 calls through CCFLikelihood.calculate
                  strong_64k   ONE 65,536-row host table through batch.likelihood_sharded at N ranks (rows split,
                               results gathered over NCCL): the metric's "batch 64K, 1/2/4/8 B200" read strong
                  sustained    the headline step repeated for >= 5 s with the clock sampler running

``--impl reference`` times the reference's own CPU implementation with all host cores on a bounded sample per step.
``--workload dense|mcmc`` print the corresponding extra as a line of its own (longer runs of the same code).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20251018
BATCH = 65536
METRIC = "likelihood evals/sec (multipoles+chi2), batch 64K"
UNIT = "evals/s"
WORKLOAD = "BOSS DR12 CMASS batched likelihood (config/boss_config.yaml), 65536 synthetic rows per GPU"
NS, NMU, NX, L, P = 30, 100, 50, 2, 60
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12
PEAK_SOURCE = ("nominal FP64 peak 148 SM x 64 lanes x 2 flop x 1.965 GHz (sm_max_mhz of MEASURED_PEAKS.json; neither it "
               "nor the profiling guide states an FP64 figure)")

# Algorithmic flops per quadrature point (convention of SURVEY.md 8(d): add / sub / mul / div / sqrt / exp = 1,
# FMA = 2, a cubic-spline look-up = 1 + 6, compares and index arithmetic = 0); derivations in DESIGN.md section 5.
#   streaming  (ccf_model.py:646-658, 681-690): 41
#   dispersion (ccf_model.py:659-671): 2 (numerator, first guess) + 5 iterations x 13 + 50 (final point) = 117
#   each further real-space multipole (:684-687): spline 7 + Legendre 3 + multiply-add 2 = 12
FLOP_POINT = {"streaming": 41, "dispersion": 117}


def flop_k1(ns, nmu, nx, npoles, rsd="streaming", n_ell=1):
    """Algorithmic flops of the multipole kernel per parameter row: quadrature points + per-(s, mu) set-up
    (dispersion: + the first guess of the coordinate map, 12) + projection + per-row table preparation."""
    per_point = FLOP_POINT[rsd] + 12 * (n_ell - 1)
    per_pair = 8 + (12 if rsd == "dispersion" else 0)
    return per_point * ns * nmu * nx + per_pair * ns * nmu + 2 * npoles * ns * nmu + 1500


def flop_k2(p):
    """chi-square + log-det part per row."""
    return (3 * p * p + 2 * p * p + 2 * p) + (3 * p * p + p ** 3 // 3 + p)


FLOP_K1_PER_EVAL = flop_k1(NS, NMU, NX, L)
FLOP_PER_EVAL = FLOP_K1_PER_EVAL + flop_k2(P)
BYTES_PER_EVAL = 80 + 8 * L * NS + 16   # parameter row in (10 doubles), theory + chi2 + lnL out

# BASELINE.json configs[3]: streaming model on a dense mu / velocity grid, l = 0, 2, 4 (multipoles only:
# the data vector has no hexadecapole).  The reference hard-codes its grids; sizes per SURVEY.md 8(d).
DENSE = {"nmu": 200, "nx": 100, "poles": [0, 2, 4], "sweep": 1048576}


def synthetic_batch(n, seed=SEED):
    """Prior-box rows (fsigma8, beta, sigma_v, aperp, apar); SURVEY.md 8(d)."""
    rng = np.random.default_rng(seed)
    rows = np.empty((n, 5))
    rows[:, 0] = rng.uniform(0.05, 1.5, n)
    rows[:, 1] = rng.uniform(0.2, 0.6, n)
    rows[:, 2] = rng.uniform(100.0, 500.0, n)
    rows[:, 3] = rng.uniform(0.9, 1.1, n)
    rows[:, 4] = rng.uniform(0.9, 1.1, n)
    return rows


def boss_blocks():
    import yaml
    with open(os.path.join(ROOT, "config", "boss_config.yaml")) as fh:
        info = yaml.full_load(fh)
    info["model"]["dir"] = ROOT
    info["data"]["dir"] = ROOT
    return info["model"], info["data"]


# ------------------------------------------------------------------------------ CPU side
_CPU = None            # per worker process: the object whose log_likelihood is timed


def reference_kind():
    """"reference" when the unmodified reference is importable on this box (baseline/_ref, or the dev
    container's read-only checkout), else "port" (the oracle restatement)."""
    from oracle import refshim
    return "reference" if refshim.find_reference() else "port"


def _cpu_init(kind, npy_dir):
    global _CPU
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"
    import warnings
    warnings.filterwarnings("ignore")
    if kind == "reference":
        # the unmodified reference through its own public API, CCFFit(model, data).log_likelihood(params)
        # (victor/ccf_fit.py:356): import shims for absent third-party modules only, its own .npy reader
        from oracle import refshim
        victor = refshim.install()
        model, data = refshim.npy_twins(boss_blocks(), npy_dir)
        _CPU = victor.CCFFit(model, data)
    else:
        from oracle.ccf_oracle import OracleFit
        model, data = boss_blocks()
        _CPU = OracleFit(model, data)


def _cpu_eval(rows):
    t0 = time.perf_counter()
    out = []
    with open(os.devnull, "w") as devnull:      # the reference prints on failed points
        saved = sys.stdout
        sys.stdout = devnull
        try:
            for row in rows:
                prm = dict(zip(("fsigma8", "beta", "sigma_v", "aperp", "apar"), map(float, row)))
                out.append(tuple(float(v) for v in _CPU.log_likelihood(prm)))
        finally:
            sys.stdout = saved
    return time.perf_counter() - t0, out


class CpuPool:
    """All host cores, one CCFFit per worker process, one BLAS thread each."""

    def __init__(self):
        import multiprocessing as mp
        self.kind = reference_kind()
        self.cores = os.cpu_count() or 1
        self.tmp = tempfile.TemporaryDirectory(prefix="vb200_npy_")
        if self.kind == "reference":
            from oracle import refshim
            refshim.npy_twins(boss_blocks(), self.tmp.name)     # written once, before the workers start
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init, initargs=(self.kind, self.tmp.name))

    def run(self, rows):
        """(seconds of the slowest worker, [(lnl, chi2)] in row order)."""
        chunks = [rows[i::self.cores] for i in range(self.cores)]
        res = self.pool.map(_cpu_eval, chunks)
        out = [None] * len(rows)
        for i, (_, vals) in enumerate(res):
            out[i::self.cores] = vals
        return max(r[0] for r in res), out

    def describe(self):
        import scipy
        if self.kind == "reference":
            return ("the UNMODIFIED reference (victor 0.1.4 from baseline/_ref or /root/reference, CCFFit.log_likelihood, "
                    f"import shims of oracle/refshim.py, scipy {scipy.__version__})")
        return f"oracle/ccf_oracle.py, the numpy/scipy port of the reference algorithm (scipy {scipy.__version__})"

    def close(self):
        self.pool.close()
        self.pool.join()
        self.tmp.cleanup()


def cpu_baseline_leg(rows_per_worker=160):
    """cpu_baseline of the GPU arm: the CPU path over a bounded sample of the same batch."""
    cp = CpuPool()
    try:
        rows = synthetic_batch(BATCH)[:rows_per_worker * cp.cores]
        cp.run(rows[:cp.cores])          # warm-up: imports, first-call caches
        t0 = time.perf_counter()
        worker, out = cp.run(rows)
        wall = time.perf_counter() - t0
        return {"value": len(rows) / max(worker, 1e-9), "unit": UNIT, "cores": cp.cores, "kind": cp.kind,
                "sample": f"first {len(rows)} rows of the 65536-row batch, {rows_per_worker} per core; {cp.describe()}; "
                          f"1 BLAS thread per process, pool wall {wall:.2f}s",
                "chi2": [o[1] for o in out]}
    finally:
        cp.close()


def cpu_table_walk(fit, rows_host, chi2_gpu, nrows=8192):
    """oracle/table_walk.c on all host cores over the first `nrows` rows of the batch (a few seconds)."""
    from oracle.table_walk import TableWalk
    tw = TableWalk(fit)
    rows = rows_host[:nrows]
    tw.likelihood(rows[:64])
    t0 = time.perf_counter()
    _, chi2, _ = tw.likelihood(rows)
    wall = time.perf_counter() - t0
    return {"value": len(rows) / wall, "unit": UNIT, "cores": os.cpu_count() or 1,
            "sample": f"first {len(rows)} rows of the batch, plain C + OpenMP walk through the same tables, wall {wall:.2f}s",
            "max_abs_chi2_difference_to_gpu": float(np.nanmax(np.abs(chi2 - chi2_gpu[:len(rows)])))}


def run_reference(args):
    """--impl reference: the reference's CPU path on all host cores; one step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cp = CpuPool()
    try:
        per = 4
        rows = synthetic_batch(BATCH)[:per * cp.cores]
        times = []
        for _ in range(max(1, args.warmup)):
            cp.run(rows[:cp.cores])
        for _ in range(args.steps):
            times.append(cp.run(rows)[0])
        total = sum(times)
        val = len(rows) * args.steps / total
        sample = (f"{len(rows)} rows per step ({per} per core) of the 65536-row batch, all {cp.cores} host cores; "
                  f"{cp.describe()}")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD + f" -- CPU arm: a {len(rows)}-row sample of that batch per step",
                           "rows_per_step": len(rows), "rows_per_core": per},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cp.cores, "kind": cp.kind, "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
    finally:
        cp.close()
    return 0


# ------------------------------------------------------------------------------ GPU side
class ClockSampler:
    """`nvidia-smi` sampled every 100 ms.  The process takes a few hundred ms to come up, so it is started before
    the warm-up; `mark()` is called where the timed region begins and `stop()` reports the samples taken from then
    on (all samples, flagged, if the region was shorter than one sampling period)."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.index = vis.split(",")[index] if vis else index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        return self

    def mark(self):
        self.t_mark = time.time()
        return self

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        rows = []
        for ln in self.tmp.read().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                stamp = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((stamp, float(f[1]), float(f[2]), float(f[3]),
                             [name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                         "sw_power_cap"), f[5:9]) if val.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.tmp.name)
        window = [r for r in rows if self.t_mark is None or r[0] >= self.t_mark - 0.05]
        if rows and not window:
            window, out["window"] = rows, "timed region shorter than one sampling period: samples include the warm-up"
        if window:
            sm = [r[1] for r in window]
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(r[2] for r in window),
                       reasons=sorted({x for r in window for x in r[4]}), samples=len(window),
                       power_w_max=max(r[3] for r in window), sm_mhz_min=min(sm))
        return out


class Job:
    """Rank / device / process-group plumbing shared by the GPU workloads."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device; victor_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def timed_steps(self, step, steps, warmup):
        """CUDA-event time of each of `steps` calls of `step` on torch's current stream (= the stream the kernels
        are launched on), L2 flushed by a 256 MiB write before each, barrier + synchronize on both sides."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            self.flush.zero_()          # evict L2 between timed iterations (outside the event pair)
            a.record()
            step()
            b.record()
        self.barrier()
        return [a.elapsed_time(b) for a, b in ev]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def roofline_block(kernel, rows, flop_per_row, kernel_ms, extra=None):
    achieved = rows * flop_per_row / (kernel_ms * 1e-3) / 1e12
    blk = {"bound": "fp64", "kernel": kernel, "achieved": achieved, "peak": FP64_NOMINAL_TFLOPS, "unit": "TFLOP/s",
           "frac": achieved / FP64_NOMINAL_TFLOPS, "peak_source": PEAK_SOURCE, "flop_per_eval": flop_per_row,
           "kernel_ms": kernel_ms, "traffic": None}
    blk.update(extra or {})
    return blk


# ---- the extra workloads: every function returns {name: seconds or ms} of THIS rank plus a closure that turns the
# ---- max-over-ranks numbers into the JSON block
def extra_dispersion(job, fit, n, steps):
    """The north star's second model, rsd_model 'dispersion' (ccf_model.py:659-671), 65,536 rows per GPU."""
    torch = job.torch
    from victor_b200.model import params_to_rows
    eng, _ = fit._fit_engine({"rsd_model": "dispersion"})
    rows = params_to_rows(synthetic_batch(n, SEED + job.rank))
    d_params = torch.from_numpy(rows).to(job.dev)
    d_chi2 = torch.empty(n, dtype=torch.float64, device=job.dev)
    d_lnl = torch.empty(n, dtype=torch.float64, device=job.dev)
    d_theory = torch.empty((n, P), dtype=torch.float64, device=job.dev)
    ms = job.timed_steps(lambda: eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr(), d_chi2.data_ptr(),
                                                    d_lnl.data_ptr(), None), steps, 1)
    k1 = job.timed_steps(lambda: eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr(), None, None, None),
                         steps, 0)
    nonfinite = int(np.count_nonzero(~np.isfinite(d_lnl.cpu().numpy())))

    def finish(step_ms, k1_ms):
        fl = flop_k1(NS, NMU, NX, L, rsd="dispersion")
        return {"workload": f"BOSS tables, rsd_model dispersion, {n} rows per GPU (weak scaling)",
                "value": n * job.world / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms, "steps": steps,
                "roofline": roofline_block("k_multipoles<dispersion>", n, fl, k1_ms,
                                           {"flop_per_point": FLOP_POINT["dispersion"]}),
                # the reference's fixed-point coordinate iteration diverges for a few prior-box rows (high growth,
                # large separation): those end in its NaN guard, (-inf, inf), on the CPU too
                "nonfinite_rows_rank0": nonfinite}
    return [sum(ms) / len(ms), sum(k1) / len(k1)], finish


def extra_dense_sweep(job, fit, total, steps=1):
    """BASELINE.json configs[3]: one `total`-row sweep sharded over the ranks (strong scaling)."""
    torch = job.torch
    from victor_b200 import tables as T
    from victor_b200.batch import shard_bounds
    from victor_b200.model import params_to_rows
    opts = fit._merged_options({"velocity_nodes": DENSE["nx"], "mu_nodes": DENSE["nmu"]})
    eng = fit._engine(opts)
    mu, W = T.mu_projection_weights(DENSE["poles"], nmu=DENSE["nmu"])
    s = np.asarray(fit.s, dtype=np.float64)
    lo, hi = shard_bounds(total, job.world)[job.rank]
    n = hi - lo
    rows = params_to_rows(synthetic_batch(total, SEED)[lo:hi])
    d_params = torch.from_numpy(rows).to(job.dev)
    d_mult = torch.empty((n, len(DENSE["poles"]), len(s)), dtype=torch.float64, device=job.dev)
    eng.theory_ptr(d_params.data_ptr(), min(n, 16384), s, mu, W, None, d_mult.data_ptr(), None)     # warm-up
    ms = job.timed_steps(lambda: eng.theory_ptr(d_params.data_ptr(), n, s, mu, W, None, d_mult.data_ptr(), None),
                         steps, 0)
    # end to end on a bounded slice of the shard: host rows in, host multipoles out
    ne = min(n, 131072)
    pinned = torch.from_numpy(rows[:ne]).pin_memory().numpy()
    kw = {"velocity_nodes": DENSE["nx"], "mu_nodes": DENSE["nmu"]}
    fit.theory_multipole_vector_batch(s, pinned[:4096], DENSE["poles"], **kw)
    job.barrier()
    t0 = time.perf_counter()
    out = fit.theory_multipole_vector_batch(s, pinned, DENSE["poles"], **kw)
    e2e_s = time.perf_counter() - t0
    job.barrier()
    del d_mult

    def finish(step_ms, e2e_ms):
        fl = flop_k1(len(s), DENSE["nmu"], DENSE["nx"], len(DENSE["poles"]))
        return {"workload": f"BASELINE configs[3]: streaming, dense grid {DENSE['nmu']} mu x {DENSE['nx']} velocity nodes, "
                            f"l = 0, 2, 4 (multipoles only), one {total}-row sweep sharded over {job.world} GPU(s)",
                "metric": "theory-vector evals/sec", "value": total / (step_ms * 1e-3), "unit": UNIT,
                "scaling": "strong", "rows_total": total, "rows_per_gpu": n, "ms_per_sweep": step_ms, "steps": steps,
                "e2e": {"value": ne * job.world / (e2e_ms * 1e-3), "unit": UNIT, "rows_per_gpu": ne,
                        "h2d_bytes": int(ne * 80), "d2h_bytes": int(out.nbytes),
                        "api": "CCFModel.theory_multipole_vector_batch(host rows) on a slice of each shard"},
                "roofline": roofline_block("k_multipoles<streaming>", n, fl, step_ms)}
    return [sum(ms) / len(ms), e2e_s * 1e3], finish


def extra_mcmc(job, calls=1500, warm=200):
    """BASELINE.json configs[4]: one Metropolis chain per GPU, every step one n = 1 call through the plugin."""
    from victor_b200.likelihoods import CCFLikelihood
    model, data = boss_blocks()
    like = CCFLikelihood({"model": model, "data": data, "device": job.local})
    rng = np.random.default_rng(SEED + job.rank)
    state = {"x": np.array([0.47, 0.37, 380.0, 1.0])}
    step_sz = np.array([0.02, 0.005, 10.0, 0.005])

    def logp(v):
        st = {}
        like.calculate(st, fsigma8=float(v[0]), beta=float(v[1]), sigma_v=float(v[2]), epsilon=float(v[3]), alpha=1)
        return st["logp"]

    state["cur"] = logp(state["x"])
    lat = []

    def chain(ncalls):
        for _ in range(ncalls):
            y = state["x"] + step_sz * rng.standard_normal(4)
            t0 = time.perf_counter()
            new = logp(y)
            lat.append(time.perf_counter() - t0)
            if np.log(rng.uniform()) < new - state["cur"]:
                state["x"], state["cur"] = y, new

    chain(warm)
    lat.clear()
    job.barrier()
    t0 = time.perf_counter()
    chain(calls)
    job.torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    job.barrier()
    lat_us = np.array(lat) * 1e6
    med, p95 = float(np.median(lat_us)), float(np.percentile(lat_us, 95))
    cur = float(state["cur"])
    like.ccf.close()

    def finish(wall_s, med_us, p95_us):
        return {"workload": f"BASELINE configs[4]: {job.world} independent Metropolis chain(s), one per GPU, every step "
                            "one n = 1 call through CCFLikelihood.calculate (replicas only)",
                "metric": "likelihood calls/sec", "value": calls * job.world / wall_s, "unit": "calls/s",
                "calls_per_chain": calls, "latency_us": {"median": med_us, "p95": p95_us},
                "h2d_bytes_per_call": 80, "d2h_bytes_per_call": 16, "final_logp_rank0": cur}
    return [wall, med, p95], finish


def extra_strong_64k(job, fit, n, steps):
    """One n-row HOST table through batch.likelihood_sharded at world ranks: rows split, results gathered."""
    from victor_b200.batch import likelihood_sharded
    from victor_b200.model import params_to_rows
    rows = job.torch.from_numpy(params_to_rows(synthetic_batch(n, SEED))).pin_memory().numpy()   # the SAME table on every rank
    for _ in range(2):
        likelihood_sharded(fit, rows, gather=True)
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        lnl, chi2, (lo, hi) = likelihood_sharded(fit, rows, gather=True)
    job.torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    job.barrier()
    ok = bool(len(lnl) == n and np.all(np.isfinite(lnl)))

    def finish(wall_s):
        return {"workload": f"ONE {n}-row host table at the BOSS configuration split over {job.world} GPU(s) through "
                            "victor_b200.batch.likelihood_sharded (H2D of each slice, device-to-device all-gather of "
                            "lnL / chi2 over NCCL, one D2H of the gathered table)",
                "value": n * steps / wall_s, "unit": UNIT, "scaling": "strong", "ms_per_table": 1e3 * wall_s / steps,
                "steps": steps, "rows_total": n, "h2d_bytes_per_gpu": int((hi - lo) * 80), "gathered_bytes": int(n * 16),
                "all_rows_finite": ok}
    return [wall], finish


def extra_sustained(job, step, n, seconds):
    """The headline step back to back for >= `seconds` (no L2 flush, no host work): does the FP64 load hold
    the clock?  The sampler of rank 0 runs over exactly this region."""
    torch = job.torch
    job.barrier()
    sampler = ClockSampler(job.local).start().mark() if job.rank == 0 else None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    count = 0
    t0 = time.perf_counter()
    a.record()
    while True:
        for _ in range(8):
            step()
        count += 8
        torch.cuda.synchronize()
        if time.perf_counter() - t0 >= seconds:
            break
    b.record()
    job.barrier()
    ms = a.elapsed_time(b)
    clocks = sampler.stop() if sampler else None

    def finish(total_ms):
        return {"workload": f"the headline step ({n} rows per GPU) repeated back to back for >= {seconds:g} s",
                "value": n * job.world * count / (total_ms * 1e-3), "unit": UNIT, "steps": count,
                "seconds": total_ms * 1e-3, "ms_per_step": total_ms / count, "clocks": clocks}
    return [ms], finish


def run_gpu(args):
    import torch
    from victor_b200 import CCFFit
    from victor_b200.model import params_to_rows

    job = Job()
    world, rank, dev = job.world, job.rank, job.dev
    model, data = boss_blocks()
    fit = CCFFit(model, data, device=job.local)
    eng, _ = fit._fit_engine({})
    n = args.batch
    rows_host = params_to_rows(synthetic_batch(n, SEED + rank))
    d_params = torch.from_numpy(rows_host).to(dev)
    d_chi2 = torch.empty(n, dtype=torch.float64, device=dev)
    d_lnl = torch.empty(n, dtype=torch.float64, device=dev)
    d_theory = torch.empty((n, P), dtype=torch.float64, device=dev)
    stream_ptr = None  # legacy default stream == torch's current stream

    def step_device():
        eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr(), d_chi2.data_ptr(), d_lnl.data_ptr(),
                           stream_ptr)

    # --- device-resident throughput (`value`) ---
    sampler = ClockSampler(job.local).start() if rank == 0 else None
    for _ in range(args.warmup):
        step_device()
    job.barrier()
    if sampler:
        sampler.mark()
    launches0 = eng.launch_count()
    step_ms = job.timed_steps(step_device, args.steps, 0)
    launches = eng.launch_count() - launches0
    total_ms = sum(step_ms)

    # --- dominant-kernel timing for the roofline (multipoles only, CUDA events, same stream) ---
    k1_ms = job.timed_steps(lambda: eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr(), None, None,
                                                       stream_ptr), min(args.steps, 5), 0)
    clocks = sampler.stop() if sampler else None
    k1_avg_ms = sum(k1_ms) / len(k1_ms)

    # --- end to end through the public API with pinned host arrays (`e2e`) ---
    pinned = torch.from_numpy(rows_host).pin_memory()   # float64[n, 10] rows, page-locked
    host_rows = pinned.numpy()
    for _ in range(max(1, min(args.warmup, 3))):
        fit.log_likelihood_batch(host_rows)
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lnl_h, chi2_h = fit.log_likelihood_batch(host_rows)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    job.barrier()
    chi2_gpu = d_chi2.cpu().numpy()

    # --- the other configurations, same run ---
    mine, finishers = [total_ms, e2e_s * 1e3, k1_avg_ms], []
    if not args.no_extra:
        for name, fn in (("dispersion", lambda: extra_dispersion(job, fit, n, 2)),
                         ("dense_sweep", lambda: extra_dense_sweep(job, fit, args.sweep or DENSE["sweep"])),
                         ("mcmc", lambda: extra_mcmc(job)),
                         ("strong_64k", lambda: extra_strong_64k(job, fit, BATCH, 5)),
                         ("sustained", lambda: extra_sustained(job, step_device, n, args.sustain))):
            if name == "sustained" and args.sustain <= 0:
                continue
            vals, finish = fn()
            finishers.append((name, len(vals), finish))
            mine.extend(vals)
    red = job.max_over_ranks(mine)
    total_ms, e2e_ms, k1_red_ms = red[:3]

    if rank == 0:
        evals = n * world * args.steps
        value = evals / (total_ms * 1e-3)
        e2e_val = evals / (e2e_ms * 1e-3)
        extra, pos = {}, 3
        for name, cnt, finish in finishers:
            extra[name] = finish(*red[pos:pos + cnt])
            pos += cnt
        # FP64 issue-rate probe on this GPU, after the timed work: a side figure
        try:
            from victor_b200 import _probes
            fp64_probe = _probes.fp64_peak(job.local)[0]
        except Exception:
            fp64_probe = None
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        warp_points = n * NS * NMU * NX / 32.0
        roofline = roofline_block("k_multipoles<streaming, isotropic>", n, FLOP_K1_PER_EVAL, k1_avg_ms, {
            "flop_per_point": FLOP_POINT["streaming"],
            "fp64_probe_tflops": fp64_probe,
            "frac_of_probe": (n * FLOP_K1_PER_EVAL / (k1_avg_ms * 1e-3) / 1e12 / fp64_probe) if fp64_probe else None,
            # cycles one warp spends per quadrature point on its SM sub-partition (148 SMs x 4 sub-partitions)
            "cycles_per_warp_point": 148 * 4 * sm_mhz * 1e6 * (k1_avg_ms * 1e-3) / warp_points,
            "hbm_gbs_algorithmic": n * BYTES_PER_EVAL / (k1_avg_ms * 1e-3) / 1e9,
        })
        prof = os.path.join(ROOT, "profiles", "k_multipoles_traffic.json")
        if os.path.isfile(prof):
            with open(prof) as fh:
                roofline["traffic"] = json.load(fh).get("dram_bytes_per_launch")
        # what the committed `ncu --set full` capture of this kernel says limits it (static figures from profiles/,
        # not measured in this run): the FP64 pipe and the shared-memory crossbar together
        prof = os.path.join(ROOT, "profiles", "r02zb_k_multipoles_streaming_ncu.json")
        if os.path.isfile(prof):
            with open(prof) as fh:
                cap = json.load(fh)
            met = cap.get("metrics", {})

            def pct(key):
                try:
                    return float(str(met.get(key, "")).split()[0])
                except (ValueError, IndexError):
                    return None
            roofline["ncu_capture"] = {
                "source": "profiles/r02zb_k_multipoles_streaming_ncu.json",
                "fp64_pipe_pct_of_peak": pct("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                "smem_wavefronts_pct_of_peak": pct("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                "issue_slots_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fp64_instructions_per_point": cap.get("fp64_per_point"),
                "smem_wavefronts_per_point": cap.get("smem_wavefronts_per_point")}
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline_leg()
            chi2_cpu = np.array(cpu.pop("chi2"))
            cpu["max_abs_chi2_difference_to_gpu"] = float(np.max(np.abs(chi2_cpu - chi2_gpu[:len(chi2_cpu)])))
            # second CPU figure: the same table-driven algebra as the kernels, as plain C + OpenMP on the host cores
            # (oracle/table_walk.c) -- also a row-by-row check of the sample against the GPU results
            try:
                cpu["table_walk_c"] = cpu_table_walk(fit, rows_host, chi2_gpu=chi2_gpu)
            except Exception as exc:   # an extra figure: never let it take the bench line down
                cpu["table_walk_c"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": n, "ns": NS, "nmu": NMU, "nx": NX, "poles": [0, 2],
                       "likelihood": "sellentin/1000", "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": f"rows sharded over {world} GPU(s), no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(rows_host.nbytes),
                    "d2h_bytes_per_step": int(n * 16), "api": "CCFFit.log_likelihood_batch(host rows)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "flop_per_eval_total": FLOP_PER_EVAL,
            "check": {"lnl_finite": bool(np.all(np.isfinite(lnl_h))), "chi2_row0": float(chi2_h[0])},
            "extra": extra,
        }
        emit(line)
    job.close()
    fit.close()
    return 0


def run_dense(args):
    """BASELINE.json configs[3] as a line of its own: dense-grid streaming multipoles (l = 0, 2, 4), either
    `--batch` rows per GPU (weak) or one `--sweep`-row table sharded over the GPUs (strong)."""
    from victor_b200 import CCFFit
    job = Job()
    model, data = boss_blocks()
    fit = CCFFit(model, data, device=job.local)
    sampler = ClockSampler(job.local).start() if job.rank == 0 else None
    total = args.sweep or args.batch * job.world
    vals, finish = extra_dense_sweep(job, fit, total, steps=args.steps)
    red = job.max_over_ranks(vals)
    clocks = sampler.stop() if sampler else None
    if job.rank == 0:
        blk = finish(*red)
        line = {"metric": "theory-vector evals/sec (streaming, dense grid, l=0,2,4)", "value": blk["value"], "unit": UNIT,
                "n_gpus": job.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": blk["ms_per_sweep"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": blk["workload"], "rows_total": total, "rows_per_gpu": blk["rows_per_gpu"],
                           "l2": "flushed between timed steps (256 MiB write)"},
                "clocks": clocks, "e2e": blk["e2e"], "gpu_launches": int(args.steps), "roofline": blk["roofline"],
                "cpu_baseline": None}
        emit(line)
    job.close()
    fit.close()
    return 0


def run_mcmc(args):
    """BASELINE.json configs[4] as a line of its own."""
    job = Job()
    calls = 200 * max(1, args.steps)
    vals, finish = extra_mcmc(job, calls=calls, warm=200 * max(1, args.warmup))
    red = job.max_over_ranks(vals)
    if job.rank == 0:
        blk = finish(*red)
        line = {"metric": "likelihood calls/sec (n=1 MCMC steps through CCFLikelihood.calculate)",
                "value": blk["value"], "unit": "calls/s", "n_gpus": job.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * red[0] / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": blk["workload"], "calls_per_step": 200},
                "latency_us": blk["latency_us"],
                "e2e": {"value": blk["value"], "unit": "calls/s", "h2d_bytes_per_step": 200 * 80,
                        "d2h_bytes_per_step": 200 * 16, "api": "CCFLikelihood.calculate"},
                "gpu_launches": None, "final_logp": blk["final_logp_rank0"]}
        emit(line)
    job.close()
    return 0


_RESULT_FD = None


def emit(line):
    """Write the one JSON result line to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # Libraries print banners on stdout (NCCL: "NCCL version ..."): from here on file descriptor 1 points at
    # stderr, and only emit() writes to the real stdout -- exactly one JSON line from rank 0.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` block (other configs, sustained run)")
    ap.add_argument("--sustain", type=float, default=5.0, help="seconds of the sustained-clock run (0 = skip)")
    ap.add_argument("--sweep", type=int, default=0,
                    help="dense workload: total rows of one sweep sharded over the GPUs (strong scaling; default 1048576 "
                         "inside `extra`)")
    ap.add_argument("--workload", default="boss", choices=["boss", "dense", "mcmc"],
                    help="boss: BASELINE metric (default); dense: configs[3]; mcmc: configs[4]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "dense":
        return run_dense(args)
    if args.workload == "mcmc":
        return run_mcmc(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
