#!/usr/bin/env python
"""Benchmark of the B200 likelihood hot path (BASELINE.json: likelihood evals/sec, batch 64K).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

One "step" = one pass of the hot path (theory multipoles + chi-square + lnL) over one batch of
65,536 synthetic parameter rows at the BOSS DR12 CMASS configuration (config/boss_config.yaml).
N > 1 is launched by torchrun, one rank per GPU; every rank owns its own 65,536-row batch (weak
scaling, no data-path collective; NCCL is only used for the barrier and the max-over-ranks time).

Printed JSON line (rank 0): see the bench contract in the task description.  Additional keys:
  roofline      FP64 CUDA-core roofline of the dominant kernel (k_multipoles): algorithmic
                flops per launch / CUDA-event duration, against the FP64 FMA rate measured on
                this very GPU by a DFMA-chain probe (MEASURED_PEAKS.json has no FP64 number)
                and against the nominal 37.2 TFLOP/s.
  cpu_baseline  the CPU oracle (numpy/scipy port of the reference algorithm, same scipy calls)
                timed on all host cores on a bounded sample of the same batch.
  e2e           same metric through CCFFit.log_likelihood_batch with HOST arrays (pinned), H2D
                of the parameter rows and D2H of chi2 / lnL inside the timed region.

``--impl reference`` times the reference's CPU path (the oracle port: the reference is pure
Python and cannot be shipped) with all host cores on a bounded sample per step.
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


def flop_k1(ns, nmu, nx, npoles):
    """Algorithmic flops of the multipole kernel per parameter row, SURVEY.md 8(d): 41 flop per
    quadrature point + per-(s, mu) set-up + projection + per-row table preparation."""
    return 41 * ns * nmu * nx + 8 * ns * nmu + 2 * npoles * ns * nmu + 1500


def flop_k2(p):
    """chi-square + log-det part per row."""
    return (3 * p * p + 2 * p * p + 2 * p) + (3 * p * p + p ** 3 // 3 + p)


FLOP_K1_PER_EVAL = flop_k1(NS, NMU, NX, L)
FLOP_PER_EVAL = FLOP_K1_PER_EVAL + flop_k2(P)
BYTES_PER_EVAL = 80 + 8 * L * NS + 16   # parameter row in (10 doubles), theory + chi2 + lnL out

# BASELINE.json configs[3]: streaming model on a dense mu / velocity grid, l = 0, 2, 4 (multipoles only:
# the data vector has no hexadecapole).  The reference hard-codes its grids; sizes per SURVEY.md 8(d).
DENSE = {"nmu": 200, "nx": 100, "poles": [0, 2, 4]}


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
_ORACLE = None


def _oracle_init():
    global _ORACLE
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"
    import warnings
    warnings.filterwarnings("ignore")
    from oracle.ccf_oracle import OracleFit
    model, data = boss_blocks()
    _ORACLE = OracleFit(model, data)


def _oracle_eval(rows):
    t0 = time.perf_counter()
    out = []
    for row in rows:
        prm = dict(zip(("fsigma8", "beta", "sigma_v", "aperp", "apar"), map(float, row)))
        out.append(_ORACLE.log_likelihood(prm))
    return time.perf_counter() - t0, out


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


def cpu_oracle_throughput(rows_per_worker=160, repeats=1):
    """Oracle port on all host cores: evals / max worker time."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    rows = synthetic_batch(BATCH)[:rows_per_worker * cores]
    chunks = [rows[i::cores] for i in range(cores)]
    ctx = mp.get_context("spawn")
    best = None
    with ctx.Pool(cores, initializer=_oracle_init) as pool:
        pool.map(_oracle_eval, [c[:1] for c in chunks])  # warm-up: imports, first-call caches
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_oracle_eval, chunks)
            wall = time.perf_counter() - t0
            worker = max(r[0] for r in res)
            val = len(rows) / max(worker, 1e-9)
            best = val if best is None else max(best, val)
    return {"value": best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {len(rows)} rows of the 65536-row batch, {rows_per_worker} per core, "
                      f"oracle/ccf_oracle.py (scipy {__import__('scipy').__version__}), "
                      f"1 BLAS thread per process, pool wall {wall:.2f}s"}


def run_reference(args):
    """--impl reference: the CPU path on all host cores; one step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per = 4
    rows = synthetic_batch(BATCH)[:per * cores]
    chunks = [rows[i::cores] for i in range(cores)]
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(cores, initializer=_oracle_init) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_oracle_eval, [c[:1] for c in chunks])
        for _ in range(args.steps):
            res = pool.map(_oracle_eval, chunks)
            times.append(max(r[0] for r in res))
    total = sum(times)
    val = len(rows) * args.steps / total
    sample = (f"{len(rows)} rows per step ({per} per core) of the 65536-row batch; oracle port of the "
              "reference algorithm (the reference is pure Python and is not shipped)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "rows_per_step": len(rows)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------ GPU side
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.index = vis.split(",")[index] if vis else index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
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
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.tmp.read().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(power))
        return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from victor_b200 import CCFFit, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; victor_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    model, data = boss_blocks()
    fit = CCFFit(model, data, device=local)
    eng, _ = fit._fit_engine({})
    n = args.batch
    from victor_b200.model import params_to_rows
    rows_host = params_to_rows(synthetic_batch(n, SEED + rank))
    d_params = torch.from_numpy(rows_host).to(dev)
    d_chi2 = torch.empty(n, dtype=torch.float64, device=dev)
    d_lnl = torch.empty(n, dtype=torch.float64, device=dev)
    d_theory = torch.empty((n, P), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream_ptr = None  # legacy default stream == torch's current stream

    def step_device():
        eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr(), d_chi2.data_ptr(), d_lnl.data_ptr(),
                           stream_ptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- device-resident throughput (`value`) ---
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()          # evict L2 between timed iterations (outside the event pair)
        a.record()
        step_device()
        b.record()
    barrier()
    launches = eng.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)

    # --- dominant-kernel timing for the roofline (multipoles only, CUDA events, same stream) ---
    k1_ms = []
    for _ in range(min(args.steps, 5)):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr(), None, None, stream_ptr)
        b.record()
        torch.cuda.synchronize()
        k1_ms.append(a.elapsed_time(b))
    clocks = sampler.stop() if rank == 0 else None
    k1_avg_ms = sum(k1_ms) / len(k1_ms)

    # --- end to end through the public API with pinned host arrays (`e2e`) ---
    pinned = torch.from_numpy(rows_host).pin_memory()   # float64[n, 10] rows, page-locked
    host_rows = pinned.numpy()
    for _ in range(max(1, min(args.warmup, 3))):
        fit.log_likelihood_batch(host_rows)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lnl_h, chi2_h = fit.log_likelihood_batch(host_rows)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        evals = n * world * args.steps
        value = evals / (total_ms * 1e-3)
        e2e_val = evals / (e2e_ms * 1e-3)
        # FP64 probe on this GPU, after the timed work
        import ctypes
        tf, ms = ctypes.c_double(), ctypes.c_double()
        rc = _lib.load().vb200_fp64_peak(local, 4096, ctypes.byref(tf), ctypes.byref(ms))
        fp64_peak = float(tf.value) if rc == 0 else None
        achieved = n * FLOP_K1_PER_EVAL / (k1_avg_ms * 1e-3) / 1e12
        peak = fp64_peak or FP64_NOMINAL_TFLOPS
        roofline = {
            "bound": "fp64", "kernel": "k_multipoles<fast>", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": ("DFMA-chain probe on this GPU in this run (vb200_fp64_peak); MEASURED_PEAKS.json "
                            "has no FP64 figure" if fp64_peak else "nominal 148 SM x 64 lanes x 2 x 1.965 GHz"),
            "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
            "flop_per_eval": FLOP_K1_PER_EVAL, "kernel_ms": k1_avg_ms,
            # how full the FP64 pipe is under the issue model measured on this chip (tools/probe_mix.py,
            # tools/sass_opcycles.py): 37 FP64 instructions per quadrature point, 10.5 of them with three
            # register operands (3 cycles each, the others 2) = 84.5 pipe cycles per warp-point; and the
            # register-file read model (tools/sass_regreads.py): 213.5 32-bit reads = 106.8 cycles
            "fp64_pipe": {"instr_per_point": 37.0, "model_cycles_per_warp_point": 84.5,
                          "regfile_cycles_per_warp_point": 106.8,
                          "frac_of_regfile": (n * NS * NMU * NX / 32.0) * 106.8
                          / (148 * 4 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6) / (k1_avg_ms * 1e-3),
                          "frac_of_pipe": (n * NS * NMU * NX / 32.0) * 84.5
                          / (148 * 4 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6) / (k1_avg_ms * 1e-3)},
            "hbm_gbs_algorithmic": n * BYTES_PER_EVAL / (k1_avg_ms * 1e-3) / 1e9,
            "traffic": None,
        }
        prof = os.path.join(ROOT, "profiles", "k_multipoles_traffic.json")
        if os.path.isfile(prof):
            with open(prof) as fh:
                roofline["traffic"] = json.load(fh).get("dram_bytes_per_launch")
        cpu = cpu_oracle_throughput() if (world == 1 and not args.no_cpu) else None
        if cpu is not None:
            # second CPU figure: the same table-driven algebra as the kernels, as plain C + OpenMP on the host cores
            # (oracle/table_walk.c) -- also a row-by-row check of the sample against the GPU results
            try:
                cpu["table_walk_c"] = cpu_table_walk(fit, rows_host, chi2_gpu=d_chi2.cpu().numpy())
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
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    fit.close()
    return 0


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; victor_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def run_dense(args):
    """BASELINE.json configs[3]: dense-grid streaming multipoles (l = 0, 2, 4), rows sharded over the
    GPUs.  Reported as evaluations (theory vectors) per second; no chi-square (no l = 4 data)."""
    import torch
    import torch.distributed as dist
    from victor_b200 import CCFFit, tables as T
    from victor_b200.model import params_to_rows

    world, rank, local = _dist_setup()
    dev = torch.device("cuda", local)
    model, data = boss_blocks()
    fit = CCFFit(model, data, device=local)
    opts = fit._merged_options({"velocity_nodes": DENSE["nx"], "mu_nodes": DENSE["nmu"]})
    eng = fit._engine(opts)
    mu, W = T.mu_projection_weights(DENSE["poles"], nmu=DENSE["nmu"])
    s = np.asarray(fit.s, dtype=np.float64)
    if args.sweep:
        # one parameter sweep of `--sweep` rows in total, sharded over the ranks (strong scaling: configs[3] reads
        # "1M-point parameter sweep sharded over 2/4/8 B200")
        from victor_b200.batch import shard_bounds
        lo, hi = shard_bounds(args.sweep, world)[rank]
        rows_host = params_to_rows(synthetic_batch(args.sweep, SEED)[lo:hi])
        n = hi - lo
    else:
        n = args.batch
        rows_host = params_to_rows(synthetic_batch(n, SEED + rank))
    d_params = torch.from_numpy(rows_host).to(dev)
    d_mult = torch.empty((n, len(DENSE["poles"]), len(s)), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        eng.theory_ptr(d_params.data_ptr(), n, s, mu, W, None, d_mult.data_ptr(), None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        step()
        b.record()
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    pinned = torch.from_numpy(rows_host).pin_memory().numpy()
    fit.theory_multipole_vector_batch(s, pinned, DENSE["poles"], velocity_nodes=DENSE["nx"], mu_nodes=DENSE["nmu"])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = fit.theory_multipole_vector_batch(s, pinned, DENSE["poles"], velocity_nodes=DENSE["nx"],
                                                mu_nodes=DENSE["nmu"])
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        evals = (args.sweep if args.sweep else n * world) * args.steps
        fl = flop_k1(len(s), DENSE["nmu"], DENSE["nx"], len(DENSE["poles"]))
        achieved = n * args.steps * fl / (total_ms * 1e-3) / 1e12
        line = {"metric": "theory-vector evals/sec (streaming, dense grid, l=0,2,4)", "value": evals / (total_ms * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.sweep else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BOSS tables, streaming model, dense grid (BASELINE.json configs[3])",
                           "sweep_rows_total": args.sweep or None,
                           "rows_per_gpu": n, "ns": len(s), "nmu": DENSE["nmu"], "nx": DENSE["nx"],
                           "poles": DENSE["poles"], "l2": "flushed between timed steps (256 MiB write)",
                           "parallelism": f"rows sharded over {world} GPU(s), no collective"},
                "clocks": clocks,
                "e2e": {"value": evals / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n * 64),
                        "d2h_bytes_per_step": int(out.nbytes), "api": "CCFModel.theory_multipole_vector_batch(host rows)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "fp64", "kernel": "k_multipoles<fast>", "achieved": achieved,
                             "peak": FP64_NOMINAL_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_NOMINAL_TFLOPS,
                             "peak_source": "nominal 148 SM x 64 lanes x 2 x 1.965 GHz", "flop_per_eval": fl,
                             "traffic": None},
                "cpu_baseline": None}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    fit.close()
    return 0


def run_mcmc(args):
    """BASELINE.json configs[4]: one Metropolis chain per GPU, every step one n = 1 likelihood call
    through the cobaya plugin's ``calculate``.  Reports calls per second (whole job) and latency."""
    import torch
    import torch.distributed as dist
    from victor_b200.likelihoods import CCFLikelihood

    world, rank, local = _dist_setup()
    dev = torch.device("cuda", local)
    model, data = boss_blocks()
    like = CCFLikelihood({"model": model, "data": data, "device": local})
    rng = np.random.default_rng(SEED + rank)
    x = np.array([0.47, 0.37, 380.0, 1.0])
    step_sz = np.array([0.02, 0.005, 10.0, 0.005])

    def logp(v):
        st = {}
        like.calculate(st, fsigma8=float(v[0]), beta=float(v[1]), sigma_v=float(v[2]), epsilon=float(v[3]), alpha=1)
        return st["logp"]

    calls_per_step = 200
    cur = logp(x)
    lat = []

    def chain(ncalls):
        nonlocal x, cur
        for _ in range(ncalls):
            y = x + step_sz * rng.standard_normal(4)
            t0 = time.perf_counter()
            new = logp(y)
            lat.append(time.perf_counter() - t0)
            if np.log(rng.uniform()) < new - cur:
                x, cur = y, new

    for _ in range(args.warmup):
        chain(calls_per_step)
    lat.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        chain(calls_per_step)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t[0])
    if rank == 0:
        calls = calls_per_step * args.steps * world
        lat_us = np.array(lat) * 1e6
        line = {"metric": "likelihood calls/sec (n=1 MCMC steps through CCFLikelihood.calculate)",
                "value": calls / wall, "unit": "calls/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "Metropolis chain per GPU, BOSS config (BASELINE.json configs[4])",
                           "calls_per_step": calls_per_step, "parallelism": f"{world} independent chain(s), replicas only"},
                "latency_us": {"median": float(np.median(lat_us)), "p95": float(np.percentile(lat_us, 95)),
                               "min": float(lat_us.min())},
                "e2e": {"value": calls / wall, "unit": "calls/s", "h2d_bytes_per_step": calls_per_step * 80,
                        "d2h_bytes_per_step": calls_per_step * 16, "api": "CCFLikelihood.calculate"},
                "gpu_launches": int(2 * calls_per_step * args.steps), "final_logp": float(cur)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    like.ccf.close()
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
    ap.add_argument("--sweep", type=int, default=0,
                    help="dense workload: total rows of one sweep sharded over the GPUs (strong scaling), e.g. 1048576")
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
