#!/usr/bin/env python
"""Short, deterministic launch sequence for ncu: a few 65,536-row likelihood passes, nothing else.

    python tools/profile_target.py [--batch 65536] [--passes 3] [--fast 1] [--threads 256]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from bench import boss_blocks, synthetic_batch, P  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--fast", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--nsplit", type=int, default=0)
    ap.add_argument("--ilp", type=int, default=0)
    ap.add_argument("--expdeg", type=int, default=None)
    ap.add_argument("--newton", type=int, default=None)
    ap.add_argument("--fuse", type=int, default=1, help="1: chi2 / lnL in the K1 epilogue; 0: separate K2 launch")
    ap.add_argument("--theory", type=int, default=1, help="0: do not ask for the theory vectors (chi2 / lnL only)")
    ap.add_argument("--sigma-v", type=float, default=None, help="override the sigma_v column (access-pattern probe)")
    ap.add_argument("--rsd", default="streaming", help="rsd_model (general kernel for anything but streaming)")
    ap.add_argument("--aniso", type=int, default=0, help="1: assume_isotropic False")
    ap.add_argument("--tuned", type=int, default=1, help="0: force the general kernel")
    ap.add_argument("--bucket", type=int, default=1, help="0: K2 row by row (k_chi2) instead of grouped by covariance bracket")
    args = ap.parse_args()
    model, data = boss_blocks()
    fit = CCFFit(model, data, device=0)
    kw = {}
    if args.rsd != "streaming":
        kw["rsd_model"] = args.rsd
    if args.aniso:
        kw["assume_isotropic"] = False
    eng, _ = fit._fit_engine(kw)
    eng.set_option("fast_math", args.fast)
    eng.set_option("threads", args.threads)
    eng.set_option("nsplit", args.nsplit)
    eng.set_option("ilp", args.ilp)
    if args.expdeg is not None:
        eng.set_option("exp_degree", args.expdeg)
    if args.newton is not None:
        eng.set_option("newton", args.newton)
    eng.set_option("fuse", args.fuse)
    eng.set_option("tuned", args.tuned)
    eng.set_option("bucket", args.bucket)
    n = args.batch
    dev = torch.device("cuda", 0)
    rows = params_to_rows(synthetic_batch(n))
    if args.sigma_v is not None:
        rows[:, 2] = args.sigma_v
    d_params = torch.from_numpy(rows).to(dev)
    d_theory = torch.empty((n, P), dtype=torch.float64, device=dev)
    d_chi2 = torch.empty(n, dtype=torch.float64, device=dev)
    d_lnl = torch.empty(n, dtype=torch.float64, device=dev)
    times = []
    for _ in range(args.passes):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr() if args.theory else None, d_chi2.data_ptr(),
                           d_lnl.data_ptr())
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    print(f"batch={n} fast={args.fast} threads={args.threads} nsplit={args.nsplit} ilp={args.ilp} expdeg={args.expdeg} newton={args.newton} fuse={args.fuse} theory={args.theory} sigma_v={args.sigma_v} rsd={args.rsd} aniso={args.aniso} tuned={args.tuned} bucket={args.bucket} lib={os.path.basename(os.environ.get('VICTOR_B200_LIB', 'default'))} "
          f"ms={['%.3f' % t for t in times]} evals/s={n / (min(times) * 1e-3):.4g} "
          f"chi2[0]={float(d_chi2[0]):.10f} "
          f"digest={__import__('hashlib').sha1(d_theory.cpu().numpy().tobytes() + d_chi2.cpu().numpy().tobytes()).hexdigest()[:12]}")
    fit.close()


if __name__ == "__main__":
    main()
