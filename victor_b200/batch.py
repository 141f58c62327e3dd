"""Sharding a parameter table over the GPUs of one box.

Every parameter row is an independent unit of work and the tables are a few MB, so the path
shards trivially: each GPU holds a replica of the tables and evaluates a contiguous slice of
the rows.  There is no exchange step on the data path; the only optional communication is a
gather of the 16 bytes per row of results (chi2, lnL).

Ways to drive N GPUs:

* ``likelihood_sharded`` -- one process per GPU under ``torch.distributed`` (torchrun, NCCL): every rank
  passes the SAME full table, evaluates its own slice into a device buffer and the per-row results are
  all-gathered device to device, then copied to the host once.
* ``evaluate_sharded`` -- the same split for any host-array evaluator (and the gloo backend of the CPU
  tests): results travel as host arrays staged through a tensor.
* ``MultiDeviceFit`` -- one process, one context per visible GPU, one host thread per context
  (the ctypes calls release the GIL, so the launches and copies of the N devices overlap).

The reference has no counterpart: its only parallelism is cobaya's one-chain-per-MPI-rank
(README.md:30), which maps to one plugin instance (one context) per rank with no traffic at all.
"""
import threading

import numpy as np


def shard_bounds(n, world):
    """Contiguous slices of ``ceil(n / world)`` rows: [(lo, hi)] * world (trailing ones may be empty)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    per = -(-int(n) // world) if n > 0 else 0
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def evaluate_sharded(evaluate, rows, gather=True, group=None):
    """Evaluate this rank's slice of ``rows`` and optionally gather every rank's results.

    ``evaluate(rows_slice) -> (lnl[m], chi2[m])`` (e.g. ``CCFFit.log_likelihood_batch``).
    Returns ``(lnl, chi2, (lo, hi))``: with ``gather`` the full-length arrays, identical on
    every rank and bit-identical to a single-process evaluation of the whole table (the same
    kernel does the same per-row arithmetic wherever a row lands); without it this rank's slice.
    """
    import torch
    import torch.distributed as dist

    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n = rows.shape[0]
    if not (dist.is_available() and dist.is_initialized()):
        lnl, chi2 = evaluate(rows)
        return np.asarray(lnl), np.asarray(chi2), (0, n)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank]
    if hi > lo:
        lnl, chi2 = evaluate(rows[lo:hi])
        lnl, chi2 = np.asarray(lnl, dtype=np.float64), np.asarray(chi2, dtype=np.float64)
    else:
        lnl, chi2 = np.empty(0), np.empty(0)
    if not gather:
        return lnl, chi2, (lo, hi)
    per = bounds[0][1] - bounds[0][0]
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    mine = torch.full((2, per), float("nan"), dtype=torch.float64)
    mine[0, :hi - lo] = torch.from_numpy(lnl)
    mine[1, :hi - lo] = torch.from_numpy(chi2)
    mine = mine.to(dev)
    out = torch.empty((world, 2, per), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=group)
    out = out.cpu().numpy()
    lnl_all = np.concatenate([out[r, 0, :b - a] for r, (a, b) in enumerate(bounds)])
    chi2_all = np.concatenate([out[r, 1, :b - a] for r, (a, b) in enumerate(bounds)])
    return lnl_all, chi2_all, (lo, hi)


def likelihood_sharded(fit, rows, gather=True, group=None, **kwargs):
    """One parameter table, N ranks, results gathered device to device.

    Every rank passes the SAME ``rows``; rank r copies its own contiguous slice to its GPU, evaluates it into its
    slot of a device buffer (``CCFFit.log_likelihood_device``: nothing comes back to the host in between) and the
    slots are exchanged with one ``all_gather_into_tensor`` over NCCL (16 B per row); one device-to-host copy of the
    gathered table follows.  Returns ``(lnl, chi2, (lo, hi))`` like ``evaluate_sharded`` -- which remains the path
    for host-array evaluators, the gloo backend and the 'likelihood' beta-interpolation mode."""
    import torch
    import torch.distributed as dist

    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n = rows.shape[0]
    distributed = dist.is_available() and dist.is_initialized()
    like_mode = fit.fit_options.get("beta_interpolation") == "likelihood" or kwargs.get("beta_interpolation") == "likelihood"
    if not distributed or dist.get_backend(group) != "nccl" or like_mode:
        return evaluate_sharded(lambda part: fit.log_likelihood_batch(part, **kwargs), rows, gather=gather, group=group)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank]
    per = bounds[0][1] - bounds[0][0]
    eng, _ = fit._fit_engine(kwargs)
    dev = torch.device("cuda", eng.device)
    if not gather:
        mine = torch.empty((2, per), dtype=torch.float64, device=dev)
        if hi > lo:
            fit.log_likelihood_device(rows[lo:hi], out=mine, **kwargs)
        host = mine[:, :hi - lo].cpu().numpy()
        return host[0], host[1], (lo, hi)
    table, mirror = _gather_buffers(fit, world, per, dev)
    mine = table[rank]
    if hi > lo:
        fit.log_likelihood_device(rows[lo:hi], out=mine, **kwargs)
    dist.all_gather_into_tensor(table.view(-1), mine.reshape(-1), group=group)   # in place: rank r's slot is table[r]
    mirror.copy_(table, non_blocking=True)                                       # one DMA into page-locked memory
    torch.cuda.current_stream(dev).synchronize()
    host = mirror.numpy()
    lnl_all = np.concatenate([host[r, 0, :b - a] for r, (a, b) in enumerate(bounds)])
    chi2_all = np.concatenate([host[r, 1, :b - a] for r, (a, b) in enumerate(bounds)])
    return lnl_all, chi2_all, (lo, hi)


def _gather_buffers(fit, world, per, dev):
    """The gather table [world][lnL | chi2][per] on the device and its page-locked host mirror, kept on the fit
    between calls of the same shape.  Slots start NaN-filled; a rank only writes the rows it owns and the padding
    behind a short last slice is never read."""
    import torch
    key = (world, per, dev.index)
    cache = fit.__dict__.setdefault("_gather_cache", {})
    if key not in cache:
        cache.clear()                                                            # one shape at a time
        table = torch.full((world, 2, per), float("nan"), dtype=torch.float64, device=dev)
        mirror = torch.empty((world, 2, per), dtype=torch.float64).pin_memory()
        cache[key] = (table, mirror)
    return cache[key]


class MultiDeviceFit:
    """N replicas of a ``CCFFit`` on N GPUs of one box, driven from one process.

    ``factory(device) -> CCFFit``; ``log_likelihood_batch(rows)`` splits the table with
    ``shard_bounds`` and evaluates the slices concurrently, one host thread per device.
    """

    def __init__(self, factory, devices):
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("MultiDeviceFit needs at least one device")
        self.fits = [factory(d) for d in self.devices]

    def log_likelihood_batch(self, params, **kwargs):
        from .model import params_to_rows
        rows = params_to_rows(params)
        bounds = shard_bounds(len(rows), len(self.fits))
        results = [None] * len(self.fits)
        errors = []

        def work(i):
            lo, hi = bounds[i]
            try:
                if hi > lo:
                    results[i] = self.fits[i].log_likelihood_batch(rows[lo:hi], **kwargs)
                else:
                    results[i] = (np.empty(0), np.empty(0))
            except BaseException as exc:  # re-raised on the calling thread
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(self.fits))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return (np.concatenate([r[0] for r in results]), np.concatenate([r[1] for r in results]))

    def close(self):
        for f in self.fits:
            f.close()
