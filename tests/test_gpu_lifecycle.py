"""GPU: lifecycle and misuse of the C ABI -- what a binding written by someone else will do wrong sooner or later.

Contexts come and go without leaking device memory, two contexts driven from two host threads do not disturb each
other ("distinct contexts are independent"), and bad arguments come back as error codes with a
message, never as a crash or a poisoned context."""
import copy
import ctypes
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHI2_ATOL = 1e-6


@pytest.fixture(scope="module")
def blocks(boss_blocks):
    model, data = boss_blocks
    return copy.deepcopy(model), copy.deepcopy(data)


def _rows(n, seed):
    from bench import synthetic_batch
    from victor_b200.model import params_to_rows
    return params_to_rows(synthetic_batch(n, seed=seed))


def test_contexts_release_their_device_memory(blocks):
    """Twelve fits built, used on every call path (one row, a graph-replayed handful, a batch, theory vectors)
    and closed: free device memory ends where it started (within the allocator's granularity)."""
    import torch
    from victor_b200 import CCFFit
    rows = _rows(600, 1)

    def cycle():
        f = CCFFit(*copy.deepcopy(blocks))
        eng, _ = f._fit_engine({})
        eng.likelihood(rows[:1])
        eng.likelihood(rows[:7])
        eng.likelihood(rows, want_theory=True)
        f.log_likelihood_batch(rows[:64], rsd_model="dispersion")
        f.close()

    cycle()                                                   # first use: CUDA module load, torch context
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(12):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 8 << 20, f"{(free0 - free1) / 2**20:.1f} MiB of device memory not returned"


def test_two_contexts_on_two_threads(blocks):
    """Two fits, each owned by one host thread, evaluating different tables at the same time (batches, MCMC-size
    calls and option changes interleaved): every result equals the single-threaded one bit for bit."""
    from victor_b200 import CCFFit
    tables = [_rows(3000, 11), _rows(3000, 12)]
    ref = CCFFit(*copy.deepcopy(blocks))
    eng, _ = ref._fit_engine({})
    want = [eng.likelihood(t)[1:] for t in tables]
    want_one = [[eng.likelihood(t[i:i + 1])[1:] for i in range(20)] for t in tables]
    ref.close()

    errors = []

    def worker(k):
        try:
            f = CCFFit(*copy.deepcopy(blocks))
            e, _ = f._fit_engine({})
            for rep in range(6):
                chi2, lnl = e.likelihood(tables[k])[1:]
                assert np.array_equal(chi2, want[k][0]) and np.array_equal(lnl, want[k][1]), f"batch, thread {k}"
                for i in range(20):
                    c1, l1 = e.likelihood(tables[k][i:i + 1])[1:]
                    assert c1[0] == want_one[k][i][0][0] and l1[0] == want_one[k][i][1][0], f"one row, thread {k}"
            f.close()
        except Exception as exc:      # noqa: BLE001 -- reported below, in the main thread
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_bad_arguments_come_back_as_error_codes(blocks):
    """NULL handles and pointers, negative sizes, unknown options, grids that do not fit: a negative VB200_E* code
    and a message from vb200_last_error() -- and the context still evaluates correctly afterwards."""
    from victor_b200 import CCFFit, _lib
    lib = _lib.load()
    f = CCFFit(*copy.deepcopy(blocks))
    eng, _ = f._fit_engine({})
    h = eng.handle
    rows = _rows(4, 3)
    out = np.empty((2, 4))
    good = lambda: lib.vb200_likelihood(h, rows.ctypes.data, 4, None, out.ctypes.data, out.ctypes.data + 32, None)  # noqa: E731
    assert good() == 0
    first = out.copy()

    def refused(rc):
        assert rc < 0
        assert _lib.last_error(), "an error code without a message"

    refused(lib.vb200_likelihood(None, rows.ctypes.data, 4, None, out.ctypes.data, out.ctypes.data + 32, None))
    refused(lib.vb200_likelihood(h, None, 4, None, out.ctypes.data, out.ctypes.data + 32, None))
    refused(lib.vb200_likelihood(h, rows.ctypes.data, -1, None, out.ctypes.data, out.ctypes.data + 32, None))
    refused(lib.vb200_likelihood(h, rows.ctypes.data, 4, None, None, None, None))        # nothing asked for
    refused(lib.vb200_set_option(h, b"no_such_option", 1))
    refused(lib.vb200_set_option(None, b"fuse", 1))
    s = np.linspace(1.0, 100.0, 8)
    mu = np.linspace(0.0, 1.0, 5)
    xi = np.empty((4, 5, 8))
    refused(lib.vb200_theory(h, rows.ctypes.data, 4, s.ctypes.data, 0, mu.ctypes.data, 5, None, 0, xi.ctypes.data,
                             None, None))                                                 # empty s grid
    refused(lib.vb200_theory(h, rows.ctypes.data, 4, s.ctypes.data, 8, None, 5, None, 0, xi.ctypes.data, None, None))
    refused(lib.vb200_theory(h, rows.ctypes.data, 4, s.ctypes.data, 8, mu.ctypes.data, 5, None, 0, None, None, None))
    refused(lib.vb200_theory_pairs(h, rows.ctypes.data, 4, s.ctypes.data, mu.ctypes.data, -3, xi.ctypes.data, None))
    bad_handle = ctypes.c_void_p()
    refused(lib.vb200_create(None, None, 0, ctypes.byref(bad_handle)))
    assert not bad_handle.value
    assert lib.vb200_synchronize(h) == 0

    assert good() == 0
    assert np.array_equal(out, first)                         # the context is as it was
    f.close()


def test_a_wrong_device_index_is_refused(blocks):
    from victor_b200 import CCFFit
    with pytest.raises((RuntimeError, ValueError)):
        f = CCFFit(*copy.deepcopy(blocks), device=4096)
        f.log_likelihood({"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0})
