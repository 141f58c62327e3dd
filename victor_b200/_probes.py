"""ctypes binding of libvictor_b200_probes.so (include/victor_b200_probes.h): self-test and measurement
hooks, kept out of the product library.  Used by tests/, bench.py and tools/probe_*.py only."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvictor_b200_probes.so")
EXPORTS = ("vb200p_last_error", "vb200p_math_selftest", "vb200p_pipe_probe", "vb200p_seed_probe", "vb200p_mix_probe",
           "vb200p_fp64_peak", "vb200p_load_probe", "vb200p_quad_probe")
SELFTEST_OUTPUTS = 10

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: run __graft_entry__.build() or make -C victor_b200/csrc")
    lib = ctypes.CDLL(LIB_PATH)
    lib.vb200p_last_error.restype = c_char_p
    lib.vb200p_math_selftest.restype = c_int
    lib.vb200p_math_selftest.argtypes = [c_int, c_void_p, c_int64, c_void_p]
    lib.vb200p_pipe_probe.restype = c_int
    lib.vb200p_pipe_probe.argtypes = [c_int, c_int, c_int, POINTER(c_double)]
    lib.vb200p_mix_probe.restype = c_int
    lib.vb200p_mix_probe.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_double)]
    lib.vb200p_seed_probe.restype = c_int
    lib.vb200p_seed_probe.argtypes = [c_int, c_void_p, c_int64, c_void_p]
    lib.vb200p_load_probe.restype = c_int
    lib.vb200p_load_probe.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_double)]
    lib.vb200p_quad_probe.restype = c_int
    lib.vb200p_quad_probe.argtypes = [c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, POINTER(c_double)]
    lib.vb200p_fp64_peak.restype = c_int
    lib.vb200p_fp64_peak.argtypes = [c_int, c_int, POINTER(c_double), POINTER(c_double)]
    _lib = lib
    return lib


def last_error():
    return load().vb200p_last_error().decode("utf-8", "replace")


def fp64_peak(device=0, iters=4096):
    """(TFLOP/s, ms) of the DFMA-chain probe on `device`; raises RuntimeError on failure."""
    tf, ms = c_double(), c_double()
    if load().vb200p_fp64_peak(int(device), int(iters), ctypes.byref(tf), ctypes.byref(ms)) != 0:
        raise RuntimeError(last_error())
    return float(tf.value), float(ms.value)
