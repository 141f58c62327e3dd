"""CPU: the C-ABI shared library loads, exports every symbol include/victor_b200.h declares,
agrees with the ctypes binding on struct layout, and reports errors without a GPU.
No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest


@pytest.fixture(scope="module")
def lib():
    from victor_b200 import _lib
    return _lib.load()


def header_functions(repo_root, header="victor_b200.h", prefix="vb200_"):
    with open(os.path.join(repo_root, "include", header)) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib, repo_root):
    from victor_b200 import _lib
    names = header_functions(repo_root)
    assert len(names) == 12 and not any("probe" in n or "selftest" in n for n in names)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/victor_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    dyn = set(re.findall(r"\bT (vb200_[a-z0-9_]+)", out))
    assert set(names) <= dyn
    # nothing but the C ABI (and no C++-mangled API) is part of the contract
    assert all(not n.startswith("_Z") for n in dyn)


def test_probe_library_is_separate(repo_root):
    """Self-test and measurement hooks live in their own library and header, not in the product ABI."""
    from victor_b200 import _lib, _probes
    plib = _probes.load()
    names = header_functions(repo_root, "victor_b200_probes.h", "vb200p_")
    assert set(names) == set(_probes.EXPORTS)
    for name in names:
        assert hasattr(plib, name)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "vb200p_" not in out and "probe" not in out


def test_struct_layout_matches(lib):
    from victor_b200 import _lib
    assert lib.vb200_abi_check(ctypes.sizeof(_lib.ModelTablesC), ctypes.sizeof(_lib.FitTablesC)) == 0
    assert lib.vb200_abi_check(1, 2) != 0
    assert b"mismatch" in lib.vb200_last_error()
    assert b"sm_100a" in lib.vb200_version()


def test_library_holds_sm100a_code_only(repo_root):
    from victor_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_error_paths_without_device(lib):
    from victor_b200 import _lib
    if lib.vb200_device_count() > 0:
        pytest.skip("a GPU is visible")
    assert lib.vb200_last_error() != b""            # why no device is visible
    handle = ctypes.c_void_p()
    rc = lib.vb200_create(None, None, 0, ctypes.byref(handle))
    assert rc == -1 and handle.value is None and b"NULL" in lib.vb200_last_error()
    assert lib.vb200_likelihood(None, None, 0, None, None, None, None) == -1
    assert lib.vb200_theory(None, None, 0, None, 0, None, 0, None, 0, None, None, None) == -1
    assert lib.vb200_set_option(None, b"threads", 128) == -1
    assert lib.vb200_launch_count(None) == 0
    lib.vb200_destroy(None)                           # no-op by contract


def test_engine_refuses_to_run_without_gpu(boss_blocks):
    import copy
    from victor_b200 import CCFFit, _lib
    if _lib.load().vb200_device_count() > 0:
        pytest.skip("a GPU is visible")
    model, data = boss_blocks
    fit = CCFFit(copy.deepcopy(model), copy.deepcopy(data))      # host-side construction works
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fit.theory_multipoles(fit.s, {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380, "epsilon": 1.0})


def test_product_never_imports_the_oracle(repo_root):
    """oracle/ is test infrastructure: nothing under victor_b200/ may reference it."""
    pkg = os.path.join(repo_root, "victor_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as fh:
                    text = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "ccf_oracle" not in text and "table_emul" not in text, fn


def test_package_surface_mirrors_the_reference():
    """Names exported by victor/__init__.py that lie on or next to the path (plotting and the excursion-set profile
    are outside it)."""
    import numpy as np
    import victor_b200
    for name in ("CCFModel", "CCFFit", "BackgroundCosmology", "InputError", "utils", "__version__"):
        assert hasattr(victor_b200, name), name
    for name in ("InputError", "multipoles_from_fn", "fn_from_multipoles"):
        assert hasattr(victor_b200.utils, name), name
    from conftest import load_golden
    iaH = float(load_golden("boss_tables")["iaH"])
    cosmo = victor_b200.BackgroundCosmology({"Omega_m": 0.31})
    assert abs((1 + 0.57) / (100 * cosmo.Ez(0.57)) - iaH) < 1e-16       # ccf_model.py:44-45
    assert abs(cosmo.Om(0.0) - 0.31) < 1e-15 and np.allclose(cosmo.H([0.0, 0.57]), [67.5, 67.5 * cosmo.Ez(0.57)])


def test_every_runtime_option_is_documented_in_the_header(repo_root):
    """vb200_set_option's keys (api.cu) and the option list of include/victor_b200.h name the same set."""
    import re
    with open(os.path.join(repo_root, "victor_b200", "csrc", "api.cu")) as fh:
        keys = set(re.findall(r'!strcmp\(key, "(\w+)"\)', fh.read()))
    with open(os.path.join(repo_root, "include", "victor_b200.h")) as fh:
        header = fh.read()
    assert len(keys) >= 10
    missing = sorted(k for k in keys if f'"{k}"' not in header)
    assert not missing, f"options without a line in the header: {missing}"
