"""CPU: the three input-file formats of the path -- HDF5 and pickled-dict ``.npy`` as in the reference
(victor/ccf_model.py:54-68, victor/ccf_fit.py:44-57, 116-129), and this repository's ``.npz`` re-encodings.

tests/golden/example_void_model.hdf5 and tests/golden/cmass_data.hdf5 are the reference's own files
(data/example_data/example_void_model.hdf5; data/BOSS_DR12_CMASS_data/CMASS_..._data.hdf5), byte for byte;
tests/golden/example_void_model.npy holds the same arrays as a pickled dict (the reference's other format)."""
import copy
import os
import shutil

import numpy as np
import pytest

from victor_b200 import io_hdf5
from victor_b200.utils import InputError, load_input_file

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
PAIRS = [("example_void_model.hdf5", os.path.join(ROOT, "data", "example", "example_void_model.npz")),
         ("cmass_data.hdf5", os.path.join(ROOT, "data", "boss_dr12_cmass", "cmass_data.npz"))]


def same_arrays(a, b):
    assert set(a) == set(b)
    for key in a:
        x, y = np.asarray(a[key]), np.asarray(b[key])
        assert x.shape == y.shape and x.dtype == y.dtype, key
        assert np.array_equal(x, y), key            # bit for bit


@pytest.mark.parametrize("h5name,npz", PAIRS)
def test_hdf5_reader_returns_the_files_arrays(h5name, npz):
    """The built-in reader (superblock v0, contiguous float64 datasets: what victor ships) against the .npz twin
    the product's configurations point at."""
    want = load_input_file(npz)
    h5 = os.path.join(GOLDEN, h5name)
    same_arrays(io_hdf5._read_native(h5), want)
    same_arrays(io_hdf5.read_hdf5(h5), want)        # h5py if importable, else the native reader
    same_arrays(load_input_file(h5), want)          # dispatch on the extension, ccf_model.py:54-59
    for arr in load_input_file(h5).values():
        assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]


def test_every_reference_hdf5_extension_is_recognised(tmp_path):
    src = os.path.join(GOLDEN, "example_void_model.hdf5")
    want = load_input_file(src)
    for ext in (".hdf", ".h4", ".hdf4", ".he2", ".h5", ".hdf5", ".he5", ".h5py"):     # ccf_model.py:55
        dst = tmp_path / f"model{ext}"
        shutil.copy(src, dst)
        same_arrays(load_input_file(str(dst)), want)


def test_npy_dict_input():
    """``np.load(fn, allow_pickle=True).item()`` -- ccf_model.py:62-63."""
    same_arrays(load_input_file(os.path.join(GOLDEN, "example_void_model.npy")),
                load_input_file(PAIRS[0][1]))


def test_bad_files_raise_input_errors(tmp_path):
    with pytest.raises(InputError):
        load_input_file(str(tmp_path / "missing.hdf5"))
    bad = tmp_path / "model.txt"
    bad.write_text("r 1 2 3")
    with pytest.raises(InputError):
        load_input_file(str(bad))
    junk = tmp_path / "junk.hdf5"
    junk.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises((ValueError, InputError, OSError)):
        load_input_file(str(junk))
    trunc = tmp_path / "trunc.hdf5"
    with open(os.path.join(GOLDEN, "example_void_model.hdf5"), "rb") as fh:
        trunc.write_bytes(fh.read()[:600])
    with pytest.raises(Exception):
        load_input_file(str(trunc))


@pytest.mark.parametrize("fmt", ["hdf5", "npy"])
def test_model_loads_identically_from_every_format(example_block, fmt):
    """A CCFModel built from the HDF5 or .npy file carries the same host state (and therefore the same packed
    tables) as one built from the .npz the shipped configuration names."""
    from victor_b200 import CCFModel, tables as T
    base = CCFModel(copy.deepcopy(example_block))
    blk = copy.deepcopy(example_block)
    blk["dir"] = GOLDEN
    blk["input_model_data_file"] = f"example_void_model.{fmt}"
    other = CCFModel(blk)
    for key in ("r", "r_for_sv", "mu_for_sv", "sv_rmu"):
        assert np.array_equal(getattr(base, key), getattr(other, key)), key
    for ell in base.real_multipoles:
        assert np.array_equal(base.real_multipoles[ell], other.real_multipoles[ell])
    a, b = T.build_model_tables(base, base.model), T.build_model_tables(other, other.model)
    for name in ("xi_tab", "v0", "d0", "sv", "origin", "bucket_base"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_fit_reads_hdf5_data_vector(boss_blocks):
    """CCFFit._load_redshiftspace_ccf on the reference's own HDF5 data file (ccf_fit.py:44-114)."""
    from victor_b200 import CCFFit
    model, data = copy.deepcopy(boss_blocks[0]), copy.deepcopy(boss_blocks[1])
    base = CCFFit(copy.deepcopy(model), copy.deepcopy(data))
    data["redshift_space_ccf"]["data_file"] = os.path.relpath(os.path.join(GOLDEN, "cmass_data.hdf5"), ROOT)
    other = CCFFit(model, data)
    assert np.array_equal(base.s, other.s) and np.array_equal(base.beta_ccf, other.beta_ccf)
    for ell in base.redshift_multipoles:
        assert np.array_equal(base.redshift_multipoles[ell], other.redshift_multipoles[ell])
