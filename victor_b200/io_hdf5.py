"""Minimal read-only HDF5 reader for the flat, contiguous files victor ships.

The reference reads its model / data / covariance inputs with ``h5py``
(``victor/ccf_model.py:64-68``, ``victor/ccf_fit.py:53-57``, ``:125-129``): every
top-level dataset is loaded whole into a dict.  ``h5py`` is not part of this image,
so this module decodes the subset of the HDF5 file format those files use:

* superblock version 0 (8-byte offsets and lengths),
* a root group stored as a version-1 B-tree of symbol-table nodes plus a local heap,
* version-1 object headers (with continuation blocks),
* simple dataspaces, fixed-point / IEEE floating-point datatypes,
* contiguous (layout class 1) or compact (class 0) data layout, no filters.

If ``h5py`` *is* importable it is preferred, so users with chunked / compressed files
are not restricted by this reader.  Anything outside the subset raises
``NotImplementedError`` naming the feature.
"""
import struct

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _Reader:
    def __init__(self, buf):
        self.buf = buf
        if buf[:8] != _SIGNATURE:
            raise ValueError("not an HDF5 file (bad signature)")
        version = buf[8]
        if version != 0:
            raise NotImplementedError(f"HDF5 superblock version {version} (only 0 is supported)")
        self.off_size, self.len_size = buf[13], buf[14]
        if (self.off_size, self.len_size) != (8, 8):
            raise NotImplementedError("HDF5 files with offsets/lengths that are not 8 bytes")
        # superblock v0: 24 bytes of versions/K values/flags, then 4 addresses, then root entry
        self.base = self.u64(24)
        root_entry = 24 + 4 * 8
        self.root = self.symbol_entry(root_entry)

    def u16(self, o):
        return struct.unpack_from("<H", self.buf, o)[0]

    def u32(self, o):
        return struct.unpack_from("<I", self.buf, o)[0]

    def u64(self, o):
        return struct.unpack_from("<Q", self.buf, o)[0]

    def symbol_entry(self, o):
        return {
            "name_off": self.u64(o),
            "header": self.u64(o + 8),
            "cache": self.u32(o + 16),
            "btree": self.u64(o + 24),
            "heap": self.u64(o + 32),
        }

    # ---- groups -------------------------------------------------------------------------
    def heap_data(self, addr):
        if self.buf[addr:addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        return self.u64(addr + 24)

    def heap_string(self, data_addr, off):
        start = data_addr + off
        end = self.buf.index(b"\x00", start)
        return self.buf[start:end].decode("utf-8")

    def group_entries(self, btree_addr, heap_addr):
        data_addr = self.heap_data(heap_addr)
        out = {}
        self._walk_btree(btree_addr, data_addr, out)
        return out

    def _walk_btree(self, addr, heap_data_addr, out):
        if self.buf[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        node_type, level, used = self.buf[addr + 4], self.buf[addr + 5], self.u16(addr + 6)
        if node_type != 0:
            raise NotImplementedError("non-group B-tree node at group level")
        o = addr + 8 + 16  # skip sibling pointers
        for i in range(used):
            child = self.u64(o + 8 + i * 16)  # key_i (8), child_i (8), ...
            if level > 0:
                self._walk_btree(child, heap_data_addr, out)
            else:
                self._read_snod(child, heap_data_addr, out)

    def _read_snod(self, addr, heap_data_addr, out):
        if self.buf[addr:addr + 4] != b"SNOD":
            raise ValueError("bad symbol-table node signature")
        n = self.u16(addr + 6)
        for i in range(n):
            e = self.symbol_entry(addr + 8 + 40 * i)
            out[self.heap_string(heap_data_addr, e["name_off"])] = e

    # ---- object headers -----------------------------------------------------------------
    def messages(self, addr):
        version = self.buf[addr]
        if version != 1:
            raise NotImplementedError(f"object header version {version} (only 1 is supported)")
        nmsg = self.u16(addr + 2)
        size = self.u32(addr + 8)
        blocks = [(addr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            o, remaining = blocks.pop(0)
            end = o + remaining
            while o + 8 <= end and len(msgs) < nmsg:
                mtype, msize = self.u16(o), self.u16(o + 2)
                body = o + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self.u64(body), self.u64(body + 8)))
                msgs.append((mtype, body, msize))
                o = body + msize
        return msgs

    def dataset(self, header_addr):
        shape = dtype = data = None
        for mtype, body, msize in self.messages(header_addr):
            if mtype == 0x01:
                ver, rank = self.buf[body], self.buf[body + 1]
                if ver == 1:
                    dims_at = body + 8
                elif ver == 2:
                    dims_at = body + 4
                else:
                    raise NotImplementedError(f"dataspace message version {ver}")
                shape = tuple(self.u64(dims_at + 8 * i) for i in range(rank))
            elif mtype == 0x03:
                cls = self.buf[body] & 0x0F
                bits0 = self.buf[body + 1]
                size = self.u32(body + 4)
                order = ">" if (bits0 & 1) else "<"
                if cls == 1:
                    dtype = np.dtype(f"{order}f{size}")
                elif cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
                else:
                    raise NotImplementedError(f"HDF5 datatype class {cls}")
            elif mtype == 0x08:
                ver = self.buf[body]
                if ver != 3:
                    raise NotImplementedError(f"data layout message version {ver}")
                lclass = self.buf[body + 1]
                if lclass == 1:
                    data = ("contiguous", self.u64(body + 2), self.u64(body + 10))
                elif lclass == 0:
                    data = ("compact", body + 4, self.u16(body + 2))
                else:
                    raise NotImplementedError("chunked HDF5 datasets (install h5py for these)")
            elif mtype == 0x0B:
                raise NotImplementedError("filtered HDF5 datasets (install h5py for these)")
        if shape is None or dtype is None or data is None:
            return None  # not a dataset (e.g. a sub-group)
        count = int(np.prod(shape)) if shape else 1
        _, addr, _ = data
        if addr == _UNDEF:
            arr = np.zeros(shape, dtype=dtype.newbyteorder("="))
        else:
            start = addr if data[0] == "compact" else addr + self.base
            arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=start).reshape(shape)
        return np.array(arr, dtype=dtype.newbyteorder("="))  # native-endian owned copy


def _read_native(path):
    with open(path, "rb") as fh:
        buf = fh.read()
    rd = _Reader(buf)
    if rd.root["cache"] != 1:
        raise NotImplementedError("HDF5 root group without a cached symbol table")
    out = {}
    for name, entry in rd.group_entries(rd.root["btree"], rd.root["heap"]).items():
        arr = rd.dataset(entry["header"])
        if arr is not None:
            out[name] = arr
    return out


def read_hdf5(path):
    """Return ``{name: ndarray}`` for every top-level dataset of ``path``.

    Same result as the reference's ``{key: f[key][:] for key in f.keys()}`` loop.
    """
    try:
        import h5py  # noqa: F401  (preferred when available)
    except ImportError:
        return _read_native(path)
    with h5py.File(path, "r") as f:
        return {key: f[key][:] for key in f.keys()}
