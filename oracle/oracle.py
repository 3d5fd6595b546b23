"""ctypes loader for the C oracle (oracle/sn_oracle.c).  TEST INFRASTRUCTURE ONLY:
import from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never
from the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libsn_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libsn_oracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "sn_oracle.c")):
            build()
        L = C.CDLL(so)
        vp, u64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int32, C.c_uint32
        L.sn_oracle_new.restype = vp
        L.sn_oracle_new.argtypes = [u64, vp, vp, vp, vp, C.c_uint, C.c_uint, C.c_uint]
        for f in ("count", "prune", "edges", "hbv", "paths"):
            getattr(L, "sn_oracle_" + f).argtypes = [vp]
        L.sn_oracle_run.argtypes = [vp, C.c_int]
        L.sn_oracle_set_ign_bc_below.argtypes = [vp, C.c_int64]
        L.sn_oracle_set_count_len_k.argtypes = [vp, C.c_int]
        L.sn_oracle_free.argtypes = [vp]
        for f, rt in (("n_occ", u64), ("n_kmers", u64), ("n_edges", u64), ("n_vert", i32), ("n_hbv_edges", i32)):
            getattr(L, "sn_oracle_" + f).restype = rt
            getattr(L, "sn_oracle_" + f).argtypes = [vp]
        for f in ("good_len", "path_offset", "path_off", "path_edges", "fwd_xlat", "rev_xlat"):
            getattr(L, "sn_oracle_" + f).restype = vp
            getattr(L, "sn_oracle_" + f).argtypes = [vp]
        L.sn_oracle_get_kmers.argtypes = [vp, vp]
        L.sn_oracle_edge_len.restype = u32
        L.sn_oracle_edge_len.argtypes = [vp, u64]
        L.sn_oracle_edge_seq.restype = vp
        L.sn_oracle_edge_seq.argtypes = [vp, u64]
        L.sn_oracle_write_hbv.argtypes = [vp, C.c_char_p]
        L.sn_oracle_write_paths.argtypes = [vp, C.c_char_p]
        L.sn_oracle_involution.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class Oracle:
    """Runs the restated reference pipeline on ragged reads given as base codes."""

    def __init__(self, bases, quals, off, bc, min_qual=7, min_freq=3, min_bc=2, ign_bc_below=0, count_len_k=False):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8).ravel()
        self.quals = np.ascontiguousarray(quals, dtype=np.uint8).ravel()
        self.off = np.ascontiguousarray(off, dtype=np.uint64)
        self.bc = None if bc is None else np.ascontiguousarray(bc, dtype=np.int32)
        self.n = len(self.off) - 1
        L = lib()
        self.h = L.sn_oracle_new(self.n, self.bases.ctypes.data, self.quals.ctypes.data, self.off.ctypes.data,
                                 None if self.bc is None else self.bc.ctypes.data, min_qual, min_freq, min_bc)
        if ign_bc_below:
            L.sn_oracle_set_ign_bc_below(self.h, int(ign_bc_below))
        if count_len_k:        # the tada variant: reads trimmed to exactly K bases are counted
            L.sn_oracle_set_count_len_k(self.h, 1)

    @classmethod
    def from_matrix(cls, bases2d, quals2d, bc, **kw):
        n, L = bases2d.shape
        off = np.arange(n + 1, dtype=np.uint64) * L
        return cls(bases2d, quals2d, off, bc, **kw)

    def run(self, with_paths=True):
        lib().sn_oracle_run(self.h, int(with_paths))
        return self

    def stage(self, name):
        getattr(lib(), "sn_oracle_" + name)(self.h)
        return self

    def close(self):
        if self.h:
            lib().sn_oracle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_occ(self):
        return lib().sn_oracle_n_occ(self.h)

    def good_len(self):
        return _arr(lib().sn_oracle_good_len(self.h), self.n, np.uint32)

    def kmers(self):
        """(n,6) u32: w0,w1,w2,count,ctx_before_prune,ctx_after_prune, sorted by k-mer."""
        n = lib().sn_oracle_n_kmers(self.h)
        out = np.zeros((n, 6), dtype=np.uint32)
        lib().sn_oracle_get_kmers(self.h, out.ctypes.data)
        return out

    def edges(self):
        L = lib()
        out = []
        for e in range(L.sn_oracle_n_edges(self.h)):
            n = L.sn_oracle_edge_len(self.h, e)
            out.append(_arr(L.sn_oracle_edge_seq(self.h, e), n, np.uint8).tobytes())
        return out

    def paths(self):
        L = lib()
        off = _arr(L.sn_oracle_path_off(self.h), self.n + 1, np.uint64)
        return (_arr(L.sn_oracle_path_offset(self.h), self.n, np.int32), off,
                _arr(L.sn_oracle_path_edges(self.h), int(off[-1]), np.int32))

    def xlat(self):
        L = lib()
        n = L.sn_oracle_n_edges(self.h)
        return _arr(L.sn_oracle_fwd_xlat(self.h), n, np.int32), _arr(L.sn_oracle_rev_xlat(self.h), n, np.int32)

    def involution(self):
        n = lib().sn_oracle_n_hbv_edges(self.h)
        inv = np.zeros(n, dtype=np.int32)
        lib().sn_oracle_involution(self.h, inv.ctypes.data)
        return inv

    def write_hbv(self, path):
        assert lib().sn_oracle_write_hbv(self.h, path.encode()) == 0

    def write_paths(self, path):
        assert lib().sn_oracle_write_paths(self.h, path.encode()) == 0
