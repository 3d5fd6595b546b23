"""ctypes driver for tests/hostsim (CPU execution of the product's host+device logic)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_L = None


def lib():
    global _L
    if _L is None:
        so = os.path.join(HERE, "libhostsim.so")
        src = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(ROOT, "supernova_b200", "csrc", f) for f in
                                                       ("sn_hbv.cpp", "sn_formats.cpp", "sn_kmer.cuh", "sn_graph.cuh", "sn_path.cuh", "sn_hbv.h", "sn_msp.cuh", "sn_dfside.cuh", "sn_synth.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-DSN_HOSTSIM", "-o", so] + src[:3] + ["-lz"])
        L = C.CDLL(so)
        vp, u64 = C.c_void_p, C.c_uint64
        L.hs_new.restype = vp
        L.hs_new.argtypes = [u64, vp, C.c_int]
        L.hs_free.argtypes = [vp]
        L.hs_prune.argtypes = [vp]
        L.hs_edges.argtypes = [vp]
        L.hs_n_edges.restype = u64
        L.hs_n_edges.argtypes = [vp]
        L.hs_edges_bytes.restype = u64
        L.hs_edges_bytes.argtypes = [vp]
        L.hs_get_edges.argtypes = [vp, vp, vp, vp]
        L.hs_get_graph_info.argtypes = [vp, vp, vp, vp]
        L.hs_hbv.argtypes = [vp, C.c_char_p]
        L.hs_paths.argtypes = [vp, u64, vp, vp, vp, vp, vp, C.c_char_p]
        L.hs_pathsx.argtypes = [vp, C.c_char_p]
        L.hs_window_bucket.argtypes = [C.c_uint32] * 4
        L.hs_window_bucket.restype = C.c_uint32
        L.hs_synth_reads.argtypes = [u64, u64, C.c_uint32, u64, vp, u64, u64, vp, vp, vp]
        L.hs_mark_dups.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.hs_extract_read.argtypes = [vp, C.c_uint32, C.c_int32, vp]
        L.hs_extract_read.restype = C.c_uint32
        L.hs_msp_read.argtypes = [vp, C.c_uint32, C.c_int32, vp, vp, vp, vp]
        L.hs_msp_read.restype = C.c_uint32
        L.hs_sk_expand.argtypes = [vp, u64, vp]
        L.hs_sk_expand.restype = u64
        _L = L
    return _L


class HostSim:
    def __init__(self, recs, bits=10):
        self.recs = np.ascontiguousarray(recs, dtype=np.uint32)
        self.n = len(self.recs)
        self.h = lib().hs_new(self.n, self.recs.ctypes.data, bits)

    def __del__(self):
        if getattr(self, "h", None):
            lib().hs_free(self.h)
            self.h = None

    def prune(self):
        lib().hs_prune(self.h)

    def edges(self):
        rc = lib().hs_edges(self.h)
        assert rc == 0, "an entry was left without an edge"
        n = lib().hs_n_edges(self.h)
        ln = np.zeros(n, np.uint32); off = np.zeros(n + 1, np.uint64); packed = np.zeros(lib().hs_edges_bytes(self.h), np.uint8)
        lib().hs_get_edges(self.h, ln.ctypes.data, off.ctypes.data, packed.ctypes.data)
        return ln, off, packed

    def graph_info(self):
        ctx = np.zeros(self.n, np.uint8); edge = np.zeros(self.n, np.uint32); off = np.zeros(self.n, np.uint32)
        lib().hs_get_graph_info(self.h, ctx.ctypes.data, edge.ctypes.data, off.ctypes.data)
        return ctx, edge, off

    def hbv(self, path):
        assert lib().hs_hbv(self.h, path.encode()) == 0

    def paths(self, bases, boff, ln, quals, qoff, path):
        rc = lib().hs_paths(self.h, len(ln), bases.ctypes.data, boff.ctypes.data, ln.ctypes.data, quals.ctypes.data, qoff.ctypes.data,
                            path.encode())
        assert rc == 0, rc

    def pathsx(self, path):
        assert lib().hs_pathsx(self.h, path.encode()) == 0

    def mark_dups(self, bases, boff, ln, quals, qoff, bc):
        n = len(ln)
        dup = np.zeros(n // 2, np.uint8); art = np.zeros(n // 2, np.uint8); counts = np.zeros(2, np.uint64)
        bc = np.ascontiguousarray(bc, dtype=np.int32)
        rc = lib().hs_mark_dups(self.h, n, bases.ctypes.data, boff.ctypes.data, ln.ctypes.data, quals.ctypes.data, qoff.ctypes.data,
                                bc.ctypes.data, dup.ctypes.data, art.ctypes.data, counts.ctypes.data)
        assert rc == 0, rc
        return dup, art, int(counts[0]), int(counts[1])
