"""ctypes binding of include/supernova_b200.h.  Mirrors the reference's C++ call
``buildReadQGraph48`` (paths/long/BuildReadQGraph48.h:28-40): load reads/quals/barcodes,
count k-mers, build edges, build the HyperBasevector, path the reads, write ``a.hbv`` and
``tmp.paths``."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsupernova_b200.so")
_LIB = None

STAGES = ("exchange", "ghosts", "edge_dict", "ingest_h2d", "ingest_parse", "h2d", "goodlen", "msp_hist", "msp_scatter", "bucket_count", "make_dict", "prune", "edges", "hbv_dev", "hbv_layout", "hbv_host", "hbv_csr", "path", "paths_index", "pathsx", "mark_dups")


class SnError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("min_qual", C.c_uint32), ("min_freq", C.c_uint32), ("min_bc", C.c_uint32), ("ign_bc_below", C.c_int64)]

    def __init__(self, min_qual=7, min_freq=3, min_bc=2, ign_bc_below=0):
        super().__init__(min_qual, min_freq, min_bc, ign_bc_below)


class SynthSpec(C.Structure):
    _fields_ = [("genome_bases", C.c_uint64), ("total_pairs", C.c_uint64), ("seed", C.c_uint64), ("n_barcodes", C.c_uint32)]


class DupStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_pairs", "n_dup_pairs", "n_dups", "n_interdups", "n_art_pairs")]


class Counts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_reads", "n_bases", "n_kmer_occurrences", "n_kmers_distinct", "n_kmers",
                                          "n_edges", "n_edge_bases", "n_hbv_vertices", "n_hbv_edges", "n_path_edges", "n_superkmers")]


def lib():
    """Loads the C-ABI library; raises if it has not been built (no silent fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            raise SnError(f"{SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(SO)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.sn_ctx_create.argtypes = [C.POINTER(vp), i32]
        L.sn_ctx_destroy.argtypes = [vp]
        L.sn_last_error.argtypes = [vp]
        L.sn_last_error.restype = C.c_char_p
        L.sn_load_reads.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp]
        L.sn_load_reads_q8.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp]
        L.sn_write_to_left_right.argtypes = [vp, C.c_char_p, C.c_char_p]
        L.sn_build_paths_index.argtypes = [vp]
        L.sn_get_paths_index.argtypes = [vp, vp, vp, vp]
        L.sn_write_paths_index.argtypes = [vp, C.c_char_p, C.c_char_p]
        L.sn_load_fasth_text.argtypes = [vp, vp, u64]
        L.sn_load_fasth_file.argtypes = [vp, C.c_char_p]
        L.sn_load_fasth_files.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32]
        L.sn_save_read_files.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p]
        L.sn_load_reads_streamed.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.sn_load_read_files.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p]
        L.sn_load_read_files_bc.argtypes = [vp, C.c_char_p, C.c_char_p, vp, u64]
        L.sn_load_read_files_range.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, u64, u64]
        L.sn_write_paths_arrays.argtypes = [C.c_char_p, u64, vp, vp, vp]
        L.sn_build_graph_from_edges.argtypes = [vp, C.c_char_p]
        L.sn_count_kmers.argtypes = [vp, C.POINTER(Params)]
        for f in ("sn_build_edges", "sn_build_hbv", "sn_path_reads"):
            getattr(L, f).argtypes = [vp]
        L.sn_get_counts.argtypes = [vp, C.POINTER(Counts)]
        L.sn_get_good_lengths.argtypes = [vp, vp]
        L.sn_get_kmers.argtypes = [vp, vp]
        L.sn_get_kmer_graph_info.argtypes = [vp, vp, vp, vp]
        L.sn_get_edges.argtypes = [vp, vp, vp, vp]
        L.sn_get_edges_bytes.argtypes = [vp, C.POINTER(u64)]
        L.sn_get_hbv.argtypes = [vp] + [vp] * 9
        L.sn_get_paths.argtypes = [vp, vp, vp, vp]
        L.sn_build_pathsx.argtypes = [vp]
        L.sn_generate_reads.argtypes = [vp, C.POINTER(SynthSpec), u64, u64, vp]
        L.sn_set_semantics.argtypes = [vp, i32]
        L.sn_get_pathsx.argtypes = [vp, C.POINTER(u64), C.POINTER(vp), C.POINTER(u64), C.POINTER(vp)]
        L.sn_mark_dups.argtypes = [vp, C.POINTER(DupStats)]
        L.sn_get_dups.argtypes = [vp, vp, vp]
        for f in ("sn_write_hbv", "sn_write_paths", "sn_write_edges_bv", "sn_write_inv", "sn_write_kmer_spectrum", "sn_write_pathsx", "sn_write_dup",
                  "sn_write_hbx", "sn_write_edges_fastb", "sn_write_kmers", "sn_write_k"):
            getattr(L, f).argtypes = [vp, C.c_char_p]
        L.sn_build_read_qgraph48.argtypes = [vp, C.c_char_p, C.POINTER(Params), i32, i32]
        L.sn_stage_ms.argtypes = [vp, C.c_char_p]
        L.sn_stage_ms.restype = C.c_double
        L.sn_kernel_launches.argtypes = [vp]
        L.sn_kernel_launches.restype = u64
        L.sn_stream.argtypes = [vp]
        L.sn_stream.restype = vp
        L.sn_pqvec_encode.argtypes = [vp, C.c_uint32, vp]
        L.sn_pqvec_encode.restype = u64
        L.sn_pqvec_decode.argtypes = [vp, u64, vp, C.c_uint32]
        L.sn_pqvec_decode.restype = C.c_uint32
        L.sn_pack_reads.argtypes = [u64, vp, vp, vp, i32] + [C.POINTER(vp)] * 5
        L.sn_free.argtypes = [vp]
        L.sn_write_read_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, u64, vp, vp, vp, vp, vp, vp]
        L.sn_device_count.restype = i32
        L.sn_msp_bucket_bits.argtypes = [u64]
        L.sn_msp_bucket_bits.restype = i32
        L.sn_nccl_unique_id.argtypes = [vp]
        L.sn_comm_init_nccl.argtypes = [vp, i32, i32, vp]
        L.sn_local_group_create.argtypes = [i32]
        L.sn_local_group_create.restype = vp
        L.sn_local_group_destroy.argtypes = [vp]
        L.sn_local_group_abort.argtypes = [vp]
        L.sn_comm_init_local.argtypes = [vp, vp, i32]
        L.sn_comm_free.argtypes = [vp]
        L.sn_mg_build_graph.argtypes = [vp, C.POINTER(Params), i32]
        L.sn_mg_dict_is_sharded.argtypes = [vp]
        _LIB = L
    return _LIB


def nccl_unique_id():
    """the 128 bytes of ncclGetUniqueId (rank 0 makes it, the launcher hands it to every rank)"""
    buf = (C.c_char * 128)()
    if lib().sn_nccl_unique_id(buf):
        raise SnError("sn_nccl_unique_id: " + lib().sn_last_error(None).decode())
    return bytes(buf)


def run_local_ranks(n_ranks, fn, device=0):
    """Test infrastructure: n ranks as n contexts of this process on ONE device, a host thread per rank
    (sn_comm_init_local).  fn(rank, ctx) runs on every rank; returns the list of its results."""
    import threading
    L = lib()
    group = L.sn_local_group_create(n_ranks)
    ctxs = [Context(device) for _ in range(n_ranks)]
    out, errs = [None] * n_ranks, [None] * n_ranks

    def work(r):
        try:
            ctxs[r].comm_init_local(group, r)
            out[r] = fn(r, ctxs[r])
        except BaseException as e:          # noqa: BLE001  (a failing rank must not leave the others waiting in a collective)
            errs[r] = e
            L.sn_local_group_abort(group)
    th = [threading.Thread(target=work, args=(r,)) for r in range(n_ranks)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in ctxs:
        c.close()
    L.sn_local_group_destroy(group)
    first = [e for e in errs if e is not None and "another rank of the local group failed" not in str(e)] or [e for e in errs if e is not None]
    if first:
        raise first[0]
    return out


def _p(a):
    return None if a is None else a.ctypes.data


def _copy(ptr, nbytes, dtype):
    if nbytes == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_char * nbytes).from_address(ptr), dtype=dtype).copy()


def pack_reads(codes, quals, off, threads=None):
    """Base codes + Phred bytes (ragged, element offsets ``off``) -> the reference's in-memory
    layout: .fastb packed bases + byte offsets + lengths, .qualp PQVec blob + byte offsets."""
    L = lib()
    codes = np.ascontiguousarray(codes, dtype=np.uint8).ravel()
    quals = np.ascontiguousarray(quals, dtype=np.uint8).ravel()
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = len(off) - 1
    threads = threads or min(64, os.cpu_count() or 1)
    outs = [C.c_void_p() for _ in range(5)]
    rc = L.sn_pack_reads(n, _p(codes), _p(quals), _p(off), threads, *[C.byref(o) for o in outs])
    if rc:
        raise SnError("sn_pack_reads failed")
    boff = _copy(outs[1].value, 8 * (n + 1), np.uint64)
    ln = _copy(outs[2].value, 4 * n, np.uint32)
    pqoff = _copy(outs[4].value, 8 * (n + 1), np.uint64)
    bases = _copy(outs[0].value, int(boff[-1]), np.uint8)
    pq = _copy(outs[3].value, int(pqoff[-1]), np.uint8)
    for o in outs:
        L.sn_free(o)
    return bases, boff, ln, pq, pqoff


def write_read_files(head, bases, boff, ln, pq, pqoff, bc):
    bc32 = None if bc is None else np.ascontiguousarray(bc, dtype=np.int32)
    rc = lib().sn_write_read_files((head + ".fastb").encode(), (head + ".qualp").encode(), (head + ".bci").encode(),
                                   len(ln), _p(bases), _p(boff), _p(ln), _p(pq), _p(pqoff), _p(bc32))
    if rc:
        raise SnError("sn_write_read_files failed: " + lib().sn_last_error(None).decode())


def pqvec_encode(q):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    out = np.zeros(2 * len(q) + 8, dtype=np.uint8)
    n = lib().sn_pqvec_encode(_p(q), len(q), _p(out))
    return out[:n].copy()


def pqvec_decode(pq, cap):
    pq = np.ascontiguousarray(pq, dtype=np.uint8)
    out = np.zeros(cap, dtype=np.uint8)
    n = lib().sn_pqvec_decode(_p(pq), len(pq), _p(out), cap)
    return out[:min(n, cap)], n


class Context:
    """One GPU's worth of the hot path (``sn_ctx``)."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.sn_ctx_create(C.byref(h), device)
        if rc:
            raise SnError(self.L.sn_last_error(None).decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.sn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise SnError(f"[{rc}] " + self.L.sn_last_error(self.h).decode())

    # ---- ingest -------------------------------------------------------------------
    def load_reads(self, bases, boff, ln, pq, pqoff, bc):
        a = [np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(boff, np.uint64), np.ascontiguousarray(ln, np.uint32),
             np.ascontiguousarray(pq, np.uint8), np.ascontiguousarray(pqoff, np.uint64),
             None if bc is None else np.ascontiguousarray(bc, np.int32)]
        self._ck(self.L.sn_load_reads(self.h, len(a[2]), *[_p(x) for x in a]))

    def load_reads_ptr(self, n, bases, boff, ln, pq, pqoff, bc):
        """Raw host addresses (e.g. pinned torch tensors)."""
        self._ck(self.L.sn_load_reads(self.h, n, bases, boff, ln, pq, pqoff, bc))

    def load_reads_streamed_ptr(self, n, bases, boff, ln, pq, pqoff, bc, params=None, with_hist=True):
        """sn_load_reads_streamed: chunked copies with the good lengths and the first MSP pass under them."""
        params = params or Params()
        self._ck(self.L.sn_load_reads_streamed(self.h, n, bases, boff, ln, pq, pqoff, bc, C.byref(params), 1 if with_hist else 0))

    def load_fasth_text(self, text):
        """sn_load_fasth_text: barcoded pseudo-FASTQ text (bytes or a uint8 array) parsed, packed and PQVec-encoded on the device."""
        a = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        self._ck(self.L.sn_load_fasth_text(self.h, _p(a), a.size))

    def load_fasth_ptr(self, ptr, n_bytes):
        self._ck(self.L.sn_load_fasth_text(self.h, ptr, n_bytes))

    def load_fasth_file(self, path):
        self._ck(self.L.sn_load_fasth_file(self.h, path.encode()))

    def load_fasth_files(self, paths):
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        self._ck(self.L.sn_load_fasth_files(self.h, arr, len(paths)))

    def save_read_files(self, head):
        self._ck(self.L.sn_save_read_files(self.h, (head + ".fastb").encode(), (head + ".qualp").encode(), (head + ".bci").encode()))

    def load_reads_q8(self, bases, boff, ln, quals, qoff, bc):
        a = [np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(boff, np.uint64), np.ascontiguousarray(ln, np.uint32),
             np.ascontiguousarray(quals, np.uint8), np.ascontiguousarray(qoff, np.uint64),
             None if bc is None else np.ascontiguousarray(bc, np.int32)]
        self._ck(self.L.sn_load_reads_q8(self.h, len(a[2]), *[_p(x) for x in a]))

    def load_read_files(self, head):
        self._ck(self.L.sn_load_read_files(self.h, (head + ".fastb").encode(), (head + ".qualp").encode(), (head + ".bci").encode()))

    def load_read_files_bc(self, head, bc):
        """reads.fastb / reads.qualp + the per-read barcode ordinals held in memory (or None)"""
        b = None if bc is None else np.ascontiguousarray(bc, np.int32)
        self._ck(self.L.sn_load_read_files_bc(self.h, (head + ".fastb").encode(), (head + ".qualp").encode(), _p(b), 0 if b is None else len(b)))

    def build_graph_from_edges(self, bv_path):
        """buildGraphFromMSP up to its pathReads call: edge file -> HBV + dictionary of the edge k-mers"""
        self._ck(self.L.sn_build_graph_from_edges(self.h, bv_path.encode()))

    # ---- stages ---------------------------------------------------------------------
    def count_kmers(self, params=None):
        p = params or Params()
        self._ck(self.L.sn_count_kmers(self.h, C.byref(p)))

    def build_edges(self):
        self._ck(self.L.sn_build_edges(self.h))

    def build_hbv(self):
        self._ck(self.L.sn_build_hbv(self.h))

    def path_reads(self):
        self._ck(self.L.sn_path_reads(self.h))

    def build_read_qgraph48(self, work_dir=None, params=None, with_paths=True, write_files=None):
        p = params or Params()
        wf = (work_dir is not None) if write_files is None else write_files
        self._ck(self.L.sn_build_read_qgraph48(self.h, None if work_dir is None else work_dir.encode(), C.byref(p),
                                               int(with_paths), int(wf)))

    # ---- multi-GPU, collectives inside the library (sn_multi.cu) -----------------------------
    def comm_init_nccl(self, rank, n_ranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.sn_comm_init_nccl(self.h, rank, n_ranks, buf))

    def comm_init_local(self, group, rank):
        self._ck(self.L.sn_comm_init_local(self.h, group, rank))

    def mg_build_graph(self, params=None, with_paths=False):
        p = params or Params()
        self._ck(self.L.sn_mg_build_graph(self.h, C.byref(p), int(with_paths)))

    def dict_is_sharded(self):
        return bool(self.L.sn_mg_dict_is_sharded(self.h))

    # ---- results ----------------------------------------------------------------------
    def counts(self):
        c = Counts()
        self._ck(self.L.sn_get_counts(self.h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in Counts._fields_}

    def good_lengths(self):
        out = np.zeros(self.counts()["n_reads"], dtype=np.uint32)
        self._ck(self.L.sn_get_good_lengths(self.h, _p(out)))
        return out

    def kmers(self):
        """(n,4) u32: w0,w1,w2,count|ctx<<24, sorted by k-mer."""
        out = np.zeros((self.counts()["n_kmers"], 4), dtype=np.uint32)
        self._ck(self.L.sn_get_kmers(self.h, _p(out)))
        return out

    def kmer_graph_info(self):
        n = self.counts()["n_kmers"]
        ctx = np.zeros(n, np.uint8); edge = np.zeros(n, np.uint32); off = np.zeros(n, np.uint32)
        self._ck(self.L.sn_get_kmer_graph_info(self.h, _p(ctx), _p(edge), _p(off)))
        return ctx, edge, off

    def edges(self):
        n = self.counts()["n_edges"]
        nb = C.c_uint64()
        self._ck(self.L.sn_get_edges_bytes(self.h, C.byref(nb)))
        ln = np.zeros(n, np.uint32); off = np.zeros(n + 1, np.uint64); packed = np.zeros(nb.value, np.uint8)
        self._ck(self.L.sn_get_edges(self.h, _p(ln), _p(off), _p(packed)))
        return ln, off, packed

    def hbv(self):
        c = self.counts()
        nv, nh, ne = c["n_hbv_vertices"], c["n_hbv_edges"], c["n_edges"]
        fs = np.zeros(nv + 1, np.uint32); ts = np.zeros(nv + 1, np.uint32)
        fv, fe, tv, te = (np.zeros(nh, np.int32) for _ in range(4))
        fwd = np.zeros(ne, np.int32); rev = np.zeros(ne, np.int32); inv = np.zeros(nh, np.int32)
        self._ck(self.L.sn_get_hbv(self.h, *[_p(x) for x in (fs, fv, fe, ts, tv, te, fwd, rev, inv)]))
        return dict(from_start=fs, from_v=fv, from_e=fe, to_start=ts, to_v=tv, to_e=te, fwd=fwd, rev=rev, inv=inv)

    def paths(self):
        c = self.counts()
        off = np.zeros(c["n_reads"], np.int32); poff = np.zeros(c["n_reads"] + 1, np.uint64); e = np.zeros(c["n_path_edges"], np.int32)
        self._ck(self.L.sn_get_paths(self.h, _p(off), _p(poff), _p(e)))
        return off, poff, e

    def write_to_left_right(self, left, right):
        self._ck(self.L.sn_write_to_left_right(self.h, left.encode(), right.encode()))

    def build_paths_index(self):
        self._ck(self.L.sn_build_paths_index(self.h))

    def paths_index(self):
        c = self.counts()
        off = np.zeros(c["n_hbv_edges"] + 1, np.uint64); ids = np.zeros(c["n_path_edges"], np.uint64); cb = np.zeros(c["n_hbv_edges"], np.int32)
        self._ck(self.L.sn_get_paths_index(self.h, _p(off), _p(ids), _p(cb)))
        return off, ids, cb

    def write_paths_index(self, paths_inv, countsb):
        self._ck(self.L.sn_write_paths_index(self.h, paths_inv.encode(), countsb.encode()))

    def generate_reads(self, G, total_pairs, n_bc, seed, first_pair=0, n_pairs=None):
        """sn_generate_reads: the counter-based synthetic reads (csrc/sn_synth.cuh; twin: synth.make_reads_cb) made on the device"""
        from . import synth
        spec = SynthSpec(int(G), int(total_pairs), int(seed), int(n_bc))
        T = np.ascontiguousarray(synth.cb_error_thresholds(), np.uint32)
        n_pairs = total_pairs - first_pair if n_pairs is None else n_pairs
        self._ck(self.L.sn_generate_reads(self.h, C.byref(spec), int(first_pair), int(n_pairs), _p(T)))

    def set_semantics(self, tada=False):
        """SN_SEM_TADA: reads trimmed to exactly K bases are counted (lib/tada/src/cmd_msp.rs:109-110)"""
        self._ck(self.L.sn_set_semantics(self.h, 1 if tada else 0))

    # ---- DF side: ReadPathVecX, MarkDups, the files next to a.hbv (10X/DF.cc:573-600, 10X/WriteFiles.cc:16-60) ----
    def build_pathsx(self):
        self._ck(self.L.sn_build_pathsx(self.h))

    def pathsx(self):
        """-> (zip_index i64[ceil(n_reads/10)], zipped_data u8[]) -- copies of the context's buffers"""
        ni, nb, pi, pd = C.c_uint64(), C.c_uint64(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.sn_get_pathsx(self.h, C.byref(ni), C.byref(pi), C.byref(nb), C.byref(pd)))
        idx = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_int64)), shape=(ni.value,)).copy() if ni.value else np.zeros(0, np.int64)
        dat = np.ctypeslib.as_array(C.cast(pd, C.POINTER(C.c_uint8)), shape=(nb.value,)).copy() if nb.value else np.zeros(0, np.uint8)
        return idx, dat

    def write_pathsx(self, path):
        self._ck(self.L.sn_write_pathsx(self.h, path.encode()))

    def mark_dups(self):
        st = DupStats()
        self._ck(self.L.sn_mark_dups(self.h, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in DupStats._fields_}

    def dups(self):
        n = self.counts()["n_reads"] // 2
        dup = np.zeros(n, np.uint8); art = np.zeros(n, np.uint8)
        self._ck(self.L.sn_get_dups(self.h, _p(dup), _p(art)))
        return dup, art

    def write_dup(self, path):
        self._ck(self.L.sn_write_dup(self.h, path.encode()))

    def write_hbx(self, path):
        self._ck(self.L.sn_write_hbx(self.h, path.encode()))

    def write_edges_fastb(self, path):
        self._ck(self.L.sn_write_edges_fastb(self.h, path.encode()))

    def write_kmers(self, path):
        self._ck(self.L.sn_write_kmers(self.h, path.encode()))

    def write_k(self, path):
        self._ck(self.L.sn_write_k(self.h, path.encode()))

    def write_hbv(self, path):
        self._ck(self.L.sn_write_hbv(self.h, path.encode()))

    def write_paths(self, path):
        self._ck(self.L.sn_write_paths(self.h, path.encode()))

    def write_edges_bv(self, path):
        self._ck(self.L.sn_write_edges_bv(self.h, path.encode()))

    def write_inv(self, path):
        self._ck(self.L.sn_write_inv(self.h, path.encode()))

    def write_kmer_spectrum(self, path):
        self._ck(self.L.sn_write_kmer_spectrum(self.h, path.encode()))

    def stage_ms(self):
        return {s: self.L.sn_stage_ms(self.h, s.encode()) for s in STAGES}

    def kernel_launches(self):
        return int(self.L.sn_kernel_launches(self.h))

    def stream_ptr(self):
        return int(self.L.sn_stream(self.h) or 0)
