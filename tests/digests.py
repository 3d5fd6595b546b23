"""Digests used by the scale parity tests and by bench.py's --check leg (test infrastructure).

kmer_digest is order independent, so a k-mer table can be compared without sorting it: the reference's
kmers.kvec is in thread-race order, the product's dictionary in (bucket, hash, k-mer) order."""
import hashlib

import numpy as np


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays"""
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def kmer_digest(w0, w1, w2, count_ctx, chunk=1 << 24):
    """{n, sum, xor} over h(entry) for entries {w0,w1,w2,count:24|ctx<<24} (all u32 arrays)."""
    n = len(w0)
    s = np.uint64(0)
    x = np.uint64(0)
    with np.errstate(over="ignore"):
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            hi = (w0[a:b].astype(np.uint64) << np.uint64(32)) | w1[a:b].astype(np.uint64)
            lo = (w2[a:b].astype(np.uint64) << np.uint64(32)) | count_ctx[a:b].astype(np.uint64)
            h = _mix64(_mix64(hi) + lo * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1))
            s = s + h.sum(dtype=np.uint64)
            x = x ^ np.bitwise_xor.reduce(h)
    return {"n": int(n), "sum": "%016x" % int(s), "xor": "%016x" % int(x)}


def file_md5(path):
    m = hashlib.md5()
    with open(path, "rb") as f:
        while True:
            blk = f.read(1 << 24)
            if not blk:
                break
            m.update(blk)
    return m.hexdigest()


def bytes_md5(*arrays):
    m = hashlib.md5()
    for a in arrays:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()


def paths_file_bytes(offset, poff, edges):
    """The feudal ReadPathVec file (tmp.paths) of n reads: FCB, per read {i32 offset, u32 lastSkip = 0, i32 edges[]},
    then n+1 absolute u64 offsets (paths/long/ReadPath.h:61-63, feudal/FeudalFileWriter.cc:100-121; header
    (n, 1, 0, 24, 4, varTab, fixedOff) as observed on the reference's files)."""
    offset = np.ascontiguousarray(offset, np.int32)
    poff = np.ascontiguousarray(poff, np.int64)
    edges = np.ascontiguousarray(edges, np.int32)
    n, m = len(offset), len(edges)
    plen = np.diff(poff)
    start = 2 * np.arange(n, dtype=np.int64) + poff[:-1]           # first int of every record
    var = np.zeros(2 * n + m, np.int32)
    var[start] = offset
    if m:
        var[np.arange(m, dtype=np.int64) + 2 * (np.repeat(np.arange(n, dtype=np.int64), plen) + 1)] = edges
    tab = np.empty(n + 1, np.uint64)
    tab[:n] = 24 + 4 * start
    tab[n] = 24 + 4 * (2 * n + m)
    var_tab = 24 + 4 * (2 * n + m)
    fcb = np.uint32(n).tobytes() + bytes([1, 0, 24, 4]) + np.uint64(var_tab).tobytes() + np.uint64(var_tab + 8 * (n + 1)).tobytes()
    return fcb + var.tobytes() + tab.tobytes()
