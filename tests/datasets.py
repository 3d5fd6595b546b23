"""Seeded data sets shared by the parity tests."""
import functools
import numpy as np
from supernova_b200 import synth


@functools.lru_cache(maxsize=None)
def get(name):
    """-> (codes, quals, off, bc, bc_ids) ragged."""
    if name.startswith("stress"):
        return synth.make_stress(int(name[6:]))
    if name == "empty_kmers":        # every read too short / too low quality: empty dictionary
        rng = np.random.default_rng(3)
        n, L = 64, 60
        b = rng.integers(0, 4, size=n * L, dtype=np.uint8)
        q = np.full(n * L, 3, np.uint8)
        return b, q, np.arange(n + 1, dtype=np.uint64) * L, np.repeat(np.arange(1, n // 2 + 1, dtype=np.int32), 2), None
    if name == "nobc":               # all reads unbarcoded: nothing passes the barcode rule
        b, q, off, bc, ids = synth.make_stress(9, n_pairs=500)
        return b, q, off, np.zeros_like(bc), ids
    G, pairs, nbc, seed = synth.CONFIGS[name]
    b, q, bc, ids = synth.make_reads(G, pairs, nbc, seed)
    n, L = b.shape
    return b.ravel(), q.ravel(), np.arange(n + 1, dtype=np.uint64) * L, bc, ids


def unpack_edges(ln, off, packed):
    out = []
    for e in range(len(ln)):
        b = packed[int(off[e]):int(off[e]) + (int(ln[e]) + 3) // 4]
        codes = np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).ravel()[:ln[e]]
        out.append(codes.astype(np.uint8).tobytes())
    return out
