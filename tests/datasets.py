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
    if name == "onebc":              # every barcoded read carries the same barcode: only ignBcBelow lets a k-mer pass min_bc = 2
        return synth.make_stress(5, n_pairs=3000, n_bc=1)
    if name == "fewbc":              # three barcodes: many k-mers are seen under one barcode only
        return synth.make_stress(6, n_pairs=3000, n_bc=3)
    if name == "repeats":
        # a repeat family: 4,000 copies of a 300-bp element, every copy diverged by 8 % substitutions, back to back with
        # 60-bp unique spacers; plus poly-A and a (CA)n microsatellite.  The element's minimizers name a handful of buckets
        # that receive ~10^5 distinct k-mers each (heavy buckets: k_bucket_count2 shares them out over many CTAs).
        rng = np.random.Generator(np.random.Philox(key=77))
        elem = rng.integers(0, 4, size=300, dtype=np.uint8)
        parts = []
        for _ in range(4000):
            e = elem.copy()
            m = rng.random(300) < 0.08
            e[m] = (e[m] + rng.integers(1, 4, size=int(m.sum()))) & 3
            parts += [e, rng.integers(0, 4, size=60, dtype=np.uint8)]
        parts += [np.zeros(400, np.uint8), rng.integers(0, 4, size=200, dtype=np.uint8), np.tile(np.array([1, 0], np.uint8), 300), rng.integers(0, 4, size=500, dtype=np.uint8)]
        hap = np.concatenate(parts)
        G = len(hap)
        synth._HAP = np.stack([hap, hap])
        n_pairs = 20 * G // 300
        blk, q = synth._gen_chunk((G, 5, 150, 0, n_pairs, 0))
        synth._HAP = None
        bcid = np.sort(rng.integers(0, 2000, size=n_pairs))
        new = np.ones(n_pairs, bool); new[1:] = bcid[1:] != bcid[:-1]
        bc = np.repeat(np.cumsum(new).astype(np.int32), 2)
        return blk.ravel(), q.ravel(), np.arange(2 * n_pairs + 1, dtype=np.uint64) * 150, bc, None
    if name == "dupes":
        return _dupes()
    if name == "polyA":
        # one k-mer (A^48) with more than 2^24 occurrences: 84,000 pairs of 150 x A give 17.3 M of them --
        # the count saturates at 16,777,215 (kmers/ReadPather.h:128-129,145) -- on top of a stress set, with
        # four barcodes on the homopolymer reads.  Every poly-A super-k-mer lands in ONE minimizer bucket.
        b, q, off, bc, ids = synth.make_stress(11, n_pairs=1500, n_bc=20)
        n_a, L = 2 * 84_000, 150
        b = np.concatenate([b, np.zeros(n_a * L, np.uint8)])
        q = np.concatenate([q, np.full(n_a * L, 37, np.uint8)])
        off = np.concatenate([off, off[-1] + np.arange(1, n_a + 1, dtype=np.uint64) * L])
        top = int(bc.max())
        bc = np.concatenate([bc, (top + 1 + (np.arange(n_a) // (n_a // 4)).clip(0, 3)).astype(np.int32)])
        return b, q, off, bc, None
    G, pairs, nbc, seed = synth.CONFIGS[name]
    b, q, bc, ids = synth.make_reads(G, pairs, nbc, seed)
    n, L = b.shape
    return b.ravel(), q.ravel(), np.arange(n + 1, dtype=np.uint64) * L, bc, ids


def _dupes():
    """Duplicate read pairs for MarkDups (10X/SecretOps.cc:599-774), on one 70-kb unique contig tiled by error-free reads
    (run with MIN_FREQ=1, MIN_BC=0: DUPES_PARAMS): exact copies (equal quality sums: the tie and 'artifactual duplicate'
    branches), copies with better or worse qualities, copies whose partner differs after its first five bases, a triple,
    and two placements 65,536 bases apart on the same edge whose partners start alike -- ReadPathX keeps the offset in 16 bits
    (10X/paths/ReadPathParser.cc:31), so the reference calls them duplicates."""
    rng = np.random.Generator(np.random.Philox(key=4242))
    C = rng.integers(0, 4, size=70_000, dtype=np.uint8)

    def rc(x):
        return (3 - x[::-1]).astype(np.uint8)

    def q_of(n):
        return np.where(rng.random(n) < 0.05, 30, 37).astype(np.uint8)
    pairs = []                                       # (r1, q1, r2, q2)
    for i in range(0, 70_000 - 150 + 1, 100):
        j = i + 200 if i + 350 <= 70_000 else int(rng.integers(0, 69_000))
        pairs.append([C[i:i + 150].copy(), q_of(150), rc(C[j:j + 150]), q_of(150)])
    extra = []
    for t in range(40):                              # exact copies: ties
        a = pairs[5 + 7 * t]
        extra.append([x.copy() for x in a])
    for t in range(20):                              # the copy is better / worse by one quality value
        a = [x.copy() for x in pairs[300 + 3 * t]]
        a[1][10] = 38 if t % 2 == 0 else 20
        extra.append(a)
    for t in range(20):                              # partner read differs after its first five bases
        a = [x.copy() for x in pairs[400 + 5 * t]]
        a[2] = np.concatenate([a[2][:5], rng.integers(0, 4, size=145, dtype=np.uint8)])   # (unrelated sequence: no branch off the contig)
        extra.append(a)
    for t in range(5):                               # a triple: two more exact copies
        a = pairs[600 + 4 * t]
        extra += [[x.copy() for x in a], [x.copy() for x in a]]
    for t in range(6):                               # 65,536 apart, partners starting with the same five bases
        p0 = 1000 + 500 * t
        head = rc(C[p0 + 200:p0 + 350])[:5]
        k = next(k for k in range(p0 + 65_536 + 160, 69_850) if np.array_equal(rc(C[k:k + 150])[:5], head))
        extra.append([C[p0:p0 + 150].copy(), q_of(150), rc(C[p0 + 200:p0 + 350]), q_of(150)])
        extra.append([C[p0 + 65_536:p0 + 65_686].copy(), q_of(150), rc(C[k:k + 150]), q_of(150)])
    n0 = len(pairs)
    bcid = np.sort(rng.integers(0, 30, size=n0))
    bcid[:25] = -1                                   # some unbarcoded pairs (they stay in front: sorted input)
    bcid = np.concatenate([bcid, 30 + np.arange(len(extra)) // 3])
    allp = pairs + extra
    new = np.ones(len(allp), bool); new[1:] = bcid[1:] != bcid[:-1]
    new &= bcid >= 0
    ordinal = np.cumsum(new).astype(np.int32); ordinal[bcid < 0] = 0
    codes = np.concatenate([x for a in allp for x in (a[0], a[2])])
    quals = np.concatenate([x for a in allp for x in (a[1], a[3])])
    off = np.arange(2 * len(allp) + 1, dtype=np.uint64) * 150
    return codes, quals, off, np.repeat(ordinal, 2), np.repeat(bcid, 2)


DUPES_PARAMS = dict(min_qual=7, min_freq=1, min_bc=0)


def unpack_edges(ln, off, packed):
    out = []
    for e in range(len(ln)):
        b = packed[int(off[e]):int(off[e]) + (int(ln[e]) + 3) // 4]
        codes = np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).ravel()[:ln[e]]
        out.append(codes.astype(np.uint8).tobytes())
    return out
