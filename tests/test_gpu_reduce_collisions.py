"""-m gpu: the reduce kernel's rare paths -- k-mers whose 32-bit hashes collide (one run of
equal hash holding several k-mers), runs crossing warp-chunk and step boundaries, very long
runs -- checked against the oracle's counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hash(w0, w1, w2):
    M = 0xFFFFFFFF
    h = (w0 * 0x9E3779B1) & M
    h = (((h ^ (h >> 15)) + w1) * 0x85EBCA77) & M
    h = (((h ^ (h >> 13)) + w2) * 0xC2B2AE3D) & M
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M
    h ^= h >> 16
    return h


def test_long_runs_and_many_kmers(built):
    """A high-copy tandem repeat (runs of thousands of records spanning many warp chunks)
    next to ordinary sequence."""
    import supernova_b200 as sb
    from oracle.oracle import Oracle
    rng = np.random.default_rng(5)
    unit = rng.integers(0, 4, size=53, dtype=np.uint8)
    rep = np.tile(unit, 40)
    uniq = rng.integers(0, 4, size=20000, dtype=np.uint8)
    genome = np.concatenate([uniq[:10000], rep, uniq[10000:]])
    n = 40000
    starts = np.concatenate([rng.integers(10000, 10000 + len(rep) - 150, size=n // 2), rng.integers(0, len(genome) - 150, size=n // 2)])
    codes = np.stack([genome[s:s + 150] for s in starts])
    flip = rng.random(n) < 0.5
    codes[flip] = 3 - codes[flip][:, ::-1]
    quals = np.full_like(codes, 37)
    bc = np.sort(rng.integers(1, 200, size=n)).astype(np.int32)
    off = np.arange(n + 1, dtype=np.uint64) * 150
    o = Oracle(codes.ravel(), quals.ravel(), off, bc).stage("count")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes.ravel(), quals.ravel(), off)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.count_kmers(sb.Params())
        km, ok = ctx.kmers(), o.kmers()
        assert ok[:, 3].max() > 3000          # the repeat really produces long runs
        assert np.array_equal(km[:, :3], ok[:, :3])
        assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
        assert ctx.counts()["n_kmers_distinct"] >= len(ok)


def test_hash_collisions_are_separated(built):
    """Reads built so that several DISTINCT canonical k-mers share one 32-bit hash: the table
    must still hold each of them with its own count."""
    import supernova_b200 as sb
    from oracle.oracle import Oracle
    rng = np.random.default_rng(17)
    # find colliding pairs among random 48-mers restricted to a small hash space: brute force on
    # the low 20 bits is not enough -- the kernel groups by the full 32-bit hash -- so search
    # directly for full collisions among 2^17.5 candidates sharing a fixed 32-base prefix
    # (birthday bound ~ 2^16 for 32 bits).
    prefix = rng.integers(0, 4, size=32, dtype=np.uint8)
    m = 400_000
    tails = rng.integers(0, 4, size=(m, 16), dtype=np.uint8)
    kmers = np.concatenate([np.broadcast_to(prefix, (m, 32)), tails], axis=1)
    # canonical form + words
    def words(c):
        w = []
        for i in range(3):
            v = np.zeros(len(c), dtype=np.uint64)
            for j in range(16):
                v = (v << np.uint64(2)) | c[:, 16 * i + j].astype(np.uint64)
            w.append(v)
        return w
    rc = 3 - kmers[:, ::-1]
    fw, rw = words(kmers), words(rc)
    less = (fw[0] < rw[0]) | ((fw[0] == rw[0]) & ((fw[1] < rw[1]) | ((fw[1] == rw[1]) & (fw[2] <= rw[2]))))
    cw = [np.where(less, f, r) for f, r in zip(fw, rw)]
    hs = np.array([_hash(int(a), int(b), int(c)) for a, b, c in zip(cw[0][:m], cw[1][:m], cw[2][:m])], dtype=np.uint64) if m <= 1000 else None
    # vectorised hash
    M = np.uint64(0xFFFFFFFF)
    h = (cw[0] * np.uint64(0x9E3779B1)) & M
    h = (((h ^ (h >> np.uint64(15))) + cw[1]) * np.uint64(0x85EBCA77)) & M
    h = (((h ^ (h >> np.uint64(13))) + cw[2]) * np.uint64(0xC2B2AE3D)) & M
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & M
    h ^= h >> np.uint64(16)
    order = np.argsort(h, kind="stable")
    hsort = h[order]
    dup = np.nonzero(hsort[1:] == hsort[:-1])[0]
    assert len(dup) >= 3, "no 32-bit collisions found among the candidates"
    pick = set()
    for d in dup[:12]:
        pick.add(int(order[d])); pick.add(int(order[d + 1]))
    pick = sorted(pick)
    # each chosen 48-mer is embedded in its own 150-base read context, repeated with 2 barcodes
    reads, bcs = [], []
    for t, ki in enumerate(pick):
        left = rng.integers(0, 4, size=51, dtype=np.uint8)
        right = rng.integers(0, 4, size=51, dtype=np.uint8)
        rd = np.concatenate([left, kmers[ki], right])
        copies = 3 + t % 4
        for c in range(copies):
            reads.append(rd if c % 2 == 0 else (3 - rd[::-1]).astype(np.uint8))
            bcs.append(1 + (c % 2))
    order2 = np.argsort(bcs, kind="stable")
    codes = np.stack(reads)[order2]
    bc = np.array(bcs, dtype=np.int32)[order2]
    quals = np.full_like(codes, 37)
    n = len(codes)
    off = np.arange(n + 1, dtype=np.uint64) * 150
    o = Oracle(codes.ravel(), quals.ravel(), off, bc).stage("count")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes.ravel(), quals.ravel(), off)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.count_kmers(sb.Params())
        km, ok = ctx.kmers(), o.kmers()
        # the colliding k-mers are really in the table
        hk = np.array([_hash(int(a), int(b), int(c)) for a, b, c in ok[:, :3]])
        assert len(np.unique(hk)) < len(hk), "the data set does not exercise a collision"
        assert np.array_equal(km[:, :3], ok[:, :3])
        assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))


@pytest.mark.parametrize("min_bc", [2, 1, 0])
def test_duplicate_reads_and_the_barcode_rule(built, min_bc):
    """Identical reads (the same super-k-mers many times over: the count kernel expands one copy
    and carries the number of copies and a summary of their barcodes) under every barcode
    situation Kmerizer::reduce distinguishes: one barcode, two, unbarcoded copies, ignored
    (-1) copies, and mixtures arriving in different orders."""
    import supernova_b200 as sb
    from oracle.oracle import Oracle
    rng = np.random.default_rng(23)
    groups = [
        [5, 5, 5, 5, 5],            # one barcode only
        [5, 5, 6, 6, 5],            # two barcodes
        [0, 0, 0, 0],               # unbarcoded only
        [0, 7, 7, 0, 7],            # unbarcoded + one barcode
        [-1, 8, 8],                 # an ignored copy decides
        [9, 9, -1, 9],
        [0, -1, 0],
        [10, 11, 12, 13, 10, 11],   # many barcodes
        [14, 14],                   # below minFreq
        [15, 16, 0, -1, 15, 16, 17],
        [18, 0, 18, 19],
    ]
    reads, bcs = [], []
    for g, barcodes in enumerate(groups):
        locus = rng.integers(0, 4, size=220, dtype=np.uint8)
        for c, b in enumerate(barcodes):
            rd = locus[:150] if c % 3 != 2 else locus[40:190]          # full copies and a shifted overlapping read
            reads.append(rd if c % 2 == 0 else (3 - rd[::-1]).astype(np.uint8))
            bcs.append(b)
    # shuffle so that copies of one read sit at different places of the bucket
    perm = rng.permutation(len(reads))
    codes = np.stack(reads)[perm]
    bc = np.array(bcs, dtype=np.int32)[perm]
    quals = np.full_like(codes, 37)
    n = len(codes)
    off = np.arange(n + 1, dtype=np.uint64) * 150
    o = Oracle(codes.ravel(), quals.ravel(), off, bc, min_bc=min_bc).stage("count")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes.ravel(), quals.ravel(), off)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.count_kmers(sb.Params(min_bc=min_bc))
        km, ok = ctx.kmers(), o.kmers()
        assert len(ok) > 0
        assert np.array_equal(km[:, :3], ok[:, :3])
        assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
