"""Seeded synthetic linked-read generator.

Test/bench infrastructure for the hot path: produces the inputs the reference's
``buildReadQGraph48`` consumes (reads as base codes, Phred quals, per-read barcode
ordinal in ``.bci`` order) plus, for the small configs, the 9-line barcode-sorted
pseudo-FASTQ that the reference's own ``ParseBarcodedFastqs`` ingests
(10X/ParseBarcodedFastqs.cc:3-9,56-146).

Model: haplotype A = i.i.d. uniform ACGT of length G; haplotype B = A with one SNP
per 1000 bp window; pairs drawn uniformly (haplotype, start, insert U[300,500),
strand flip p=0.5); R1 = first L bases of the fragment, R2 = first L bases of its
reverse complement; substitution errors with probability 0.001 + 0.02 (j/L)^3 at
read position j, an erroneous base gets Q in {2,12,20}, a correct one Q37 (5 % Q30);
each pair gets a uniform barcode id, ``unbarcoded_frac`` of the pairs carry no
barcode.  Record order = unbarcoded pairs first, then pairs sorted by barcode, which
is the order ParseBarcodedFastqs writes (``:284-303``).
"""
from __future__ import annotations

import gzip
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genome(G: int, seed: int):
    rng = np.random.Generator(np.random.Philox(key=seed))
    hap_a = rng.integers(0, 4, size=G, dtype=np.uint8)
    hap_b = hap_a.copy()
    nwin = G // 1000
    if nwin:
        pos = np.arange(nwin, dtype=np.int64) * 1000 + rng.integers(0, 1000, size=nwin)
        hap_b[pos] = (hap_a[pos] + rng.integers(1, 4, size=nwin, dtype=np.uint8)) & 3
    return hap_a, hap_b


_HAP = None


def _gen_chunk(args):
    """Reads of pairs [c0,c1): an independent Philox stream keyed by (seed+2, c0)."""
    G, seed, L, c0, c1, shard = args
    hap = _HAP
    perr = (0.001 + 0.02 * (np.arange(L) / L) ** 3).astype(np.float32)
    ar = np.arange(L, dtype=np.int64)
    m = c1 - c0
    crng = np.random.Generator(np.random.Philox(key=[seed + 2 + 7919 * shard, c0]))
    h = crng.integers(0, 2, size=m)
    ins = crng.integers(300, 500, size=m)
    start = (crng.random(m) * (G - ins)).astype(np.int64)
    flip = crng.random(m) < 0.5
    fwd = hap[h[:, None], start[:, None] + ar[None, :]]
    rev = 3 - hap[h[:, None], (start + ins - 1)[:, None] - ar[None, :]]
    r1 = np.where(flip[:, None], rev, fwd)
    r2 = np.where(flip[:, None], fwd, rev)
    blk = np.empty((2 * m, L), dtype=np.uint8)
    blk[0::2] = r1
    blk[1::2] = r2
    err = crng.random((2 * m, L), dtype=np.float32) < perr[None, :]
    sub = crng.integers(1, 4, size=(2 * m, L), dtype=np.uint8)
    blk = np.where(err, (blk + sub) & 3, blk).astype(np.uint8)
    q = np.where(crng.random((2 * m, L), dtype=np.float32) < 0.05, 30, 37).astype(np.uint8)
    qerr = np.array([2, 12, 20], dtype=np.uint8)[crng.integers(0, 3, size=(2 * m, L))]
    q = np.where(err, qerr, q).astype(np.uint8)
    return blk, q


def make_reads(G: int, n_pairs: int, n_bc: int, seed: int, L: int = 150,
               unbarcoded_frac: float = 0.02, chunk: int = 100_000, workers: int = 1, shard: int = 0):
    """Returns (bases[n_reads,L] u8 codes, quals[n_reads,L] u8, bc[n_reads] i32 ordinal,
    bc_ids[n_reads] i64 raw barcode id or -1).  `workers` > 1 generates the chunks in forked
    processes; the result does not depend on it.  `shard` > 0 draws a different, independent
    set of pairs (and barcodes) from the SAME genome (multi-GPU read shards)."""
    global _HAP
    _HAP = np.stack(make_genome(G, seed))
    rng = np.random.Generator(np.random.Philox(key=seed + 1 + 7919 * shard))
    bcid = rng.integers(0, n_bc, size=n_pairs, dtype=np.int64)
    unb = rng.random(n_pairs) < unbarcoded_frac
    bcid[unb] = -1
    order = np.argsort(bcid, kind="stable")          # -1 (unbarcoded) first, then by barcode
    bcid = bcid[order]
    # dense ordinal: 0 = unbarcoded, 1.. in order of first appearance (ParseBarcodedFastqs.cc:107-114)
    new = np.ones(n_pairs, dtype=bool)
    new[1:] = bcid[1:] != bcid[:-1]
    new &= bcid >= 0
    ordinal = np.cumsum(new).astype(np.int32)
    ordinal[bcid < 0] = 0

    n_reads = 2 * n_pairs
    bases = np.empty((n_reads, L), dtype=np.uint8)
    quals = np.empty((n_reads, L), dtype=np.uint8)
    jobs = [(G, seed, L, c0, min(n_pairs, c0 + chunk), shard) for c0 in range(0, n_pairs, chunk)]
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            for (g, s, l, c0, c1, sh), (blk, q) in zip(jobs, pool.imap(_gen_chunk, jobs)):
                bases[2 * c0:2 * c1] = blk
                quals[2 * c0:2 * c1] = q
    else:
        for job in jobs:
            blk, q = _gen_chunk(job)
            bases[2 * job[3]:2 * job[4]] = blk
            quals[2 * job[3]:2 * job[4]] = q
    _HAP = None
    bc = np.repeat(ordinal, 2)
    return bases, quals, bc, np.repeat(bcid, 2)


def barcode_string(i: int) -> str:
    s = []
    for _ in range(16):
        s.append("ACGT"[i & 3])
        i >>= 2
    return "".join(reversed(s))


def write_fasth(path: str, bases, quals, bc_ids):
    """9-line barcode-sorted pseudo-FASTQ (ParseBarcodedFastqs.cc:3-9)."""
    n_pairs = bases.shape[0] // 2
    L = bases.shape[1]
    with gzip.open(path, "wb", compresslevel=1) as f:
        bq = b"I" * 16
        for p in range(n_pairs):
            r1 = BASES[bases[2 * p]].tobytes()
            r2 = BASES[bases[2 * p + 1]].tobytes()
            q1 = (quals[2 * p] + 33).astype(np.uint8).tobytes()
            q2 = (quals[2 * p + 1] + 33).astype(np.uint8).tobytes()
            b = bc_ids[2 * p]
            bcs = (barcode_string(int(b)) + "-1").encode() if b >= 0 else b"NNNNNNNNNNNNNNNN"
            f.write(b"@p%d\n" % p + r1 + b"\n" + q1 + b"\n" + r2 + b"\n" + q2 + b"\n" +
                    bcs + b"\n" + bq + b"\nACGTACGT\nIIIIIIII\n")


CONFIGS = {
    # name: (G, pairs, nBC, seed)   -- SURVEY.md §8(d) table
    "tiny": (5_000, 1_000, 50, 7),
    "C1": (50_000, 10_000, 500, 1234),
    "mid": (2_000_000, 373_333, 50_000, 20261017),
    "C2": (63_000_000, 4_000_000, 1_000_000, 20261017),
}


def make_stress(seed: int, n_pairs: int = 6000, n_bc: int = 40):
    """Small adversarial data set for parity tests: repeats, inverted repeats (even-length
    palindromes), homopolymers, tandem repeats, a circular contig (perfect circle of
    k-mers), ragged read lengths (20..150), low-quality tails and unbarcoded reads.
    Returns ragged (bases, quals, off, bc, bc_ids) with reads as concatenated codes."""
    rng = np.random.Generator(np.random.Philox(key=seed))

    def rnd(n):
        return rng.integers(0, 4, size=n, dtype=np.uint8)

    def rc(s):
        return (3 - s[::-1]).astype(np.uint8)

    unit = rnd(700)
    pal_half = rnd(60)
    contigs = []
    g = [rnd(3000), unit, rnd(500), unit, rnd(800), rc(unit), rnd(400),
         pal_half, rc(pal_half), rnd(600),                      # exact even palindrome of 120
         np.zeros(90, np.uint8), rnd(300), np.full(70, 3, np.uint8), rnd(300),   # homopolymers
         np.tile(rnd(2), 60), rnd(300), np.tile(rnd(7), 30), rnd(300), np.tile(rnd(31), 6), rnd(900),
         np.tile(rnd(48), 4), rnd(500), np.tile(rnd(24), 8), rnd(1000)]
    contigs.append(np.concatenate(g))
    circ = rnd(400)
    contigs.append(np.concatenate([circ, circ, circ]))           # reads only ever see the circle
    contigs.append(np.concatenate([np.tile(rnd(5), 80)]))        # pure tandem repeat contig
    small = rnd(130)
    contigs.append(small)                                        # contig shorter than a read
    lens = np.array([len(c) for c in contigs])
    weights = np.array([0.75, 0.15, 0.05, 0.05])
    bases, quals = [], []
    bcid = rng.integers(0, n_bc, size=n_pairs, dtype=np.int64)
    bcid[rng.random(n_pairs) < 0.05] = -1
    order = np.argsort(bcid, kind="stable")
    bcid = bcid[order]
    new = np.ones(n_pairs, dtype=bool)
    new[1:] = bcid[1:] != bcid[:-1]
    new &= bcid >= 0
    ordinal = np.cumsum(new).astype(np.int32)
    ordinal[bcid < 0] = 0
    for p in range(n_pairs):
        ci = rng.choice(4, p=weights)
        c = contigs[ci]
        for _ in range(2):
            L = int(rng.choice([150, 150, 150, 150, 120, 97, 60, 49, 48, 47, 20]))
            L = min(L, len(c))
            if ci == 1:
                s = int(rng.integers(0, 400))
            else:
                s = int(rng.integers(0, len(c) - L + 1))
            r = c[s:s + L].copy()
            if rng.random() < 0.5:
                r = rc(r)
            perr = 0.002 + 0.03 * (np.arange(L) / 150.0) ** 3
            err = rng.random(L) < (perr if ci != 1 else 0.0)   # circle contig stays error-free
            r[err] = (r[err] + rng.integers(1, 4, size=int(err.sum()))) & 3
            q = np.where(rng.random(L) < 0.05, 30, 37).astype(np.uint8)
            q[err] = np.array([2, 12, 20], dtype=np.uint8)[rng.integers(0, 3, size=int(err.sum()))]
            if rng.random() < 0.15:                      # low-quality tail
                t = int(rng.integers(1, L + 1))
                q[L - t:] = rng.integers(2, 7, size=t)
            if rng.random() < 0.05:                      # low-quality spot in the middle
                q[int(rng.integers(0, L))] = 3
            bases.append(r.astype(np.uint8))
            quals.append(q)
    off = np.zeros(len(bases) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in bases])
    return (np.concatenate(bases), np.concatenate(quals), off, np.repeat(ordinal, 2), np.repeat(bcid, 2))


def write_fasth_ragged(path: str, bases, quals, off, bc_ids):
    n_pairs = (len(off) - 1) // 2
    with gzip.open(path, "wb", compresslevel=1) as f:
        bq = b"I" * 16
        for p in range(n_pairs):
            a0, a1, a2 = int(off[2 * p]), int(off[2 * p + 1]), int(off[2 * p + 2])
            r1 = BASES[bases[a0:a1]].tobytes()
            r2 = BASES[bases[a1:a2]].tobytes()
            q1 = (quals[a0:a1] + 33).astype(np.uint8).tobytes()
            q2 = (quals[a1:a2] + 33).astype(np.uint8).tobytes()
            b = bc_ids[2 * p]
            bcs = (barcode_string(int(b)) + "-1").encode() if b >= 0 else b"NNNNNNNNNNNNNNNN"
            f.write(b"@p%d\n" % p + r1 + b"\n" + q1 + b"\n" + r2 + b"\n" + q2 + b"\n" +
                    bcs + b"\n" + bq + b"\nACGTACGT\nIIIIIIII\n")


def fasth_text(bases, quals, bc_ids):
    """The 9-line pseudo-FASTQ text of fixed-length read pairs as one uint8 array (vectorised: the
    bench builds gigabytes of it).  Same content as write_fasth, names zero-padded to a fixed width."""
    n_pairs, L = bases.shape[0] // 2, bases.shape[1]
    name = np.frombuffer(b"@p", np.uint8)
    digits = 10
    rec_len = 2 + digits + 1 + 4 * (L + 1) + 19 + 17 + 9 + 9
    out = np.empty((n_pairs, rec_len), np.uint8)
    o = 0
    out[:, 0:2] = name; o = 2
    idx = np.arange(n_pairs, dtype=np.int64)
    for d in range(digits):
        out[:, o + digits - 1 - d] = 48 + (idx // 10 ** d) % 10
    o += digits
    out[:, o] = 10; o += 1
    for m, src, add in ((0, bases, None), (0, quals, 33), (1, bases, None), (1, quals, 33)):
        blk = src[m::2]
        out[:, o:o + L] = BASES[blk] if add is None else (blk + add).astype(np.uint8)
        o += L
        out[:, o] = 10; o += 1
    b = np.asarray(bc_ids[0::2], dtype=np.int64)
    for k in range(16):
        out[:, o + 15 - k] = np.frombuffer(b"ACGT", np.uint8)[(b >> (2 * k)) & 3]
    unb = b < 0
    if unb.any():
        out[unb, o:o + 16] = ord("N")
    o += 16
    out[:, o] = ord("-"); out[:, o + 1] = ord("1"); out[:, o + 2] = 10
    if unb.any():
        out[unb, o] = ord("N"); out[unb, o + 1] = ord("N")
    o += 3
    out[:, o:o + 16] = ord("I"); out[:, o + 16] = 10; o += 17
    out[:, o:o + 9] = np.frombuffer(b"ACGTACGT\n", np.uint8); o += 9
    out[:, o:o + 9] = np.frombuffer(b"IIIIIIII\n", np.uint8); o += 9
    assert o == rec_len
    return out.ravel()


# ---------------------------------------------------------------------------------------------------------------------
# Counter-based generator: the numpy twin of csrc/sn_synth.cuh -- every value is a pure function of
# (seed, stream, index), all integer arithmetic, so the device generator (sn_generate_reads) and this one agree bit for bit.
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _key(seed, stream):
    return _mix64(np.array([(int(seed) ^ ((stream * 0xD6E8FEB86659FD93) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]


def _h(key, a, b):
    with np.errstate(over="ignore"):
        return _mix64(_mix64(key ^ a) ^ b)


def cb_error_thresholds(L=150):
    """substitution probability at position j in units of 2^-24: floor(2^24 * (0.001 + 0.02 (j/L)^3)), exact integers"""
    return np.array([(2 ** 24 * (L ** 3 * 100 + 2 * 1000 * j ** 3)) // (1000 * 100 * L ** 3) for j in range(L)], dtype=np.uint32)


def _hap_base(kg, ks, hap, i):
    """base i of haplotype hap (arrays broadcast)"""
    with np.errstate(over="ignore"):
        b = (_h(kg, i >> np.uint64(5), np.uint64(0)) >> (np.uint64(2) * (i & np.uint64(31)))) & np.uint64(3)
        w = i // np.uint64(1000)
        s = _h(ks, w, np.uint64(0))
        snp = (hap != 0) & (i == w * np.uint64(1000) + (s & np.uint64(0xFFFFFFFF)) % np.uint64(1000))
        alt = (b + np.uint64(1) + (s >> np.uint64(32)) % np.uint64(3)) & np.uint64(3)
    return np.where(snp, alt, b)


def make_reads_cb(G, total_pairs, n_bc, seed, first_pair=0, n_pairs=None, L=150):
    """pairs [first_pair, first_pair + n_pairs) of the job -> (bases u8 [2n, L], quals u8 [2n, L], bc i32 [2n])"""
    assert L == 150
    n_pairs = total_pairs - first_pair if n_pairs is None else n_pairs
    with np.errstate(over="ignore"):
        kg, ks, kp, ke = (_key(seed, k) for k in (1, 2, 3, 4))
        p = np.arange(first_pair, first_pair + n_pairs, dtype=np.uint64)
        u = _h(kp, p, np.uint64(0))
        hap = (u & np.uint64(1))[:, None]
        flip = ((u >> np.uint64(1)) & np.uint64(1)).astype(bool)
        insert = (np.uint64(300) + ((u >> np.uint64(8)) & np.uint64(0xFFFFFF)) % np.uint64(200))
        start = _h(kp, p, np.uint64(1)) % (np.uint64(G) - insert)
        j = np.arange(L, dtype=np.uint64)[None, :]
        fwd = _hap_base(kg, ks, hap, start[:, None] + j)
        rev = np.uint64(3) - _hap_base(kg, ks, hap, (start + insert - np.uint64(1))[:, None] - j)
        r1 = np.where(flip[:, None], rev, fwd)
        r2 = np.where(flip[:, None], fwd, rev)
        bases = np.empty((2 * n_pairs, L), np.uint64)
        bases[0::2], bases[1::2] = r1, r2
        rid = np.arange(2 * first_pair, 2 * (first_pair + n_pairs), dtype=np.uint64)
        kr = _mix64(ke ^ rid)[:, None]
        e = _mix64(kr ^ j)
        T = cb_error_thresholds(L).astype(np.uint64)[None, :]
        err = (e & np.uint64(0xFFFFFF)) < T
        sub = (bases + np.uint64(1) + ((e >> np.uint64(24)) & np.uint64(0xFF)) % np.uint64(3)) & np.uint64(3)
        bases = np.where(err, sub, bases).astype(np.uint8)
        eq = np.array([2, 12, 20], np.uint8)[(((e >> np.uint64(32)) & np.uint64(0xFF)) % np.uint64(3)).astype(np.int64)]
        good = np.where(((e >> np.uint64(40)) & np.uint64(0xFFFF)) % np.uint64(100) < np.uint64(5), 30, 37).astype(np.uint8)
        quals = np.where(err, eq, good).astype(np.uint8)
    bc = (1 + (np.arange(first_pair, first_pair + n_pairs, dtype=object) * n_bc) // total_pairs).astype(np.int32)
    return bases, quals, np.repeat(bc, 2)
