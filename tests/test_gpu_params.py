"""-m gpu: the parameters and limits of Kmerizer::reduce the default sets do not reach
(BuildReadQGraph48.cc:91-137,155-181; kmers/ReadPather.h:128-145): ignBcBelow > 0, thresholds other than the
pipeline's, the 24-bit count saturation, and the MSPEDGES file writer.  CUDA path through the C ABI against the C
oracle and against the reference's own binary.  Bar: bit-exact."""
import os

import numpy as np
import pytest

import datasets
import refrun

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


def _run(sb, name, wd, **prm):
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get(name)
    o = Oracle(codes, quals, off, bc, **prm).run()
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    ctx = sb.Context(0)
    ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
    ctx.build_read_qgraph48(wd, sb.Params(**prm), with_paths=True)
    return ctx, o, (pb, boff, ln, pq, pqoff, bc)


def _check(sb, ctx, o, wd, packed, extra, overflow_expected=False):
    km, ok = ctx.kmers(), o.kmers()
    assert km.shape[0] == ok.shape[0]
    assert np.array_equal(km[:, :3], ok[:, :3])
    assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
    o.write_hbv(wd + "/oracle.hbv"); o.write_paths(wd + "/oracle.paths")
    assert open(wd + "/a.hbv", "rb").read() == open(wd + "/oracle.hbv", "rb").read()
    assert open(wd + "/tmp.paths", "rb").read() == open(wd + "/oracle.paths", "rb").read()
    if refrun.have_ref():
        rd = wd + "/ref"
        os.makedirs(rd)
        sb.write_read_files(rd + "/reads", *packed)
        _, log = refrun.run_probe(rd, extra=extra)
        ref = refrun.read_kvec(rd + "/kmers.kvec")
        mine = np.stack([km[:, 0], km[:, 1], km[:, 2], km[:, 3] & 0xFFFFFF, km[:, 3] >> 24], axis=1)
        if "buffer overflows" in log:
            # on a MapReduce buffer overflow the reference pre-summarises groups and LOSES their
            # per-occurrence barcodes (BuildReadQGraph48.cc:183-184, MapReduceEngine.h:538-541), so it drops k-mers
            # that pass the barcode rule -- how many depends on its thread count.  What it keeps must be ours, with
            # the same count and context; the oracle above (no overflow mode) pins the rest.
            assert overflow_expected, "unexpected MapReduce buffer overflow in the reference run"
            def key(a):
                return (a[:, 0].astype(np.uint64) << np.uint64(32)) | a[:, 1].astype(np.uint64), a[:, 2]
            mh, ml = key(mine); rh, rl = key(ref)
            mv = np.empty(len(mine), dtype=[("h", np.uint64), ("l", np.uint32)]); mv["h"], mv["l"] = mh, ml
            rv = np.empty(len(ref), dtype=[("h", np.uint64), ("l", np.uint32)]); rv["h"], rv["l"] = rh, rl
            at = np.searchsorted(mv, rv)
            assert bool((at < len(mv)).all()) and np.array_equal(mv[at], rv)
            assert np.array_equal(mine[at, 3:], ref[:, 3:])
            assert 0 < len(ref) <= len(mine)
            return
        assert np.array_equal(mine, ref)
        for f in ("a.hbv", "tmp.paths", "stats/histogram_kmer_count.json"):
            assert open(wd + "/" + f, "rb").read() == open(rd + "/" + f, "rb").read(), f


@pytest.mark.parametrize("name,ign", [("onebc", 1), ("onebc", 2500), ("fewbc", 3001), ("fewbc", 6000), ("C1", 9000), ("nobc", 300)])
def test_ign_bc_below(sb, name, ign, tmp_path):
    """reads with id < ignBcBelow count as barcode -1: a k-mer they touch passes the barcode rule whatever its barcodes
    (BuildReadQGraph48.cc:110-112,158-159,176-178)"""
    wd = str(tmp_path)
    ctx, o, packed = _run(sb, name, wd, ign_bc_below=ign)
    try:
        _check(sb, ctx, o, wd, packed, ("IGN_BC_BELOW=%d" % ign,))
        # the parameter changes the result (the sets hold k-mers seen under a single barcode)
        with sb.Context(0) as c0:
            c0.load_reads(*packed)
            c0.count_kmers(sb.Params())
            assert c0.counts()["n_kmers"] <= ctx.counts()["n_kmers"]
    finally:
        ctx.close()


@pytest.mark.parametrize("prm", [dict(min_qual=10, min_freq=2, min_bc=1), dict(min_qual=2, min_freq=5, min_bc=0), dict(min_qual=20, min_freq=2, min_bc=2), dict(min_qual=20, min_freq=1, min_bc=0)])
def test_other_thresholds_against_reference(sb, prm, tmp_path):
    # (MIN_FREQ=1 together with the barcode rule is left out: the reference binary itself dies on it with
    # ForceAssert(result) in EdgeBuilder::lookup, BuildReadQGraph48.cc:470-474)
    wd = str(tmp_path)
    ctx, o, packed = _run(sb, "stress3", wd, **prm)
    try:
        _check(sb, ctx, o, wd, packed, ("MIN_QUAL=%d" % prm["min_qual"], "MIN_FREQ=%d" % prm["min_freq"], "MIN_BC=%d" % prm["min_bc"]))
    finally:
        ctx.close()


def test_count_saturates_at_24_bits(sb, tmp_path):
    """A^48 occurs 17.3 M times: KDef::setCount clips at 2^24 - 1 (kmers/ReadPather.h:128-129,145).  All of it sits in
    one minimizer bucket (one CTA of k_bucket_count2), whose records are nearly all copies of three super-k-mers."""
    wd = str(tmp_path)
    ctx, o, packed = _run(sb, "polyA", wd)
    try:
        km = ctx.kmers()
        assert km[0, 0] == 0 and km[0, 1] == 0 and km[0, 2] == 0            # A^48 sorts first
        assert (km[0, 3] & 0xFFFFFF) == 0xFFFFFF
        assert ctx.counts()["n_kmer_occurrences"] > (1 << 24)
        _check(sb, ctx, o, wd, packed, (), overflow_expected=True)
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["tiny", "stress2", "C1"])
def test_write_edges_bv_round_trip(sb, name, tmp_path):
    """sn_write_edges_bv writes the vec<basevector> MSPEDGES file (feudal/BinaryStream.h:486-493, feudal/FieldVec.h:596-598,762;
    tada's writer: debruijn.rs:895-929): parsed back, it holds the device edge set; the edge set equals the oracle's."""
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    p = str(tmp_path / "edges.bv")
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.count_kmers(sb.Params()); ctx.build_edges()
        ctx.write_edges_bv(p)
        eln, eoff, packed = ctx.edges()
    d = open(p, "rb").read()
    assert d[:8] == b"BINWRITE"
    n = int(np.frombuffer(d, "<u8", 1, 8)[0])
    assert n == len(eln)
    at, got = 16, []
    for e in range(n):
        l = int(np.frombuffer(d, "<u4", 1, at)[0]); at += 4
        nb = (l + 3) // 4
        b = np.frombuffer(d, np.uint8, nb, at); at += nb
        got.append(np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).ravel()[:l].astype(np.uint8).tobytes())
    assert at == len(d)
    assert got == datasets.unpack_edges(eln, eoff, packed)
    o = Oracle(codes, quals, off, bc).run(with_paths=False)
    assert sorted(got) == sorted(o.edges())


def test_repeat_family_heavy_buckets(sb, tmp_path):
    """Low-complexity / repeat-family input: a few minimizer buckets receive 10^5 distinct k-mers.  k_bucket_count2 shares
    such a bucket out over 2^d CTAs by hash prefix (sn_msp.cuh); the table must equal the oracle's, and sharing out with
    other thresholds (SN_BC_HEAVY) must give the same bytes.  Prints the count time (-s shows it)."""
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get("repeats")
    o = Oracle(codes, quals, off, bc).run(with_paths=False)
    ok = o.kmers()
    packed = sb.pack_reads(codes, quals, off)
    ref = None
    for heavy in (None, "64", "1000000"):
        if heavy:
            os.environ["SN_BC_HEAVY"] = heavy
        try:
            with sb.Context(0) as ctx:
                ctx.load_reads(*packed, bc)
                ctx.build_read_qgraph48(None, sb.Params(), with_paths=False)
                ctx.build_read_qgraph48(None, sb.Params(), with_paths=False)
                km = ctx.kmers(); ms = ctx.stage_ms()["bucket_count"]; h = ctx.hbv()
        finally:
            os.environ.pop("SN_BC_HEAVY", None)
        print("repeats: SN_BC_HEAVY=%s bucket_count %.2f ms, %d k-mers, %d occurrences" % (heavy, ms, km.shape[0], o.n_occ))
        assert np.array_equal(km[:, :3], ok[:, :3]) and np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
        if ref is None:
            ref = h
        else:
            assert all(np.array_equal(ref[x], h[x]) for x in ref)


@pytest.mark.parametrize("name", ["stress1", "stress4", "C1"])
@pytest.mark.parametrize("streamed", [False, True])
def test_tada_semantics_count_reads_trimmed_to_k(sb, name, streamed, tmp_path):
    """SN_SEM_TADA (SURVEY a14/a15 and the equivalence note): a read whose trimmed length is exactly K gives its one k-mer,
    without neighbours (lib/tada/src/cmd_msp.rs:109-110; the C++ path drops it, BuildReadQGraph48.cc:160).  The oracle carries
    the same switch (pinned on the CPU against an independent restatement of the Rust rule, tests/test_oracle_golden.py);
    table, a.hbv and tmp.paths must equal the oracle's.  The stress sets hold reads of 48 bases."""
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get(name)
    if name == "C1":          # (fixed-length reads: trim some of them to exactly K by a low quality right behind base K)
        quals = quals.copy()
        rng = np.random.default_rng(5)
        for r in rng.choice(len(off) - 1, 400, replace=False):
            quals[int(off[r]) + 48:int(off[r + 1])] = 2
            quals[int(off[r]):int(off[r]) + 48] = 37
    o = Oracle(codes, quals, off, bc, count_len_k=True).run()
    o0 = Oracle(codes, quals, off, bc).stage("count")
    gl = o.good_len()
    assert int((gl == 48).sum()) > 0 and o.kmers().shape[0] >= o0.kmers().shape[0]
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    wd = str(tmp_path)
    with sb.Context(0) as ctx:
        ctx.set_semantics(tada=True)
        if streamed:
            import ctypes as C
            a = [np.ascontiguousarray(x) for x in (pb, boff, ln, pq, pqoff, bc.astype(np.int32))]
            ctx.load_reads_streamed_ptr(len(ln), *[x.ctypes.data_as(C.c_void_p) for x in a])
        else:
            ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
        km, ok = ctx.kmers(), o.kmers()
        assert ctx.counts()["n_kmer_occurrences"] == int(np.where(gl >= 48, gl - 47, 0).sum())
        assert km.shape[0] == ok.shape[0] and np.array_equal(km[:, :3], ok[:, :3])
        assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
        o.write_hbv(wd + "/oracle.hbv"); o.write_paths(wd + "/oracle.paths")
        assert open(wd + "/a.hbv", "rb").read() == open(wd + "/oracle.hbv", "rb").read()
        assert open(wd + "/tmp.paths", "rb").read() == open(wd + "/oracle.paths", "rb").read()
        # and back: the default rule on the same context
        ctx.set_semantics(tada=False)
        ctx.count_kmers(sb.Params())
        k0 = o0.kmers()
        km = ctx.kmers()
        assert km.shape[0] == k0.shape[0] and np.array_equal(km[:, :3], k0[:, :3])
