"""-m gpu: parity beyond the small sets.

* "mid" (112 Mbp, 56x of a 2 Mbp diploid): the CUDA path through the C ABI against
  the reference's own buildReadQGraph48 binary (oracle/_ref) on the same fastb/qualp/bci --
  k-mer table, a.hbv bytes, tmp.paths bytes, k-mer spectrum.  Falls back to the C oracle when the
  reference binary did not travel.
* C2 (1.2 Gbp, BASELINE.json configs[1]) and C2b (the same genome at 56x, 3.53 Gbp): EXACT parity against the
  reference's own buildReadQGraph48 through committed golden digests (tests/golden/scale_digests.json, written by
  tests/golden/make_scale_digests.py from a run of oracle/_ref/OracleProbe on the same seeded reads): the
  order-independent digest of the {k-mer, count, ctx} table against kmers.kvec, md5 of a.hbv, of tmp.paths and
  of the k-mer spectrum; the inputs' md5 first, so a generator drift cannot pass for a parity failure.
* C2 also through the size-independent properties the domain offers: sorted distinct canonical k-mers above the
  thresholds, every valid k-mer in exactly one unipath, edge/HBV bookkeeping, involution,
  graph-consistent ReadPaths, and idempotence (a second run gives the same bytes).
Bar: bit-exact."""
import json
import os

import numpy as np
import pytest

import datasets
import digests
import refrun

pytestmark = pytest.mark.gpu
K = 48
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_digests.json")))


def against_golden(sb, ctx, name, wd, packed, bc):
    g = GOLD[name]
    sb.write_read_files(wd + "/reads", *packed, bc)
    for f in ("fastb", "qualp", "bci"):                      # same inputs as the reference run that made the digests
        assert digests.file_md5(wd + "/reads." + f) == g["inputs"][f], "generator drift: reads." + f
        os.remove(wd + "/reads." + f)
    km = ctx.kmers()
    assert km.shape[0] == g["n_kmers"]
    assert digests.kmer_digest(km[:, 0], km[:, 1], km[:, 2], km[:, 3]) == g["kmers"]
    del km
    assert os.path.getsize(wd + "/a.hbv") == g["a.hbv_bytes"] and digests.file_md5(wd + "/a.hbv") == g["a.hbv"]
    assert os.path.getsize(wd + "/tmp.paths") == g["tmp.paths_bytes"] and digests.file_md5(wd + "/tmp.paths") == g["tmp.paths"]
    assert digests.file_md5(wd + "/stats/histogram_kmer_count.json") == g["histogram_kmer_count.json"]


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


def test_mid_matches_reference(sb, tmp_path):
    codes, quals, off, bc, _ = datasets.get("mid")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    wd = str(tmp_path)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
        km = ctx.kmers()
        c = ctx.counts()
    assert c["n_bases"] == 373_333 * 2 * 150
    if refrun.have_ref():
        rd = wd + "/ref"
        os.makedirs(rd)
        sb.write_read_files(rd + "/reads", pb, boff, ln, pq, pqoff, bc)
        refrun.run_probe(rd)
        ref = refrun.read_kvec(rd + "/kmers.kvec")
        mine = np.stack([km[:, 0], km[:, 1], km[:, 2], km[:, 3] & 0xFFFFFF, km[:, 3] >> 24], axis=1)
        assert np.array_equal(mine, ref)
        assert open(wd + "/a.hbv", "rb").read() == open(rd + "/a.hbv", "rb").read()
        assert open(wd + "/tmp.paths", "rb").read() == open(rd + "/tmp.paths", "rb").read()
        assert open(wd + "/stats/histogram_kmer_count.json").read() == open(rd + "/stats/histogram_kmer_count.json").read()
    else:
        from oracle.oracle import Oracle
        o = Oracle(codes, quals, off, bc).run()
        ok = o.kmers()
        assert np.array_equal(km[:, :3], ok[:, :3]) and np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
        o.write_hbv(wd + "/o.hbv"); o.write_paths(wd + "/o.paths")
        assert open(wd + "/a.hbv", "rb").read() == open(wd + "/o.hbv", "rb").read()
        assert open(wd + "/tmp.paths", "rb").read() == open(wd + "/o.paths", "rb").read()


def _rev2(x):
    """reverse the sixteen 2-bit fields of u32 words"""
    x = ((x >> 2) & 0x33333333) | ((x & 0x33333333) << 2)
    x = ((x >> 4) & 0x0F0F0F0F) | ((x & 0x0F0F0F0F) << 4)
    return x.byteswap()


@pytest.fixture(scope="module")
def c2(sb, tmp_path_factory):
    codes, quals, off, bc, _ = datasets.get("C2")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    n_bases = int(codes.size)
    del codes, quals
    wd = str(tmp_path_factory.mktemp("C2"))
    ctx = sb.Context(0)
    ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
    ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
    yield dict(ctx=ctx, n_bases=n_bases, n_reads=len(ln), wd=wd, packed=(pb, boff, ln, pq, pqoff), bc=bc)
    ctx.close()


def test_c2_matches_the_reference_exactly(c2, sb):
    """BASELINE.json configs[1] at full size against the reference's own run (golden digests)."""
    against_golden(sb, c2["ctx"], "C2", c2["wd"], c2["packed"], c2["bc"])


def test_c2_kmer_table_properties(c2):
    ctx = c2["ctx"]
    c = ctx.counts()
    assert c["n_bases"] == c2["n_bases"] == 1_200_000_000
    km = ctx.kmers()
    assert km.shape[0] == c["n_kmers"] > 0
    key_hi = (km[:, 0].astype(np.uint64) << np.uint64(32)) | km[:, 1].astype(np.uint64)
    lo = km[:, 2]
    # strictly increasing: sorted, no duplicate k-mer
    assert np.all((key_hi[1:] > key_hi[:-1]) | ((key_hi[1:] == key_hi[:-1]) & (lo[1:] > lo[:-1])))
    cnt = km[:, 3] & 0xFFFFFF
    assert int(cnt.min()) >= 3                                   # minFreq
    assert int(cnt.astype(np.uint64).sum()) <= c["n_kmer_occurrences"]
    assert c["n_kmers"] <= c["n_kmers_distinct"] <= c["n_kmer_occurrences"]
    # canonical: k-mer <= its reverse complement (words are MSB-first, 16 bases each)
    r0, r1, r2 = ~_rev2(km[:, 2].copy()), ~_rev2(km[:, 1].copy()), ~_rev2(km[:, 0].copy())
    le = (km[:, 0] < r0) | ((km[:, 0] == r0) & ((km[:, 1] < r1) | ((km[:, 1] == r1) & (km[:, 2] <= r2))))
    assert bool(le.all())


def test_c2_every_kmer_in_exactly_one_unipath(c2):
    ctx = c2["ctx"]
    c = ctx.counts()
    ln, off, packed = ctx.edges()
    assert int((ln.astype(np.int64) - (K - 1)).sum()) == c["n_kmers"]
    assert int(ln.min()) >= K and int(ln.astype(np.uint64).sum()) == c["n_edge_bases"]
    assert np.array_equal(np.diff(off.astype(np.int64)), (ln.astype(np.int64) + 3) // 4)
    pruned, edge, eoff = ctx.kmer_graph_info()
    assert int(edge.max()) == c["n_edges"] - 1
    per_edge = np.bincount(edge, minlength=c["n_edges"])
    assert np.array_equal(per_edge, ln.astype(np.int64) - (K - 1))
    # offsets inside an edge are a permutation of 0..len-K: their sum per edge is n(n-1)/2
    s = np.bincount(edge, weights=eoff.astype(np.float64), minlength=c["n_edges"])
    n = per_edge.astype(np.float64)
    assert np.array_equal(s, n * (n - 1) / 2)
    assert bool((eoff < per_edge[edge]).all())


def test_c2_hbv_properties(c2):
    ctx = c2["ctx"]
    c = ctx.counts()
    h = ctx.hbv()
    ln, _, _ = ctx.edges()
    nh, nv = c["n_hbv_edges"], c["n_hbv_vertices"]
    inv, fwd, rev = h["inv"], h["fwd"], h["rev"]
    assert np.array_equal(inv[inv], np.arange(nh, dtype=np.int32))
    n_pal = int((fwd == rev).sum())
    assert nh == 2 * c["n_edges"] - n_pal
    assert np.array_equal(inv[fwd], rev)
    # every HBV edge id is used exactly once by fwd/rev (palindromes by both)
    used = np.bincount(np.concatenate([fwd, rev[fwd != rev]]), minlength=nh)
    assert bool((used == 1).all())
    # CSR: each HBV edge leaves one vertex and enters one vertex; lists sorted by (neighbour, edge)
    assert h["from_start"][-1] == nh and h["to_start"][-1] == nh
    assert np.array_equal(np.sort(h["from_e"]), np.arange(nh)) and np.array_equal(np.sort(h["to_e"]), np.arange(nh))
    left = np.empty(nh, np.int32); right = np.empty(nh, np.int32)
    left[h["from_e"]] = np.repeat(np.arange(nv, dtype=np.int32), np.diff(h["from_start"].astype(np.int64)))
    right[h["to_e"]] = np.repeat(np.arange(nv, dtype=np.int32), np.diff(h["to_start"].astype(np.int64)))
    assert np.array_equal(right[h["from_e"]], h["from_v"]) and np.array_equal(left[h["to_e"]], h["to_v"])
    # the reverse complement of an edge v->w runs rc(w)->rc(v): one vertex pairing serves every edge
    pair = np.full(nv, -1, np.int64)
    pair[left] = right[inv]
    pair[right] = left[inv]
    assert bool((pair >= 0).all())
    assert np.array_equal(pair[left], right[inv]) and np.array_equal(pair[right], left[inv])
    assert np.array_equal(pair[pair], np.arange(nv))
    # ReadPaths follow the graph
    poffset, poff, pe = ctx.paths()
    assert len(poff) == c2["n_reads"] + 1 and int(poff[-1]) == c["n_path_edges"] == len(pe) > 0
    assert int(pe.min()) >= 0 and int(pe.max()) < nh
    plen = np.diff(poff.astype(np.int64))
    inner = np.ones(len(pe), bool)
    inner[(poff[1:][plen > 0] - 1).astype(np.int64)] = False          # last edge of every path
    a = np.nonzero(inner)[0]
    assert np.array_equal(right[pe[a]], left[pe[a + 1]])
    placed = plen > 0
    assert 0.3 < placed.mean() <= 1.0
    elen = np.empty(nh, np.int64); elen[fwd] = ln; elen[rev] = ln      # HBV edge lengths via fwd/rev
    first = pe[poff[:-1][placed].astype(np.int64)]
    assert bool((poffset[placed] < elen[first]).all())


def test_c2_idempotent(c2, sb):
    ctx = c2["ctx"]
    h0, p0, k0 = ctx.hbv(), ctx.paths(), ctx.kmers()
    e0 = ctx.edges()
    ctx.build_read_qgraph48(None, sb.Params(), with_paths=True)
    h1, p1, k1 = ctx.hbv(), ctx.paths(), ctx.kmers()
    e1 = ctx.edges()
    assert np.array_equal(k0, k1)
    assert all(np.array_equal(h0[x], h1[x]) for x in h0)
    assert all(np.array_equal(a, b) for a, b in zip(p0, p1))
    # the edge ORDER may differ between runs (atomics); the set and the HBV built from it do not
    assert np.array_equal(np.sort(e0[0]), np.sort(e1[0]))


def test_streamed_load_matches_plain_load(sb, tmp_path):
    """sn_load_reads_streamed (chunked copies, good lengths + first MSP pass under them) leaves the
    context exactly where sn_load_reads + the same stages would: identical k-mers, HBV and paths;
    a second build on the same context recomputes everything and agrees again."""
    import torch
    codes, quals, off, bc, _ = datasets.get("mid")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    with sb.Context(0) as ref:
        ref.load_reads(pb, boff, ln, pq, pqoff, bc)
        ref.build_read_qgraph48(None, sb.Params(), with_paths=True)
        k0, h0, p0, g0 = ref.kmers(), ref.hbv(), ref.paths(), ref.good_lengths()
    host = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (pb, boff, ln, pq, pqoff, np.ascontiguousarray(bc, np.int32))]
    ptrs = [t.data_ptr() for t in host]
    for with_hist in (True, False):
        with sb.Context(0) as ctx:
            ctx.load_reads_streamed_ptr(len(ln), *ptrs, params=sb.Params(), with_hist=with_hist)
            for rep in range(2):
                ctx.build_read_qgraph48(None, sb.Params(), with_paths=True)
                assert np.array_equal(ctx.good_lengths(), g0)
                assert np.array_equal(ctx.kmers(), k0)
                h1, p1 = ctx.hbv(), ctx.paths()
                assert all(np.array_equal(h0[x], h1[x]) for x in h0)
                assert all(np.array_equal(a, b) for a, b in zip(p0, p1))
            # a different min_qual must not reuse the good lengths computed at load time
        with sb.Context(0) as ctx:
            ctx.load_reads_streamed_ptr(len(ln), *ptrs, params=sb.Params(min_qual=20), with_hist=with_hist)
            ctx.build_read_qgraph48(None, sb.Params(), with_paths=False)
            assert np.array_equal(ctx.good_lengths(), g0)
            assert np.array_equal(ctx.kmers(), k0)


@pytest.mark.parametrize("name,passes", [("C1", 2), ("C1", 7), ("stress2", 3), ("mid", 3)])
def test_count_in_several_passes(sb, name, passes):
    """More than 2^31 k-mer occurrences are counted in several passes over consecutive bucket
    ranges (forced here with SN_COUNT_PASSES): same dictionary, same graph, same paths."""
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(None, sb.Params(), with_paths=True)
        k0, h0, p0, c0 = ctx.kmers(), ctx.hbv(), ctx.paths(), ctx.counts()
        os.environ["SN_COUNT_PASSES"] = str(passes)
        try:
            ctx.build_read_qgraph48(None, sb.Params(), with_paths=True)
        finally:
            del os.environ["SN_COUNT_PASSES"]
        k1, h1, p1, c1 = ctx.kmers(), ctx.hbv(), ctx.paths(), ctx.counts()
    assert np.array_equal(k0, k1)
    assert all(np.array_equal(h0[x], h1[x]) for x in h0)
    assert all(np.array_equal(a, b) for a, b in zip(p0, p1))
    for f in ("n_kmers", "n_kmers_distinct", "n_superkmers", "n_kmer_occurrences", "n_edges", "n_hbv_edges"):
        assert c0[f] == c1[f], f


@pytest.mark.parametrize("name", ["tiny", "stress1", "stress2", "stress3", "C1", "mid"])
def test_hbv_numbering_over_laid_out_records(sb, name, tmp_path):
    """A giant component is numbered over records laid out along the graph (multi-source labelling +
    sort on the device, sn_hbvdev.cuh k_lay_*); forced here for every component size: the same a.hbv,
    xlat tables and paths as with the records in unipath order."""
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    out = {}
    for mode in ("plain", "laid_out"):
        if mode == "laid_out":
            os.environ["SN_HBV_LAYOUT_MIN"] = "1"
        else:
            os.environ["SN_HBV_NO_LAYOUT"] = "1"
        try:
            with sb.Context(0) as ctx:
                ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
                ctx.build_read_qgraph48(None, sb.Params(), with_paths=True)
                ctx.write_hbv(str(tmp_path / (mode + ".hbv")))
                out[mode] = (ctx.hbv(), ctx.paths(), ctx.stage_ms().get("hbv_layout"))
        finally:
            os.environ.pop("SN_HBV_LAYOUT_MIN", None); os.environ.pop("SN_HBV_NO_LAYOUT", None)
    assert out["laid_out"][2] > 0 and out["plain"][2] == 0
    assert open(tmp_path / "plain.hbv", "rb").read() == open(tmp_path / "laid_out.hbv", "rb").read()
    assert all(np.array_equal(out["plain"][0][x], out["laid_out"][0][x]) for x in out["plain"][0])
    assert all(np.array_equal(a, b) for a, b in zip(out["plain"][1], out["laid_out"][1]))
