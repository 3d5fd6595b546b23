"""CPU: the product's host+device per-item logic (sn_kmer.cuh / sn_graph.cuh / sn_path.cuh,
the code the CUDA kernels execute per k-mer / per read) run in loops on the CPU and compared
with the oracle and the golden vectors.  tests/hostsim is a test harness, not a fallback."""
import gzip
import os

import numpy as np
import pytest

import datasets
from oracle.oracle import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


@pytest.mark.parametrize("name", ["tiny", "stress1", "stress2", "C1"])
def test_graph_and_paths_logic(sb, name, tmp_path):
    from hostsim import HostSim
    codes, quals, off, bc, _ = datasets.get(name)
    o = Oracle(codes, quals, off, bc).run()
    km = o.kmers()
    recs = np.stack([km[:, 0], km[:, 1], km[:, 2], km[:, 3] | (km[:, 4] << 24)], axis=1).astype(np.uint32)
    hs = HostSim(recs)
    hs.prune()
    ctx, _, _ = hs.graph_info()
    assert np.array_equal(ctx, km[:, 5].astype(np.uint8))
    ln, eo, packed = hs.edges()
    assert sorted(datasets.unpack_edges(ln, eo, packed)) == sorted(o.edges())
    # every k-mer sits at (edge, off) where the edge really spells it (either strand)
    _, edge, eoff = hs.graph_info()
    assert (edge < len(ln)).all() and (eoff + 48 <= ln[edge]).all()
    hs.hbv(str(tmp_path / "h.hbv"))
    o.write_hbv(str(tmp_path / "o.hbv"))
    assert open(tmp_path / "h.hbv", "rb").read() == open(tmp_path / "o.hbv", "rb").read()
    pb, boff, pl, pq, pqoff = sb.pack_reads(codes, quals, off, threads=2)
    hs.paths(pb, boff, pl, quals, off, str(tmp_path / "h.paths"))
    o.write_paths(str(tmp_path / "o.paths"))
    assert open(tmp_path / "h.paths", "rb").read() == open(tmp_path / "o.paths", "rb").read()
    if name in ("tiny", "stress1"):
        with gzip.open(os.path.join(GOLD, name, "tmp.paths.gz"), "rb") as f:
            assert open(tmp_path / "h.paths", "rb").read() == f.read()


def test_extract_logic_matches_oracle_occurrences(sb):
    """k_extract's per-occurrence logic: the multiset of (canonical k-mer, ctx, bc) records."""
    import ctypes as C
    from hostsim import lib
    codes, quals, off, bc, _ = datasets.get("stress3")
    o = Oracle(codes, quals, off, bc).stage("count")
    gl = o.good_len()
    pb, boff, pl, pq, pqoff = sb.pack_reads(codes, quals, off, threads=2)
    padded = np.concatenate([pb, np.zeros(32, np.uint8)])
    recs = []
    buf = np.zeros((256, 4), np.uint32)
    for r in range(len(pl)):
        n = lib().hs_extract_read(padded.ctypes.data + int(boff[r]), int(gl[r]), int(bc[r]), buf.ctypes.data)
        if n:
            recs.append(buf[:n].copy())
    recs = np.concatenate(recs)
    assert len(recs) == o.n_occ
    # reduce on the CPU with numpy and compare with the oracle's table (count >= 3, >= 2 barcodes)
    order = np.lexsort((recs[:, 2], recs[:, 1], recs[:, 0]))
    recs = recs[order]
    key = recs[:, :3]
    head = np.ones(len(recs), bool)
    head[1:] = (key[1:] != key[:-1]).any(axis=1)
    gid = np.cumsum(head) - 1
    cnt = np.bincount(gid)
    ctx = np.zeros(gid[-1] + 1, np.uint32)
    np.bitwise_or.at(ctx, gid, recs[:, 3] >> 24)
    bcv = (recs[:, 3] & 0xFFFFFF).astype(np.int64)
    mn = np.full(gid[-1] + 1, 1 << 30)
    mx = np.zeros(gid[-1] + 1, np.int64)
    pos = bcv > 0
    np.minimum.at(mn, gid[pos], bcv[pos])
    np.maximum.at(mx, gid[pos], bcv[pos])
    valid = (cnt >= 3) & (mx > 0) & (mn != mx)
    ok = o.kmers()
    assert np.array_equal(key[head][valid], ok[:, :3])
    assert np.array_equal(cnt[valid], ok[:, 3])
    assert np.array_equal(ctx[valid], ok[:, 4])


@pytest.mark.parametrize("name", ["stress1", "stress3", "tiny"])
def test_msp_superkmers_expand_to_the_extract_records(sb, name):
    """k_msp_scan + sk_occurrence: per read, the super-k-mer records expand to exactly the records
    Kmerizer::map emits (same order), a run holds <= 47 k-mers, and every occurrence of a
    canonical k-mer carries the same bucket hash (tada's check_consistent_shard,
    lib/tada/src/kmer/mod.rs:1102-1150)."""
    from hostsim import lib
    codes, quals, off, bc, _ = datasets.get(name)
    gl = Oracle(codes, quals, off, bc).stage("count").good_len()
    pb, boff, pl, pq, pqoff = sb.pack_reads(codes, quals, off, threads=2)
    padded = np.concatenate([pb, np.zeros(64, np.uint8)])
    a = np.zeros((256, 4), np.uint32); b = np.zeros((256, 4), np.uint32); bh = np.zeros(256, np.uint32)
    sk = np.zeros((256, 8), np.uint32); nsk = np.zeros(1, np.uint32)
    all_recs, all_bh, tot_sk = [], [], 0
    for r in range(len(pl)):
        p = padded.ctypes.data + int(boff[r])
        n = lib().hs_extract_read(p, int(gl[r]), int(bc[r]), a.ctypes.data)
        m = lib().hs_msp_read(p, int(gl[r]), int(bc[r]), b.ctypes.data, bh.ctypes.data, sk.ctypes.data, nsk.ctypes.data)
        assert n == m
        assert np.array_equal(a[:n], b[:n]), r
        if n:
            k = int(nsk[0]); tot_sk += k
            nk = ((sk[:k, 0] >> 24) & 0x3F) + 1
            assert nk.sum() == n and nk.max() <= 47
            all_recs.append(b[:n, :3].copy()); all_bh.append(bh[:n].copy())
    recs = np.concatenate(all_recs); bhs = np.concatenate(all_bh)
    order = np.lexsort((recs[:, 2], recs[:, 1], recs[:, 0]))
    recs, bhs = recs[order], bhs[order]
    same = (recs[1:] == recs[:-1]).all(axis=1)
    assert np.array_equal(bhs[1:][same], bhs[:-1][same]), "a canonical k-mer was sent to two buckets"
    assert tot_sk * 6 < len(recs)            # super-k-mers are much fewer than k-mers (~17 k-mers each on clean reads)


@pytest.mark.parametrize("name", ["tiny", "stress1", "dupes"])
def test_df_side_logic_matches_golden(sb, name, tmp_path):
    """The per-read logic of sn_dfside.cuh (ReadPathX records, MarkDups), applied kernel by kernel as sn_dfside.cu does,
    against the files the reference's ReadPathVecX and MarkDups wrote (tests/golden: a.pathsX, a.dup, dup_stats.json)."""
    import json
    from hostsim import HostSim
    from oracle import dfside
    codes, quals, off, bc, _ = datasets.get(name)
    P = datasets.DUPES_PARAMS if name == "dupes" else {}
    o = Oracle(codes, quals, off, bc, **P).run()
    km = o.kmers()
    recs = np.stack([km[:, 0], km[:, 1], km[:, 2], km[:, 3] | (km[:, 4] << 24)], axis=1).astype(np.uint32)
    hs = HostSim(recs)
    hs.prune()
    hs.edges()
    hs.hbv(str(tmp_path / "h.hbv"))
    pb, boff, pl, pq, pqoff = sb.pack_reads(codes, quals, off, threads=2)
    hs.paths(pb, boff, pl, quals, off, str(tmp_path / "h.paths"))
    g = os.path.join(GOLD, name)
    gz = lambda f: gzip.open(os.path.join(g, f + ".gz"), "rb").read()
    assert open(tmp_path / "h.paths", "rb").read() == gz("tmp.paths")
    hs.pathsx(str(tmp_path / "h.pathsX"))
    assert open(tmp_path / "h.pathsX", "rb").read() == gz("a.pathsX")
    dup, art, ndups, inter = hs.mark_dups(pb, boff, pl, quals, off, bc)
    assert dfside.dup_file(dup.tolist()) == gz("a.dup")
    assert dfside.dup_percentages(dup.tolist(), ndups, inter, art.tolist()) == json.load(open(os.path.join(g, "dup_stats.json")))


@pytest.mark.parametrize("bits,lnr,lp", [(6, 1, 1), (8, 2, 2), (10, 3, 3), (9, 0, 2), (7, 3, 1)])
def test_interleaved_pass_bucket_mapping(sb, bits, lnr, lp):
    """msp_window_bucket (sn_msp.cuh), the mapping behind the sharded count in passes (sn_multi.cu: mg_count_sharded): the global
    bucket is owner | within; the top lp bits of `within` name the pass.  Every bucket belongs to exactly one pass; inside a pass
    the renumbered buckets are owner-major (an owner's share is one contiguous range, owner o at o << (wb - lp)) and keep their
    order; a rank's slices of the passes, in pass order, are its buckets in bucket order."""
    from hostsim import lib
    L = lib()
    nb, wb, P = 1 << bits, bits - lnr, 1 << lp
    per_pass = nb >> lp
    seen = np.zeros(nb, np.int64)
    for ps in range(P):
        pcfg = wb | (lp << 8) | (ps << 16)
        comp = np.array([L.hs_window_bucket(b, 0, per_pass, pcfg) for b in range(nb)], np.int64)
        mine = comp != 0xFFFFFFFF
        seen += mine
        b = np.nonzero(mine)[0]
        assert len(b) == per_pass and np.array_equal(np.sort(comp[b]), np.arange(per_pass))
        assert (np.diff(comp[b]) > 0).all()                                   # order kept
        owner = b >> wb
        assert np.array_equal(comp[b] >> (wb - lp), owner)                    # owner-major: owner o holds [o << (wb-lp), (o+1) << (wb-lp))
        assert np.array_equal((b & ((1 << wb) - 1)) >> (wb - lp), np.full(len(b), ps))
    assert (seen == 1).all()
    # plain windows (one GPU in passes) and no window at all
    assert L.hs_window_bucket(5, 4, 8, 0) == 1 and L.hs_window_bucket(3, 4, 8, 0) == 0xFFFFFFFF and L.hs_window_bucket(12, 4, 8, 0) == 0xFFFFFFFF
    assert L.hs_window_bucket(123456, 0, 0xFFFFFFFF, 0) == 123456
