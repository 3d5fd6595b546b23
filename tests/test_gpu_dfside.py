"""-m gpu: the DF-side steps after the hot path: writePathsIndex
(10X/PathsIndex.cc:23-143, called at 10X/DF.cc:588) -- a.paths.inv (edge -> reads, feudal
VecULongVec), a.countsb, and a.inv -- against the reference's own code run by oracle/_ref/OracleProbe
INDEX=True on the same reads; HyperBasevectorX, ReadPathVecX and MarkDups (10X/DF.cc:573-600) -- a.hbx,
a.fastb, a.kmers, a.pathsX, a.dup -- against the golden files the reference wrote.  Bar: byte-identical
files; on the mid set the properties and a numpy restatement."""
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


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("name", ["tiny", "stress1", "stress2", "C1"])
def test_paths_index_files_match_the_reference(sb, name, tmp_path):
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    wd = str(tmp_path)
    rd = wd + "/ref"
    os.makedirs(rd)
    sb.write_read_files(rd + "/reads", pb, boff, ln, pq, pqoff, bc)
    _, log = refrun.run_probe(rd, extra=("INDEX=True",))
    if not os.path.exists(rd + "/a.paths.inv"):
        pytest.skip("this OracleProbe was built without INDEX support")
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
        ctx.build_paths_index()
        ctx.write_paths_index(wd + "/a.paths.inv", wd + "/a.countsb")
        ctx.write_inv(wd + "/a.inv")
        ctx.write_to_left_right(wd + "/a.to_left", wd + "/a.to_right")
    for f in ("a.inv", "a.countsb", "a.paths.inv") + (("a.to_left", "a.to_right") if os.path.exists(rd + "/a.to_left") else ()):
        assert open(wd + "/" + f, "rb").read() == open(rd + "/" + f, "rb").read(), f


@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_dfside_files_match_golden(sb, name, tmp_path):
    """The same files against the committed golden ones (written by the reference, tests/golden/make_golden.py)
    and against the python restatement (oracle/dfside.py): needs no reference binary on the box."""
    import gzip
    from oracle import dfside
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    wd = str(tmp_path)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
        ctx.build_paths_index()
        ctx.write_paths_index(wd + "/a.paths.inv", wd + "/a.countsb")
        ctx.write_inv(wd + "/a.inv")
        ctx.write_to_left_right(wd + "/a.to_left", wd + "/a.to_right")
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    for f in ("a.inv", "a.to_left", "a.to_right", "a.countsb", "a.paths.inv"):
        assert open(wd + "/" + f, "rb").read() == gzip.open(g + "/" + f + ".gz", "rb").read(), f
    pi, cb = dfside.paths_index(dfside.read_paths(open(wd + "/tmp.paths", "rb").read()), dfside.read_vec_int(open(wd + "/a.inv", "rb").read()))
    assert pi == open(wd + "/a.paths.inv", "rb").read() and cb == open(wd + "/a.countsb", "rb").read()


@pytest.mark.parametrize("name", ["tiny", "stress1", "dupes"])
@pytest.mark.parametrize("raw_quals", [False, True])
def test_pathsx_dups_and_graph_files_match_golden(sb, name, raw_quals, tmp_path):
    """a.hbx, a.fastb, a.kmers, a.k, a.pathsX, a.dup and MarkDups' three percentages: byte-identical to what the reference's
    HyperBasevectorX, ReadPathVecX and MarkDups wrote for the same reads (tests/golden/make_golden.py; 10X/WriteFiles.cc:16-60,
    10X/DF.cc:573-600).  `dupes` holds ties, artifactual duplicates and two placements that only collide in the 16-bit offset of
    the ReadPathX record.  With the quals as PQVec streams and as one byte per base."""
    import gzip
    import json
    from oracle import dfside
    codes, quals, off, bc, _ = datasets.get(name)
    P = sb.Params(**datasets.DUPES_PARAMS) if name == "dupes" else sb.Params()
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    wd = str(tmp_path)
    with sb.Context(0) as ctx:
        if raw_quals:
            ctx.load_reads_q8(pb, boff, ln, quals, off, bc)
        else:
            ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(wd, P, with_paths=True)
        ctx.build_pathsx(); ctx.write_pathsx(wd + "/a.pathsX")
        st = ctx.mark_dups(); ctx.write_dup(wd + "/a.dup")
        dup, art = ctx.dups()
        ctx.write_hbx(wd + "/a.hbx"); ctx.write_edges_fastb(wd + "/a.fastb"); ctx.write_kmers(wd + "/a.kmers"); ctx.write_k(wd + "/a.k")
        idx, dat = ctx.pathsx()
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    for f in ("a.hbv", "tmp.paths", "a.hbx", "a.fastb", "a.kmers", "a.pathsX", "a.dup"):
        assert open(wd + "/" + f, "rb").read() == gzip.open(g + "/" + f + ".gz", "rb").read(), f
    assert open(wd + "/a.k").read() == "48\n"
    assert st["n_pairs"] == len(dup) and st["n_dup_pairs"] == int(dup.sum()) and st["n_art_pairs"] == int(art.sum())
    assert dfside.dup_percentages(dup.tolist(), st["n_dups"], st["n_interdups"], art.tolist()) == json.load(open(g + "/dup_stats.json"))
    px = open(wd + "/a.pathsX", "rb").read()
    assert px[40:40 + 8 * len(idx)] == idx.tobytes() and px[40 + 8 * len(idx):] == dat.tobytes()
    if name == "dupes":
        assert st["n_art_pairs"] > 0 and st["n_dup_pairs"] > st["n_art_pairs"]


def test_pathsx_and_dups_properties_mid(sb):
    """mid set (112 Mbp): every ReadPathX record decodes back to its path (edge count, 16-bit offset, first edge, branch ids
    through the HBV adjacency), the index points at every 10th record; MarkDups' flags agree with the restatement."""
    from oracle import dfside
    codes, quals, off, bc, _ = datasets.get("mid")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(None, sb.Params(), with_paths=True)
        ctx.build_pathsx()
        st = ctx.mark_dups()
        dup, art = ctx.dups()
        idx, dat = ctx.pathsx()
        poffset, poff, pe = ctx.paths()
        H = ctx.hbv()
    n = len(poffset)
    plen = np.diff(poff).astype(np.int64)
    assert plen.max() < 256
    size = np.where(plen > 0, (plen - 1 + 3) // 4 + 7, 1)
    zoff = np.concatenate([[0], np.cumsum(size)])
    assert zoff[-1] == len(dat) and np.array_equal(idx, zoff[:-1][::10])
    assert np.array_equal(dat[zoff[:-1]], plen.astype(np.uint8))
    placed = np.nonzero(plen > 0)[0]
    z = zoff[placed]
    off16 = dat[z + 1].astype(np.uint16) | (dat[z + 2].astype(np.uint16) << 8)
    assert np.array_equal(off16.view(np.int16), poffset[placed].astype(np.int16))
    e0 = dat[z + 3].astype(np.uint32) | (dat[z + 4].astype(np.uint32) << 8) | (dat[z + 5].astype(np.uint32) << 16) | (dat[z + 6].astype(np.uint32) << 24)
    assert np.array_equal(e0, pe[poff[placed]].astype(np.uint32))
    # branch ids of a sample of multi-edge paths, decoded through From(ToRight(e))
    to_right = np.zeros(len(H["from_e"]), np.int64)
    to_right[H["from_e"]] = H["from_v"]
    multi = placed[plen[placed] > 1][:20000]
    for r in multi:
        edges = pe[poff[r]:poff[r + 1]]
        cur, out = int(edges[0]), [int(edges[0])]
        for i in range(len(edges) - 1):
            b = (int(dat[zoff[r] + 7 + i // 4]) >> (2 * (i % 4))) & 3
            w = int(to_right[cur])
            cur = int(H["from_e"][H["from_start"][w] + b])
            out.append(cur)
        assert out == edges.tolist(), r
    # MarkDups (10X/SecretOps.cc:599-774) restated with numpy on the same paths: groups of (first edge, 16-bit offset, first five
    # bases of the partner), winner = first member with the highest quality sum of its pair
    L = 150
    B = codes.reshape(n, L).astype(np.int64); Q = quals.reshape(n, L).astype(np.int64)
    mate = np.arange(n) ^ 1
    head = ((((B[mate, 0] * 4 + B[mate, 1]) * 4 + B[mate, 2]) * 4 + B[mate, 3]) * 4 + B[mate, 4])
    first = pe[np.minimum(poff[:-1], len(pe) - 1).astype(np.int64)].astype(np.int64)
    key = (first << 26) | ((poffset.astype(np.int16).astype(np.int64) + 32768) << 10) | head
    ids = np.nonzero(plen > 0)[0]
    ids = ids[np.lexsort((ids, key[ids]))]
    k = key[ids]
    start = np.nonzero(np.concatenate([[True], k[1:] != k[:-1]]))[0]
    end = np.concatenate([start[1:], [len(ids)]])
    qpair = Q.sum(1); qpair = qpair + qpair[mate]
    d2 = np.zeros(n // 2, np.uint8); a2 = np.zeros(n // 2, np.uint8)
    nd = ni = 0
    for s0, e0 in zip(start[end - start > 1], end[end - start > 1]):
        g = ids[s0:e0]
        q = qpair[g]
        best, cur, tie = 0, q[0], False
        for l in range(1, len(g)):
            if q[l] == cur:
                tie = True
            elif q[l] > cur:
                cur, best = q[l], l
        d2[np.delete(g, best) // 2] = 1
        nd += len(g) - 1
        b = int(bc[g[0]]); inter = False
        for x in g[1:]:
            if b == 0:
                b = int(bc[x])
            elif int(bc[x]) != b:
                inter = True
        ni += (len(g) - 1) if inter else 0
        if tie:
            for i, x in enumerate(g):
                if any(np.array_equal(B[x], B[y]) and np.array_equal(Q[x], Q[y]) for y in g[:i]):
                    a2[x // 2] = 1
    assert np.array_equal(dup, d2) and np.array_equal(art, a2)
    assert (st["n_dups"], st["n_interdups"]) == (nd, ni) and d2.sum() > 0


def test_paths_index_properties_mid(sb):
    codes, quals, off, bc, _ = datasets.get("mid")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.build_read_qgraph48(None, sb.Params(), with_paths=True)
        ctx.build_paths_index()
        ioff, ids, cb = ctx.paths_index()
        poffset, poff, pe = ctx.paths()
        inv = ctx.hbv()["inv"]
        nh = ctx.counts()["n_hbv_edges"]
    assert int(ioff[-1]) == len(pe) == len(ids) > 0
    per_edge = np.bincount(pe, minlength=nh)
    assert np.array_equal(np.diff(ioff.astype(np.int64)), per_edge)
    # ids ascending inside every edge's list
    d = np.diff(ids.astype(np.int64))
    starts = ioff[1:-1].astype(np.int64)
    d[starts[(starts > 0) & (starts < len(ids))] - 1] = 0
    assert (d >= 0).all()
    # every (read, edge) pair of the paths is in the index exactly as often as it occurs
    read_of = np.repeat(np.arange(len(poff) - 1, dtype=np.int64), np.diff(poff.astype(np.int64)))
    edge_of = np.repeat(np.arange(nh, dtype=np.int64), per_edge)
    a = np.sort(pe.astype(np.int64) * (len(poff)) + read_of)
    b = np.sort(edge_of * (len(poff)) + ids.astype(np.int64))
    assert np.array_equal(a, b)
    tot = per_edge + np.where(inv == np.arange(nh), 0, per_edge[inv])
    assert np.array_equal(cb, tot)
