"""-m gpu: the sharded multi-GPU path (sn_mg_build_graph, csrc/sn_multi.cu + sn_edges2.cuh) on ONE device: N ranks = N
contexts of this process, a host thread per rank, collectives through the in-process communicator (sn_comm_init_local;
same C++ code path as NCCL apart from who moves the bytes).  The dictionary stays sharded by minimizer-bucket range;
neighbours on other ranks are ghosts, unipath chains are cut at rank borders and stitched over the gathered stop table.
Every rank must end with the single-GPU result: a.hbv bytes, its slice of tmp.paths, and the union of the shards = the
k-mer table (count, context before and after recomputeAdjacencies).  Bar: bit-exact.
The same stop-indexed edge stage with one rank (SN_EDGES2=1) against the first implementation."""
import os

import numpy as np
import pytest

import datasets

pytestmark = pytest.mark.gpu
SETS = ["tiny", "stress1", "stress2", "stress3", "C1"]


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


def _single(sb, name, wd):
    codes, quals, off, bc, _ = datasets.get(name)
    packed = sb.pack_reads(codes, quals, off)
    with sb.Context(0) as ctx:
        ctx.load_reads(*packed, bc)
        ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
        out = dict(kmers=ctx.kmers(), info=ctx.kmer_graph_info(), hbv=open(wd + "/a.hbv", "rb").read(), paths=ctx.paths(), counts=ctx.counts(),
                   edges=sorted(datasets.unpack_edges(*ctx.edges())))
    return out, (codes, quals, off, bc)


def _slice(sb, data, rank, n_ranks):
    codes, quals, off, bc = data
    n = len(off) - 1
    lo = (n * rank // n_ranks) & ~1
    hi = n if rank == n_ranks - 1 else (n * (rank + 1) // n_ranks) & ~1
    o = off[lo:hi + 1] - off[lo]
    a, b = int(off[lo]), int(off[hi])
    return sb.pack_reads(codes[a:b], quals[a:b], o), np.ascontiguousarray(bc[lo:hi], np.int32)


@pytest.mark.parametrize("name", SETS)
def test_stop_indexed_edge_stage_on_one_rank(sb, name, tmp_path):
    """the default edge stage (stop-indexed, shared with the sharded path) against the first implementation (SN_EDGES2=0)"""
    ref, data = _single(sb, name, str(tmp_path))
    os.environ["SN_EDGES2"] = "0"
    try:
        wd = str(tmp_path / "v2"); os.makedirs(wd)
        with sb.Context(0) as ctx:
            ctx.load_reads(*sb.pack_reads(*data[:3]), data[3])
            ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
            assert np.array_equal(ctx.kmers(), ref["kmers"])
            i1 = ctx.kmer_graph_info()
            assert np.array_equal(i1[0], ref["info"][0])
            assert sorted(datasets.unpack_edges(*ctx.edges())) == ref["edges"]
            assert open(wd + "/a.hbv", "rb").read() == ref["hbv"]
            assert all(np.array_equal(a, b) for a, b in zip(ctx.paths(), ref["paths"]))
    finally:
        del os.environ["SN_EDGES2"]


@pytest.mark.parametrize("name,n_ranks", [(s, n) for s in SETS for n in (2, 3, 8)] + [("mid", 4), ("empty_kmers", 2), ("nobc", 2)])
def test_local_ranks_match_the_single_gpu_run(sb, name, n_ranks, tmp_path):
    _check_local_ranks(sb, name, n_ranks, tmp_path)


@pytest.mark.parametrize("name,n_ranks,passes", [("stress1", 2, 2), ("stress3", 8, 2), ("C1", 4, 4), ("C1", 2, 8), ("mid", 4, 2), ("empty_kmers", 2, 2)])
def test_sharded_count_in_passes(sb, name, n_ranks, passes, tmp_path, monkeypatch):
    """A rank that would receive more than 2^32 k-mer occurrences (BASELINE config 3: 15 G per GPU) counts in passes; every
    pass takes a slice of EVERY owner's bucket range (interleaved passes, msp_window_bucket).  Forced here on small sets:
    the result must not change."""
    monkeypatch.setenv("SN_MG_PASSES", str(passes))
    _check_local_ranks(sb, name, n_ranks, tmp_path)


def test_sharded_passes_need_a_power_of_two_ranks(sb, tmp_path, monkeypatch):
    monkeypatch.setenv("SN_MG_PASSES", "2")
    _, data = _single(sb, "C1", str(tmp_path))

    def fn(rank, ctx):
        packed, bc = _slice(sb, data, rank, 3)
        ctx.load_reads(*packed, bc)
        ctx.mg_build_graph(sb.Params(), with_paths=False)
    with pytest.raises(sb.SnError, match="power-of-two number of ranks"):
        sb.run_local_ranks(3, fn)


def _check_local_ranks(sb, name, n_ranks, tmp_path):
    ref, data = _single(sb, name, str(tmp_path))

    def rank_fn(with_paths):
        def fn(rank, ctx):
            packed, bc = _slice(sb, data, rank, n_ranks)
            ctx.load_reads(*packed, bc)
            ctx.mg_build_graph(sb.Params(), with_paths=with_paths)
            p = str(tmp_path / f"r{rank}_{int(with_paths)}.hbv")
            ctx.write_hbv(p)
            res = dict(hbv=open(p, "rb").read(), sharded=ctx.dict_is_sharded(), kmers=ctx.kmers(), info=ctx.kmer_graph_info(), counts=ctx.counts(),
                       edges=sorted(datasets.unpack_edges(*ctx.edges())))
            if with_paths:
                res["paths"] = ctx.paths()
            return res
        return fn
    # (a) graph only: the k-mer table stays sharded
    out = sb.run_local_ranks(n_ranks, rank_fn(False))
    for r in out:
        assert r["hbv"] == ref["hbv"]
        assert r["edges"] == ref["edges"]
        assert r["sharded"]
    km = np.concatenate([r["kmers"] for r in out])
    order = np.lexsort((km[:, 2], km[:, 1], km[:, 0]))
    assert np.array_equal(km[order], ref["kmers"])                                     # every k-mer on exactly one rank
    pruned = np.concatenate([r["info"][0] for r in out])[order]
    assert np.array_equal(pruned, ref["info"][0])                                      # contexts after recomputeAdjacencies
    assert sum(r["counts"]["n_superkmers"] for r in out) == ref["counts"]["n_superkmers"]
    # (b) with ReadPaths: every rank paths its own reads against the gathered table
    out = sb.run_local_ranks(n_ranks, rank_fn(True))
    for r in out:
        assert r["hbv"] == ref["hbv"]
        assert not r["sharded"]
        assert np.array_equal(r["kmers"], ref["kmers"])
    offs = np.concatenate([r["paths"][0] for r in out])
    plen = np.concatenate([np.diff(r["paths"][1].astype(np.int64)) for r in out])
    pe = np.concatenate([r["paths"][2] for r in out])
    assert np.array_equal(offs, ref["paths"][0])
    assert np.array_equal(np.concatenate([[0], np.cumsum(plen)]).astype(np.uint64), ref["paths"][1])
    assert np.array_equal(pe, ref["paths"][2])


def test_ghost_table_overflow_is_loud(sb, tmp_path):
    _, data = _single(sb, "C1", str(tmp_path))
    os.environ["SN_GHOST_CAP"] = "64"

    def fn(rank, ctx):
        packed, bc = _slice(sb, data, rank, 2)
        ctx.load_reads(*packed, bc)
        ctx.mg_build_graph(sb.Params(), with_paths=False)
    try:
        with pytest.raises(sb.SnError, match="ghost table overflowed"):
            sb.run_local_ranks(2, fn)
    finally:
        del os.environ["SN_GHOST_CAP"]
