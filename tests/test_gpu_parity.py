"""-m gpu: the CUDA path, called through the C ABI, against the oracle (and against the
reference's own binaries when oracle/_ref is present) on the same seeded inputs.
Bar: bit-exact (integer / byte / index work)."""
import os

import numpy as np
import pytest

import datasets
import refrun

pytestmark = pytest.mark.gpu

SETS = ["tiny", "stress1", "stress2", "stress3", "C1", "empty_kmers", "nobc"]


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


@pytest.fixture(scope="module", params=SETS)
def run(request, sb, tmp_path_factory):
    from oracle.oracle import Oracle
    name = request.param
    codes, quals, off, bc, ids = datasets.get(name)
    o = Oracle(codes, quals, off, bc).run()
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    wd = str(tmp_path_factory.mktemp(name))
    ctx = sb.Context(0)
    ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
    ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
    yield dict(name=name, o=o, ctx=ctx, wd=wd, packed=(pb, boff, ln, pq, pqoff), bc=bc, codes=codes, quals=quals, off=off)
    ctx.close()


def test_good_lengths(run):
    assert np.array_equal(run["ctx"].good_lengths(), run["o"].good_len())


def test_kmer_table(run):
    km, ok = run["ctx"].kmers(), run["o"].kmers()
    assert run["ctx"].counts()["n_kmer_occurrences"] == run["o"].n_occ
    assert km.shape[0] == ok.shape[0]
    assert np.array_equal(km[:, :3], ok[:, :3])
    assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))


def test_pruned_contexts(run):
    ctx, edge, off = run["ctx"].kmer_graph_info()
    assert np.array_equal(ctx, run["o"].kmers()[:, 5].astype(np.uint8))


def test_edge_set(run):
    ln, off, packed = run["ctx"].edges()
    assert sorted(datasets.unpack_edges(ln, off, packed)) == sorted(run["o"].edges())


def test_hbv_bytes(run):
    run["o"].write_hbv(run["wd"] + "/oracle.hbv")
    assert open(run["wd"] + "/a.hbv", "rb").read() == open(run["wd"] + "/oracle.hbv", "rb").read()


def test_paths_bytes(run):
    run["o"].write_paths(run["wd"] + "/oracle.paths")
    assert open(run["wd"] + "/tmp.paths", "rb").read() == open(run["wd"] + "/oracle.paths", "rb").read()


def test_involution(run):
    assert np.array_equal(run["ctx"].hbv()["inv"], run["o"].involution())


def test_q8_ingest_matches_pqvec_ingest(run, sb):
    pb, boff, ln, pq, pqoff = run["packed"]
    with sb.Context(0) as c2:
        c2.load_reads_q8(pb, boff, ln, run["quals"], run["off"], run["bc"])
        c2.build_read_qgraph48(None, sb.Params(), with_paths=True)
        assert np.array_equal(c2.kmers(), run["ctx"].kmers())
        a, b = c2.paths(), run["ctx"].paths()
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref binaries not built")
def test_against_reference_binary(run, sb):
    """The reference's own buildReadQGraph48 on the same fastb/qualp/bci files."""
    if run["name"] in ("empty_kmers", "nobc"):
        pytest.skip("the reference exits on an empty k-mer set")
    wd = run["wd"] + "/ref"
    os.makedirs(wd, exist_ok=True)
    pb, boff, ln, pq, pqoff = run["packed"]
    sb.write_read_files(wd + "/reads", pb, boff, ln, pq, pqoff, run["bc"])
    refrun.run_probe(wd)
    ref = refrun.read_kvec(wd + "/kmers.kvec")
    km = run["ctx"].kmers()
    mine = np.stack([km[:, 0], km[:, 1], km[:, 2], km[:, 3] & 0xFFFFFF, km[:, 3] >> 24], axis=1)
    assert np.array_equal(mine, ref)
    assert open(run["wd"] + "/a.hbv", "rb").read() == open(wd + "/a.hbv", "rb").read()
    assert open(run["wd"] + "/tmp.paths", "rb").read() == open(wd + "/tmp.paths", "rb").read()
    spec_ref = open(wd + "/stats/histogram_kmer_count.json").read()
    assert open(run["wd"] + "/stats/histogram_kmer_count.json").read() == spec_ref


def test_trim_rule_against_the_references_known_answers(sb):
    """k_pqvec_goodlen / k_q8_goodlen on the reference's own known-answer vectors for the trim rule (lib/tada/src/cmd_msp.rs:329-350)"""
    from test_oracle_golden import _qv_trim_vectors
    codes, quals, off, expect = _qv_trim_vectors()
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    for raw in (False, True):
        with sb.Context(0) as ctx:
            if raw:
                ctx.load_reads_q8(pb, boff, ln, quals, off, None)
            else:
                ctx.load_reads(pb, boff, ln, pq, pqoff, None)
            ctx.count_kmers(sb.Params(min_qual=10))
            assert np.array_equal(ctx.good_lengths(), expect)
