"""-m gpu: the first DF-side step after the hot path: writePathsIndex
(10X/PathsIndex.cc:23-143, called at 10X/DF.cc:588) -- a.paths.inv (edge -> reads, feudal
VecULongVec), a.countsb, and a.inv -- against the reference's own code run by oracle/_ref/OracleProbe
INDEX=True on the same reads.  Bar: byte-identical files; on the mid set also the properties."""
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
