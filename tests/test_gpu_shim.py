"""-m gpu: the drop-in boundary, compiled.  oracle/_ref/OracleProbe_b200 is the reference's own harness main and
its 97-object closure with ONE translation unit replaced: paths/long/BuildReadQGraph48.o by
integration/BuildReadQGraph48_b200.cc, the reference-side binding of libsupernova_b200.so (same signatures as
BuildReadQGraph48.h:24-40, compiled against the reference's headers by oracle/build_ref.sh).  Its files must equal
the stock binary's byte for byte: a.hbv (through the reference's own BinaryReader/Writer round trip of the
HyperBasevector the shim filled), tmp.paths, the k-mer spectrum -- for buildReadQGraph48 and for
buildGraphFromMSP (edges from an MSPEDGES file)."""
import os
import shutil

import numpy as np
import pytest

import datasets
import refrun

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (refrun.have_ref() and refrun.have_shim()), reason="oracle/_ref harnesses not built")]


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


def _files(sb, name, wd):
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    os.makedirs(wd, exist_ok=True)
    sb.write_read_files(wd + "/reads", pb, boff, ln, pq, pqoff, bc)
    return pb, boff, ln, pq, pqoff, bc


def _same(a, b, names):
    for f in names:
        assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read(), f


@pytest.mark.parametrize("name", ["tiny", "stress1", "stress3", "C1", "mid"])
def test_shim_build_read_qgraph48(sb, name, tmp_path):
    ref, mine = str(tmp_path / "ref"), str(tmp_path / "b200")
    _files(sb, name, ref)
    _files(sb, name, mine)
    refrun.run_probe(ref, keep_kvec=False)
    _, log = refrun.run_probe(mine, keep_kvec=False, binary="OracleProbe_b200")
    assert "supernova_b200" in log
    _same(ref, mine, ["a.hbv", "tmp.paths", "stats/histogram_kmer_count.json"])


def test_shim_without_paths_and_with_thresholds(sb, tmp_path):
    ref, mine = str(tmp_path / "ref"), str(tmp_path / "b200")
    _files(sb, "stress2", ref)
    _files(sb, "stress2", mine)
    extra = ("MIN_QUAL=10", "MIN_FREQ=2", "MIN_BC=1", "IGN_BC_BELOW=700")
    refrun.run_probe(ref, paths=False, keep_kvec=False, extra=extra)
    refrun.run_probe(mine, paths=False, keep_kvec=False, extra=extra, binary="OracleProbe_b200")
    _same(ref, mine, ["a.hbv", "stats/histogram_kmer_count.json"])
    assert not os.path.exists(mine + "/tmp.paths")


def _scramble_bv(sb, src, dst, seed):
    """the same edge set as tada would hand it over: arbitrary order, arbitrary orientation"""
    d = open(src, "rb").read()
    n = int(np.frombuffer(d, "<u8", 1, 8)[0])
    p, edges = 16, []
    for _ in range(n):
        l = int(np.frombuffer(d, "<u4", 1, p)[0]); p += 4
        nb = (l + 3) // 4
        b = np.frombuffer(d, np.uint8, nb, p); p += nb
        edges.append(np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).ravel()[:l].astype(np.uint8))
    rng = np.random.default_rng(seed)
    order = rng.permutation(n)
    out = [b"BINWRITE", np.uint64(n).tobytes()]
    for e in order:
        s = edges[e]
        if rng.random() < 0.5:
            s = (3 - s[::-1]).astype(np.uint8)
        pad = np.zeros((len(s) + 3) // 4 * 4, np.uint8); pad[:len(s)] = s
        q = pad.reshape(-1, 4)
        out += [np.uint32(len(s)).tobytes(), (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8).tobytes()]
    open(dst, "wb").write(b"".join(out))


@pytest.mark.parametrize("name,scramble", [("tiny", False), ("stress1", True), ("stress3", True), ("C1", True), ("mid", True)])
def test_build_graph_from_msp_edges(sb, name, scramble, tmp_path):
    """buildGraphFromMSP: the reference's own (stock harness, MSPEDGES=) against sn_build_graph_from_edges through the
    ABI and through the compiled shim, on the edge file sn_write_edges_bv wrote (re-ordered and re-oriented at random)."""
    ref, mine, api = str(tmp_path / "ref"), str(tmp_path / "b200"), str(tmp_path / "api")
    pb, boff, ln, pq, pqoff, bc = _files(sb, name, ref)
    _files(sb, name, mine)
    os.makedirs(api)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.count_kmers(sb.Params()); ctx.build_edges()
        ctx.write_edges_bv(ref + "/edges0.bv")
        e0 = ctx.edges()
    # the written file holds exactly the device edges
    d = open(ref + "/edges0.bv", "rb").read()
    assert d[:8] == b"BINWRITE" and int(np.frombuffer(d, "<u8", 1, 8)[0]) == len(e0[0])
    assert len(d) == 16 + 4 * len(e0[0]) + int(((e0[0].astype(np.int64) + 3) // 4).sum())
    if scramble:
        _scramble_bv(sb, ref + "/edges0.bv", ref + "/edges.bv", 7)
    else:
        shutil.copy(ref + "/edges0.bv", ref + "/edges.bv")
    msp = ("MSPEDGES=" + ref + "/edges.bv",)
    refrun.run_probe(ref, keep_kvec=False, extra=msp)
    refrun.run_probe(mine, keep_kvec=False, extra=msp, binary="OracleProbe_b200")
    _same(ref, mine, ["a.hbv", "tmp.paths"])
    with sb.Context(0) as ctx:
        ctx.load_read_files_bc(ref + "/reads", None)
        ctx.build_graph_from_edges(ref + "/edges.bv")
        ctx.path_reads()
        ctx.write_hbv(api + "/a.hbv"); ctx.write_paths(api + "/tmp.paths")
    _same(ref, api, ["a.hbv", "tmp.paths"])
    if not scramble:
        # canonical edges in: the graph equals the one buildReadQGraph48 builds from the reads
        refrun.run_probe(mine, keep_kvec=False, binary="OracleProbe_b200")
        assert open(ref + "/a.hbv", "rb").read() == open(mine + "/a.hbv", "rb").read()
