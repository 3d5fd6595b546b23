"""CPU: the reference's production boundary buildGraphFromMSP (paths/long/BuildReadQGraph48.h:24-26, .cc:1631-1684), run
through the stock harness (oracle/_ref/OracleProbe MSPEDGES=...), on the C oracle's edge set: it must rebuild the graph
and the ReadPaths buildReadQGraph48 itself produces.  Pins the oracle's edges through a second route and the harness
mode tests/test_gpu_shim.py relies on.  Also: the helpers the bench's reference arm uses."""
import os

import numpy as np
import pytest

import datasets
import digests
import refrun


def write_bv(path, seqs):
    out = [b"BINWRITE", np.uint64(len(seqs)).tobytes()]
    for e in seqs:
        s = np.frombuffer(e, np.uint8)
        pad = np.zeros((len(s) + 3) // 4 * 4, np.uint8); pad[:len(s)] = s
        q = pad.reshape(-1, 4)
        out += [np.uint32(len(s)).tobytes(), (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8).tobytes()]
    open(path, "wb").write(b"".join(out))


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_reference_graph_from_msp_edges_equals_reference_graph_from_reads(name, tmp_path, built):
    import supernova_b200 as sb
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    a, b = str(tmp_path / "reads"), str(tmp_path / "msp")
    for d in (a, b):
        os.makedirs(d)
        sb.write_read_files(d + "/reads", pb, boff, ln, pq, pqoff, bc)
    refrun.run_probe(a, keep_kvec=False)
    o = Oracle(codes, quals, off, bc).run(with_paths=False)
    write_bv(b + "/edges.bv", o.edges())
    _, log = refrun.run_probe(b, keep_kvec=False, extra=("MSPEDGES=" + b + "/edges.bv",))
    assert "reading MSP edge file" in log
    for f in ("a.hbv", "tmp.paths"):
        assert open(a + "/" + f, "rb").read() == open(b + "/" + f, "rb").read(), f


def test_phase_split_parses_the_reference_log():
    log = """Sat Oct 17 12:24:10 2026: loading reads.
Sat Oct 17 12:24:13 2026: MapReduce needs 1 passes.
.
Sat Oct 17 12:24:20 2026: MapReduceEngine::run complete
pVec->size() = 52372
Sat Oct 17 12:24:20 2026: MapReduce needs 1 passes.
.Sat Oct 17 12:24:21 2026 Pass 1 parse, swizzle, reduce
Sat Oct 17 12:24:30 2026: MapReduceEngine::run complete
Sat Oct 17 12:24:30 2026: computing spectrum
Sat Oct 17 12:24:32 2026: recomputing adjacencies
Sat Oct 17 12:24:35 2026: finding edge sequences.
Sat Oct 17 12:24:39 2026: building from edges 2
Sat Oct 17 12:24:44 2026: pathing reads
Sat Oct 17 12:24:50 2026: pathing iteration 1 of 1
ORACLE_SECONDS 40.5
"""
    ph = refrun.phase_split(log, 40.5)
    assert ph["qual_scan"] == 3 and ph["mapreduce_1"] == 7 and ph["mapreduce_2"] == 10 and ph["spectrum_kvec_dict"] == 2
    assert ph["recompute_adjacencies"] == 3 and ph["build_edges"] == 4 and ph["hbv_from_edges"] == 5 and ph["path_reads"] == 6.5
    assert ph["count_once_seconds"] == 33.5


def test_kmer_digest_is_order_independent_and_sensitive():
    rng = np.random.default_rng(1)
    w = rng.integers(0, 2 ** 32, size=(5000, 4), dtype=np.uint64).astype(np.uint32)
    d0 = digests.kmer_digest(w[:, 0], w[:, 1], w[:, 2], w[:, 3], chunk=777)
    p = rng.permutation(len(w))
    assert digests.kmer_digest(w[p, 0], w[p, 1], w[p, 2], w[p, 3]) == d0
    w2 = w.copy(); w2[17, 3] ^= 1 << 24                      # one context bit
    assert digests.kmer_digest(w2[:, 0], w2[:, 1], w2[:, 2], w2[:, 3]) != d0
    w3 = w.copy(); w3[[3, 4], 2] = w3[[4, 3], 2]             # two k-mers swap a word
    assert digests.kmer_digest(w3[:, 0], w3[:, 1], w3[:, 2], w3[:, 3]) != d0
