"""-m gpu: sn_generate_reads (csrc/sn_synth.cuh, SURVEY.md §8(d): counter-based synthetic linked reads made on the device)
against its numpy twin (supernova_b200/synth.py:make_reads_cb): the reads a context holds after generating are, as
.fastb/.qualp/.bci files, byte-identical to the twin's reads packed on the host; the hot path on them equals the oracle on the
twin's reads.  Bar: bit-exact."""
import os

import numpy as np
import pytest

from supernova_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


def _twin_files(sb, G, total, nbc, seed, first, n, head):
    b, q, bc = synth.make_reads_cb(G, total, nbc, seed, first, n)
    off = np.arange(2 * n + 1, dtype=np.uint64) * 150
    pb, boff, ln, pq, pqoff = sb.pack_reads(b.ravel(), q.ravel(), off)
    # the barcode ordinals of a slice start where the slice starts: a .bci holds ordinals from 0, so compare the arrays too
    return (b, q, bc), (pb, boff, ln, pq, pqoff)


@pytest.mark.parametrize("G,total,nbc,seed,first,n,chunk", [(50_000, 10_000, 500, 1234, 0, 10_000, None), (2_000_000, 373_333, 50_000, 5, 100_000, 20_000, "3000"),
                                                               (3_200_000_000, 600_000_000, 4_000_000, 20261017, 599_990_000, 10_000, "4097")])
def test_generated_reads_equal_the_numpy_twin(sb, G, total, nbc, seed, first, n, chunk, tmp_path, monkeypatch):
    if chunk:
        monkeypatch.setenv("SN_SYN_CHUNK", chunk)
    (b, q, bc), packed = _twin_files(sb, G, total, nbc, seed, first, n, str(tmp_path))
    with sb.Context(0) as ctx:
        ctx.generate_reads(G, total, nbc, seed, first, n)
        c = ctx.counts()
        assert c["n_reads"] == 2 * n and c["n_bases"] == 300 * n
        ctx.save_read_files(str(tmp_path / "dev"))
        gl = None
        ctx.count_kmers(sb.Params())
        gl = ctx.good_lengths()
    pb, boff, ln, pq, pqoff = packed
    # .bci needs ordinals from 0/1 upward without gaps: rebase the slice's ordinals for the host-written twin
    uniq = np.unique(bc)
    rebased = (np.searchsorted(uniq, bc) + 1).astype(np.int32)
    sb.write_read_files(str(tmp_path / "twin"), pb, boff, ln, pq, pqoff, rebased)
    for ext in (".fastb", ".qualp"):
        assert open(str(tmp_path / "dev") + ext, "rb").read() == open(str(tmp_path / "twin") + ext, "rb").read(), ext
    dev_bci = np.frombuffer(open(str(tmp_path / "dev") + ".bci", "rb").read(), "<i8", offset=16)
    twin_bci = np.frombuffer(open(str(tmp_path / "twin") + ".bci", "rb").read(), "<i8", offset=16)
    # same barcode boundaries (the device file keeps the job-wide ordinals, so it may start with empty barcodes)
    assert np.array_equal(np.unique(dev_bci), np.unique(twin_bci))
    # good lengths from the PQVec stream the device wrote = the trim rule on the twin's quals
    ok = q >= 7
    run = np.zeros(2 * n, np.int64); glt = np.zeros(2 * n, np.int64)
    for j in range(150):
        run = np.where(ok[:, j], run + 1, 0)
        glt = np.where(run >= 48, j + 1, glt)
    assert np.array_equal(gl, glt.astype(np.uint32))


@pytest.mark.parametrize("streams", [1, 3])
def test_hot_path_on_generated_reads_equals_the_oracle_on_the_twin(sb, streams, tmp_path):
    """C1-shaped job (50 kbp genome, 10,000 pairs, 60x): the device-generated reads through count + graph + paths against the C
    oracle on the twin's reads; with the job cut into slices (what ranks do) the union of the slices is the job."""
    from oracle.oracle import Oracle
    G, total, nbc, seed = 50_000, 10_000, 500, 4321
    b, q, bc = synth.make_reads_cb(G, total, nbc, seed)
    off = np.arange(2 * total + 1, dtype=np.uint64) * 150
    o = Oracle(b.ravel(), q.ravel(), off, bc).run()
    wd = str(tmp_path)
    if streams == 1:
        with sb.Context(0) as ctx:
            ctx.generate_reads(G, total, nbc, seed)
            ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
            km, ok = ctx.kmers(), o.kmers()
        assert km.shape[0] == ok.shape[0] > 40_000 and np.array_equal(km[:, :3], ok[:, :3]) and np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
        o.write_hbv(wd + "/o.hbv"); o.write_paths(wd + "/o.paths")
        assert open(wd + "/a.hbv", "rb").read() == open(wd + "/o.hbv", "rb").read()
        assert open(wd + "/tmp.paths", "rb").read() == open(wd + "/o.paths", "rb").read()
    else:
        def fn(rank, ctx):
            lo, hi = total * rank // streams, total * (rank + 1) // streams
            ctx.generate_reads(G, total, nbc, seed, lo, hi - lo)
            ctx.mg_build_graph(sb.Params(), with_paths=True)
            if rank == 0:
                ctx.write_hbv(wd + "/a.hbv")
            return ctx.paths()
        res = sb.run_local_ranks(streams, fn)
        o.write_hbv(wd + "/o.hbv")
        assert open(wd + "/a.hbv", "rb").read() == open(wd + "/o.hbv", "rb").read()
        ooff, opoff, oe = o.paths()
        assert np.array_equal(np.concatenate([r[0] for r in res]), ooff) and np.array_equal(np.concatenate([r[2] for r in res]), oe)


def test_bucket_passes_equal_one_pass_on_a_generated_sub_sample(sb, tmp_path, monkeypatch):
    """The 22.5 Gbp run of tools/big_run.py (BASELINE config 3's per-GPU share, profiles/r02_c3_share_one_gpu_22p5gbp.json) counts in
    five bucket passes.  Same generator, a sub-sample (6.3 Mbp genome, 120 Mbp of reads): five forced passes give the table, the
    a.hbv and the tmp.paths of the single pass."""
    import hashlib
    G, total, nbc, seed = 6_300_000, 400_000, 100_000, 20261017
    res = []
    for passes in (None, "5"):
        if passes:
            monkeypatch.setenv("SN_COUNT_PASSES", passes)
        wd = str(tmp_path / ("p" + (passes or "1")))
        os.makedirs(wd)
        with sb.Context(0) as ctx:
            ctx.generate_reads(G, total, nbc, seed)
            ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
            km = ctx.kmers()
            c = ctx.counts()
        res.append((hashlib.md5(km.tobytes()).hexdigest(), open(wd + "/a.hbv", "rb").read(), open(wd + "/tmp.paths", "rb").read(), c["n_kmers"], c["n_edges"]))
    assert res[0] == res[1] and res[0][3] > 5_000_000
