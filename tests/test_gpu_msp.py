"""-m gpu: the minimizer-sharded count (k_msp_scan / k_bucket_count) on its rare paths.
SN_MSP_BITS forces a handful of huge buckets, so that one bucket holds far more distinct k-mers
than the shared-memory table: the pass must split (by further hash bits) again and again and
still emit every k-mer exactly once; buckets also span many TMA chunks."""
import os

import numpy as np
import pytest

import datasets

pytestmark = pytest.mark.gpu


@pytest.fixture
def msp_bits():
    old = os.environ.get("SN_MSP_BITS")
    def set_bits(b):
        if b is None:
            os.environ.pop("SN_MSP_BITS", None)
        else:
            os.environ["SN_MSP_BITS"] = str(b)
    yield set_bits
    set_bits(old)


@pytest.mark.parametrize("name", ["stress1", "C1", "nobc"])
@pytest.mark.parametrize("bits", [1, 3, 7])
def test_split_passes_match_the_oracle(built, msp_bits, name, bits):
    import supernova_b200 as sb
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get(name)
    o = Oracle(codes, quals, off, bc).stage("count")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    res = {}
    for b in (None, bits):
        msp_bits(b)
        with sb.Context(0) as ctx:
            ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
            ctx.count_kmers(sb.Params())
            res[b] = (ctx.kmers(), ctx.counts())
    km, ok = res[bits][0], o.kmers()
    assert np.array_equal(km[:, :3], ok[:, :3])
    assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
    assert np.array_equal(res[None][0], km)
    for key in ("n_kmer_occurrences", "n_kmers_distinct", "n_kmers"):
        assert res[None][1][key] == res[bits][1][key], key


@pytest.mark.parametrize("min_freq,min_bc", [(1, 0), (2, 1), (5, 2)])
def test_thresholds(built, min_freq, min_bc):
    """Kmerizer::reduce's filter for other thresholds than the pipeline's (3, 2)."""
    import supernova_b200 as sb
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get("stress2")
    o = Oracle(codes, quals, off, bc, min_freq=min_freq, min_bc=min_bc).stage("count")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    with sb.Context(0) as ctx:
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
        ctx.count_kmers(sb.Params(min_freq=min_freq, min_bc=min_bc))
        km, ok = ctx.kmers(), o.kmers()
        assert np.array_equal(km[:, :3], ok[:, :3])
        assert np.array_equal(km[:, 3], ok[:, 3] | (ok[:, 4] << 24))
