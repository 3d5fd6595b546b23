"""The counter-based synthetic read generator: supernova_b200/csrc/sn_synth.cuh (device, host+device per-read
logic) and its numpy twin supernova_b200/synth.py:make_reads_cb must agree bit for bit.
CPU: the twin against the per-read logic of the kernel run on the CPU (tests/hostsim), slices against the whole, and the
statistics the spec asks for.  GPU (tests/test_gpu_synth.py): the same against sn_generate_reads."""
import numpy as np
import pytest

from supernova_b200 import synth


def _hostsim(G, total, nbc, seed, first, n):
    from hostsim import lib
    T = np.ascontiguousarray(synth.cb_error_thresholds(), np.uint32)
    b = np.zeros((2 * n, 150), np.uint8); q = np.zeros((2 * n, 150), np.uint8); bc = np.zeros(2 * n, np.int32)
    lib().hs_synth_reads(G, total, nbc, seed, T.ctypes.data, first, n, b.ctypes.data, q.ctypes.data, bc.ctypes.data)
    return b, q, bc


@pytest.mark.parametrize("G,total,nbc,seed,first,n", [(50_000, 10_000, 500, 1234, 0, 3000), (63_000_000, 4_000_000, 1_000_000, 20261017, 3_999_000, 1000),
                                                         (3_200_000_000, 600_000_000, 4_000_000, 7, 599_999_500, 500), (1000, 10, 1, 0, 0, 10)])
def test_numpy_twin_equals_the_kernel_logic(built, G, total, nbc, seed, first, n):
    b, q, bc = synth.make_reads_cb(G, total, nbc, seed, first, n)
    hb, hq, hbc = _hostsim(G, total, nbc, seed, first, n)
    assert np.array_equal(b, hb) and np.array_equal(q, hq) and np.array_equal(bc, hbc)


def test_slices_are_the_whole_and_the_statistics_hold():
    G, total, nbc, seed = 200_000, 20_000, 300, 99
    b, q, bc = synth.make_reads_cb(G, total, nbc, seed)
    parts = [synth.make_reads_cb(G, total, nbc, seed, f, 5000) for f in range(0, total, 5000)]
    assert np.array_equal(b, np.concatenate([x[0] for x in parts])) and np.array_equal(q, np.concatenate([x[1] for x in parts]))
    assert np.array_equal(bc, np.concatenate([x[2] for x in parts]))
    assert bc.min() == 1 and bc.max() == nbc and (np.diff(bc) >= 0).all()
    assert set(np.unique(q)) <= {2, 12, 20, 30, 37}
    err = np.isin(q, (2, 12, 20))
    T = synth.cb_error_thresholds().astype(np.float64) / 2 ** 24
    assert abs(T[0] - 0.001) < 1e-6 and abs(T[149] - (0.001 + 0.02 * (149 / 150) ** 3)) < 1e-6
    assert abs(err.mean() - T.mean()) < 5e-4                       # substitution rate 0.001 + 0.02 (j/150)^3 averaged over the read
    good = q[~err]
    assert abs((good == 30).mean() - 0.05) < 0.003
    assert abs((b == 0).mean() - 0.25) < 0.01
    # the two reads of a pair are the two ends of one fragment of 300..499 bases: R2's reverse complement ends where the fragment ends
    clean = (~err[0::2]).all(1) & (~err[1::2]).all(1)
    assert clean.any()
    # both haplotypes and both strands occur: identical R1 prefixes never dominate
    assert len({bytes(x[:40]) for x in b[0::2][:2000]}) > 1990
