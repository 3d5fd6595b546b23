"""CPU: host format layer (supernova_b200/csrc/sn_formats.cpp) against files written by the
reference's ParseBarcodedFastqs (golden) -- byte-for-byte, including the PQVec encoder."""
import gzip
import os

import numpy as np
import pytest

import datasets

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gz(path):
    with gzip.open(path, "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_read_files_match_reference_writer(sb, name, tmp_path):
    codes, quals, off, bc, _ = datasets.get(name)
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off, threads=3)
    head = str(tmp_path / "mine")
    sb.write_read_files(head, pb, boff, ln, pq, pqoff, bc)
    for ext in ("fastb", "qualp", "bci"):
        assert open(head + "." + ext, "rb").read() == gz(os.path.join(GOLD, name, "reads." + ext + ".gz")), ext


def test_pqvec_round_trip(sb):
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 47, 150, 255, 256, 700):
        for kind in range(4):
            if kind == 0:
                q = rng.integers(0, 64, size=n)
            elif kind == 1:
                q = np.full(n, 37)
            elif kind == 2:
                q = np.where(rng.random(n) < 0.03, 2, 37)
            else:
                q = np.repeat(rng.integers(0, 64, size=n // 10 + 1), 10)[:n]
            q = q.astype(np.uint8)
            enc = sb.pqvec_encode(q)
            assert enc[-1] == 0
            dec, m = sb.pqvec_decode(enc, max(n, 1))
            assert m == n and np.array_equal(dec[:n], q)


def test_pack_reads_layout(sb):
    codes, quals, off, bc, _ = datasets.get("stress1")
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off, threads=2)
    assert np.array_equal(ln, np.diff(off).astype(np.uint32))
    assert np.array_equal(np.diff(boff), (ln.astype(np.uint64) + 3) // 4)
    r = 17
    b = pb[int(boff[r]):int(boff[r + 1])]
    dec = np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).ravel()[:ln[r]]
    assert np.array_equal(dec, codes[int(off[r]):int(off[r + 1])])
