"""CPU: the C oracle (oracle/sn_oracle.c) against the committed golden vectors produced by
the reference's own binaries (tests/golden/make_golden.py), and -- when oracle/_ref is
present -- against the reference run live on more inputs.  This is what pins the oracle."""
import gzip
import os

import numpy as np
import pytest

import datasets
import refrun
from oracle.oracle import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gz(path):
    with gzip.open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_oracle_matches_golden(name, tmp_path):
    codes, quals, off, bc, _ = datasets.get(name)
    o = Oracle(codes, quals, off, bc).run()
    kv = np.load(os.path.join(GOLD, name, "kvec_sorted.npy"))
    km = o.kmers()
    assert np.array_equal(km[:, :5], kv)                       # k-mer, count, context (pre-prune)
    o.write_hbv(str(tmp_path / "o.hbv"))
    o.write_paths(str(tmp_path / "o.paths"))
    assert open(tmp_path / "o.hbv", "rb").read() == gz(os.path.join(GOLD, name, "a.hbv.gz"))
    assert open(tmp_path / "o.paths", "rb").read() == gz(os.path.join(GOLD, name, "tmp.paths.gz"))


@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_golden_spectrum(name):
    import json
    codes, quals, off, bc, _ = datasets.get(name)
    o = Oracle(codes, quals, off, bc).stage("count")
    spec = json.load(open(os.path.join(GOLD, name, "histogram_kmer_count.json")))["vals"]
    counts = np.bincount(o.kmers()[:, 3].astype(np.int64), minlength=len(spec))
    assert counts.tolist() == spec


def test_oracle_invariants():
    """Invariants the reference's own (Rust) tests state for this path:
    every k-mer of an edge is a dictionary k-mer, edges partition the dictionary,
    rc(rc(x)) == x through the involution."""
    codes, quals, off, bc, _ = datasets.get("stress2")
    o = Oracle(codes, quals, off, bc).run(with_paths=False)
    km = o.kmers()
    edges = o.edges()
    assert sum(len(e) - 47 for e in edges) == km.shape[0]
    inv = o.involution()
    assert np.array_equal(inv[inv], np.arange(len(inv)))


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("name", ["stress2", "stress3", "C1"])
def test_oracle_matches_reference_live(name, tmp_path):
    from supernova_b200 import synth
    codes, quals, off, bc, ids = datasets.get(name)
    wd = str(tmp_path)
    synth.write_fasth_ragged(wd + "/reads.fastq.gz", codes, quals, off, ids)
    refrun.parse_fastqs(wd, wd + "/reads.fastq.gz")
    refrun.run_probe(wd)
    o = Oracle(codes, quals, off, bc).run()
    assert np.array_equal(o.kmers()[:, :5], refrun.read_kvec(wd + "/kmers.kvec"))
    o.write_hbv(wd + "/o.hbv")
    o.write_paths(wd + "/o.paths")
    assert open(wd + "/o.hbv", "rb").read() == open(wd + "/a.hbv", "rb").read()
    assert open(wd + "/o.paths", "rb").read() == open(wd + "/tmp.paths", "rb").read()


@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_paths_index_restatement_matches_golden(name):
    """oracle/dfside.py (writePathsIndex, 10X/PathsIndex.cc:23-143) on the golden tmp.paths + a.inv
    reproduces the golden a.paths.inv / a.countsb the reference wrote."""
    from oracle import dfside
    g = os.path.join(GOLD, name)
    paths = dfside.read_paths(gz(g + "/tmp.paths.gz"))
    inv = dfside.read_vec_int(gz(g + "/a.inv.gz"))
    pi, cb = dfside.paths_index(paths, inv)
    assert pi == gz(g + "/a.paths.inv.gz")
    assert cb == gz(g + "/a.countsb.gz")


@pytest.mark.parametrize("name", ["tiny", "stress1", "dupes"])
def test_df_files_restatement_matches_golden(name):
    """oracle/dfside.py on the golden a.hbv / tmp.paths / reads.* reproduces what the reference's own HyperBasevectorX,
    vecbvec / vec<int> writers, ReadPathVecX and MarkDups wrote next (10X/WriteFiles.cc:16-60, 10X/DF.cc:573-600):
    a.hbx, a.fastb, a.kmers, a.pathsX, a.dup and the three percentages MarkDups prints."""
    import json
    from oracle import dfside
    g = os.path.join(GOLD, name)
    h = dfside.read_hbv(gz(g + "/a.hbv.gz"))
    paths = dfside.read_paths(gz(g + "/tmp.paths.gz"))
    assert dfside.hbx_file(h) == gz(g + "/a.hbx.gz")
    assert dfside.edges_fastb_file(h) == gz(g + "/a.fastb.gz")
    assert dfside.kmers_file(h) == gz(g + "/a.kmers.gz")
    assert dfside.pathsx_file(paths, h) == gz(g + "/a.pathsX.gz")
    bases = dfside.read_fastb(gz(g + "/reads.fastb.gz"))
    quals = dfside.read_qualp(gz(g + "/reads.qualp.gz"))
    bc = dfside.expand_bci(gz(g + "/reads.bci.gz"))
    dup, ndups, interdups, art = dfside.mark_dups(paths, bases, quals, bc)
    assert dfside.dup_file(dup) == gz(g + "/a.dup.gz")
    assert dfside.dup_percentages(dup, ndups, interdups, art) == json.load(open(g + "/dup_stats.json"))
    if name == "dupes":          # the set has what it is for: ties, artifactual duplicates, and offsets that only collide in 16 bits
        assert sum(art) > 0 and max(o for o, _ in paths) > 65536


@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_ingest_restatement_matches_golden(name, tmp_path):
    """oracle/dfside.py (ParseBarcodedFastqs, 10X/ParseBarcodedFastqs.cc:56-146,284-303; PQVecEncoder,
    feudal/PQVec.cc:17-127) on the data set's pseudo-FASTQ text reproduces the golden reads.fastb/.qualp/.bci."""
    from oracle import dfside
    from supernova_b200 import synth
    codes, quals, off, bc, ids = datasets.get(name)
    p = str(tmp_path / "x.fastq.gz")
    synth.write_fasth_ragged(p, codes, quals, off, ids)
    fb, qp, bci = dfside.parse_fasth(gz(p))
    g = os.path.join(GOLD, name)
    assert fb == gz(g + "/reads.fastb.gz") and qp == gz(g + "/reads.qualp.gz") and bci == gz(g + "/reads.bci.gz")


def test_two_file_ingest_restatement_matches_golden(tmp_path):
    """FASTQS={a,b}: the barcode ordinal runs on across the files, the comparison string starts empty in each
    (ParseBarcodedFastqs.cc:66-67,258-264) -- oracle/dfside.py against what the reference wrote for the tiny set cut in two."""
    import sys
    sys.path.insert(0, os.path.join(GOLD))
    from make_golden import split_text
    from oracle import dfside
    from supernova_b200 import synth
    codes, quals, off, bc, ids = datasets.get("tiny")
    p = str(tmp_path / "x.fastq.gz")
    synth.write_fasth_ragged(p, codes, quals, off, ids)
    a, b = split_text(gz(p))
    fb, qp, bci = dfside.parse_fasth([a, b])
    g = os.path.join(GOLD, "tiny")
    assert fb == gz(g + "/split.reads.fastb.gz") and qp == gz(g + "/split.reads.qualp.gz") and bci == gz(g + "/split.reads.bci.gz")
    assert bci != gz(g + "/reads.bci.gz")                  # the cut really opens one more barcode


@pytest.mark.parametrize("name", ["tiny", "stress1", "stress4"])
def test_oracle_tada_switch_against_an_independent_restatement(name):
    """count_len_k (the tada rule, lib/tada/src/cmd_msp.rs:109-146, utils.rs:322-408): every k-mer of every read trimmed to
    >= K bases by find_trim_len, canonical (min of k-mer and reverse complement), valid iff seen >= min_kmer_obs times and
    under more than one barcode (bc 0 = none).  Restated here with python dicts, independently of the C oracle's sort.
    PARITY UNPINNED against the Rust binary (no rustc in this image); pinned against this restatement only."""
    codes, quals, off, bc, _ = datasets.get(name)
    K = 48
    obs, bcs = {}, {}
    n_len_k = 0
    for r in range(len(off) - 1):
        b = codes[int(off[r]):int(off[r + 1])]
        q = quals[int(off[r]):int(off[r + 1])]
        good, trim = 0, 0
        for i in range(len(q) - 1, -1, -1):                     # find_trim_len
            if q[i] < 7:
                good = 0
            else:
                good += 1
                if good == K:
                    trim = i + K
                    break
        if trim < K:
            continue
        n_len_k += trim == K
        s = bytes(b[:trim])
        rc = bytes(3 - x for x in reversed(s))
        for i in range(trim - K + 1):
            f, v = s[i:i + K], rc[trim - K - i:trim - i]
            k = min(f, v)
            obs[k] = obs.get(k, 0) + 1
            if bc[r] > 0:
                bcs.setdefault(k, set()).add(int(bc[r]))
    valid = {k: c for k, c in obs.items() if c >= 3 and len(bcs.get(k, ())) > 1}
    o = Oracle(codes, quals, off, bc, count_len_k=True).stage("count")
    km = o.kmers()
    mine = {}
    for row in km:
        w = (int(row[0]) << 64) | (int(row[1]) << 32) | int(row[2])
        mine[bytes((w >> (2 * (K - 1 - i))) & 3 for i in range(K))] = int(row[3])
    assert mine == valid
    assert n_len_k > 0 or name == "tiny"


def _qv_trim_vectors():
    """The reference's own known-answer test for the trim rule, lib/tada/src/cmd_msp.rs:329-350 (test_qv_trim_read): an 88-base
    read whose quals all pass min_qual = 10; one qual at a time is set to Q1; expected length = 88 if i < 88 - K, else i if
    i >= K, else 0.  Returned as reads (codes, quals, off) with the expected lengths."""
    seq = "TAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAACCCTAAC"
    q = "FFFFFFFFIFFFFFFFFFFIIIFFBFIFFFFIFBFFIBFIFBBFFIFFIFFFFFFFFFFFBBBBBBBBBB07BB7BB<BBBBBBBBBB"
    assert len(seq) == len(q) == 88
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    b = np.array([code[c] for c in seq], np.uint8)
    base_q = np.array([ord(c) - 33 for c in q], np.uint8)
    n = len(seq)
    codes = np.tile(b, n)
    quals = np.tile(base_q, n).reshape(n, n).copy()
    quals[np.arange(n), np.arange(n)] = 1                      # (myquals[i] = 34 as u8)
    expect = np.array([n if i < n - 48 else (i if i >= 48 else 0) for i in range(n)], np.uint32)
    return codes, quals.ravel(), np.arange(n + 1, dtype=np.uint64) * n, expect


def test_trim_rule_against_the_references_known_answers():
    codes, quals, off, expect = _qv_trim_vectors()
    o = Oracle(codes, quals, off, None, min_qual=10).stage("count")
    assert np.array_equal(o.good_len(), expect)
