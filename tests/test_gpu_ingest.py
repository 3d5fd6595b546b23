"""-m gpu: ingest on the device -- the barcoded pseudo-FASTQ of the pipeline
parsed, 2-bit packed and PQVec-encoded by CUDA kernels (sn_load_fasth_text) against
  * the committed golden reads.fastb/.qualp/.bci the reference's ParseBarcodedFastqs wrote, and
  * ParseBarcodedFastqs itself (oracle/_ref) run on the same text, including the odd records:
    unbarcoded reads between barcoded ones, ',raw' suffixes, N and lower-case bases, N inside a
    barcode, a barcode that comes back later, ragged read lengths.
Bar: byte-identical files."""
import gzip
import io
import os

import numpy as np
import pytest

import datasets
import refrun
from supernova_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


def fasth_text(name, tmp):
    codes, quals, off, bc, ids = datasets.get(name)
    p = os.path.join(tmp, name + ".fastq.gz")
    synth.write_fasth_ragged(p, codes, quals, off, ids)
    return p, gzip.open(p, "rb").read()


def files(head):
    return [open(head + e, "rb").read() for e in (".fastb", ".qualp", ".bci")]


@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_golden_read_files(sb, name, tmp_path):
    _, text = fasth_text(name, str(tmp_path))
    with sb.Context(0) as ctx:
        ctx.load_fasth_text(text)
        ctx.save_read_files(str(tmp_path / "mine"))
    mine = files(str(tmp_path / "mine"))
    for got, f in zip(mine, ("reads.fastb", "reads.qualp", "reads.bci")):
        want = gzip.open(os.path.join(HERE, "golden", name, f + ".gz"), "rb").read()
        assert got == want, f


def odd_text():
    rng = np.random.default_rng(11)

    def rec(i, n1, n2, barcode, lower=False, with_n=False):
        def seq(n):
            s = "".join("ACGT"[x] for x in rng.integers(0, 4, size=n))
            if with_n and n > 10:
                s = s[:3] + "N" + s[4:9] + "n" + s[10:]
            return s.lower() if lower else s

        def qual(n):
            q = "".join(chr(33 + int(x)) for x in rng.choice([2, 12, 20, 30, 37, 41], size=n))
            if with_n and n > 20:       # 'N' (Q45) and 'n' (Q77) in a QUALITY line: the reference folds them to 'A' = Q32 like everywhere else (:84-85)
                q = q[:5] + "N" + q[6:12] + "n" + q[13:]
            return q
        return "@r%d some name\n%s\n%s\n%s\n%s\n%s\n%s\nACGTACGT\nIIIIIIII\n" % (i, seq(n1), qual(n1), seq(n2), qual(n2), barcode, "I" * 16)
    recs = [
        rec(0, 150, 150, "NNNNNNNNNNNNNNNN"),                 # no gem group: unbarcoded
        rec(1, 150, 151, "AAACCCGGGTTTAAAC-1,AAACCCGGGTTTAAAT"),
        rec(2, 98, 150, "AAACCCGGGTTTAAAC-1,AAACCCGGGTTTAAAG"),   # same corrected barcode, other raw one
        rec(3, 150, 150, "-1"),                               # gem group without a barcode: unbarcoded
        rec(4, 47, 48, "AAACCCGGGTTTAAAC-1"),                 # still the same barcode (unbarcoded records do not reset it)
        rec(5, 150, 150, "AANCCCGGGTTTAAAC-1", with_n=True),  # N -> A inside the barcode too: the same barcode again
        rec(6, 150, 150, "CCCCCCCCCCCCCCCC-1", lower=True),
        rec(7, 150, 150, "CCCCCCCCCCCCCCCC-2"),               # another gem group = another barcode
        rec(8, 1, 2, "GGGGGGGGGGGGGGGG-1"),
        rec(9, 150, 150, "AAACCCGGGTTTAAAC-1"),               # a barcode that comes back: counted again
        rec(10, 150, 150, "TTTT"),
        rec(11, 256, 200, "TTTTTTTTTTTTTTTT-1,x"),
    ]
    return "".join(recs).encode()


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("which", ["odd", "C1", "stress3"])
def test_against_parse_barcoded_fastqs(sb, which, tmp_path):
    wd = str(tmp_path)
    if which == "odd":
        text = odd_text()
        p = os.path.join(wd, "odd.fastq.gz")
        with gzip.open(p, "wb") as f:
            f.write(text)
    else:
        p, text = fasth_text(which, wd)
    refrun.parse_fastqs(wd, p)                      # -> wd/reads.{fastb,qualp,bci}
    with sb.Context(0) as ctx:
        ctx.load_fasth_file(p)                      # the .gz path (zlib on the host, the rest on the device)
        ctx.save_read_files(wd + "/mine")
        c = ctx.counts()
    assert files(wd + "/mine") == files(wd + "/reads")
    assert c["n_reads"] * 9 == 2 * text.count(b"\n")


def test_ingest_then_build_matches_packed_load(sb, tmp_path):
    codes, quals, off, bc, ids = datasets.get("C1")
    _, text = fasth_text("C1", str(tmp_path))
    pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
    with sb.Context(0) as a, sb.Context(0) as b:
        a.load_reads(pb, boff, ln, pq, pqoff, bc)
        a.build_read_qgraph48(None, sb.Params(), with_paths=True)
        b.load_fasth_text(text)
        b.build_read_qgraph48(None, sb.Params(), with_paths=True)
        assert np.array_equal(a.kmers(), b.kmers())
        ha, hb = a.hbv(), b.hbv()
        assert all(np.array_equal(ha[x], hb[x]) for x in ha)
        assert all(np.array_equal(x, y) for x, y in zip(a.paths(), b.paths()))


def test_ingest_errors_are_loud(sb):
    good = odd_text()
    with sb.Context(0) as ctx:
        i = good.index(b"\n") + 5
        for bad, what in ((good[:-1], "newline"), (good + b"@x\nACGT\n", "9 per record"), (good.replace(b"@r3", b"#r3"), "'@'"),
                          (good[:i] + b"R" + good[i + 1:], "ACGTN"), (good.replace(b"#", b"~", 1), "quality character")):
            with pytest.raises(sb.SnError) as e:
                ctx.load_fasth_text(bad)
            assert what in str(e.value), (what, str(e.value))
        ctx.load_fasth_text(good)                   # and the context still works afterwards
        assert ctx.counts()["n_reads"] == 24


def test_odd_text_matches_the_restatement(sb, tmp_path):
    """The hand-made text against oracle/dfside.py (the python restatement of ParseBarcodedFastqs, pinned to the
    golden files in tests/test_oracle_golden.py): needs no reference binary on the box."""
    from oracle import dfside
    text = odd_text()
    with sb.Context(0) as ctx:
        ctx.load_fasth_text(text)
        ctx.save_read_files(str(tmp_path / "mine"))
    assert files(str(tmp_path / "mine")) == list(dfside.parse_fasth(text))


def test_two_input_files(sb, tmp_path):
    """sn_load_fasth_files = ParseBarcodedFastqs FASTQS={a,b}: golden files (the tiny set cut in two inside a barcode),
    the python restatement, and the reference binary when it is on the box."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden import split_text
    from oracle import dfside
    wd = str(tmp_path)
    _, text = fasth_text("tiny", wd)
    a, b = split_text(text)
    pa, pb_ = wd + "/a.fastq.gz", wd + "/b.fasth"
    with gzip.open(pa, "wb") as f:
        f.write(a)
    with open(pb_, "wb") as f:                       # (the second one not compressed)
        f.write(b)
    with sb.Context(0) as ctx:
        ctx.load_fasth_files([pa, pb_])
        ctx.save_read_files(wd + "/mine")
    mine = files(wd + "/mine")
    g = os.path.join(HERE, "golden", "tiny")
    assert mine == [gzip.open(g + "/split." + f + ".gz", "rb").read() for f in ("reads.fastb", "reads.qualp", "reads.bci")]
    assert mine == list(dfside.parse_fasth([a, b]))
    if refrun.have_ref():
        pb2 = wd + "/b.fastq.gz"
        with gzip.open(pb2, "wb") as f:
            f.write(b)
        refrun.parse_fastq_set(wd, [pa, pb2])
        assert mine == files(wd + "/reads")
