"""Generates the committed golden fixtures by running the REFERENCE's own binaries
(oracle/_ref, built from /root/reference by oracle/build_ref.sh) in this container.
For each data set: the input files written by the reference's ParseBarcodedFastqs
(reads.fastb/.qualp/.bci) and the outputs of its buildReadQGraph48 (kmers.kvec reduced to
the sorted {k-mer,count,ctx} records, a.hbv, tmp.paths, histogram_kmer_count.json) and of the DF-side
step that follows (Involution, ToLeft/ToRight, writePathsIndex: a.inv, a.to_left, a.to_right, a.paths.inv, a.countsb;
HyperBasevectorX, the edge files, ReadPathVecX and MarkDups: a.hbx, a.fastb, a.kmers, a.pathsX, a.dup + the three
percentages MarkDups prints).

    python tests/golden/make_golden.py        # needs oracle/_ref
"""
import gzip
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402
import refrun  # noqa: E402
from supernova_b200 import synth  # noqa: E402

SETS = ["tiny", "stress1", "dupes"]
DF_FILES = ("a.hbx", "a.fastb", "a.kmers", "a.pathsX", "a.dup")


def split_text(text):
    """two files out of one: cut at the record boundary just after the middle (inside a barcode's run of records)"""
    lines = text.split(b"\n")[:-1]
    cut = 9 * ((len(lines) // 9) // 2 + 1)
    return b"\n".join(lines[:cut]) + b"\n", b"\n".join(lines[cut:]) + b"\n"


def main():
    assert refrun.have_ref(), "build oracle/_ref first (bash oracle/build_ref.sh)"
    for name in SETS:
        codes, quals, off, bc, ids = datasets.get(name)
        wd = tempfile.mkdtemp()
        synth.write_fasth_ragged(wd + "/reads.fastq.gz", codes, quals, off, ids)
        refrun.parse_fastqs(wd, wd + "/reads.fastq.gz")
        extra = ("INDEX=True", "DFSIDE=True")             # + a.inv, a.to_left/right, a.paths.inv, a.countsb, a.hbx, ... (DF side, 10X/DF.cc:573-600)
        if name == "dupes":
            P = datasets.DUPES_PARAMS
            extra += ("MIN_QUAL=%d" % P["min_qual"], "MIN_FREQ=%d" % P["min_freq"], "MIN_BC=%d" % P["min_bc"])
        _, log = refrun.run_probe(wd, extra=extra)
        out = os.path.join(HERE, name)
        os.makedirs(out, exist_ok=True)
        stats = {}
        for key, pat in (("dup_perc", r"([0-9.]+|-?nan)% of pairs appear to be duplicates"), ("interdup_perc", r"([0-9.]+|-?nan)% of duplicates involve"),
                         ("art_dup_perc", r"([0-9.]+|-?nan)% of pairs appear to be artifactual")):
            stats[key] = re.search(pat, log).group(1)
        with open(out + "/dup_stats.json", "w") as f:
            json.dump(stats, f)
        for f in ("reads.fastb", "reads.qualp", "reads.bci", "a.hbv", "tmp.paths", "a.inv", "a.to_left", "a.to_right", "a.paths.inv", "a.countsb") + DF_FILES:
            with open(os.path.join(wd, f), "rb") as src, gzip.GzipFile(os.path.join(out, f + ".gz"), "wb", mtime=0) as dst:
                dst.write(src.read())
        shutil.copy(wd + "/stats/histogram_kmer_count.json", out + "/histogram_kmer_count.json")
        if name == "tiny":      # the same text as two input files, cut inside a barcode (FASTQS={a,b}: ParseBarcodedFastqs.cc:258-264)
            a, b = split_text(gzip.open(wd + "/reads.fastq.gz", "rb").read())
            sd = wd + "/split"
            os.makedirs(sd)
            for nm, t in (("a", a), ("b", b)):
                with gzip.open(sd + "/" + nm + ".fastq.gz", "wb") as f:
                    f.write(t)
            refrun.parse_fastq_set(sd, [sd + "/a.fastq.gz", sd + "/b.fastq.gz"])
            for f in ("reads.fastb", "reads.qualp", "reads.bci"):
                with open(os.path.join(sd, f), "rb") as src, gzip.GzipFile(os.path.join(out, "split." + f + ".gz"), "wb", mtime=0) as dst:
                    dst.write(src.read())
        kv = refrun.read_kvec(wd + "/kmers.kvec")
        if name != "dupes":                # (the DF-side set: its dictionary is not a fixture)
            np.save(out + "/kvec_sorted.npy", kv)
        shutil.rmtree(wd)
        print(name, "->", out, kv.shape)


if __name__ == "__main__":
    main()
