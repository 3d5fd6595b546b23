"""The C++ host driver over the C ABI (supernova_b200/csrc/sn_cli.cpp -> supernova_b200/sn_build_graph).
CPU: it builds, prints its usage, and fails LOUDLY without a CUDA device (no CPU fallback).
GPU: fasth.gz in -> every file the reference leaves behind, against the golden ones."""
import gzip
import os
import subprocess

import pytest

import datasets
from supernova_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "supernova_b200", "sn_build_graph")
GOLD = os.path.join(ROOT, "tests", "golden")


def test_cli_usage_and_loud_failure(built, tmp_path):
    assert os.access(EXE, os.X_OK)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: sn_build_graph" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([EXE, "HEAD=" + str(tmp_path / "reads"), "OUT=" + str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_scale_tool_usage_and_loud_failure(built):
    """supernova_b200/sn_scale (csrc/sn_scale.cpp: scale runs on device-generated reads): builds, and refuses to run without a GPU"""
    exe = os.path.join(ROOT, "supernova_b200", "sn_scale")
    assert os.access(exe, os.X_OK)
    r = subprocess.run([exe, "bogus"], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: sn_scale" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "NGPU=1", "MULT=0.01"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_scale_tool_one_gpu_matches_the_library(built, tmp_path):
    """sn_scale NGPU=1 on a small job: its a.hbv is the one the library writes for the same generated reads through Python"""
    import json
    import supernova_b200 as sb
    exe = os.path.join(ROOT, "supernova_b200", "sn_scale")
    hbv = str(tmp_path / "scale.hbv")
    r = subprocess.run([exe, "NGPU=1", "G=500000", "PAIRS=40000", "NBC=2000", "SEED=11", "HBV=" + hbv, "PASSES=3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"] and out["reads"] == 80000 and out["kmers"] > 400000
    with sb.Context(0) as ctx:
        ctx.generate_reads(500000, 40000, 2000, 11)
        ctx.build_read_qgraph48(str(tmp_path), sb.Params(), with_paths=False)
        assert ctx.counts()["n_kmers"] == out["kmers"] and ctx.counts()["n_edges"] == out["unipaths"]
    assert open(hbv, "rb").read() == open(str(tmp_path / "a.hbv"), "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_cli_reproduces_the_reference_files(built, name, tmp_path):
    codes, quals, off, bc, ids = datasets.get(name)
    wd = str(tmp_path)
    fq = wd + "/in.fastq.gz"
    synth.write_fasth_ragged(fq, codes, quals, off, ids)
    r = subprocess.run([EXE, "FASTH=" + fq, "HEAD=" + wd + "/reads", "OUT=" + wd, "INDEX=True"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "k-mers" in r.stdout
    g = os.path.join(GOLD, name)
    for f in ("reads.fastb", "reads.qualp", "reads.bci", "a.hbv", "tmp.paths", "a.inv", "a.to_left", "a.to_right", "a.paths.inv", "a.countsb",
              "a.hbx", "a.fastb", "a.kmers", "a.pathsX", "a.dup"):
        assert open(wd + "/" + f, "rb").read() == gzip.open(g + "/" + f + ".gz", "rb").read(), f
    assert open(wd + "/a.k").read() == "48\n"
    import json
    st = json.load(open(g + "/dup_stats.json"))
    assert st["dup_perc"] + "% of pairs appear to be duplicates" in r.stdout
    assert st["art_dup_perc"] + "% of pairs appear to be artifactual duplicates" in r.stdout
    assert st["interdup_perc"] + "% of duplicates involve multiple barcodes" in r.stdout
    assert open(wd + "/stats/histogram_kmer_count.json").read() == open(g + "/histogram_kmer_count.json").read()
    # and from the files alone (no FASTH): the same graph
    out2 = wd + "/again"
    os.makedirs(out2)
    r = subprocess.run([EXE, "HEAD=" + wd + "/reads", "OUT=" + out2, "PATHS=False"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out2 + "/a.hbv", "rb").read() == open(wd + "/a.hbv", "rb").read()
    assert not os.path.exists(out2 + "/tmp.paths")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "stress1"])
def test_cli_multi_gpu_without_python(built, name, tmp_path):
    """NGPU=2: the sharded path driven by the C++ host alone -- a thread and a context per GPU, NCCL inside the library,
    no launcher, no torch.  The files equal the golden ones of the reference."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    codes, quals, off, bc, ids = datasets.get(name)
    wd = str(tmp_path)
    fq = wd + "/in.fastq.gz"
    synth.write_fasth_ragged(fq, codes, quals, off, ids)
    r = subprocess.run([EXE, "FASTH=" + fq, "HEAD=" + wd + "/reads", "OUT=" + wd, "NGPU=2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "2 GPUs:" in r.stdout, r.stdout + r.stderr
    g = os.path.join(GOLD, name)
    for f in ("a.hbv", "tmp.paths"):
        assert open(wd + "/" + f, "rb").read() == gzip.open(g + "/" + f + ".gz", "rb").read(), f
    assert open(wd + "/stats/histogram_kmer_count.json").read() == open(g + "/histogram_kmer_count.json").read()
    r = subprocess.run([EXE, "HEAD=" + wd + "/reads", "OUT=" + wd, "NGPU=2", "PATHS=False"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(wd + "/a.hbv", "rb").read() == gzip.open(g + "/a.hbv.gz", "rb").read()
