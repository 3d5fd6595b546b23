"""The Martian exec-stage adapter of the three graph stages (supernova_b200/csrc/sn_martian.cpp -> supernova_b200/sn_martian;
SURVEY §8(b) boundary B1: lib/tada/mro/_asm_stages.mro:20-47, protocol external/martian/src/lib.rs).
CPU: the protocol -- what is read from and written to the chunk's metadata directory, journal files, the stages that need
no device, the loud failure of the one that does.  GPU: MSP -> SHARD_ASM -> MAIN_ASM_SN chained the way Martian chains
them (split, main per chunk, join); `asm_graph` must be the edge set of the tada rule -- through the reference's own
buildGraphFromMSP (oracle/_ref) and through sn_build_graph_from_edges it gives the HyperBasevector of the oracle."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import datasets
import refrun
from supernova_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "supernova_b200", "sn_martian")


class Chunk:
    """one invocation: a metadata directory with the files Martian puts there"""

    def __init__(self, base, name, stage, kind, args, outs=None, chunk_defs=None, chunk_outs=None):
        self.md = os.path.join(base, name)
        self.files = os.path.join(self.md, "files")
        os.makedirs(self.files)
        self.run_file = os.path.join(base, "journal", name)
        os.makedirs(os.path.dirname(self.run_file), exist_ok=True)
        self.stage, self.kind = stage, kind

        def put(n, v):
            with open(os.path.join(self.md, "_" + n), "w") as f:
                json.dump(v, f)
        put("args", args)
        put("jobinfo", {"name": name, "threads": 1, "memGB": 1, "monitor_flag": "disable"})
        if outs is not None:
            put("outs", outs)
        if chunk_defs is not None:
            put("chunk_defs", chunk_defs)
        if chunk_outs is not None:
            put("chunk_outs", chunk_outs)

    def run(self):
        return subprocess.run([EXE, "martian", self.stage, self.kind, self.md, self.files, self.run_file], cwd=self.files, capture_output=True, text=True)

    def get(self, n):
        return json.load(open(os.path.join(self.md, "_" + n)))

    def has(self, n):
        return os.path.exists(os.path.join(self.md, "_" + n))

    def journal(self, n):
        return os.path.exists(self.run_file + "." + (n if self.kind == "main" else self.kind + "_" + n))


def _ok(c):
    r = c.run()
    assert r.returncode == 0, r.stderr + (open(os.path.join(c.md, "_errors")).read() if c.has("errors") else "")
    assert c.has("complete") and not c.has("errors")
    assert c.journal("complete") and c.journal("heartbeat") and c.journal("log") and c.journal("jobinfo")
    ji = c.get("jobinfo")
    assert ji["cwd"] == c.files and ji["pid"] > 0 and "wallclock" in ji and "rusage" in ji and ji["name"]      # (what was there stays)
    log = open(os.path.join(c.md, "_log")).read()
    assert "__start__" in log and "__end__" in log
    return c


def _front_stages(base, fastqs, trim_min_qual=7, min_kmer_obs=3):
    """MSP and SHARD_ASM the way Martian runs them; returns SHARD_ASM's outs"""
    msp_args = {"trim_min_qual": trim_min_qual, "fastqs": fastqs, "barcode_whitelist": "/nonexistent/whitelist.txt"}
    sp = _ok(Chunk(base, "msp_split", "msp", "split", msp_args))
    defs = sp.get("stage_defs")["chunks"]
    assert sp.journal("stage_defs") and [f for d in defs for f in d["chunk"]] == fastqs and all(len(d["chunk"]) <= 8 for d in defs)
    assert all("__mem_gb" in d and "__threads" in d and os.path.exists(d["permutation"]) for d in defs)
    couts = []
    for i, d in enumerate(defs):
        a = dict(msp_args); a.update(d)
        c = Chunk(base, "msp_chnk%d" % i, "msp", "main", a, outs={"chunks": None})
        c_out = os.path.join(c.files, "chunks.msp")
        json.dump({"chunks": c_out}, open(os.path.join(c.md, "_outs"), "w"))
        _ok(c)
        assert c.get("outs") == {"chunks": c_out} and os.path.exists(c_out)
        couts.append(c.get("outs"))
    jn = _ok(Chunk(base, "msp_join", "msp", "join", msp_args, outs={"chunks": []}, chunk_defs=defs, chunk_outs=couts))
    msp_out = jn.get("outs")
    assert msp_out["chunks"] == [o["chunks"] for o in couts]
    sh_args = {"min_kmer_obs": min_kmer_obs, "chunks": msp_out["chunks"]}
    sp = _ok(Chunk(base, "shard_split", "shard-asm", "split", sh_args))
    defs = sp.get("stage_defs")["chunks"]
    assert len(defs) >= 1 and all(d["total_chunks"] == len(defs) for d in defs)
    couts = []
    for i, d in enumerate(defs):
        a = dict(sh_args); a.update(d)
        c = Chunk(base, "shard_chnk%d" % i, "shard-asm", "main", a)
        o = {"sedge_asm": os.path.join(c.files, "sedge_asm.sedge_asm"), "sedge_bcs": os.path.join(c.files, "sedge_bcs.sedge_bcs")}
        json.dump(o, open(os.path.join(c.md, "_outs"), "w"))
        _ok(c)
        couts.append(c.get("outs"))
    jn = _ok(Chunk(base, "shard_join", "shard-asm", "join", sh_args, outs={"sedge_asm": [], "sedge_bcs": []}, chunk_defs=defs, chunk_outs=couts))
    out = jn.get("outs")
    assert len(out["sedge_asm"]) == len(out["sedge_bcs"]) == len(defs) and all(os.path.exists(f) for f in out["sedge_asm"] + out["sedge_bcs"])
    assert all(os.path.dirname(f) == jn.files for f in out["sedge_asm"])          # renamed into the stage directory (cmd_shard_asm.rs:143-158)
    return out


def _fastqs(base, name, n_files):
    codes, quals, off, bc, ids = datasets.get(name)
    whole = os.path.join(base, "all.fastq.gz")
    synth.write_fasth_ragged(whole, codes, quals, off, ids)
    if n_files == 1:
        return [whole]
    lines = gzip.open(whole, "rb").read().split(b"\n")[:-1]
    nrec = len(lines) // 9
    out = []
    for i in range(n_files):
        a, b = 9 * (nrec * i // n_files), 9 * (nrec * (i + 1) // n_files)
        p = os.path.join(base, "part%d.fastq.gz" % i)
        with gzip.open(p, "wb") as f:
            f.write(b"\n".join(lines[a:b]) + b"\n")
        out.append(p)
    return out


def test_protocol_of_the_stages_without_a_device(built, tmp_path):
    assert os.access(EXE, os.X_OK)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: sn_martian martian" in r.stderr
    base = str(tmp_path)
    fq = [os.path.join(base, "f%02d.fastq.gz" % i) for i in range(11)]           # 11 files: chunks of 8 and 3
    out = _front_stages(base, fq, trim_min_qual=9, min_kmer_obs=4)
    text = open(out["sedge_asm"][0]).read()
    assert "trim_min_qual 9" in text and "min_kmer_obs 4" in text and all("fastq " + f in text for f in fq)
    # an unknown stage, a missing _args: exit 1 and _errors, never a silent success
    c = Chunk(base, "bad_stage", "sort-bcs", "main", {})
    r = c.run()
    assert r.returncode == 1 and c.has("errors") and not c.has("complete") and "unknown stage" in open(os.path.join(c.md, "_errors")).read()
    c = Chunk(base, "bad_args", "msp", "main", {"chunk": ["x"]}, outs={"chunks": "/nonexistent/dir/x.msp"})
    r = c.run()
    assert r.returncode == 1 and "trim_min_qual" in open(os.path.join(c.md, "_errors")).read()
    # MAIN_ASM_SN: split needs no device; main does and says so
    sp = _ok(Chunk(base, "main_split", "main-asm-sn", "split", out))
    assert len(sp.get("stage_defs")["chunks"]) == 1
    import torch
    if not torch.cuda.is_available():
        c = Chunk(base, "main_chnk0", "main-asm-sn", "main", out, outs={"asm_graph": os.path.join(base, "asm_graph.bv")})
        r = c.run()
        assert r.returncode == 1 and c.has("errors") and not c.has("complete")
        assert "no CPU fallback" in open(os.path.join(c.md, "_errors")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_files", [("tiny", 1), ("stress1", 3), ("stress4", 10)])
def test_the_three_stages_produce_the_asm_graph(built, name, n_files, tmp_path):
    import supernova_b200 as sb
    from oracle.oracle import Oracle
    base = str(tmp_path)
    fq = _fastqs(base, name, n_files)
    out = _front_stages(base, fq)
    sp = _ok(Chunk(base, "main_split", "main-asm-sn", "split", out))
    defs = sp.get("stage_defs")["chunks"]
    a = dict(out); a.update(defs[0])
    c = Chunk(base, "main_chnk0", "main-asm-sn", "main", a)
    chunk_graph = os.path.join(c.files, "asm_graph.bv")
    json.dump({"asm_graph": chunk_graph}, open(os.path.join(c.md, "_outs"), "w"))
    _ok(c)
    final = os.path.join(base, "asm_graph.bv")
    jn = _ok(Chunk(base, "main_join", "main-asm-sn", "join", out, outs={"asm_graph": final}, chunk_defs=defs, chunk_outs=[c.get("outs")]))
    assert jn.get("outs") == {"asm_graph": final} and os.path.exists(final) and not os.path.exists(chunk_graph)
    # the edge set is the oracle's under the tada rule (reads trimmed to exactly K count) ...
    codes, quals, off, bc, _ = datasets.get(name)
    o = Oracle(codes, quals, off, bc, count_len_k=True).run()
    d = open(final, "rb").read()
    assert d[:8] == b"BINWRITE"
    n = int(np.frombuffer(d, "<u8", 1, 8)[0])
    edges, p = [], 16
    for _ in range(n):
        nb = int(np.frombuffer(d, "<u4", 1, p)[0]); p += 4
        by = np.frombuffer(d, np.uint8, (nb + 3) // 4, p); p += (nb + 3) // 4
        edges.append(np.stack([(by >> (2 * j)) & 3 for j in range(4)], axis=1).ravel()[:nb].astype(np.uint8).tobytes())
    assert p == len(d) and sorted(edges) == sorted(o.edges())
    # ... and what DF builds from the file (buildGraphFromMSP) is the oracle's HyperBasevector
    wd = os.path.join(base, "from_edges")
    os.makedirs(wd)
    o.write_hbv(wd + "/oracle.hbv")
    with sb.Context(0) as ctx:
        ctx.build_graph_from_edges(final)
        ctx.write_hbv(wd + "/a.hbv")
    assert open(wd + "/a.hbv", "rb").read() == open(wd + "/oracle.hbv", "rb").read()
    if refrun.have_ref():
        pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
        rd = os.path.join(base, "ref")
        os.makedirs(rd)
        sb.write_read_files(rd + "/reads", pb, boff, ln, pq, pqoff, bc)
        refrun.run_probe(rd, extra=("MSPEDGES=" + final,), keep_kvec=False)
        assert open(rd + "/a.hbv", "rb").read() == open(wd + "/oracle.hbv", "rb").read()
