#!/usr/bin/env python
"""A C3-shaped share on ONE B200: `mult` x the C2 workload (1.2 Gbp each: 63 Mbp x mult genome at 19x) generated ON the device
(sn_generate_reads) and pushed through count (bucket passes when the occurrences exceed 2^32) -> unipaths -> HyperBasevector.
    python tools/big_run.py <mult> [with_paths=0] [out.json]
mult = 18.75 is BASELINE config 3's per-GPU share (22.5 Gbp).  Prints one JSON line: sizes, stage times, device memory."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def mem_used_mb():
    try:
        return int(subprocess.check_output(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits", "-i", "0"]).decode().split()[0])
    except Exception:
        return None


def main():
    import supernova_b200 as sb
    mult = float(sys.argv[1]); with_paths = len(sys.argv) > 2 and sys.argv[2] == "1"
    G = int(63_000_000 * mult); pairs = int(4_000_000 * mult); nbc = min(int(1_000_000 * mult), 4_000_000); seed = 20261017
    out = {"mult": mult, "genome_bases": G, "pairs": pairs, "gbp": pairs * 300 / 1e9, "n_barcodes": nbc, "seed": seed}
    with sb.Context(0) as ctx:
        t0 = time.time()
        ctx.generate_reads(G, pairs, nbc, seed)
        out["generate_s"] = round(time.time() - t0, 3); out["mem_after_generate_mb"] = mem_used_mb()
        t0 = time.time()
        ctx.count_kmers(sb.Params())
        out["count_s"] = round(time.time() - t0, 3); out["mem_after_count_mb"] = mem_used_mb()
        t0 = time.time()
        ctx.build_edges()
        out["edges_s"] = round(time.time() - t0, 3); out["mem_after_edges_mb"] = mem_used_mb()
        t0 = time.time()
        ctx.build_hbv()
        out["hbv_s"] = round(time.time() - t0, 3)
        if with_paths:
            t0 = time.time()
            ctx.path_reads()
            out["paths_s"] = round(time.time() - t0, 3)
        out["mem_end_mb"] = mem_used_mb()
        out["counts"] = ctx.counts()
        out["stage_ms"] = {k: round(v, 2) for k, v in ctx.stage_ms().items() if v >= 0}
        total = out["count_s"] + out["edges_s"] + out["hbv_s"]
        out["gbp_per_s_count_to_hbv_wall"] = round(out["gbp"] / total, 2)
        c = out["counts"]
        out["passes"] = int(c["n_kmer_occurrences"] // 3_600_000_000 + 1)
        # what must hold at any size: every valid k-mer lies on exactly one unipath; the HBV has two oriented edges per unipath
        # but for the palindromic ones
        out["kmers_on_edges"] = int(c["n_edge_bases"] - 47 * c["n_edges"])
        out["ok"] = bool(out["kmers_on_edges"] == c["n_kmers"] and c["n_edges"] <= c["n_hbv_edges"] <= 2 * c["n_edges"])
    line = json.dumps(out)
    print(line)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(line + "\n")


if __name__ == "__main__":
    main()
