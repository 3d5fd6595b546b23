#!/usr/bin/env python
"""Writes the reference's input files (reads.fastb / reads.qualp / reads.bci) of a named synthetic
workload into a directory.  Bench/test infrastructure: bench.py --impl reference runs this in a
subprocess so that the process timing the reference never maps the product library.

    python tools/prep_workload.py <workload> <scale_div> <out_dir> [shard]
prints one JSON line: {"gbp": ..., "G": ..., "pairs": ..., "n_reads": ...}"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    name, div, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    shard = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    import bench
    import supernova_b200 as sb
    from supernova_b200 import synth
    G, pairs, nbc, seed = bench.WORKLOADS[name]
    G, pairs, nbc = G // div, pairs // div, max(2, nbc // div)
    b, q, bc, _ = synth.make_reads(G, pairs, nbc, seed, workers=min(os.cpu_count() or 1, 32), shard=shard)
    n, L = b.shape
    off = np.arange(n + 1, dtype=np.uint64) * L
    pb, boff, ln, pq, pqoff = sb.pack_reads(b.ravel(), q.ravel(), off)
    os.makedirs(out, exist_ok=True)
    sb.write_read_files(out + "/reads", pb, boff, ln, pq, pqoff, bc)
    print(json.dumps({"gbp": n * L / 1e9, "G": G, "pairs": pairs, "n_bc": nbc, "seed": seed, "n_reads": int(n), "read_len": int(L)}))


if __name__ == "__main__":
    main()
