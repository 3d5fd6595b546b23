#!/usr/bin/env python
"""Profiling aid: N ranks of the sharded path as N contexts on ONE GPU (in-process communicator), on the bench's
genome-scaled weak-scaling workload.  Not a measurement of multi-GPU speed (the ranks share the device): it exposes the
kernels of the replicated, whole-graph stages at N x scale to ncu on a single-GPU box.
    python tools/local_scale.py <n_ranks> [workload=C2] [steps=2]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import bench
    import supernova_b200 as sb
    n = int(sys.argv[1]); wl = sys.argv[2] if len(sys.argv) > 2 else "C2"; steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    data = []
    for r in range(n):
        codes, quals, off, bc, meta = bench.gen_workload(wl, r, genome_mult=n)
        bc = np.where(bc > 0, bc + r * meta["n_bc"], 0).astype(np.int32)
        data.append((sb.pack_reads(codes, quals, off), bc))
        del codes, quals
    print("generated", n, "x", wl, flush=True)

    def fn(rank, ctx):
        packed, bc = data[rank]
        ctx.load_reads(*packed, bc)
        out = []
        for s in range(steps):
            t0 = time.time()
            ctx.mg_build_graph(sb.Params(), with_paths=False)
            out.append((time.time() - t0, {k: round(v, 2) for k, v in ctx.stage_ms().items() if v >= 0}, ctx.counts()))
        return out
    res = sb.run_local_ranks(n, fn)
    for r, o in enumerate(res):
        print("rank", r, "wall %.1f ms" % (1e3 * o[-1][0]), o[-1][1], "kmers", o[-1][2]["n_kmers"], "edges", o[-1][2]["n_edges"], flush=True)


if __name__ == "__main__":
    main()
