"""Development helper: times the count stage of the C2 workload under kernel variants
(SN_BC_VARIANT / SN_MSP_OCC).  usage: python tools/sweep_count.py "variant:occ" ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import supernova_b200 as sb

codes, quals, off, bc, meta = bench.gen_workload(os.environ.get("SWEEP_WORKLOAD", "C2"), 0)
pb, boff, ln, pq, pqoff = sb.pack_reads(codes, quals, off)
del codes, quals
ctx = sb.Context(0)
ctx.load_reads(pb, boff, ln, pq, pqoff, bc)
for spec in sys.argv[1:]:
    v, occ = spec.split(":")
    os.environ["SN_BC_VARIANT"] = v
    os.environ["SN_MSP_OCC"] = occ
    best = None
    for rep in range(3):
        ctx.count_kmers(sb.Params())
        st = ctx.stage_ms()
        t = {k: round(st[k], 3) for k in ("msp_hist", "msp_scatter", "bucket_count", "make_dict")}
        if best is None or t["bucket_count"] < best["bucket_count"]:
            best = t
    c = ctx.counts()
    print(spec, best, "n_kmers", c["n_kmers"], "distinct", c["n_kmers_distinct"], "sk", c["n_superkmers"], flush=True)
