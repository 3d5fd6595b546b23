#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Golden digests of the reference's own buildReadQGraph48 (oracle/_ref/OracleProbe,
compiled from /root/reference by oracle/build_ref.sh) at the BASELINE sizes the oracle cannot reach inside a
test run: C2 (1.2 Gbp, BASELINE.json configs[1]) and C2b (the same genome at 56x).

    python tests/golden/make_scale_digests.py C2 [C2b ...]      # merges into tests/golden/scale_digests.json

Per data set: the order-independent digest of kmers.kvec ({k-mer, count, ctx} per entry, digests.kmer_digest),
md5 of a.hbv, of tmp.paths and of stats/histogram_kmer_count.json, plus the sizes.  The inputs come from the
seeded generator (supernova_b200/synth.py), so the GPU box regenerates byte-identical reads and
tests/test_gpu_scale.py compares the CUDA path's outputs against these digests."""
import hashlib
import json
import os
import struct
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import digests  # noqa: E402
import refrun  # noqa: E402

OUT = os.path.join(HERE, "scale_digests.json")


def main():
    import supernova_b200 as sb
    from supernova_b200 import synth
    import bench
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in sys.argv[1:]:
        G, pairs, nbc, seed = bench.WORKLOADS[name]
        t0 = time.time()
        b, q, bc, ids = synth.make_reads(G, pairs, nbc, seed, workers=min(os.cpu_count() or 1, 32))
        n, L = b.shape
        off = np.arange(n + 1, dtype=np.uint64) * L
        pb, boff, ln, pq, pqoff = sb.pack_reads(b.ravel(), q.ravel(), off)
        del b, q
        wd = tempfile.mkdtemp(prefix="sn_gold_", dir=os.environ.get("SN_GOLD_TMP", "/tmp"))
        sb.write_read_files(wd + "/reads", pb, boff, ln, pq, pqoff, bc)
        inputs_md5 = {f: digests.file_md5(wd + "/reads." + f) for f in ("fastb", "qualp", "bci")}
        del pb, pq
        print(name, "inputs ready", round(time.time() - t0, 1), "s", flush=True)
        secs, out = refrun.run_probe(wd, paths=True, keep_kvec=True)
        print(name, "reference ran", secs, "s", flush=True)
        if "buffer overflows" in out:
            raise SystemExit("the reference reported MapReduce buffer overflows: its barcode rule is lossy in that mode (SURVEY 8(a) a4)")
        d = open(wd + "/kmers.kvec", "rb").read()
        nk, = struct.unpack_from("<Q", d, 8)
        e = np.frombuffer(d, dtype="<u4", offset=16).reshape(nk, 6)
        kd = digests.kmer_digest(e[:, 0], e[:, 1], e[:, 2], e[:, 5])
        res[name] = {"G": G, "pairs": pairs, "n_bc": nbc, "seed": seed, "n_kmers": int(nk), "kmers": kd,
                     "a.hbv": digests.file_md5(wd + "/a.hbv"), "a.hbv_bytes": os.path.getsize(wd + "/a.hbv"),
                     "tmp.paths": digests.file_md5(wd + "/tmp.paths"), "tmp.paths_bytes": os.path.getsize(wd + "/tmp.paths"),
                     "histogram_kmer_count.json": digests.file_md5(wd + "/stats/histogram_kmer_count.json"),
                     "inputs": inputs_md5, "reference_seconds": secs, "cores": os.cpu_count()}
        print(name, json.dumps(res[name]), flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
        for root, _, files in os.walk(wd, topdown=False):
            for f in files:
                os.remove(os.path.join(root, f))
            os.rmdir(root)


if __name__ == "__main__":
    main()
