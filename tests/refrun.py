"""Helpers that run the reference's own binaries (oracle/_ref, built by
oracle/build_ref.sh) on a data set.  Test infrastructure."""
import os
import re
import struct
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def have_ref():
    return os.access(os.path.join(REF, "OracleProbe"), os.X_OK) and os.access(os.path.join(REF, "ParseBarcodedFastqs"), os.X_OK)


def parse_fastqs(workdir, fastq):
    subprocess.check_call([os.path.join(REF, "ParseBarcodedFastqs"), "FASTQS=" + fastq,
                           "OUT_HEAD=" + os.path.join(workdir, "reads")], cwd=workdir,
                          stdout=subprocess.DEVNULL)


def parse_fastq_set(workdir, fastqs):
    """FASTQS={a,b,...}: several barcode-sorted files in one call (ParseBarcodedFastqs.cc:258-264)."""
    subprocess.check_call([os.path.join(REF, "ParseBarcodedFastqs"), "FASTQS={" + ",".join(fastqs) + "}",
                           "OUT_HEAD=" + os.path.join(workdir, "reads")], cwd=workdir, stdout=subprocess.DEVNULL)


def host_threads():
    """threads the reference will use: processorsOnline (system/SysConf.cc:132-137)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def have_shim():
    """OracleProbe_b200: the reference closure with BuildReadQGraph48.o replaced by integration/BuildReadQGraph48_b200.cc"""
    return os.access(os.path.join(REF, "OracleProbe_b200"), os.X_OK)


def run_probe(workdir, paths=True, keep_kvec=True, extra=(), binary="OracleProbe", head=None):
    env = dict(os.environ)
    # the reference sizes its worker pool from processorsOnline, but its OpenMP regions (parallel sorts,
    # parallelFor) obey OMP_NUM_THREADS -- torchrun exports OMP_NUM_THREADS=1 to its children, which
    # would throttle those regions only.  Give the reference every host thread, explicitly.
    env["OMP_NUM_THREADS"] = str(host_threads())
    env.pop("OMP_THREAD_LIMIT", None)
    if keep_kvec:
        env["SN_KEEP_KVEC"] = "1"
    out = subprocess.check_output([os.path.join(REF, binary), "HEAD=" + (head or os.path.join(workdir, "reads")),
                                   "OUT=" + workdir, "PATHS=" + ("True" if paths else "False"), *extra],
                                  cwd=workdir, env=env, stderr=subprocess.STDOUT).decode()
    secs = None
    for line in out.splitlines():
        if line.startswith("ORACLE_SECONDS"):
            secs = float(line.split()[1])
    return secs, out


def read_kvec(path):
    """kmers.kvec -> (n,5) u32 sorted by k-mer: w0,w1,w2,count,ctx (bc field ignored: SURVEY §8(c))."""
    d = open(path, "rb").read()
    assert d[:8] == b"BINWRITE"
    n, = struct.unpack_from("<Q", d, 8)
    e = np.frombuffer(d, dtype="<u4", offset=16).reshape(n, 6)
    ref = np.stack([e[:, 0], e[:, 1], e[:, 2], e[:, 5] & 0xFFFFFF, e[:, 5] >> 24], axis=1)
    idx = np.lexsort((ref[:, 2], ref[:, 1], ref[:, 0]))
    return ref[idx]


_MONTHS = {m: i + 1 for i, m in enumerate("Jan Feb Mar Apr May Jun Jul Aug Sep Oct Nov Dec".split())}


def phase_split(log, total_seconds=None):
    """Seconds per phase of one buildReadQGraph48 run from the reference's own Date() stamps (1 s
    resolution; BuildReadQGraph48.cc:227-321,1688-1774): qual scan (GoodLenTailFinder), MapReduce #1
    (count only, :267-277), MapReduce #2 (fill, :280-286), spectrum + kmers.kvec + dictionary,
    recomputeAdjacencies, buildEdges, buildHBVFromEdges, pathReads.  `count_once` = total minus
    MapReduce #1: what the run would cost if the k-mer MapReduce ran once instead of twice."""
    import calendar
    stamps = []
    for line in log.splitlines():
        m = re.match(r"^\.?\w{3} (\w{3}) +(\d+) (\d+):(\d+):(\d+) (\d{4})[: ]+(.*)$", line)
        if m and m.group(1) in _MONTHS:
            t = calendar.timegm((int(m.group(6)), _MONTHS[m.group(1)], int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))))
            stamps.append((t, m.group(7).strip()))

    def first(text, after=0, nth=0):
        k = 0
        for i, (t, msg) in enumerate(stamps):
            if i >= after and text in msg:
                if k == nth:
                    return i
                k += 1
        return None
    marks = [("qual_scan", first("loading reads")), ("mapreduce_1", first("MapReduce needs")),
             ("mapreduce_2", first("MapReduce needs", nth=1)), ("spectrum_kvec_dict", first("computing spectrum")),
             ("recompute_adjacencies", first("recomputing adjacencies")), ("build_edges", first("finding edge sequences")),
             ("hbv_from_edges", first("building from edges")), ("path_reads", first("pathing reads"))]
    marks = [(n, i) for n, i in marks if i is not None]
    out = {}
    for k, (name, i) in enumerate(marks):
        if k + 1 < len(marks):
            out[name] = float(stamps[marks[k + 1][1]][0] - stamps[i][0])
        elif total_seconds is not None:                # the last phase runs to the end of the call
            out[name] = max(0.0, float(total_seconds) - float(stamps[i][0] - stamps[marks[0][1]][0]))
        else:
            out[name] = float(stamps[-1][0] - stamps[i][0])
    if total_seconds is not None and "mapreduce_1" in out:
        out["count_once_seconds"] = max(0.0, float(total_seconds) - out["mapreduce_1"])
    return out
