"""Helpers that run the reference's own binaries (oracle/_ref, built by
oracle/build_ref.sh) on a data set.  Test infrastructure."""
import os
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


def run_probe(workdir, paths=True, keep_kvec=True, extra=()):
    env = dict(os.environ)
    if keep_kvec:
        env["SN_KEEP_KVEC"] = "1"
    out = subprocess.check_output([os.path.join(REF, "OracleProbe"), "HEAD=" + os.path.join(workdir, "reads"),
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
