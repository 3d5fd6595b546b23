"""oracle/dfside.py -- TEST INFRASTRUCTURE (never imported by the product).
numpy / pure-Python restatement of the two steps either side of the hot path that the product also
implements on the device, each function citing the reference lines it
follows.  Pinned by tests/test_oracle_golden.py against the golden files the reference's own binaries
wrote (tests/golden/make_golden.py)."""
import struct

import numpy as np

MAGIC = b"BINWRITE"


# ---- feudal / BINWRITE readers (feudal/FeudalControlBlock.h:159-165, feudal/BinaryStream.h:486-493) ----
def read_paths(data):
    """feudal ReadPathVec (paths/long/ReadPath.h:61-63) -> list of (offset, [edges])."""
    n, _, _, _, _, vt, fo = struct.unpack_from("<IBBBBQQ", data, 0)
    offs = np.frombuffer(data, dtype="<u8", count=n + 1, offset=vt)
    out = []
    for i in range(n):
        a, b = int(offs[i]), int(offs[i + 1])
        if b - a < 8:
            out.append((0, []))
            continue
        off = struct.unpack_from("<i", data, a)[0]
        edges = np.frombuffer(data, dtype="<i4", count=(b - a - 8) // 4, offset=a + 8)
        out.append((off, edges.tolist()))
    return out


def read_vec_int(data):
    assert data[:8] == MAGIC
    n, = struct.unpack_from("<Q", data, 8)
    return np.frombuffer(data, dtype="<i4", count=n, offset=16)


# ---- writePathsIndex (10X/PathsIndex.cc:23-143) ----------------------------------------------------------
def paths_index(paths, inv):
    """-> (a.paths.inv bytes, a.countsb bytes).  paths: list of (offset, edges); inv: involution."""
    num_edges = len(inv)
    pairs = sorted((e, rid) for rid, (_, edges) in enumerate(paths) for e in edges)      # :51-57, ParallelSort :86
    lists = [[] for _ in range(num_edges)]
    for e, rid in pairs:
        lists[e].append(rid)                                                              # :99-107
    counts = [len(x) for x in lists]                                                      # :110
    countsb = list(counts)
    for e in range(num_edges):                                                            # :123-131
        re = int(inv[e])
        if e < re:
            countsb[e] = countsb[re] = counts[e] + counts[re]
    # IncrementalWriter<ULongVec>: FCB {n, flags 1, sizeofFixed 0, sizeofX 16, sizeofA 8, varTab, fixedOff},
    # var data = the ids as u64, then n+1 absolute offsets
    var = b"".join(struct.pack("<%dQ" % len(x), *x) for x in lists)
    offs, pos = [], 24
    for x in lists:
        offs.append(pos)
        pos += 8 * len(x)
    offs.append(pos)
    vt = 24 + len(var)
    fcb = struct.pack("<IBBBBQQ", num_edges, 1, 0, 16, 8, vt, vt + 8 * (num_edges + 1))
    inv_file = fcb + var + struct.pack("<%dQ" % len(offs), *offs)
    cb_file = MAGIC + struct.pack("<QQ", 1, num_edges) + struct.pack("<%di" % num_edges, *countsb)   # vec<vec<int>>, :133
    return inv_file, cb_file


# ---- ParseBarcodedFastqs (10X/ParseBarcodedFastqs.cc:56-146, 284-303) ---------------------------------------
def _ceil_lg2(x):                        # math/PowerOf2.h ceilLg2
    return 0 if x <= 1 else (x - 1).bit_length()


def _block_size(nqs, nbits):             # feudal/PQVec.h:57-58
    return (nqs * nbits + 17 + 7) >> 3


def pqvec_encode(q):
    """PQVecEncoder (feudal/PQVec.cc:17-127) on a list of quals -> bytes."""
    blocks, costs = [], [1]
    for i, qi in enumerate(q):                                   # init :17-85
        minv, maxv = min(63, qi), qi
        bits, nqs = _ceil_lg2(maxv + 1 - minv), 1
        best_cost, best = costs[i] + _block_size(nqs, bits), [1, bits, minv]
        j = i
        while j != 0 and nqs < 255:
            j -= 1
            v = q[j]
            maxv, minv = max(maxv, v), min(minv, v)
            bits = _ceil_lg2(maxv + 1 - minv)
            nqs += 1
            cur = costs[j] + _block_size(nqs, bits)
            if cur < best_cost:
                best_cost, best = cur, [nqs, bits, minv]
        costs.append(best_cost)
        to_remove = best[0] - 1
        if not to_remove:
            blocks.append(best)
        else:
            while to_remove > blocks[-1][0]:
                to_remove -= blocks[-1][0]
                blocks.pop()
            if to_remove == blocks[-1][0]:
                blocks[-1] = best
            else:
                blocks[-1][0] -= to_remove
                blocks.append(best)
    out, it = bytearray(), 0
    for nqs, nbits, minq in blocks:                              # encode :87-127
        out.append(nqs)
        bits = nbits | (minq << 3)
        out.append(bits & 0xFF)
        bits >>= 8
        if not nbits:
            out.append(bits & 0xFF)
            it += nqs
        else:
            off = 1
            for _ in range(nqs):
                bits |= (q[it] - minq) << off
                it += 1
                off += nbits
                if off >= 8:
                    out.append(bits & 0xFF)
                    off -= 8
                    bits >>= 8
            if off:
                out.append(bits & 0xFF)
    out.append(0)
    return bytes(out)


def parse_fasth(text):
    """newUnpackBarcodeSortedFastq + main (:56-146, :284-303) -> (reads.fastb, reads.qualp, reads.bci) bytes.
    `text`: the content of one file, or a list of contents (FASTQS={a,b,...}, :258-264: the barcode ordinal runs on
    across the files, :66, the comparison string starts empty in each, :67)."""
    texts = [text] if isinstance(text, (bytes, bytearray)) else list(text)
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
    b0, bc_reads, bcs = [], [], []
    barc = 0
    for one in texts:
      lines = one.split(b"\n")
      assert lines[-1] == b"" and (len(lines) - 1) % 9 == 0, "out of sync"
      lastb = None
      for r in range((len(lines) - 1) // 9):
          L = [x.replace(b"n", b"A").replace(b"N", b"A") for x in lines[9 * r:9 * r + 9]]       # :84-85
          assert L[0].startswith(b"@")
          pair = []
          for m in (0, 1):
              bases = [code[c & 0xDF] for c in L[1 + 2 * m]]
              quals = [c - 33 for c in L[2 + 2 * m]]                                       # convertPhred :37-44
              packed = bytearray((len(bases) + 3) // 4)
              for i, v in enumerate(bases):
                  packed[i >> 2] |= v << (2 * (i & 3))                                     # feudal/FieldVec.h:596-598
              pair.append((len(bases), bytes(packed), pqvec_encode(quals)))
          buf = L[5]
          if b"-" in buf and not buf.startswith(b"-"):                                     # :107
              key = buf.split(b",")[0]                                                     # SafeBefore(",") :108
              if key != lastb:
                  barc += 1
                  lastb = key
              bcs += [barc, barc]
              bc_reads += pair
          else:
              b0 += pair
    reads = b0 + bc_reads
    n = len(reads)
    bci = [0]
    cur = 0
    for i, b in enumerate(bcs):                                                          # :284-293
        if b != cur:
            cur = b
            bci.append(i + len(b0))
    bci.append(len(b0) + len(bcs))

    def feudal(var_list, size_fixed, size_x, size_a, fixed=b""):
        var = b"".join(var_list)
        offs, pos = [], 24
        for v in var_list:
            offs.append(pos)
            pos += len(v)
        offs.append(pos)
        vt = 24 + len(var)
        return struct.pack("<IBBBBQQ", n, 1, size_fixed, size_x, size_a, vt, vt + 8 * (n + 1)) + var + struct.pack("<%dQ" % (n + 1), *offs) + fixed
    fastb = feudal([x[1] for x in reads], 4, 16, 1, struct.pack("<%dI" % n, *[x[0] for x in reads]))
    qualp = feudal([x[2] for x in reads], 0, 8, 1)
    bci_file = MAGIC + struct.pack("<Q", len(bci)) + struct.pack("<%dq" % len(bci), *bci)
    return fastb, qualp, bci_file
