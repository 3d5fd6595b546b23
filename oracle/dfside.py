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


def pqvec_decode(b):
    """PQVec blocks -> list of quals (feudal/PQVec.cc:129-187): [nQs u8][nBits:3 | minQ:6 | first data bits ...], 0 ends."""
    out, i = [], 0
    while b[i] != 0:
        nqs = b[i]
        hdr = b[i + 1] | (b[i + 2] << 8)
        nb, minq = hdr & 7, (hdr >> 3) & 63
        total = _block_size(nqs, nb)
        blk = int.from_bytes(b[i:i + total], "little") >> 17
        for _ in range(nqs):
            out.append((blk & ((1 << nb) - 1)) + minq)
            blk >>= nb
        i += total
    return out


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


# ---- the files DF writes next to a.hbv (10X/WriteFiles.cc:16-60, 10X/DF.cc:573-600) ----------------------------
def read_hbv(data):
    """a.hbv (paths/HyperBasevector.cc:121-125) -> dict K, from_v, from_e, to_e (lists per vertex), edges [(nbases, packed bytes)]."""
    assert data[:8] == MAGIC
    K, = struct.unpack_from("<i", data, 8)
    pos = 12

    def vecvec(pos):
        n, = struct.unpack_from("<Q", data, pos)
        pos += 8
        out = []
        for _ in range(n):
            m, = struct.unpack_from("<Q", data, pos)
            pos += 8
            out.append(list(struct.unpack_from("<%di" % m, data, pos)))
            pos += 4 * m
        return out, pos
    from_v, pos = vecvec(pos)
    from_e, pos = vecvec(pos)
    to_e, pos = vecvec(pos)
    ne, = struct.unpack_from("<Q", data, pos)
    pos += 8
    edges = []
    for _ in range(ne):
        nb, = struct.unpack_from("<I", data, pos)
        pos += 4
        edges.append((nb, data[pos:pos + (nb + 3) // 4]))
        pos += (nb + 3) // 4
    assert pos == len(data)
    return {"K": K, "from_v": from_v, "from_e": from_e, "to_e": to_e, "edges": edges}


def hbx_tables(h):
    """what HyperBasevectorX(hbv) adds (graph/DigraphTemplate.h:3281-3307; digraph to_ is rebuilt from from_ on read,
    graph/Digraph.cc:1421-1433): to_ lists and to_left_/to_right_."""
    n = len(h["from_v"])
    ne = len(h["edges"])
    to_left, to_right = [0] * ne, [0] * ne
    for v in range(n):
        for j, e in enumerate(h["from_e"][v]):
            to_left[e] = v
            to_right[e] = h["from_v"][v][j]
    to_v = [[to_left[e] for e in h["to_e"][v]] for v in range(n)]
    return to_v, to_left, to_right


def _serfvecs(lists):                    # MasterVec<SerfVec<int>> (feudal/OuterVec.h:377-379, feudal/SmallVec.h:355-357)
    return struct.pack("<Q", len(lists)) + b"".join(struct.pack("<I%di" % len(x), len(x), *x) for x in lists)


def _vec_int(v):
    return struct.pack("<Q", len(v)) + struct.pack("<%di" % len(v), *v)


def hbx_file(h):
    """a.hbx: K, digraphX {from_, to_}, from_edge_obj_, to_edge_obj_, edges_, to_left_, to_right_
    (paths/HyperBasevector.cc:133-137, graph/Digraph.h:435-437, graph/DigraphTemplate.h:3107-3113)."""
    to_v, to_left, to_right = hbx_tables(h)
    edges = struct.pack("<Q", len(h["edges"])) + b"".join(struct.pack("<I", nb) + bytes(p) for nb, p in h["edges"])
    return (MAGIC + struct.pack("<i", h["K"]) + _serfvecs(h["from_v"]) + _serfvecs(to_v) + _serfvecs(h["from_e"]) + _serfvecs(h["to_e"])
            + edges + _vec_int(to_left) + _vec_int(to_right))


def edges_fastb_file(h):
    """a.fastb: the HBV edges as a feudal vecbvec (WriteFiles.cc:46-47)."""
    n = len(h["edges"])
    var = b"".join(bytes(p) for _, p in h["edges"])
    offs, pos = [], 24
    for _, p in h["edges"]:
        offs.append(pos)
        pos += len(p)
    offs.append(pos)
    vt = 24 + len(var)
    return (struct.pack("<IBBBBQQ", n, 1, 4, 16, 1, vt, vt + 8 * (n + 1)) + var + struct.pack("<%dQ" % (n + 1), *offs)
            + struct.pack("<%dI" % n, *[nb for nb, _ in h["edges"]]))


def kmers_file(h):
    """a.kmers: vec<int> of hb.Kmers(e) = bases - K + 1 (WriteFiles.cc:48-51)."""
    return MAGIC + _vec_int([nb - h["K"] + 1 for nb, _ in h["edges"]])


# ---- ReadPathVecX (10X/paths/ReadPathParser.cc:17-50,217-229; ReadPathVecX.cc:309-312,684-691,976-996) -----------
def pathsx_file(paths, h):
    """a.pathsX of InitializePathsXFromPaths (10X/DfTools.cc:24-78: the reads appended in order).  One record per read:
    [numEdges u8] and, when that byte is not 0, [offset i16][first edge u32][2 bits per following edge = its index in
    From(ToRight(previous))]; ZipIndex holds the byte position of every 10th record."""
    _, _, to_right = hbx_tables(h)
    data = bytearray()
    index = []
    for rid, (off, edges) in enumerate(paths):
        if rid % 10 == 0:                                       # updateZipIndex :684-691 (skip = 10, start_rid = 0)
            index.append(len(data))
        n = len(edges)
        rec = bytearray(((n - 1 + 3) // 4 + 7) if n else 1)     # LLzip :19
        rec[0] = n & 0xFF
        if rec[0]:
            struct.pack_into("<hI", rec, 1, ((off + 0x8000) & 0xFFFF) - 0x8000, edges[0] & 0xFFFFFFFF)
            idx, sub = 7, 0
            for a, b in zip(edges, edges[1:]):
                fr = h["from_e"][to_right[a]]
                if b in fr:
                    rec[idx] = (rec[idx] + ((fr.index(b) & 0xFF) << sub)) & 0xFF       # LLencodeBranchId :217-229
                    sub += 2
                    if sub > 7:
                        idx, sub = idx + 1, 0
        data += rec
    return (struct.pack("<qqqqq", 10, 0, len(paths), len(index), len(data)) + struct.pack("<%dq" % len(index), *index) + bytes(data))


# ---- MarkDups, version 2 (10X/SecretOps.cc:599-774) ------------------------------------------------------------------
def mark_dups(paths, bases, quals, bc):
    """-> (dup per PAIR as 0/1 list, ndups, interdups, art per pair).  paths: list of (offset, edges); bases: list of base
    code lists; quals: list of qual lists; bc: per-read barcode ordinal."""
    n = len(paths)
    X = []
    for id1 in range(n):                                        # :617-627
        id2 = id1 + 1 if id1 % 2 == 0 else id1 - 1
        off, edges = paths[id1]
        # the path comes back out of the ReadPathVecX (paths.unzip, :620): the edge count went through one byte and
        # the offset through an int16 (ReadPathParser.cc:27,31,119,126)
        off = ((off + 0x8000) & 0xFFFF) - 0x8000
        if len(edges) & 0xFF == 0:
            X.append((-1, -1, -1, -1))
        else:
            head = 0
            for j in range(5):
                head = head * 4 + bases[id2][j]
            X.append((edges[0], off, head, id1))
    X.sort()                                                     # :638
    dup = [0] * (n // 2)
    art = [0] * (n // 2)
    ndups = interdups = 0
    j = 0
    while j < n:
        k = j + 1
        while k < n and X[k][:3] == X[j][:3]:
            k += 1
        if X[j][0] >= 0:
            if k - j > 1:                                        # :649-659
                ndups += k - j - 1
                inter = False
                b = bc[X[j][3]]
                for l in range(j + 1, k):
                    if b == 0:
                        b = bc[X[l][3]]
                    elif bc[X[l][3]] != b:
                        inter = True
                if inter:
                    interdups += k - j - 1
            qsum = {}
            if k - j > 1:                                        # :691-704 (members of a group only; else 0)
                for l in range(j, k):
                    i1 = X[l][3]
                    i2 = i1 + 1 if i1 % 2 == 0 else i1 - 1
                    qsum[i1] = sum(quals[i1]) + sum(quals[i2])
            best, q, tie = j, qsum.get(X[j][3], 0), False        # :725-737
            for l in range(j + 1, k):
                ql = qsum.get(X[l][3], 0)
                if ql == q:
                    tie = True
                    if X[l][3] < X[best][3]:
                        best = l
                elif ql > q:
                    q, best = ql, l
            if tie:                                              # :738-752
                qb = sorted((tuple(bases[X[l][3]]), tuple(quals[X[l][3]]), X[l][3] // 2) for l in range(j, k))
                m = 0
                while m < len(qb):
                    nn = m + 1
                    while nn < len(qb) and qb[nn][:2] == qb[m][:2]:
                        nn += 1
                    for x in range(m + 1, nn):
                        art[qb[x][2]] = 1
                    m = nn
            for l in range(j, k):                                # :753-754
                if l != best:
                    dup[X[l][3] // 2] = 1
        j = k
    return dup, ndups, interdups, art


def dup_file(dup):
    """a.dup: vec<Bool> (Bool = unsigned char, system/Types.h:121) as BINWRITE."""
    return MAGIC + struct.pack("<Q", len(dup)) + bytes(dup)


def read_fastb(data):
    """feudal vecbvec -> list of base-code lists (feudal/FieldVec.h:586-607)."""
    n, _, _, _, _, vt, fo = struct.unpack_from("<IBBBBQQ", data, 0)
    offs = np.frombuffer(data, dtype="<u8", count=n + 1, offset=vt)
    lens = np.frombuffer(data, dtype="<u4", count=n, offset=fo)
    out = []
    for i in range(n):
        p = np.frombuffer(data, dtype=np.uint8, count=(int(lens[i]) + 3) // 4, offset=int(offs[i]))
        out.append(((p[:, None] >> np.array([0, 2, 4, 6])) & 3).reshape(-1)[:int(lens[i])].tolist())
    return out


def read_qualp(data):
    """feudal VecPQVec -> list of qual lists."""
    n, _, _, _, _, vt, _ = struct.unpack_from("<IBBBBQQ", data, 0)
    offs = np.frombuffer(data, dtype="<u8", count=n + 1, offset=vt)
    return [pqvec_decode(data[int(offs[i]):int(offs[i + 1])]) for i in range(n)]


def expand_bci(data):
    """reads.bci -> per-read barcode ordinal (10X/DF.cc:464-469)."""
    assert data[:8] == MAGIC
    bci = np.frombuffer(data, dtype="<i8", offset=16)
    bc = np.full(int(bci[-1]), -1, np.int64)
    for b in range(len(bci) - 1):
        bc[int(bci[b]):int(bci[b + 1])] = b
    return bc.tolist()


def dup_percentages(dup, ndups, interdups, art):
    """the three figures MarkDups prints (SecretOps.cc:757-770), formatted as it prints them (fixed, 2 decimals)."""
    npairs = len(dup)
    f = lambda x: "%.2f" % x
    return {"dup_perc": f(100.0 * sum(dup) / npairs), "interdup_perc": f(100.0 * interdups / ndups) if ndups else "-nan",
            "art_dup_perc": f(100.0 * sum(art) / npairs)}
