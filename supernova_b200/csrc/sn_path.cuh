// sn_path.cuh -- ReadPath threading of one read onto the graph; host+device so the
// same code runs in the CUDA kernel (one read per thread) and in tests/hostsim.
// References: paths/long/BuildReadQGraph48.cc:605-747 (PathPart, Pather::path),
// :810-816 (isJoinable), :1217-1336 (HBVPather::algorithmTwo), :1385-1420
// (pathPartsToReadPath); paths/long/ExtendReadPath.cc:15-358.
#pragma once
#include "sn_kmer.cuh"

namespace sn {

#define SN_MAX_READ_LEN 256
#define SN_MAX_PARTS (SN_MAX_READ_LEN - SN_K + 4)
#define SN_MAX_PATH 384

// Unipath edges as produced by the edge stage (fastb packing, byte-aligned per edge).
struct EdgeStore {
    const uint8_t* bases;
    const uint64_t* off;     // byte offset of each edge
    const uint32_t* len;     // bases
};
// HyperBasevector adjacency in CSR form (graph/Digraph.h from_/to_ + edge objects).
struct HbvView {
    const int32_t* fwd_xlat; const int32_t* rev_xlat;   // unipath id -> HBV edge id
    const int32_t* to_left;  const int32_t* to_right;   // HBV edge -> vertex
    const uint32_t* src;                                 // HBV edge -> unipath id<<1 | rc
    const uint32_t* from_start; const int32_t* from_v; const int32_t* from_e;
    const uint32_t* to_start;   const int32_t* to_v;   const int32_t* to_e;
};

struct Part { uint32_t edge, off, len, elen_rc; };      // elen_rc = edgeKmers<<1 | rc ; 0 => gap
SN_HD bool part_gap(const Part& p) { return (p.elen_rc >> 1) == 0; }
SN_HD uint32_t part_rc(const Part& p) { return p.elen_rc & 1u; }
SN_HD uint32_t part_elen(const Part& p) { return p.elen_rc >> 1; }
SN_HD Part mk_gap(uint32_t len) { Part p; p.edge = SN_NULL_EDGE; p.off = 0; p.len = len; p.elen_rc = 0; return p; }
SN_HD bool part_same_edge(const Part& a, const Part& b) { return a.edge == b.edge && part_rc(a) == part_rc(b); }

// base p of unipath u read in orientation rc
SN_HD uint32_t edge_base(const EdgeStore& es, uint32_t u, uint32_t rc, uint32_t p)
{
    const uint8_t* b = es.bases + es.off[u];
    return rc ? (packed_base(b, es.len[u] - 1 - p) ^ 3u) : packed_base(b, p);
}
SN_HD uint32_t hbv_len(const EdgeStore& es, const HbvView& h, int32_t e) { return es.len[h.src[e] >> 1]; }
SN_HD uint32_t hbv_base(const EdgeStore& es, const HbvView& h, int32_t e, uint32_t p)
{ uint32_t s = h.src[e]; return edge_base(es, s >> 1, s & 1u, p); }

// PQVec block stream -> one Phred byte per base (feudal/PQVec.cc:129-187; block format
// [nQs u8][nBits:3 | minQ:6][nQs x nBits packed, LSB first] ... 0).  Returns the number of
// quals the stream holds (writes at most `cap`).
SN_HD uint32_t pqvec_decode(const uint8_t* p, const uint8_t* pend, uint8_t* out, uint32_t cap)
{
    uint32_t i = 0;
    while (p < pend) {
        uint32_t nq = *p++;
        if (!nq) break;
        uint32_t b0 = *p++;
        uint32_t nbits = b0 & 7u, minq = b0 >> 3;
        uint64_t acc = *p++;
        minq |= (uint32_t)(acc & 1u) << 5; acc >>= 1;
        uint32_t have = 7, mask = (1u << nbits) - 1u;
        for (uint32_t k = 0; k < nq; ++k) {
            uint32_t q = minq;
            if (nbits) {
                if (have < nbits) { acc |= (uint64_t)(*p++) << have; have += 8; }
                q += (uint32_t)acc & mask; acc >>= nbits; have -= nbits;
            }
            if (i < cap) out[i] = (uint8_t)q;
            ++i;
        }
    }
    return i;
}

// the seed at read position itr hit dictionary entry `ent`: extend the exact match along its edge (matchLen), -> the part
SN_HD Part seed_part(const DictView& d, const EdgeStore& es, const uint8_t* rd, uint32_t n, uint32_t itr, const Kmer& kmer, uint32_t ent)
{
    const DictEntry& e = d.tab[ent];
    uint32_t u = e.edge, esz = es.len[u];
    int32_t offset = (int32_t)e.off;
    // CF<K>::isRC (dna/CanonicalForm.h:84-91): is the read k-mer the RC of the edge
    // k-mer at `offset`?  The edge k-mer is the stored canonical k-mer or its RC.
    // Palindromes compare equal to themselves => not RC.
    const uint8_t* eb = es.bases + es.off[u];
    Kmer ek = kmer_from_packed_w(eb, (uint32_t)offset);
    bool rc = !(ek == kmer);
    uint32_t len = 1;
    // matchLen: 16 bases per step (both sequences packed 2 bits per base)
    if (!rc) {
        uint32_t a = itr + SN_K, b = (uint32_t)offset + SN_K;
        while (a < n && b < esz) {
            uint32_t m = n - a < esz - b ? n - a : esz - b; if (m > 16) m = 16;
            const uint32_t eq = window_match(packed_window16(rd, a), packed_window16(eb, b), m);
            const uint32_t adv = eq < m ? eq : m;
            len += adv; a += adv; b += adv;
            if (eq < m) break;
        }
    } else {
        offset = (int32_t)esz - offset;
        uint32_t a = itr + SN_K, b = (uint32_t)offset;
        // the edge read backwards and complemented: base b of the RC is 3 - edge[esz - 1 - b]
        while (a < n && b < esz) {
            uint32_t m = n - a < esz - b ? n - a : esz - b; if (m > 16) m = 16;
            const uint32_t w = packed_window16(eb, esz - b - m);                 // edge[esz-b-m .. esz-b-1] in the low 2m bits
            const uint32_t r = ~(rev2(w) >> (2 * (16 - m)));                     // reversed and complemented: RC bases b .. b+m-1
            const uint32_t eq = window_match(packed_window16(rd, a), r, m);
            const uint32_t adv = eq < m ? eq : m;
            len += adv; a += adv; b += adv;
            if (eq < m) break;
        }
        offset -= SN_K;
    }
    Part p; p.edge = u; p.off = (uint32_t)offset; p.len = len; p.elen_rc = ((esz - SN_K + 1) << 1) | (rc ? 1u : 0u);
    return p;
}

// a10 Pather::path (:705-747).  Returns the number of parts.
SN_HD uint32_t path_parts(const DictView& d, const EdgeStore& es, const uint8_t* rd, uint32_t n, Part* parts)
{
    uint32_t np = 0;
    if (n < SN_K) { parts[np++] = mk_gap(n); return np; }
    uint32_t itr = 0, end = n - SN_K + 1;
    while (itr != end) {
        Kmer kmer = kmer_from_packed_w(rd, itr);
        bool was_rc;
        MinState ms = min_state_init(kmer);             // the k-mer's minimizer names its dictionary bucket
        uint32_t ent = dict_find_min(d, kmer, ms.minval, &was_rc);
        if (ent == SN_NULL_EDGE) {
            uint32_t gap = 1, itr2 = itr + SN_K; ++itr;
            while (itr2 != n) {
                const uint32_t nb = packed_base(rd, itr2);
                kmer = kmer_succ(kmer, nb); ++itr2;
                min_state_slide(ms, kmer, nb);
                ent = dict_find_min(d, kmer, ms.minval, &was_rc);
                if (ent != SN_NULL_EDGE) break;
                ++gap; ++itr;
            }
            parts[np++] = mk_gap(gap);
        }
        if (ent != SN_NULL_EDGE) {
            const Part p = seed_part(d, es, rd, n, itr, kmer, ent);
            parts[np++] = p;
            itr += p.len;
        }
    }
    return np;
}

#if defined(__CUDACC__)
// Pather::path for the 32 reads of a warp together.  A read whose seed k-mer is not in the dictionary (a sequencing error
// under it) has to try the following positions one by one -- up to K look-ups per error, while the lanes of reads without
// errors wait.  Here a lane that misses asks the WARP: its next 32 positions are looked up side by side, the first hit
// (ballot) ends the gap.  Same parts as path_parts (the first position that hits is the same); 32 lanes busy instead of one.
// Every lane of the warp must call this (live = false for a lane without a read).
__device__ __forceinline__ uint32_t path_parts_warp(const DictView& d, const EdgeStore& es, const uint8_t* rd, uint32_t n, Part* parts, bool live)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t np = 0, itr = 0, end = 0;
    bool done = !live;
    if (live) { if (n < SN_K) { parts[np++] = mk_gap(n); done = true; } else end = n - SN_K + 1; }
    while (!__all_sync(0xFFFFFFFFu, done)) {
        uint32_t ent = SN_NULL_EDGE; Kmer kmer; kmer.w0 = kmer.w1 = kmer.w2 = 0;
        if (!done) { kmer = kmer_from_packed_w(rd, itr); bool rc; ent = dict_find(d, kmer, &rc); }
        const bool need = !done && ent == SN_NULL_EDGE;
        unsigned m = __ballot_sync(0xFFFFFFFFu, need);
        uint32_t my_pos = 0xFFFFFFFFu, my_ent = SN_NULL_EDGE;
        while (m) {                                              // serve the lanes that missed, one after the other
            const int src = __ffs((int)m) - 1; m &= m - 1u;
            const uint8_t* rrd = reinterpret_cast<const uint8_t*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)reinterpret_cast<uintptr_t>(rd), src));
            const uint32_t ritr = __shfl_sync(0xFFFFFFFFu, itr, src), rend = __shfl_sync(0xFFFFFFFFu, end, src);
            uint32_t hit_pos = 0xFFFFFFFFu, hit_ent = SN_NULL_EDGE;
            for (uint32_t base = ritr + 1; base < rend; base += 32) {
                const uint32_t q = base + lane;
                // the minimizers of the 32 k-mers at base .. base+31 from 64 p-mer values, two per lane: k-mer base+l holds the
                // p-mers base+l .. base+l+32 = a suffix of the first 32 values and a prefix of the second 32
                // (W = 33; the read buffer is padded, positions past the read feed only k-mers past its end)
                const uint32_t wa = packed_window16(rrd, q), wb = packed_window16(rrd, q + 32);
                uint32_t sa = pmer_value(rev2(wa), ~wa), pb = pmer_value(rev2(wb), ~wb);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t ta = __shfl_down_sync(0xFFFFFFFFu, sa, o), tb = __shfl_up_sync(0xFFFFFFFFu, pb, o);
                    if (lane + (uint32_t)o < 32u) sa = ta < sa ? ta : sa;
                    if (lane >= (uint32_t)o) pb = tb < pb ? tb : pb;
                }
                const uint32_t minimizer = sa < pb ? sa : pb;
                uint32_t e = SN_NULL_EDGE;
                if (q < rend) { const Kmer k = kmer_from_packed_w(rrd, q); bool rc; e = dict_find_min(d, k, minimizer, &rc); }
                const unsigned hm = __ballot_sync(0xFFFFFFFFu, e != SN_NULL_EDGE);
                if (hm) { const int hl = __ffs((int)hm) - 1; hit_pos = base + (uint32_t)hl; hit_ent = __shfl_sync(0xFFFFFFFFu, e, hl); break; }
            }
            if ((int)lane == src) { my_pos = hit_pos; my_ent = hit_ent; }
        }
        if (!done) {
            if (need) {
                if (my_pos == 0xFFFFFFFFu) { parts[np++] = mk_gap(end - itr); itr = end; }       // nothing hits up to the end of the read
                else { parts[np++] = mk_gap(my_pos - itr); itr = my_pos; ent = my_ent; kmer = kmer_from_packed_w(rd, itr); }
            }
            if (ent != SN_NULL_EDGE) { const Part p = seed_part(d, es, rd, n, itr, kmer, ent); parts[np++] = p; itr += p.len; }
            if (itr == end) done = true;
        }
    }
    return np;
}
#endif

// PathPart::isConformingCapturedGap (:669-676)
SN_HD bool conforming_gap(const Part* p, uint32_t max_jitter)
{
    const Part& prev = p[-1]; const Part& next = p[1];
    uint32_t dist = next.off - (prev.off + prev.len);
    if (!part_same_edge(prev, next)) dist += part_elen(prev);
    int32_t dd = (int32_t)(p->len - dist);
    return (uint32_t)(dd < 0 ? -dd : dd) <= max_jitter;
}
// Pather::isJoinable (:810-816): compares the LAST K-1 bases of both oriented edges.
SN_HD bool joinable(const EdgeStore& es, const Part& a, const Part& b)
{
    if (a.edge == b.edge) return true;
    uint32_t la = es.len[a.edge], lb = es.len[b.edge];
    for (uint32_t i = 0; i < SN_K - 1; ++i)
        if (edge_base(es, a.edge, part_rc(a), la - (SN_K - 1) + i) != edge_base(es, b.edge, part_rc(b), lb - (SN_K - 1) + i)) return false;
    return true;
}
SN_HD int32_t part_hbv_edge(const HbvView& h, const Part& p) { return part_rc(p) ? h.rev_xlat[p.edge] : h.fwd_xlat[p.edge]; }

struct RPath { int32_t offset; uint32_t n; int32_t e[SN_MAX_PATH]; bool overflow; };
SN_HD void rp_push(RPath& p, int32_t e) { if (p.n < SN_MAX_PATH) p.e[p.n++] = e; else p.overflow = true; }

// pathPartsToReadPath (:1385-1420)
SN_HD void parts_to_path(const HbvView& h, const Part* parts, uint32_t np, RPath& path)
{
    path.n = 0;
    int last = -1;
    for (uint32_t i = 0; i < np; ++i) {
        if (part_gap(parts[i])) continue;
        if (last >= 0 && part_same_edge(parts[last], parts[i])) continue;
        rp_push(path, part_hbv_edge(h, parts[i]));
        last = (int)i;
    }
    if (path.n == 0) path.offset = 0;
    else if (!part_gap(parts[0])) path.offset = (int32_t)parts[0].off;
    else path.offset = (int32_t)parts[1].off - (int32_t)parts[0].len;
}

// The running penalty is an `unsigned` decayed through a double
// (`penalty -= 0.2*penalty`, ExtendReadPath.cc:50,102): two separately rounded double
// operations, then truncation.  On the device the explicit _rn intrinsics stop the
// compiler from contracting them into one FMA.
SN_HD uint32_t decay_penalty(uint32_t penalty)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__dsub_rn((double)penalty, __dmul_rn(0.2, (double)penalty));
#else
    volatile double t = 0.2 * (double)penalty;
    return (uint32_t)((double)penalty - t);
#endif
}
// scoreRightOverlap (:15-62)
SN_HD uint32_t score_right(const EdgeStore& es, const HbvView& h, const uint8_t* rd, const uint8_t* q, uint32_t n,
                           uint32_t start, int32_t e)
{
    uint32_t esz = hbv_len(es, h, e);
    uint32_t bi = n - start, ei = SN_K - 1, qsum = 0, penalty = 0;
    while (bi != n && ei != esz) {
        if (packed_base(rd, bi) != hbv_base(es, h, e, ei)) { uint32_t qs = (q[bi] == 2) ? 20u : q[bi]; penalty += qs; qsum += penalty; }
        else if (penalty > 0) penalty = decay_penalty(penalty);
        ++bi; ++ei;
    }
    qsum += 10u * (n - bi);
    return qsum;
}
// scoreLeftOverlap (:65-115)
SN_HD uint32_t score_left(const EdgeStore& es, const HbvView& h, const uint8_t* rd, const uint8_t* q,
                          uint32_t start, int32_t e)
{
    uint32_t esz = hbv_len(es, h, e);
    int32_t bi = (int32_t)start - 1, ei = (int32_t)esz - SN_K;
    uint32_t qsum = 0, penalty = 0;
    while (bi >= 0 && ei >= 0) {
        if (packed_base(rd, (uint32_t)bi) != hbv_base(es, h, e, (uint32_t)ei)) { uint32_t qs = (q[bi] == 2) ? 20u : q[bi]; penalty += qs; qsum += penalty; }
        else if (penalty > 0) penalty = decay_penalty(penalty);
        --bi; --ei;
    }
    if (bi >= 0) qsum += 10u * (uint32_t)(bi + 1);
    return qsum;
}

// attemptLeftwardExtension (:133-236) / attemptRightwardExtension (:239-358).
// `left` selects which; the two differ only in which adjacency (to_/from_) is walked.
// `qs.get()` hands out the read's Phred bytes; it is only called when an extension is really scored (a PQVec stream is
// decoded then, not for every read: most reads are placed end to end and never get here)
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class QS>
SN_HD bool extend_once(const EdgeStore& es, const HbvView& h, RPath& p, const uint8_t* rd, QS& qs, uint32_t n, bool left)
{
    if (!p.n) return false;
    uint32_t last_gap;
    if (left) {
        if (p.offset >= 0) return false;
        last_gap = (uint32_t)(-p.offset);
    } else {
        int32_t il = (int32_t)n + p.offset;
        for (uint32_t i = 0; i < p.n; ++i) il -= (int32_t)hbv_len(es, h, p.e[i]) - SN_K + 1;
        il -= (SN_K - 1);
        if (il < 10) return false;
        last_gap = (uint32_t)il;
    }
    if (last_gap < 10) return false;
    const uint32_t* in_start = left ? h.to_start : h.from_start;     // adjacency walked
    const int32_t* in_v = left ? h.to_v : h.from_v;
    const int32_t* in_e = left ? h.to_e : h.from_e;
    const uint32_t* far_in = left ? h.to_start : h.from_start;       // ToSize / FromSize of the far vertex
    const uint32_t* far_out = left ? h.from_start : h.to_start;
    int32_t v = left ? h.to_left[p.e[0]] : h.to_right[p.e[p.n - 1]];
    uint32_t beg = in_start[v], ne = in_start[v + 1] - beg;
    int32_t short_dest[8]; uint32_t ns = 0, nlong = 0; bool hanging[8];
    if (ne > 8) ne = 8;                                              // HBVFromEdges.cc:83 MAX_EDGES
    for (uint32_t i = 0; i < ne; ++i) {
        int32_t vd = in_v[beg + i];
        hanging[i] = (far_in[vd + 1] - far_in[vd] == 0) && (far_out[vd + 1] - far_out[vd] == 1);
        bool elong = hbv_len(es, h, in_e[beg + i]) - (SN_K - 1) >= last_gap;
        nlong += elong ? 1u : 0u;
        if (!elong && !hanging[i]) short_dest[ns++] = vd;
    }
    if (ne != 1 && ns > 0) {
        if (nlong > 0) return false;
        for (uint32_t i = 1; i < ns; ++i) if (short_dest[i] != short_dest[0]) return false;   // UniqueSort + solo
        int32_t sd = short_dest[0];
        if (far_in[sd + 1] - far_in[sd] != 1) return false;
    }
    int32_t least_edge = -1; uint32_t least = 0xFFFFFFFFu;
    const uint8_t* q = qs.get();
    for (uint32_t i = 0; i < ne; ++i) if (!hanging[i] || ne == 1) {
        int32_t e = in_e[beg + i];
        uint32_t sc = left ? score_left(es, h, rd, q, last_gap, e) : score_right(es, h, rd, q, n, last_gap, e);
        if (sc < least) { least_edge = e; least = sc; }
    }
    if (least_edge == -1 || least > last_gap * 10u) return false;
    if (left) {
        p.offset += (int32_t)hbv_len(es, h, least_edge) - SN_K + 1;
        if (p.n >= SN_MAX_PATH) { p.overflow = true; return false; }
        for (uint32_t i = p.n; i > 0; --i) p.e[i] = p.e[i - 1];
        p.e[0] = least_edge; ++p.n;
    } else {
        if (p.n >= SN_MAX_PATH) { p.overflow = true; return false; }
        p.e[p.n++] = least_edge;
    }
    return true;
}

// HBVPather::algorithmTwo (:1217-1336).  `parts` is scratch of SN_MAX_PARTS.
struct PlainQuals { const uint8_t* q; SN_HD const uint8_t* get() { return q; } };
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
// everything of algorithmTwo after Pather::path: `parts[0..np)` -> the ReadPath
template <class QS>
SN_HD void path_from_parts(const EdgeStore& es, const HbvView& h, const uint8_t* rd, QS& qs, uint32_t n, Part* parts, uint32_t np, RPath& path)
{
    path.n = 0; path.offset = 0; path.overflow = false;
    // seeds on short hanging edges become gaps; adjacent gaps merge (:1236-1258); in place
    uint32_t nn = 0;
    for (uint32_t i = 0; i < np; ++i) {
        Part part = parts[i];
        if (!part_gap(part)) {
            int32_t e = part_hbv_edge(h, part);
            int32_t vl = h.to_left[e], vr = h.to_right[e];
            if (h.to_start[vl + 1] - h.to_start[vl] == 0 && h.to_start[vr + 1] - h.to_start[vr] > 1 &&
                h.from_start[vr + 1] - h.from_start[vr] > 0 && part_elen(part) <= 100)
                part = mk_gap(part.len);
        }
        if (part_gap(part) && nn && part_gap(parts[nn - 1])) parts[nn - 1].len += part.len;
        else parts[nn++] = part;
    }
    np = nn;
    // first non-conforming captured gap truncates the placement (:1264-1288)
    if (np >= 3) {
        uint32_t seeds = part_gap(parts[0]) ? 0u : 1u;
        for (uint32_t i = 1; i + 1 < np; ++i) {
            if (!part_gap(parts[i])) { ++seeds; continue; }
            if (!conforming_gap(&parts[i], 3) || !joinable(es, parts[i - 1], parts[i + 1])) {
                if (seeds > 1) {
                    Part tmp = mk_gap(parts[i - 1].len);
                    for (uint32_t j = i; j < np; ++j) tmp.len += parts[j].len;
                    np = i - 1; parts[np++] = tmp;
                } else {
                    for (uint32_t j = i + 1; j < np; ++j) parts[i].len += parts[j].len;
                    np = i + 1;
                }
                break;
            }
        }
    }
    // back off terminal seeds that reach <= 5 k-mers onto an edge (:1293-1307)
    if (part_gap(parts[np - 1]) && np > 1) {
        const Part& last2 = parts[np - 2];
        if (last2.off == 0 && last2.len <= 5) {
            Part last = parts[np - 1];
            last.len += last2.len;
            np -= 2; parts[np++] = last;
        }
    } else if (!part_gap(parts[np - 1])) {
        Part& last = parts[np - 1];
        if (last.off == 0 && last.len <= 5) last = mk_gap(last.len);
    }
    parts_to_path(h, parts, np, path);
    // truncate at the first graph discontinuity (:1313-1320)
    if (path.n >= 2)
        for (uint32_t i = 0; i + 1 < path.n; ++i)
            if (h.to_right[path.e[i]] != h.to_left[path.e[i + 1]]) { path.n = i + 1; break; }
    // ExtendReadPath::attemptLeftRightExtension (ExtendReadPath.cc:121-129)
    while (extend_once(es, h, path, rd, qs, n, true)) {}
    while (extend_once(es, h, path, rd, qs, n, false)) {}
}
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class QS>
SN_HD void path_one_read_q(const DictView& d, const EdgeStore& es, const HbvView& h,
                           const uint8_t* rd, QS& qs, uint32_t n, Part* parts, RPath& path)
{
    const uint32_t np = path_parts(d, es, rd, n, parts);
    path_from_parts(es, h, rd, qs, n, parts, np, path);
}
SN_HD void path_one_read(const DictView& d, const EdgeStore& es, const HbvView& h,
                         const uint8_t* rd, const uint8_t* q, uint32_t n, Part* parts, RPath& path)
{
    PlainQuals qs; qs.q = q;
    path_one_read_q(d, es, h, rd, qs, n, parts, path);
}

}  // namespace sn
