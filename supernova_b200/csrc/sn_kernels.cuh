// sn_kernels.cuh -- the CUDA kernels of the hot path (sm_100a), in pipeline order:
//   a1  k_pqvec_goodlen / k_q8_goodlen   PQVec decode + good-length scan
//   a2-a5, a14-a15  sn_msp.cuh            super-k-mers by minimizer bucket, per-bucket count + filter,
//                                         survivors sorted by hash (sn_prims.cuh) into the dictionary
//   a6  k_prune                           adjacency prune
//   a7  k_classify, k_seg_walk, k_end_hop, k_owner_hop, k_seg_emit, k_circle_count, k_edge_form, k_fix_offsets, k_pack_edges
//   a10-a12 k_path_reads                  ReadPath threading + extension
// Reference citations live with the per-item logic in sn_kmer.cuh / sn_graph.cuh /
// sn_path.cuh; this file is thread mapping, staging and memory layout.
#pragma once
#include "sn_prims.cuh"
#include "sn_kmer.cuh"
#include "sn_graph.cuh"
#include "sn_path.cuh"

namespace sn {

// ---------------------------------------------------------------------------
// a1. PQVec decode (feudal/PQVec.cc:129-187 block format: [nQs u8][nBits:3|minQ:6]
// [nQs x nBits packed, LSB first] ... 0) fused with GoodLenTailFinder
// (BuildReadQGraph48.cc:65-89).  One read per thread; nothing but the good length leaves the
// thread (the pathing kernel decodes the quals of its read again on demand -- 0.31 B/base of
// PQVec is cheaper to re-read than 1 B/base of unpacked quals is to write and read).
// goodLen = right end of the right-most run of >= K quals >= minQual (what the backwards scan
// of the reference finds first).  Also accumulates the number of k-mer occurrences
// Kmerizer::map will emit.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_pqvec_goodlen(uint64_t n_reads, const uint8_t* __restrict__ pq, const uint64_t* __restrict__ pq_off,
                                                       const uint32_t* __restrict__ len, uint32_t min_qual, uint32_t min_gl /* reads trimmed below it give no k-mer */,
                                                       uint32_t* __restrict__ goodlen, unsigned long long* occ_total, uint32_t* bad_reads)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t occ = 0;
    if (r < n_reads) {
        // Block by block.  A block whose base quality is already >= minQual holds only good quals (every qual is
        // base + a non-negative delta): its packed deltas are skipped unread -- on sequencing data that is nearly every
        // block (a block boundary sits wherever a low quality starts or ends), so a read costs a few header parses instead
        // of one decode step per base.  Only blocks that start below minQual are unpacked qual by qual.
        const uint8_t* p = pq + pq_off[r];
        const uint8_t* pend = pq + pq_off[r + 1];
        const uint32_t L = len[r];
        uint32_t run = 0, gl = 0, i = 0, rem = 0;
        bool bad = false;
        while (i < L) {
            if (p + 3 > pend) { bad = true; break; }            // [nQs][nBits:3|minQ low 5][minQ bit 5 | 7 data bits]...
            rem = *p++;
            if (!rem || rem > L - i) { bad = true; break; }      // terminator before L quals / more quals than bases
            const uint32_t b0 = *p++, a = *p++;
            const uint32_t nbits = b0 & 7u, minq = (b0 >> 3) | ((a & 1u) << 5);
            if (minq >= min_qual) {
                run += rem; i += rem;
                if (run >= SN_K) gl = i;
                p += ((1u + rem * nbits + 7u) >> 3) - 1u;        // the block's data bytes beyond the one the header shares
                rem = 0;
                if (p > pend) { bad = true; break; }
                continue;
            }
            uint32_t acc = a >> 1, have = 7;
            const uint32_t mask = (1u << nbits) - 1u;
            for (; rem; --rem) {
                if (have < nbits) { acc |= (uint32_t)(*p++) << have; have += 8; }
                const uint32_t q = minq + (acc & mask); acc >>= nbits; have -= nbits;
                run = q >= min_qual ? run + 1 : 0;
                ++i;
                if (run >= SN_K) gl = i;
            }
        }
        if (rem != 0 || (p < pend && *p != 0)) bad = true;      // the PQVec holds more quals than the read has bases
        if (bad) { atomicAdd(bad_reads, 1u); gl = 0; }
        goodlen[r] = gl;
        occ = gl >= min_gl ? gl - SN_K + 1 : 0;
    }
    // block reduce, one atomic per block
    __shared__ uint32_t sm[8];
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_down_sync(SN_FULL, occ, o);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = occ;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 8; ++w) t += sm[w]; if (t) atomicAdd(occ_total, (unsigned long long)t); }
}

// out[i] = in[i] + add
static __global__ void __launch_bounds__(256) k_add_u64(const uint64_t* __restrict__ in, uint64_t n, uint64_t add, uint64_t* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + add;
}
// input limits checked on the device: total bases, longest read, largest barcode ordinal
static __global__ void __launch_bounds__(256) k_read_stats(uint64_t n_reads, const uint32_t* __restrict__ len, const int32_t* __restrict__ bc,
                                                    unsigned long long* total_bases, uint32_t* max_len, int32_t* max_bc)
{
    uint64_t sum = 0; uint32_t ml = 0; int32_t mb = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = len[r]; sum += l; ml = max(ml, l);
        if (bc) mb = max(mb, bc[r]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(SN_FULL, sum, o); ml = max(ml, __shfl_down_sync(SN_FULL, ml, o)); mb = max(mb, __shfl_down_sync(SN_FULL, mb, o));
    }
    if (lane_id() == 0) { atomicAdd(total_bases, (unsigned long long)sum); atomicMax(max_len, ml); atomicMax(max_bc, mb); }
}

// same, for callers that already hold one u8 per base
static __global__ void __launch_bounds__(256) k_q8_goodlen(uint64_t n_reads, const uint8_t* __restrict__ quals, const uint64_t* __restrict__ qoff,
                                                    const uint32_t* __restrict__ len, uint32_t min_qual, uint32_t min_gl,
                                                    uint32_t* __restrict__ goodlen, unsigned long long* occ_total)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t occ = 0;
    if (r < n_reads) {
        const uint8_t* q = quals + qoff[r];
        uint32_t L = len[r], run = 0, gl = 0;
        for (uint32_t i = 0; i < L; ++i) { run = q[i] >= min_qual ? run + 1 : 0; if (run >= SN_K) gl = i + 1; }
        goodlen[r] = gl;
        occ = gl >= min_gl ? gl - SN_K + 1 : 0;
    }
    __shared__ uint32_t sm[8];
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_down_sync(SN_FULL, occ, o);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = occ;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 8; ++w) t += sm[w]; if (t) atomicAdd(occ_total, (unsigned long long)t); }
}

__device__ __forceinline__ bool same_kmer(const uint4& a, const uint4& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// ---------------------------------------------------------------------------
// a6. recomputeAdjacencies
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_prune(DictEntry* tab, DictView d, Link2* __restrict__ cand)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = d.n;
    if (i >= n) return;
    Link2 l;
    tab[i].ctx = prune_ctx(d, i, &l);
    tab[i].edge = SN_NULL_EDGE; tab[i].off = 0;        // the edge stage below starts from a clean table (it may be run again)
    cand[i] = l;
}

// ---------------------------------------------------------------------------
// a7. unipath edges.  own_n[i] = number of k-mers of the edge that entry i owns
// (0 = owns none).  Singles own themselves; an edge with >= 2 k-mers is owned by the one of
// its two end entries with the smaller index; a circle is owned by its smallest k-mer.
//
// An edge is a chain of unipath links, and following a chain is one dependent random load
// per k-mer: walking a whole edge from its end costs (edge length) x (DRAM latency), and the
// longest edge alone would set the run time of the stage.  The chains are therefore cut at
// STOPS -- the edge ends plus every interior k-mer whose index hashes to 0 mod 16 (a segment then mostly stays inside one dictionary bucket, whose lines the neighbouring threads share):
//   k_seg_walk   every stop walks to the next stop on each side (~16 steps, all segments in parallel)
//                and records {next stop, arrival orientation, steps}
//   k_end_hop    every edge end hops from stop to stop (L/16 steps over a small, L2-resident table)
//                to the far end -> length and owner of the edge
//   k_owner_hop  after allocation the owner hops again and leaves {edge, offset, orientation} at
//                every stop of its edge
//   k_seg_emit   every stop of an edge writes its own k-mer and the k-mers up to the next stop:
//                bases into the edge scratch, (edge, offset) into the dictionary
// Interior k-mers that no edge end reaches are circles (k_circle_count, after the edges above).
// ---------------------------------------------------------------------------
#define SN_STOP_SHIFT 28          // interior entry i is a stop iff (i * 0x9E3779B1) >> 28 == 0 (1 in 16)
__device__ __forceinline__ bool stop_sampled(uint32_t i) { return ((i * 0x9E3779B1u) >> SN_STOP_SHIFT) == 0u; }
struct Seg { uint32_t next; uint32_t steps_o; };        // next stop (stop id, SN_NO_LINK = none on this side); steps << 1 | arrival orientation

static __global__ void __launch_bounds__(256) k_classify(DictView d,
                                                  Link2* links, uint8_t* __restrict__ etype, uint32_t* __restrict__ own_n, uint32_t* __restrict__ is_stop)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    Link2 l;
    int t = classify_links(d, i, links[i], &l);          // links[] holds prune's candidates on entry
    links[i] = l;
    etype[i] = (uint8_t)t;
    own_n[i] = t == T_SINGLE ? 1u : 0u;
    is_stop[i] = (t == T_END_DOWN || t == T_END_UP || (t == T_INTERIOR && stop_sampled(i))) ? 1u : 0u;
}
static __global__ void __launch_bounds__(256) k_scatter_flagged(const uint32_t* __restrict__ flag, const uint64_t* __restrict__ pos, uint32_t n, uint32_t* __restrict__ list)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) list[pos[i]] = i;
}
// thread per (stop, side): side 0 leaves through the down link, side 1 through the up link
static __global__ void __launch_bounds__(128) k_seg_walk(const Link2* __restrict__ links, const uint32_t* __restrict__ stops, uint32_t n_stops,
                                                  const uint64_t* __restrict__ stop_pos, Seg* __restrict__ segs)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_stops) return;
    const uint32_t s = stops[t >> 1];
    uint32_t o = t & 1u, cur = s, steps = 0;
    Seg out; out.next = SN_NO_LINK; out.steps_o = 0;
    Link2 lk = links[cur];
    for (;;) {
        const uint32_t l = o ? lk.y : lk.x;
        if (l == SN_NO_LINK) break;                       // only at the start: this side of an edge end is closed
        cur = l >> 1; o = l & 1u; ++steps;
        lk = links[cur];
        const uint32_t cont = o ? lk.y : lk.x;
        if (cont == SN_NO_LINK || stop_sampled(cur) || cur == s) {      // an edge end, a sampled interior, or once around a circle
            out.next = (uint32_t)stop_pos[cur]; out.steps_o = (steps << 1) | o;
            break;
        }
    }
    segs[t] = out;
}
// thread per stop that is an edge end: hop to the far end of the edge
static __global__ void __launch_bounds__(128) k_end_hop(const Seg* __restrict__ segs, const uint32_t* __restrict__ stops, uint32_t n_stops,
                                                 const uint8_t* __restrict__ etype, uint32_t* __restrict__ own_n)
{
    const uint32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
    if (sid >= n_stops) return;
    const uint32_t i = stops[sid];
    const int t = etype[i];
    if (t != T_END_DOWN && t != T_END_UP) return;
    uint32_t cur = sid, o = t == T_END_UP ? 1u : 0u, nk = 1;
    for (;;) {
        const Seg sg = segs[2 * cur + o];
        if (sg.next == SN_NO_LINK) break;
        nk += sg.steps_o >> 1; cur = sg.next; o = sg.steps_o & 1u;
    }
    if (i <= stops[cur]) own_n[i] = nk;                  // the other end walks the same edge; the smaller index owns it
}
static __global__ void __launch_bounds__(128) k_circle_count(const DictEntry* __restrict__ tab, const Link2* __restrict__ links, uint32_t n,
                                                     uint8_t* __restrict__ etype, uint32_t* __restrict__ own_n, uint32_t* n_circle_members)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (etype[i] != T_INTERIOR || tab[i].edge != SN_NULL_EDGE) return;     // on an edge with ends: done
    atomicAdd(n_circle_members, 1u);
    // the walker with the smallest table index completes the loop; the circle is then owned by
    // its smallest K-MER, where canonicalizeCircle (BuildReadQGraph48.cc:375-397) starts it
    uint32_t m = i; Kmer mk = entry_kmer(tab[i]);
    uint32_t nk = walk_circle_links(links, i, true, [&](uint32_t, uint32_t j, uint32_t) { Kmer q = entry_kmer(tab[j]); if (q < mk) { mk = q; m = j; } });
    if (nk) { own_n[m] = nk; etype[m] = T_CIRCLE; }
}
// counts the interior entries no edge end reached (cheap: one byte + one sector per interior entry)
static __global__ void __launch_bounds__(256) k_count_unreached(const DictEntry* __restrict__ tab, const uint8_t* __restrict__ etype, uint32_t n, uint32_t* count)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool u = i < n && etype[i] == T_INTERIOR && tab[i].edge == SN_NULL_EDGE;
    uint32_t m = __ballot_sync(SN_FULL, u);
    if (m && lane_id() == 0) atomicAdd(count, (uint32_t)__popc(m));
}
// phase 0: every owner so far (singles, edges with ends); phase 1: the circle owners only
static __global__ void __launch_bounds__(256) k_edge_sizes(const uint32_t* __restrict__ own_n, const uint8_t* __restrict__ etype, int circles_only, uint32_t n,
                                                    uint32_t* __restrict__ ebases, uint32_t* __restrict__ eflag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = own_n[i];
    if (circles_only && etype[i] != T_CIRCLE) k = 0;
    ebases[i] = k ? k + (SN_K - 1) : 0u;
    eflag[i] = k ? 1u : 0u;
}
// thread per edge (owner list): singles and circles are written here in one go (a circle is walked by
// its owner -- circles are rare); the owner of an edge with ends hops over its stops and leaves
// {edge, offset, walk orientation} at each for k_seg_emit.
struct StopInfo { uint32_t edge, off_o; };               // off_o = offset << 1 | orientation
static __global__ void __launch_bounds__(128) k_owner_hop(DictEntry* tab, const Link2* __restrict__ links, const Seg* __restrict__ segs,
                                                   const uint64_t* __restrict__ stop_pos, const uint32_t* __restrict__ owners, uint32_t n_owners, uint32_t edge0,
                                                   const uint8_t* __restrict__ etype, const uint32_t* __restrict__ own_n, const uint64_t* __restrict__ base_off, uint64_t base_shift,
                                                   uint8_t* __restrict__ tmp, uint32_t* __restrict__ elen, uint64_t* __restrict__ etmp_off, StopInfo* __restrict__ sinfo)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_owners) return;
    const uint32_t i = owners[k], e = edge0 + k;
    const int t = etype[i];
    const uint32_t nk = own_n[i];
    const uint64_t boff = base_off[i] + base_shift;
    elen[e] = nk + SN_K - 1;
    etmp_off[e] = boff;
    if (t == T_END_DOWN || t == T_END_UP) {
        uint32_t cur = (uint32_t)stop_pos[i], o = t == T_END_UP ? 1u : 0u, off = 0;
        for (;;) {
            StopInfo si; si.edge = e; si.off_o = (off << 1) | o;
            sinfo[cur] = si;
            const Seg sg = segs[2 * cur + o];
            if (sg.next == SN_NO_LINK) break;
            off += sg.steps_o >> 1; cur = sg.next; o = sg.steps_o & 1u;
        }
        return;
    }
    uint8_t* s = tmp + boff;
    Kmer km = entry_kmer(tab[i]);
    for (int b = 0; b < SN_K; ++b) s[b] = (uint8_t)kmer_base(km, b);
    tab[i].edge = e; tab[i].off = 0;
    if (t == T_CIRCLE)
        walk_circle_links(links, i, false, [&](uint32_t step, uint32_t j, uint32_t o) { s[SN_K - 1 + step] = (uint8_t)step_base(tab[j], o); tab[j].edge = e; tab[j].off = step; });
}
// thread per stop on an edge with ends: its own k-mer, then the k-mers up to (not including) the next stop
static __global__ void __launch_bounds__(128) k_seg_emit(DictEntry* tab, const Link2* __restrict__ links, const uint32_t* __restrict__ stops, uint32_t n_stops,
                                                  const StopInfo* __restrict__ sinfo, const Seg* __restrict__ segs, const uint64_t* __restrict__ etmp_off, uint8_t* __restrict__ tmp)
{
    const uint32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
    if (sid >= n_stops) return;
    const StopInfo si = sinfo[sid];
    if (si.edge == SN_NULL_EDGE) return;                  // a stop on a circle
    uint32_t cur = stops[sid], o = si.off_o & 1u, off = si.off_o >> 1;
    uint8_t* s = tmp + etmp_off[si.edge];
    if (off == 0) {                                       // the owner end: all K bases of its k-mer, in walk orientation
        Kmer km = entry_kmer(tab[cur]);
        if (o) km = kmer_rc(km);
        for (int b = 0; b < SN_K; ++b) s[b] = (uint8_t)kmer_base(km, b);
    } else s[SN_K - 1 + off] = (uint8_t)step_base(tab[cur], o);
    tab[cur].edge = si.edge; tab[cur].off = off;
    const Seg sg = segs[2 * sid + o];
    if (sg.next == SN_NO_LINK) return;
    const uint32_t steps = sg.steps_o >> 1;
    for (uint32_t k = 1; k < steps; ++k) {
        const Link2 lk = links[cur];
        const uint32_t l = o ? lk.y : lk.x;
        cur = l >> 1; o = l & 1u;
        s[SN_K - 1 + off + k] = (uint8_t)step_base(tab[cur], o);
        tab[cur].edge = si.edge; tab[cur].off = off + k;
    }
}
// whole-edge canonical form (EdgeBuilder::addEdge :480-485 / extend :457-464): is the edge stored reverse-complemented?
static __global__ void __launch_bounds__(128) k_edge_form(const uint8_t* __restrict__ tmp, const uint64_t* __restrict__ etmp_off, const uint32_t* __restrict__ elen,
                                                   uint32_t n_edges, uint8_t* __restrict__ eflip)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) eflip[e] = seq_form_u8(tmp + etmp_off[e], elen[e]) == REV ? 1 : 0;
}
static __global__ void __launch_bounds__(256) k_fix_offsets(DictEntry* tab, uint32_t n, const uint32_t* __restrict__ elen, const uint8_t* __restrict__ eflip)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t e = tab[i].edge;
    if (e != SN_NULL_EDGE && eflip[e]) tab[i].off = (elen[e] - (SN_K - 1)) - 1 - tab[i].off;
}
static __global__ void __launch_bounds__(256) k_edge_bytes(const uint32_t* __restrict__ elen, uint32_t n_edges, uint32_t* __restrict__ ebytes)
{
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) ebytes[e] = (elen[e] + 3) >> 2;
}
// one thread per output byte of the packed edge store (fastb layout)
static __global__ void __launch_bounds__(256) k_pack_edges(const uint8_t* __restrict__ tmp, const uint64_t* __restrict__ etmp_off,
                                                    const uint32_t* __restrict__ elen, const uint8_t* __restrict__ eflip,
                                                    const uint64_t* __restrict__ eoff, uint32_t n_edges, uint64_t total_bytes, uint8_t* __restrict__ packed)
{
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= total_bytes) return;
    uint32_t lo = 0, hi = n_edges;                     // largest e with eoff[e] <= x
    while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (eoff[m] <= x) lo = m; else hi = m; }
    uint32_t e = lo, len = elen[e];
    uint32_t b0 = (uint32_t)(x - eoff[e]) * 4;
    const uint8_t* s = tmp + etmp_off[e];
    uint32_t v = 0;
    for (uint32_t j = 0; j < 4 && b0 + j < len; ++j) {
        uint32_t c = eflip[e] ? (s[len - 1 - (b0 + j)] ^ 3u) : s[b0 + j];
        v |= c << (2 * j);
    }
    packed[x] = (uint8_t)v;
}

// ---------------------------------------------------------------------------
// a8. vertex discovery for buildHBVFromEdges (paths/long/HBVFromEdges.cc:124-168,277):
// per edge, the four (K-1)-mer end keys {fwd,rc} x {start,end} (two for a palindromic
// edge) as sortable records, plus an order record (length desc, first 32 bases) for the
// canonical edge order.  Both record sets go through radix_sort_kmers; k_hbv_mark /
// k_hbv_assign turn runs of equal keys into vertex groups.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int seq_form_packed(const uint8_t* s, uint32_t len)
{
    if (len & 1) return (packed_base(s, len / 2) & 2) ? REV : FWD;
    uint32_t i = 0, j = len;
    while (i != j) {
        uint32_t f = packed_base(s, i), r = packed_base(s, --j) ^ 3u;
        if (f < r) return FWD;
        if (r < f) return REV;
        ++i;
    }
    return PAL;
}
static __global__ void __launch_bounds__(128) k_hbv_keys(const uint8_t* __restrict__ ebases, const uint64_t* __restrict__ eoff, const uint32_t* __restrict__ elen,
                                                  uint32_t n_edges, uint4* __restrict__ order_rec, uint4* __restrict__ end_rec, uint8_t* __restrict__ pal, uint32_t* n_pal)
{
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const uint8_t* s = ebases + eoff[e];
    uint32_t len = elen[e];
    bool p = seq_form_packed(s, len) == PAL;
    pal[e] = p ? 1 : 0;
    if (p) atomicAdd(n_pal, 1u);
    Kmer kf = kmer_from_packed(s, 0), kl = kmer_from_packed(s, len - SN_K);
    order_rec[e] = make_uint4(~len, kf.w0, kf.w1, e);
    Kmer a = kf; a.w2 &= ~3u;                          // first K-1 bases
    Kmer b = kmer_succ(kl, 0);                          // last K-1 bases
    Kmer rl = kmer_rc(kl); rl.w2 &= ~3u;                // first K-1 bases of the reverse complement
    Kmer rf = kmer_succ(kmer_rc(kf), 0);                // last K-1 bases of the reverse complement
    const uint4 inval = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    end_rec[4ull * e + 0] = make_uint4(a.w0, a.w1, a.w2, (e << 2) | 0u);
    end_rec[4ull * e + 1] = make_uint4(b.w0, b.w1, b.w2, (e << 2) | 1u);
    end_rec[4ull * e + 2] = p ? inval : make_uint4(rl.w0, rl.w1, rl.w2, (e << 2) | 2u);
    end_rec[4ull * e + 3] = p ? inval : make_uint4(rf.w0, rf.w1, rf.w2, (e << 2) | 3u);
}
static __global__ void __launch_bounds__(256) k_hbv_mark(const uint4* __restrict__ rec, uint32_t n, uint32_t* __restrict__ flag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = rec[i];
    bool valid = r.w != 0xFFFFFFFFu;
    flag[i] = (valid && (i == 0 || !same_kmer(rec[i - 1], r))) ? 1u : 0u;
}
static __global__ void __launch_bounds__(256) k_hbv_assign(const uint4* __restrict__ rec, uint32_t n, const uint32_t* __restrict__ flag, const uint64_t* __restrict__ pos,
                                                    int32_t* __restrict__ end_group, uint32_t* __restrict__ items, uint32_t* __restrict__ gstart)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = rec[i];
    if (r.w == 0xFFFFFFFFu) return;
    uint32_t g = (uint32_t)pos[i] + flag[i] - 1u;
    end_group[r.w] = (int32_t)g;
    items[i] = r.w;
    if (flag[i]) gstart[g] = i;
}

// ---------------------------------------------------------------------------
// a10-a13. ReadPath threading, one read per thread.  k_path_reads computes every path once:
// length and offset always, the first SN_PATH_INLINE edges into a fixed-stride scratch.  After
// the scan of the lengths k_path_finish copies the short paths (nearly all) to their final
// place and re-threads only the reads whose path did not fit the scratch.  Quals come either
// unpacked (quals/qoff) or as the PQVec stream (pq/pq_off), decoded into a per-thread buffer.
// ---------------------------------------------------------------------------
#define SN_PATH_INLINE 4
struct PathInputs {
    uint64_t n_reads;
    const uint8_t* bases; const uint64_t* boff; const uint32_t* len;
    const uint8_t* quals; const uint64_t* qoff;          // unpacked quals, or
    const uint8_t* pq; const uint64_t* pq_off;           // PQVec stream
};
// the quals of a read, decoded from its PQVec stream on first use
struct LazyQuals {
    const uint8_t* q8; const uint8_t* pq; const uint8_t* pq_end; uint8_t* buf; bool ready;
    __device__ __forceinline__ const uint8_t* get()
    {
        if (q8) return q8;
        if (!ready) { pqvec_decode(pq, pq_end, buf, SN_MAX_READ_LEN); ready = true; }
        return buf;
    }
};
__device__ __forceinline__ void thread_path(const PathInputs& in, uint64_t r, const DictView& d, const EdgeStore& es, const HbvView& h,
                                            Part* parts, RPath& path, uint8_t* qbuf)
{
    LazyQuals qs;
    qs.q8 = in.pq ? nullptr : in.quals + in.qoff[r];
    qs.pq = in.pq ? in.pq + in.pq_off[r] : nullptr; qs.pq_end = in.pq ? in.pq + in.pq_off[r + 1] : nullptr;
    qs.buf = qbuf; qs.ready = false;
    path_one_read_q(d, es, h, in.bases + in.boff[r], qs, in.len[r], parts, path);
}
static __global__ void __launch_bounds__(128) k_path_reads(PathInputs in, DictView d, EdgeStore es, HbvView h,
                                                    uint32_t* __restrict__ plen, int32_t* __restrict__ poffset, int32_t* __restrict__ scratch, uint32_t* overflow)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = r < in.n_reads;                            // (no early return: the warp threads its 32 reads together, path_parts_warp)
    Part parts[SN_MAX_PARTS];
    RPath path;
    uint8_t qbuf[SN_MAX_READ_LEN];
    const uint8_t* rd = live ? in.bases + in.boff[r] : in.bases;
    const uint32_t n = live ? in.len[r] : 0u;
    const uint32_t np = path_parts_warp(d, es, rd, n, parts, live);
    if (!live) return;
    LazyQuals qs;
    qs.q8 = in.pq ? nullptr : in.quals + in.qoff[r];
    qs.pq = in.pq ? in.pq + in.pq_off[r] : nullptr; qs.pq_end = in.pq ? in.pq + in.pq_off[r + 1] : nullptr;
    qs.buf = qbuf; qs.ready = false;
    path_from_parts(es, h, rd, qs, n, parts, np, path);
    if (path.overflow) atomicAdd(overflow, 1u);
    plen[r] = path.n; poffset[r] = path.offset;
    int4 v = make_int4(path.n > 0 ? path.e[0] : 0, path.n > 1 ? path.e[1] : 0, path.n > 2 ? path.e[2] : 0, path.n > 3 ? path.e[3] : 0);
    reinterpret_cast<int4*>(scratch)[r] = v;
}
static __global__ void __launch_bounds__(128) k_path_finish(PathInputs in, DictView d, EdgeStore es, HbvView h,
                                                     const uint32_t* __restrict__ plen, const int32_t* __restrict__ scratch,
                                                     const uint64_t* __restrict__ path_off, int32_t* __restrict__ pedges)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= in.n_reads) return;
    uint32_t n = plen[r];
    if (!n) return;
    int32_t* o = pedges + path_off[r];
    if (n <= SN_PATH_INLINE) {
        int4 v = reinterpret_cast<const int4*>(scratch)[r];
        o[0] = v.x; if (n > 1) o[1] = v.y; if (n > 2) o[2] = v.z; if (n > 3) o[3] = v.w;
        return;
    }
    Part parts[SN_MAX_PARTS];
    RPath path;
    uint8_t qbuf[SN_MAX_READ_LEN];
    thread_path(in, r, d, es, h, parts, path, qbuf);
    for (uint32_t i = 0; i < path.n; ++i) o[i] = path.e[i];
}

// ---------------------------------------------------------------------------
// SURVEY §8(f) row 1 (DF side): writePathsIndex (10X/PathsIndex.cc:23-143) -- the edge -> reads index of the
// ReadPaths.  One {edge, 0, read, 0} record per path entry, sorted by the device radix sort on (edge, read)
// (the reference sorts pair<int, unsigned long> the same way: a read that crosses an edge twice is listed
// twice), per-edge counts by atomics, and countsb[e] = reads on e plus reads on its reverse complement.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_pi_records(const int32_t* __restrict__ pedges, const uint64_t* __restrict__ path_off, uint64_t n_reads,
                                                    uint4* __restrict__ rec, uint32_t* __restrict__ cnt)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    for (uint64_t k = path_off[r]; k < path_off[r + 1]; ++k) {
        const uint32_t e = (uint32_t)pedges[k];
        rec[k] = make_uint4(e, 0u, (uint32_t)r, 0u);
        atomicAdd(cnt + e, 1u);
    }
}
static __global__ void __launch_bounds__(256) k_pi_ids(const uint4* __restrict__ rec, uint64_t m, unsigned long long* __restrict__ ids)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) ids[i] = rec[i].z;
}
static __global__ void __launch_bounds__(256) k_pi_countsb(const uint32_t* __restrict__ cnt, const int32_t* __restrict__ inv, uint32_t n_h, int32_t* __restrict__ countsb)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_h) return;
    const int32_t re = inv[e];
    countsb[e] = (int32_t)(cnt[e] + ((uint32_t)re != e ? cnt[re] : 0u));     // (:123-131; a palindromic edge keeps its own count)
}

}  // namespace sn
