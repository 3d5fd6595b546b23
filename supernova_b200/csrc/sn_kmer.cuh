// sn_kmer.cuh -- K=48 k-mer, context and dictionary primitives shared by every
// kernel of the hot path.  Everything here is `SN_HD` (host+device) so the exact
// per-item logic the kernels run can also be unit-tested on a CPU box
// (tests/hostsim); the product library only ever calls it from device code.
//
// Representation (bit-exact with the reference):
//   k-mer   : 3 x u32, base i at bits 2*(15 - i%16) of word i/16, A=0 C=1 G=2 T=3
//             (kmers/KMer.h:154-160,344-350); lexicographic order == integer order.
//   context : (predMask << 4) | succMask, bit b <=> base code b
//             (kmers/KMerContext.h:27-28,103-107); RC = bit reversal of the byte
//             (kmers/KMerContext.cc:18-36).
//   reads / edges in HBM: fastb packing, 4 bases per byte, base j at bits 2*(j%4)
//             (feudal/FieldVec.h:596-598,762).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SN_HD __host__ __device__ __forceinline__
#define SN_D __device__ __forceinline__
#else
#define SN_HD inline
#define SN_D inline
#endif

#define SN_K 48

namespace sn {

struct Kmer { uint32_t w0, w1, w2; };

SN_HD bool operator==(const Kmer& a, const Kmer& b) { return a.w0 == b.w0 && a.w1 == b.w1 && a.w2 == b.w2; }
SN_HD bool operator<(const Kmer& a, const Kmer& b)
{ return a.w0 != b.w0 ? a.w0 < b.w0 : (a.w1 != b.w1 ? a.w1 < b.w1 : a.w2 < b.w2); }

SN_HD uint32_t brev32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
// reverse the order of the sixteen 2-bit fields of a word
SN_HD uint32_t rev2(uint32_t x)
{
    x = brev32(x);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}
// KMer::rc (kmers/KMer.h:203-225)
SN_HD Kmer kmer_rc(const Kmer& k) { Kmer r; r.w0 = ~rev2(k.w2); r.w1 = ~rev2(k.w1); r.w2 = ~rev2(k.w0); return r; }
// KMer::toSuccessor (kmers/KMer.h:189-201)
SN_HD Kmer kmer_succ(const Kmer& k, uint32_t c)
{ Kmer r; r.w0 = (k.w0 << 2) | (k.w1 >> 30); r.w1 = (k.w1 << 2) | (k.w2 >> 30); r.w2 = (k.w2 << 2) | (c & 3u); return r; }
// KMer::toPredecessor (kmers/KMer.h:174-187)
SN_HD Kmer kmer_pred(const Kmer& k, uint32_t c)
{ Kmer r; r.w2 = (k.w2 >> 2) | (k.w1 << 30); r.w1 = (k.w1 >> 2) | (k.w0 << 30); r.w0 = (k.w0 >> 2) | ((c & 3u) << 30); return r; }
SN_HD uint32_t kmer_base(const Kmer& k, int i)
{ uint32_t w = i < 16 ? k.w0 : (i < 32 ? k.w1 : k.w2); return (w >> (2 * (15 - (i & 15)))) & 3u; }

// CF<48>::getForm (dna/CanonicalForm.h:57-67): the outside-in comparison of base i
// with the complement of base K-1-i is the lexicographic comparison k vs rc(k).
enum Form { FWD = 0, REV = 1, PAL = 2 };
SN_HD int kmer_form(const Kmer& k, Kmer* rc_out)
{
    Kmer r = kmer_rc(k);
    if (rc_out) *rc_out = r;
    if (k == r) return PAL;
    return k < r ? FWD : REV;
}
SN_HD bool kmer_is_pal(const Kmer& k) { Kmer r = kmer_rc(k); return k == r; }

// --- context ---------------------------------------------------------------
SN_HD uint32_t ctx_rc(uint32_t c) { return brev32(c) >> 24; }
SN_HD uint32_t ctx_pred(uint32_t c) { return (c >> 4) & 0xFu; }
SN_HD uint32_t ctx_succ(uint32_t c) { return c & 0xFu; }
SN_HD uint32_t low_bit_index(uint32_t m)       // m != 0
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)m) - 1u;
#else
    return (uint32_t)__builtin_ctz(m);
#endif
}
SN_HD bool mask_single(uint32_t m) { return m != 0 && (m & (m - 1)) == 0; }
SN_HD uint32_t mask_code(uint32_t m) { return (m >> 1) - (m >> 3); }   // 1,2,4,8 -> 0,1,2,3

// --- packed (fastb layout) base access ----------------------------------------
SN_HD uint32_t packed_base(const uint8_t* p, uint64_t i) { return (p[i >> 2] >> (2 * (i & 3))) & 3u; }

// K-mer starting at base `pos` of a fastb-packed sequence (byte pointer, any alignment).
SN_HD Kmer kmer_from_packed(const uint8_t* p, uint64_t pos)
{
    // gather 13 bytes = 104 bits, shift, then convert LSB-first fields to MSB-first words
    const uint8_t* q = p + (pos >> 2);
    uint32_t sh = 2 * (uint32_t)(pos & 3);
    uint64_t lo = 0, hi = 0;
    for (int i = 0; i < 8; ++i) lo |= (uint64_t)q[i] << (8 * i);
    for (int i = 0; i < 5; ++i) hi |= (uint64_t)q[8 + i] << (8 * i);
    if (sh) { lo = (lo >> sh) | (hi << (64 - sh)); hi >>= sh; }
    Kmer k;
    k.w0 = rev2((uint32_t)lo); k.w1 = rev2((uint32_t)(lo >> 32)); k.w2 = rev2((uint32_t)hi);
    return k;
}

// 16 bases starting at base `pos` of a fastb-packed sequence, base pos in bits 1..0 (the buffers are padded: reading a
// few bytes past the sequence is allowed).  Device: two aligned 32-bit loads + one funnel shift.
SN_HD uint32_t packed_window16(const uint8_t* p, uint64_t pos)
{
#if defined(__CUDA_ARCH__)
    const uint8_t* q = p + (pos >> 2);
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3u);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(q - mis);
    const uint32_t sh = 8u * mis + 2u * (uint32_t)(pos & 3u);          // <= 30
    return __funnelshift_r(w[0], w[1], sh);
#else
    const uint8_t* q = p + (pos >> 2);
    uint64_t v = 0;
    for (int i = 0; i < 5; ++i) v |= (uint64_t)q[i] << (8 * i);
    return (uint32_t)(v >> (2 * (pos & 3)));
#endif
}
SN_HD Kmer kmer_from_packed_w(const uint8_t* p, uint64_t pos)
{
    Kmer k;
    k.w0 = rev2(packed_window16(p, pos)); k.w1 = rev2(packed_window16(p, pos + 16)); k.w2 = rev2(packed_window16(p, pos + 32));
    return k;
}
// number of leading equal bases (<= m <= 16) of two 16-base windows
SN_HD uint32_t window_match(uint32_t a, uint32_t b, uint32_t m)
{
    uint32_t x = a ^ b;
    if (m < 16) x |= 1u << (2 * m);                                    // a sentinel difference right after the m-th base
    if (!x) return 16;
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffs((int)x) - 1) >> 1;
#else
    return (uint32_t)__builtin_ctz(x) >> 1;
#endif
}

// --- minimizers (MSP, lib/tada/src/msp/mod.rs) ---------------------------------------------------
// The minimizer of a k-mer is the smallest of its W = K-P+1 canonical p-mers under a hashed order;
// k-mer and reverse complement share it, so it names one bucket for every occurrence of a
// canonical k-mer (check_consistent_shard, lib/tada/src/kmer/mod.rs:1102-1150).
#define SN_P 16
#define SN_W (SN_K - SN_P + 1)          // p-mers per k-mer window

// order of the p-mers: a bijective mix of the canonical 16-mer (equal value <=> equal p-mer)
SN_HD uint32_t pmer_order(uint32_t canon)
{
    uint32_t m = canon * 0x9E3779B1u;
    m ^= m >> 15; m *= 0x85EBCA77u; m ^= m >> 13;
    return m;
}
// bucket hash of a minimizer: the minimum of W order values is small, so mix again
SN_HD uint32_t bucket_hash(uint32_t minval)
{
    uint32_t h = minval ^ 0x5bd1e995u;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// p-mers are kept MSB-first in 32 bits (P = 16): fwd = bases j..j+15, rc = its reverse complement
SN_HD uint32_t pmer_rc(uint32_t fwd) { return ~rev2(fwd); }
SN_HD uint32_t pmer_value(uint32_t fwd, uint32_t rc) { return pmer_order(fwd < rc ? fwd : rc); }
// minimum order value over the W p-mers of a k-mer
SN_HD uint32_t kmer_minimizer(const Kmer& k)
{
    uint32_t fwd = k.w0, rc = pmer_rc(k.w0);
    uint32_t m = pmer_value(fwd, rc);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = SN_P; j < SN_K; ++j) {
        const uint32_t b = kmer_base(k, j);
        fwd = (fwd << 2) | b; rc = (rc >> 2) | ((3u - b) << 30);
        const uint32_t v = pmer_value(fwd, rc);
        m = v < m ? v : m;
    }
    return m;
}
// the same, prepared for the 8 neighbours of k: a successor drops the first p-mer and gains one,
// a predecessor drops the last and gains one
struct KmerMin { uint32_t min_wo_first, min_wo_last, first_fwd, first_rc, last_fwd, last_rc; };
SN_HD KmerMin kmer_minimizer_nb(const Kmer& k)
{
    KmerMin r;
    uint32_t fwd = k.w0, rc = pmer_rc(k.w0);
    r.first_fwd = fwd; r.first_rc = rc;
    uint32_t v = pmer_value(fwd, rc);
    r.min_wo_last = v; r.min_wo_first = 0xFFFFFFFFu;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = SN_P; j < SN_K; ++j) {
        const uint32_t b = kmer_base(k, j);
        fwd = (fwd << 2) | b; rc = (rc >> 2) | ((3u - b) << 30);
        v = pmer_value(fwd, rc);
        r.min_wo_first = v < r.min_wo_first ? v : r.min_wo_first;
        if (j < SN_K - 1) r.min_wo_last = v < r.min_wo_last ? v : r.min_wo_last;
    }
    r.last_fwd = fwd; r.last_rc = rc;
    return r;
}
SN_HD uint32_t succ_minimizer(const KmerMin& m, uint32_t c)
{ const uint32_t v = pmer_value((m.last_fwd << 2) | c, (m.last_rc >> 2) | ((3u - c) << 30)); return v < m.min_wo_first ? v : m.min_wo_first; }
SN_HD uint32_t pred_minimizer(const KmerMin& m, uint32_t c)
{ const uint32_t v = pmer_value((m.first_fwd >> 2) | (c << 30), (m.first_rc << 2) | (3u - c)); return v < m.min_wo_last ? v : m.min_wo_last; }

// minimizer of a k-mer that slides along a sequence one base at a time (read threading): the
// full pass over the W p-mers is only repeated when the minimum leaves the window
struct MinState { uint32_t minval; int pos; uint32_t last_fwd, last_rc; };
SN_HD MinState min_state_init(const Kmer& k)
{
    MinState st;
    uint32_t fwd = k.w0, rc = pmer_rc(k.w0);
    st.minval = pmer_value(fwd, rc); st.pos = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = SN_P; j < SN_K; ++j) {
        const uint32_t b = kmer_base(k, j);
        fwd = (fwd << 2) | b; rc = (rc >> 2) | ((3u - b) << 30);
        const uint32_t v = pmer_value(fwd, rc);
        if (v <= st.minval) { st.minval = v; st.pos = j - (SN_P - 1); }
    }
    st.last_fwd = fwd; st.last_rc = rc;
    return st;
}
// k = the k-mer after the slide (already holding the new last base c)
SN_HD void min_state_slide(MinState& st, const Kmer& k, uint32_t c)
{
    st.last_fwd = (st.last_fwd << 2) | c; st.last_rc = (st.last_rc >> 2) | ((3u - c) << 30);
    const uint32_t v = pmer_value(st.last_fwd, st.last_rc);
    --st.pos;
    if (v <= st.minval) { st.minval = v; st.pos = SN_W - 1; }
    else if (st.pos < 0) st = min_state_init(k);
}

// --- dictionary ------------------------------------------------------------------
// 32-bit mix of the 96-bit k-mer: orders the k-mers inside a bucket and picks the slot in the
// shared-memory table of k_bucket_count.
SN_HD uint32_t kmer_hash(const Kmer& k)
{
    uint32_t h = k.w0 * 0x9E3779B1u;
    h = ((h ^ (h >> 15)) + k.w1) * 0x85EBCA77u;
    h = ((h ^ (h >> 13)) + k.w2) * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// Dictionary entry (32 B = one DRAM sector).  The valid-k-mer table doubles as the
// reference's KmerDict (kmers/ReadPather.h:222-388).  It is ordered by (minimizer bucket, hash,
// k-mer): the order k_bucket_count produces, and the one that keeps the graph stages local --
// neighbouring k-mers of a read or a unipath share their minimizer ~94% of the time, so a
// k-mer's neighbours, and the next k-mers of a walk, sit in the same few KB.  w0..w2 + cc
// (count:24 | ctx<<24, == the kmers.kvec KDef word, context BEFORE recomputeAdjacencies) and h are
// immutable after counting; ctx (after pruning), edge and off are filled by the graph stages.
struct __attribute__((aligned(32))) DictEntry {
    uint32_t w0, w1, w2, cc;
    uint32_t edge, off, ctx, h;
};
#define SN_NULL_EDGE 0xFFFFFFFFu

struct DictView {
    const DictEntry* tab;
    const uint32_t* boff;     // (b_n << sub_bits) + 1: first entry of every (minimizer bucket of the window, top sub_bits of the hash)
    uint32_t n;
    int bits, sub_bits;
    // A rank of the sharded multi-GPU path holds the buckets [b_lo, b_lo + b_n) of the 2^bits; one GPU holds all
    // (b_lo = 0, b_n = 2^bits).  k-mers of other buckets that are neighbours of local ones are GHOSTS: an
    // open-addressed table of g_cap (power of two, 0 = none) entries right behind the local ones, tab[n + slot].
    uint32_t b_lo, b_n;
    uint32_t g_cap;
    // the hashes of the n local entries on their own (hs[i] == tab[i].h), 4 bytes each: a lookup walks these -- the
    // ~32 hashes of a cell are one or two cache lines -- and touches a 32-byte entry only where the hash matches
    const uint32_t* hs;
};
// ghost entry (a DictEntry behind the table): w0..w2 = k-mer, h = its hash, cc = owner rank, edge = index in the
// owner's table, ctx = its context there AFTER recomputeAdjacencies, off = state
enum GhostState { GH_EMPTY = 0, GH_PENDING = 1, GH_PRESENT = 2, GH_ABSENT = 3, GH_WRITING = 4 };

// (h,k) < entry ?  /  == entry ?
SN_HD int dict_cmp(uint32_t h, const Kmer& k, const DictEntry& e)
{
    if (h != e.h) return h < e.h ? -1 : 1;
    if (k.w0 != e.w0) return k.w0 < e.w0 ? -1 : 1;
    if (k.w1 != e.w1) return k.w1 < e.w1 ? -1 : 1;
    if (k.w2 != e.w2) return k.w2 < e.w2 ? -1 : 1;
    return 0;
}
// lookup of a canonical k-mer whose minimizer is known.  Inside the bucket the entries are sorted
// by a uniform hash; the offsets are kept per (bucket, top sub_bits of the hash) -- ~32 entries --
// and the search starts where the remaining hash bits interpolate, then walks a step or two.
SN_HD uint32_t ghost_find(const DictView& d, uint32_t h, const Kmer& k)
{
    uint32_t s = h & (d.g_cap - 1u);
    for (uint32_t probes = 0; probes < d.g_cap; ++probes) {
        const DictEntry& e = d.tab[d.n + s];
        if (e.off == GH_EMPTY) return SN_NULL_EDGE;
        if (e.h == h && e.w0 == k.w0 && e.w1 == k.w1 && e.w2 == k.w2) return e.off == GH_PRESENT ? d.n + s : SN_NULL_EDGE;
        s = (s + 1u) & (d.g_cap - 1u);
    }
    return SN_NULL_EDGE;
}
SN_HD uint32_t dict_find_in_bucket(const DictView& d, uint32_t minimizer, const Kmer& k)
{
    const uint32_t h = kmer_hash(k);
    const uint32_t b = (bucket_hash(minimizer) >> (32 - d.bits)) - d.b_lo;
    if (b >= d.b_n) return d.g_cap ? ghost_find(d, h, k) : SN_NULL_EDGE;       // another rank's bucket
    const uint32_t cell = d.sub_bits ? ((b << d.sub_bits) | (h >> (32 - d.sub_bits))) : b;
    const uint32_t lo = d.boff[cell], hi = d.boff[cell + 1];
    if (lo >= hi) return SN_NULL_EDGE;
    const uint32_t rem = h << d.sub_bits;
    uint32_t i = lo + (uint32_t)(((uint64_t)rem * (hi - lo)) >> 32);      // < hi
    const uint32_t* H = d.hs;
    if (H[i] < h) { do ++i; while (i < hi && H[i] < h); }
    else while (i > lo && H[i - 1] >= h) --i;
    // i = first entry of the cell with hash >= h; equal hashes are ordered by k-mer
    for (; i < hi && H[i] == h; ++i) {
        const int c = dict_cmp(h, k, d.tab[i]);
        if (c == 0) return i;
        if (c < 0) break;
    }
    return SN_NULL_EDGE;
}
// KmerDict::findEntryCanonical : returns index or SN_NULL_EDGE
SN_HD uint32_t dict_find_canonical(const DictView& d, const Kmer& k) { return dict_find_in_bucket(d, kmer_minimizer(k), k); }
// the same with the minimizer already known (it is the same for k and its reverse complement)
SN_HD uint32_t dict_find_min(const DictView& d, const Kmer& k, uint32_t minimizer, bool* was_rc)
{
    Kmer r;
    int f = kmer_form(k, &r);
    if (was_rc) *was_rc = (f == REV);
    return dict_find_in_bucket(d, minimizer, f == REV ? r : k);
}
// KmerDict::findEntry (kmers/ReadPather.h:238-241): canonicalise, then look up.
// *was_rc tells whether the stored k-mer is the RC of the query.
SN_HD uint32_t dict_find(const DictView& d, const Kmer& k, bool* was_rc)
{
    Kmer r;
    int f = kmer_form(k, &r);
    if (was_rc) *was_rc = (f == REV);
    return dict_find_canonical(d, f == REV ? r : k);
}

}  // namespace sn
