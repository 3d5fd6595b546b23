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

// --- dictionary ------------------------------------------------------------------
// 32-bit mix of the 96-bit k-mer.  The k-mer stream is SORTED BY THIS HASH (4 radix digit
// passes instead of 12 for the full key); equal k-mers share a hash, so they land in one run of
// equal hashes and k_reduce separates the (rare) distinct k-mers that collide inside a run.
SN_HD uint32_t kmer_hash(const Kmer& k)
{
    uint32_t h = k.w0 * 0x9E3779B1u;
    h = ((h ^ (h >> 15)) + k.w1) * 0x85EBCA77u;
    h = ((h ^ (h >> 13)) + k.w2) * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// Dictionary entry (32 B = one DRAM sector).  The valid-k-mer table doubles as the
// reference's KmerDict (kmers/ReadPather.h:222-388).  It is ordered by (hash, k-mer); a
// prefix index over the top SN_IDX_BITS bits of the hash turns a lookup into one index load
// plus an interpolated probe (dict_find_canonical).  w0..w2 + cc (count:24 | ctx<<24, == the kmers.kvec KDef word,
// context BEFORE recomputeAdjacencies) and h are immutable after counting; ctx (after
// pruning), edge and off are filled by the graph stages.
struct __attribute__((aligned(32))) DictEntry {
    uint32_t w0, w1, w2, cc;
    uint32_t edge, off, ctx, h;
};
#define SN_NULL_EDGE 0xFFFFFFFFu
#define SN_IDX_BITS 24        // prefix index over the top bits of the hash

struct DictView {
    const DictEntry* tab;
    const uint32_t* idx;      // (1<<SN_IDX_BITS)+1 lower bounds by top bits of h
    uint32_t n;
};

// (h,k) < entry ?  /  == entry ?
SN_HD int dict_cmp(uint32_t h, const Kmer& k, const DictEntry& e)
{
    if (h != e.h) return h < e.h ? -1 : 1;
    if (k.w0 != e.w0) return k.w0 < e.w0 ? -1 : 1;
    if (k.w1 != e.w1) return k.w1 < e.w1 ? -1 : 1;
    if (k.w2 != e.w2) return k.w2 < e.w2 ? -1 : 1;
    return 0;
}
// KmerDict::findEntryCanonical : returns index or SN_NULL_EDGE.
// The table is sorted by a uniform hash, so inside the bucket of the top SN_IDX_BITS bits the
// position of h is close to where its remaining bits interpolate: the search starts there and
// walks a step or two.  A random DRAM access moves a whole 128-byte line (4 entries), and the
// small prefix index stays L2-resident, so a lookup costs little more than one line.
SN_HD uint32_t dict_find_canonical(const DictView& d, const Kmer& k)
{
    const uint32_t h = kmer_hash(k);
    const uint32_t b = h >> (32 - SN_IDX_BITS);
    const uint32_t lo = d.idx[b], hi = d.idx[b + 1];
    if (lo >= hi) return SN_NULL_EDGE;
    const uint32_t rem = h & ((1u << (32 - SN_IDX_BITS)) - 1u);
    uint32_t i = lo + (uint32_t)(((uint64_t)rem * (hi - lo)) >> (32 - SN_IDX_BITS));      // < hi
    if (d.tab[i].h < h) { do ++i; while (i < hi && d.tab[i].h < h); }
    else while (i > lo && d.tab[i - 1].h >= h) --i;
    // i = first entry of the bucket with hash >= h; equal hashes are ordered by k-mer
    for (; i < hi; ++i) {
        const int c = dict_cmp(h, k, d.tab[i]);
        if (c == 0) return i;
        if (c < 0) break;
    }
    return SN_NULL_EDGE;
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
