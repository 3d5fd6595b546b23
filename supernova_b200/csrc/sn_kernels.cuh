// sn_kernels.cuh -- the CUDA kernels of the hot path (sm_100a), in pipeline order:
//   a1  k_pqvec_goodlen / k_q8_goodlen   PQVec decode + good-length scan
//   a2  k_extract                         canonical (k,k+1)-mer records from 2-bit reads
//   a4  radix_sort_kmers (sn_prims.cuh)   128-bit records by 96-bit k-mer
//   a5  k_reduce                          run-length count / ctx OR / barcode rule / filter
//   a6  k_build_index, k_prune            dictionary prefix index, adjacency prune
//   a7  k_classify, k_walk_count, k_circle_count, k_walk_emit, k_fix_offsets, k_pack_edges
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
__global__ void __launch_bounds__(256) k_pqvec_goodlen(uint64_t n_reads, const uint8_t* __restrict__ pq, const uint64_t* __restrict__ pq_off,
                                                       const uint32_t* __restrict__ len, uint32_t min_qual,
                                                       uint32_t* __restrict__ goodlen, unsigned long long* occ_total, uint32_t* bad_reads)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t occ = 0;
    if (r < n_reads) {
        const uint8_t* p = pq + pq_off[r];
        const uint8_t* pend = pq + pq_off[r + 1];
        uint32_t L = len[r], i = 0, run = 0, gl = 0;
        while (p < pend) {
            uint32_t nq = *p++;
            if (!nq) break;
            uint32_t b0 = *p++;
            uint32_t nbits = b0 & 7u, minq = b0 >> 3;
            uint64_t acc = *p++;
            minq |= (uint32_t)(acc & 1u) << 5; acc >>= 1;
            uint32_t have = 7;
            uint32_t mask = (1u << nbits) - 1u;
            if (!nbits) {                                   // a block of nq equal quals
                if (minq >= min_qual) { run += nq; if (run >= SN_K) gl = i + nq; } else run = 0;
                i += nq;
            } else {
                for (uint32_t k = 0; k < nq; ++k) {
                    if (have < nbits) { acc |= (uint64_t)(*p++) << have; have += 8; }
                    uint32_t q = minq + ((uint32_t)acc & mask); acc >>= nbits; have -= nbits;
                    run = q >= min_qual ? run + 1 : 0;
                    ++i;
                    if (run >= SN_K) gl = i;
                }
            }
        }
        if (i != L) atomicAdd(bad_reads, 1u);      // PQVec length disagrees with the fastb length
        goodlen[r] = gl;
        occ = gl >= SN_K + 1 ? gl - SN_K + 1 : 0;
    }
    // block reduce, one atomic per block
    __shared__ uint32_t sm[8];
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_down_sync(SN_FULL, occ, o);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = occ;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 8; ++w) t += sm[w]; if (t) atomicAdd(occ_total, (unsigned long long)t); }
}

// same, for callers that already hold one u8 per base
__global__ void __launch_bounds__(256) k_q8_goodlen(uint64_t n_reads, const uint8_t* __restrict__ quals, const uint64_t* __restrict__ qoff,
                                                    const uint32_t* __restrict__ len, uint32_t min_qual,
                                                    uint32_t* __restrict__ goodlen, unsigned long long* occ_total)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t occ = 0;
    if (r < n_reads) {
        const uint8_t* q = quals + qoff[r];
        uint32_t L = len[r], run = 0, gl = 0;
        for (uint32_t i = 0; i < L; ++i) { run = q[i] >= min_qual ? run + 1 : 0; if (run >= SN_K) gl = i + 1; }
        goodlen[r] = gl;
        occ = gl >= SN_K + 1 ? gl - SN_K + 1 : 0;
    }
    __shared__ uint32_t sm[8];
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_down_sync(SN_FULL, occ, o);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = occ;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 8; ++w) t += sm[w]; if (t) atomicAdd(occ_total, (unsigned long long)t); }
}

// ---------------------------------------------------------------------------
// a2. Kmerizer::map (BuildReadQGraph48.cc:155-172).  A CTA stages the packed bases of
// EX_READS consecutive reads in shared memory with 16-byte loads, then its threads walk
// the tile's k-mer occurrences in order: occurrence x of the tile -> thread x % 256, so a
// warp writes 32 consecutive 16-byte records (512 B, fully coalesced).  The output
// range of the tile is reserved with one atomicAdd per CTA (record order does not
// matter: the sort follows and the reduction is order independent).
// record = {w0,w1,w2 of the canonical k-mer, ctx<<24 | bc24} ; bc24 = 0xFFFFFF for "-1".
// ---------------------------------------------------------------------------
#define SN_EX_READS 128
#define SN_EX_BYTES (SN_EX_READS * (SN_MAX_READ_LEN / 4) + 48)

__global__ void __launch_bounds__(256) k_extract(uint64_t n_reads, const uint8_t* __restrict__ bases, const uint64_t* __restrict__ boff,
                                                 const uint32_t* __restrict__ goodlen, const int32_t* __restrict__ bc,
                                                 int64_t ign_bc_below, uint4* __restrict__ out, unsigned long long* cursor)
{
    __shared__ __align__(16) uint8_t sb[SN_EX_BYTES];
    __shared__ uint32_t pref[SN_EX_READS + 1];
    __shared__ uint32_t s_gl[SN_EX_READS];
    __shared__ uint32_t s_rel[SN_EX_READS];
    __shared__ uint32_t s_bc[SN_EX_READS];
    __shared__ uint32_t wsum[8];
    __shared__ unsigned long long s_base;
    const uint32_t tid = threadIdx.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * SN_EX_READS;
    const uint32_t nr = (uint32_t)min((uint64_t)SN_EX_READS, n_reads - r0);
    const uint64_t lo = boff[r0], hi = boff[r0 + nr];
    const uint64_t lo_al = lo & ~15ull;
    const uint32_t shift = (uint32_t)(lo - lo_al);
    // stage packed bases (16-byte vector loads; the allocation is padded)
    {
        const uint4* src = reinterpret_cast<const uint4*>(bases + lo_al);
        uint4* dst = reinterpret_cast<uint4*>(sb);
        uint32_t nv = (uint32_t)((hi - lo_al + 13 + 15) >> 4);      // +13: kmer_from_packed may touch 13 bytes
        for (uint32_t i = tid; i < nv; i += 256) dst[i] = src[i];
    }
    // per-read counts and their exclusive prefix (128 reads: threads 0..127)
    uint32_t cnt = 0;
    if (tid < nr) {
        uint32_t gl = goodlen[r0 + tid];
        s_gl[tid] = gl;
        s_rel[tid] = (uint32_t)(boff[r0 + tid] - lo) + shift;
        int32_t b = -1;
        if (bc && (int64_t)(r0 + tid) >= ign_bc_below) b = bc[r0 + tid];
        s_bc[tid] = b < 0 ? 0xFFFFFFu : (uint32_t)b;
        cnt = gl >= SN_K + 1 ? gl - SN_K + 1 : 0;
    }
    {
        uint32_t x = cnt, lane = tid & 31u, w = tid >> 5;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, x, o); if (lane >= (uint32_t)o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (uint32_t k = 0; k < w; ++k) wb += wsum[k];
        if (tid < SN_EX_READS) pref[tid] = wb + x - cnt;
        if (tid == SN_EX_READS - 1) pref[SN_EX_READS] = wb + x;
    }
    __syncthreads();
    const uint32_t T = pref[SN_EX_READS];
    if (tid == 0) s_base = T ? atomicAdd(cursor, (unsigned long long)T) : 0ull;
    __syncthreads();
    const uint64_t base = s_base;
    for (uint32_t x = tid; x < T; x += 256) {
        // read of occurrence x: largest rr with pref[rr] <= x
        uint32_t a = 0, b = SN_EX_READS;
        while (b - a > 1) { uint32_t m = (a + b) >> 1; if (pref[m] <= x) a = m; else b = m; }
        const uint32_t rr = a, i = x - pref[rr], gl = s_gl[rr];
        const uint8_t* rp = sb + s_rel[rr];
        Kmer k = kmer_from_packed(rp, i);
        uint32_t ctx = 0;
        if (i > 0) ctx |= 16u << packed_base(rp, i - 1);
        if (i + SN_K < gl) ctx |= 1u << packed_base(rp, i + SN_K);
        Kmer rc;
        if (kmer_form(k, &rc) == REV) { k = rc; ctx = ctx_rc(ctx); }
        out[base + x] = make_uint4(k.w0, k.w1, k.w2, (ctx << 24) | s_bc[rr]);
    }
}

// ---------------------------------------------------------------------------
// a5. Kmerizer::reduce / summarizeEntries / areIgnoredBarcodes / areEnoughBarcodes
// (BuildReadQGraph48.cc:91-137,174-181) over the records sorted by kmer_hash, fused with
// the ordered compaction of the surviving k-mers into the dictionary (tile look-back).
// The first record of every run of EQUAL HASH owns the run.  Almost always the run is one
// k-mer: count (saturating 2^24-1), OR of contexts, min/max barcode > 0 (>= 2 distinct <=>
// min != max), "ignored" flag in one walk.  When distinct k-mers collide in a run the owner
// re-walks it once per k-mer in increasing k-mer order, so the dictionary comes out ordered
// by (hash, k-mer).
// ---------------------------------------------------------------------------
#define SN_RD_THREADS 256
#define SN_RD_ITEMS 8
#define SN_RD_TILE (SN_RD_THREADS * SN_RD_ITEMS)

__device__ __forceinline__ bool same_kmer(const uint4& a, const uint4& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
__device__ __forceinline__ bool kmer_lt(const uint4& a, const uint4& b)
{ return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); }

struct RunStat { uint32_t count, ctx, minbc, maxbc; bool ign; };
__device__ __forceinline__ void stat_init(RunStat& s) { s.count = 0; s.ctx = 0; s.minbc = 0xFFFFFFFFu; s.maxbc = 0; s.ign = false; }
__device__ __forceinline__ void stat_add(RunStat& s, uint32_t aux)
{
    ++s.count; s.ctx |= aux >> 24;
    uint32_t b = aux & 0xFFFFFFu;
    if (b == 0xFFFFFFu) s.ign = true;
    else if (b) { s.minbc = min(s.minbc, b); s.maxbc = max(s.maxbc, b); }
}
__device__ __forceinline__ bool stat_valid(const RunStat& s, uint32_t min_freq, uint32_t min_bc, int has_bc)
{
    bool enough = min_bc == 0 || (min_bc == 1 ? s.maxbc != 0 : (s.maxbc != 0 && s.minbc != s.maxbc));
    return s.count >= min_freq && (!has_bc || s.ign || enough);
}
__device__ __forceinline__ DictEntry make_entry(const uint4& k, const RunStat& s, uint32_t h)
{
    DictEntry e;
    e.w0 = k.x; e.w1 = k.y; e.w2 = k.z; e.cc = min(s.count, 0xFFFFFFu) | (s.ctx << 24);
    e.edge = SN_NULL_EDGE; e.off = 0; e.ctx = s.ctx; e.h = h;
    return e;
}
// Slow path: the run [p0,p1) of equal hash holds more than one k-mer.  Visits the distinct
// k-mers in increasing order; emits the valid ones at out[pos...] when out != nullptr.
// Returns the number of valid k-mers.
__device__ __noinline__ uint32_t reduce_mixed_run(const uint4* __restrict__ keys, uint64_t p0, uint64_t p1, uint32_t h,
                                                  uint32_t min_freq, uint32_t min_bc, int has_bc, DictEntry* out, uint64_t pos, uint32_t* n_distinct)
{
    uint32_t nvalid = 0, ndist = 0;
    uint4 cur = keys[p0];
    for (uint64_t p = p0 + 1; p < p1; ++p) { uint4 r = keys[p]; if (kmer_lt(r, cur)) cur = r; }     // smallest k-mer
    for (;;) {
        RunStat st; stat_init(st);
        bool have_next = false; uint4 next = cur;
        for (uint64_t p = p0; p < p1; ++p) {
            uint4 r = keys[p];
            if (same_kmer(r, cur)) stat_add(st, r.w);
            else if (kmer_lt(cur, r) && (!have_next || kmer_lt(r, next))) { next = r; have_next = true; }
        }
        ++ndist;
        if (stat_valid(st, min_freq, min_bc, has_bc)) { if (out) out[pos + nvalid] = make_entry(cur, st, h); ++nvalid; }
        if (!have_next) break;
        cur = next;
    }
    if (n_distinct) *n_distinct = ndist;
    return nvalid;
}

// Aggregate of a (partial) run, combined with shuffles.
struct Agg { uint32_t count, ctx_flags, minbc, maxbc; };      // ctx_flags: ctx | ign<<8 | mixed<<9
__device__ __forceinline__ Agg agg_combine(const Agg& a, const Agg& b)
{ Agg r; r.count = a.count + b.count; r.ctx_flags = a.ctx_flags | b.ctx_flags; r.minbc = min(a.minbc, b.minbc); r.maxbc = max(a.maxbc, b.maxbc); return r; }
__device__ __forceinline__ Agg agg_shfl_up(const Agg& a, int d)
{ Agg r; r.count = __shfl_up_sync(SN_FULL, a.count, d); r.ctx_flags = __shfl_up_sync(SN_FULL, a.ctx_flags, d);
  r.minbc = __shfl_up_sync(SN_FULL, a.minbc, d); r.maxbc = __shfl_up_sync(SN_FULL, a.maxbc, d); return r; }
__device__ __forceinline__ Agg agg_bcast(const Agg& a, int src)
{ Agg r; r.count = __shfl_sync(SN_FULL, a.count, src); r.ctx_flags = __shfl_sync(SN_FULL, a.ctx_flags, src);
  r.minbc = __shfl_sync(SN_FULL, a.minbc, src); r.maxbc = __shfl_sync(SN_FULL, a.maxbc, src); return r; }
__device__ __forceinline__ bool agg_valid(const Agg& s, uint32_t min_freq, uint32_t min_bc, int has_bc)
{
    bool enough = min_bc == 0 || (min_bc == 1 ? s.maxbc != 0 : (s.maxbc != 0 && s.minbc != s.maxbc));
    return s.count >= min_freq && (!has_bc || ((s.ctx_flags >> 8) & 1u) || enough);
}

// Warp-streaming reduce-by-key ("warp-ballot run-length counting").  Each warp owns the runs of
// equal hash whose FIRST record lies in its chunk of SN_RD_CHUNK records; it streams 32 records
// per step (one coalesced 512-byte load), finds run heads with a ballot, reduces every run with
// a segmented shuffle scan, carries the open run across steps (and past the end of the chunk),
// and writes the surviving k-mers, in order, to its private staging slots; per-warp counts are
// scanned and k_reduce_gather compacts the stage into the dictionary.
#define SN_RD_CHUNK 1024
#define SN_RD_WARPS 8
template <int MINB>
__global__ void __launch_bounds__(SN_RD_WARPS * 32, MINB) k_reduce(const uint4* __restrict__ keys, uint32_t n, uint32_t min_freq, uint32_t min_bc, int has_bc,
                                                             DictEntry* __restrict__ stage, uint32_t cap_per_warp, uint32_t* __restrict__ warp_count,
                                                             unsigned long long* n_distinct, uint32_t* overflow)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t w = (uint64_t)blockIdx.x * SN_RD_WARPS + (threadIdx.x >> 5);
    const uint64_t a = w * SN_RD_CHUNK;
    if (a >= n) return;
    const uint64_t b = min((uint64_t)n, a + SN_RD_CHUNK);
    DictEntry* out = stage + w * cap_per_warp;
    const uint4 SENT = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);   // never a canonical k-mer
    uint4 prev_last = a > 0 ? keys[a - 1] : SENT;
    uint32_t prev_h = rs_hash(prev_last);
    bool started = false;          // an owned run has begun
    bool open = false;             // carry holds an owned, unfinished run
    Agg carry; carry.count = 0; carry.ctx_flags = 0; carry.minbc = 0xFFFFFFFFu; carry.maxbc = 0;
    uint4 carry_key = SENT; uint64_t carry_start = 0; uint32_t carry_h = 0;
    uint32_t cursor = 0, distinct = 0;
    // software pipeline: the records of the next two steps are already in flight
    uint4 nx1 = (a + lane < n) ? keys[a + lane] : SENT;
    uint4 nx2 = (a + 32 + lane < n) ? keys[a + 32 + lane] : SENT;
    for (uint64_t pos = a;; pos += 32) {
        const uint64_t idx = pos + lane;
        const bool inb = idx < n;
        uint4 r = nx1;
        nx1 = nx2;
        nx2 = (idx + 64 < n) ? keys[idx + 64] : SENT;
        uint4 pv;
        pv.x = __shfl_up_sync(SN_FULL, r.x, 1); pv.y = __shfl_up_sync(SN_FULL, r.y, 1); pv.z = __shfl_up_sync(SN_FULL, r.z, 1); pv.w = 0;
        if (lane == 0) pv = prev_last;
        const uint32_t h = rs_hash(r);                                  // one hash per record, shared by every later use
        uint32_t ph = __shfl_up_sync(SN_FULL, h, 1);
        if (lane == 0) ph = prev_h;
        const bool kh = !same_kmer(pv, r);                              // first record of its k-mer
        bool hh = kh && (idx == 0 || !inb || ph != h);                  // first record of its hash run
        if (idx > n) hh = false;                                       // only the first sentinel closes the last run
        const uint32_t hmask = __ballot_sync(SN_FULL, hh);
        // ownership window of this step: from the first owned head on (skip the tail of a
        // foreign run at the start of the chunk), up to the first head at or past `b`
        const uint32_t fmask = __ballot_sync(SN_FULL, hh && idx >= b);
        const uint32_t f = fmask ? (uint32_t)__ffs(fmask) - 1u : 32u;   // lanes >= f belong to the next warp
        uint32_t first = 0;
        if (!started) { if (!hmask) { prev_last.x = __shfl_sync(SN_FULL, r.x, 31); prev_last.y = __shfl_sync(SN_FULL, r.y, 31); prev_last.z = __shfl_sync(SN_FULL, r.z, 31);
                                      prev_h = __shfl_sync(SN_FULL, h, 31);
                                      if (pos + 32 >= n + 1) break; continue; }
                        first = (uint32_t)__ffs(hmask) - 1u; }
        const bool active = lane >= first && lane < f;
        Agg v; v.count = active ? 1u : 0u;
        v.ctx_flags = active ? ((r.w >> 24) | ((r.w & 0xFFFFFFu) == 0xFFFFFFu ? 0x100u : 0u) | ((kh && !hh) ? 0x200u : 0u)) : 0u;
        { uint32_t bcv = r.w & 0xFFFFFFu; bool pos_bc = active && bcv != 0 && bcv != 0xFFFFFFu;
          v.minbc = pos_bc ? bcv : 0xFFFFFFFFu; v.maxbc = pos_bc ? bcv : 0u; }
        // segmented inclusive scan; segments start at run heads
        const uint32_t seg_heads = hmask & ((2u << lane) - 1u);          // heads at or before this lane
        const int seg_start = seg_heads ? 31 - __clz(seg_heads) : -1;     // -1: the run continues from the carry
        const uint32_t dist = (uint32_t)((int)lane - (seg_start < 0 ? 0 : seg_start));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { Agg u = agg_shfl_up(v, d); if ((uint32_t)d <= dist) v = agg_combine(u, v); }
        if (seg_start < 0 && open) v = agg_combine(carry, v);
        // a run ends at lane l when lane l+1 is a head (or the ownership window closes there)
        const uint32_t nextmask = (hmask >> 1) | (f < 32 && f > 0 ? (1u << (f - 1)) : 0u);
        const bool ends_here = active && lane < 31 && ((nextmask >> lane) & 1u);
        // the carried run ends when lane 0 is a head (or is already foreign)
        const bool carry_ends = open && (((hmask | (f == 0 ? 1u : 0u)) & 1u) != 0);
        // --- emission, in order: carried run first, then the runs that end inside this step ---
        if (carry_ends) {
            bool mixed = (carry.ctx_flags >> 9) & 1u;
            if (!mixed) {
                ++distinct;
                if (agg_valid(carry, min_freq, min_bc, has_bc)) {
                    if (cursor < cap_per_warp) { if (lane == 0) { RunStat st; st.count = carry.count; st.ctx = carry.ctx_flags & 0xFFu; out[cursor] = make_entry(carry_key, st, carry_h); } }
                    else if (lane == 0) atomicAdd(overflow, 1u);
                    ++cursor;
                }
            } else {
                uint32_t nd = 0, nvld = 0;
                if (lane == 0) {
                    uint32_t room = cursor < cap_per_warp ? cap_per_warp - cursor : 0;
                    nvld = reduce_mixed_run(keys, carry_start, pos, carry_h, min_freq, min_bc, has_bc, nullptr, 0, &nd);
                    if (nvld <= room) reduce_mixed_run(keys, carry_start, pos, carry_h, min_freq, min_bc, has_bc, out, cursor, nullptr);
                    else atomicAdd(overflow, 1u);
                }
                cursor += __shfl_sync(SN_FULL, nvld, 0); distinct += __shfl_sync(SN_FULL, nd, 0);
            }
            open = false;
        }
        {
            const bool mixed = (v.ctx_flags >> 9) & 1u;
            const bool emit_ok = ends_here && !mixed && agg_valid(v, min_freq, min_bc, has_bc);
            const uint32_t emask = __ballot_sync(SN_FULL, emit_ok);
            const uint32_t mmask = __ballot_sync(SN_FULL, ends_here && mixed);
            distinct += __popc(__ballot_sync(SN_FULL, ends_here && !mixed));
            if (!mmask) {
                if (emit_ok) {
                    uint32_t p = cursor + __popc(emask & lanemask_lt());
                    if (p < cap_per_warp) { RunStat st; st.count = v.count; st.ctx = v.ctx_flags & 0xFFu; out[p] = make_entry(r, st, h); }
                    else atomicAdd(overflow, 1u);
                }
                cursor += __popc(emask);
            } else {
                // rare: a run with colliding k-mers ends in this step -> go through the ending lanes one by one
                uint32_t todo = emask | mmask;
                while (todo) {
                    const uint32_t l = (uint32_t)__ffs(todo) - 1u; todo &= todo - 1u;
                    uint32_t nvld = 0, nd = 0;
                    if (lane == l) {
                        uint32_t room = cursor < cap_per_warp ? cap_per_warp - cursor : 0;
                        if (!mixed) { nvld = 1; if (room) { RunStat st; st.count = v.count; st.ctx = v.ctx_flags & 0xFFu; out[cursor] = make_entry(r, st, h); } else atomicAdd(overflow, 1u); }
                        else {
                            const uint64_t start = seg_start < 0 ? carry_start : pos + (uint32_t)seg_start;
                            nvld = reduce_mixed_run(keys, start, idx + 1, h, min_freq, min_bc, has_bc, nullptr, 0, &nd);
                            if (nvld <= room) reduce_mixed_run(keys, start, idx + 1, h, min_freq, min_bc, has_bc, out, cursor, nullptr);
                            else atomicAdd(overflow, 1u);
                        }
                    }
                    cursor += __shfl_sync(SN_FULL, nvld, l); distinct += __shfl_sync(SN_FULL, nd, l);
                }
            }
        }
        started = true;
        if (f < 32) break;                                              // the next warp's first run starts here
        // the run open at lane 31 is carried into the next step
        {
            const uint32_t last_head = hmask ? 31u - (uint32_t)__clz(hmask) : 32u;
            const bool had_open = open;                                   // still true only if the carried run did not end
            Agg c31 = agg_bcast(v, 31);
            if (last_head < 32) {
                carry = c31; open = true;
                carry_key.x = __shfl_sync(SN_FULL, r.x, last_head); carry_key.y = __shfl_sync(SN_FULL, r.y, last_head); carry_key.z = __shfl_sync(SN_FULL, r.z, last_head);
                carry_h = __shfl_sync(SN_FULL, h, last_head);
                carry_start = pos + last_head;
            } else if (had_open) carry = c31;                            // the carried run swallowed the whole step
        }
        prev_last.x = __shfl_sync(SN_FULL, r.x, 31); prev_last.y = __shfl_sync(SN_FULL, r.y, 31); prev_last.z = __shfl_sync(SN_FULL, r.z, 31);
        prev_h = __shfl_sync(SN_FULL, h, 31);
        if (pos + 32 >= (uint64_t)n + 1) break;                           // the sentinel lane has been processed
    }
    if (lane == 0) { warp_count[w] = cursor; if (distinct) atomicAdd(n_distinct, (unsigned long long)distinct); }
}
// compacts the per-warp staging slots into the dictionary: one warp per source warp
__global__ void __launch_bounds__(256) k_reduce_gather(const DictEntry* __restrict__ stage, uint32_t cap_per_warp, const uint32_t* __restrict__ warp_count,
                                                       const uint64_t* __restrict__ warp_off, uint64_t n_warps, DictEntry* __restrict__ dict)
{
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_warps) return;
    const uint32_t lane = threadIdx.x & 31u, cnt = warp_count[w];
    const uint4* src = reinterpret_cast<const uint4*>(stage + w * cap_per_warp);
    uint4* dst = reinterpret_cast<uint4*>(dict + warp_off[w]);
    for (uint32_t i = lane; i < 2 * cnt; i += 32) dst[i] = src[i];
}

// ---------------------------------------------------------------------------
// a6. dictionary prefix index + recomputeAdjacencies
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_index(const DictEntry* __restrict__ tab, uint32_t n, uint32_t* __restrict__ idx)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > (1u << SN_IDX_BITS)) return;
    if (b == (1u << SN_IDX_BITS)) { idx[b] = n; return; }
    uint32_t key = b << (32 - SN_IDX_BITS);
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (tab[mid].h < key) lo = mid + 1; else hi = mid; }
    idx[b] = lo;
}
__global__ void __launch_bounds__(256) k_prune(DictEntry* tab, const uint32_t* __restrict__ idx, uint32_t n, Link2* __restrict__ cand)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DictView d; d.tab = tab; d.idx = idx; d.n = n;
    Link2 l;
    tab[i].ctx = prune_ctx(d, i, &l);
    cand[i] = l;
}

// ---------------------------------------------------------------------------
// a7. unipath edges.  own_n[i] = number of k-mers of the edge that entry i owns
// (0 = owns none).  Singles own themselves; an edge with >= 2 k-mers is walked from
// both of its end entries and owned by the end with the smaller index; a circle is
// owned by its smallest entry.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_classify(const DictEntry* __restrict__ tab, const uint32_t* __restrict__ idx, uint32_t n,
                                                  Link2* links, uint8_t* __restrict__ etype, uint32_t* __restrict__ own_n, uint32_t* __restrict__ is_end)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DictView d; d.tab = tab; d.idx = idx; d.n = n;
    Link2 l;
    int t = classify_links(d, i, links[i], &l);          // links[] holds prune's candidates on entry
    links[i] = l;
    etype[i] = (uint8_t)t;
    own_n[i] = t == T_SINGLE ? 1u : 0u;
    is_end[i] = (t == T_END_DOWN || t == T_END_UP) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_scatter_flagged(const uint32_t* __restrict__ flag, const uint64_t* __restrict__ pos, uint32_t n, uint32_t* __restrict__ list)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) list[pos[i]] = i;
}
__global__ void __launch_bounds__(128) k_walk_count(const Link2* __restrict__ links, const uint32_t* __restrict__ ends, uint32_t n_ends,
                                                    const uint8_t* __restrict__ etype, uint32_t* __restrict__ own_n, uint8_t* __restrict__ visited)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ends) return;
    uint32_t i = ends[t], last = i;
    visited[i] = 1;
    uint32_t nk = walk_links(links, i, etype[i] == T_END_UP ? 1u : 0u, [&](uint32_t, uint32_t j, uint32_t) { visited[j] = 1; last = j; });
    if (i <= last) own_n[i] = nk;       // the other end walks the same edge; the smaller index owns it
}
__global__ void __launch_bounds__(128) k_circle_count(const DictEntry* __restrict__ tab, const Link2* __restrict__ links, uint32_t n,
                                                     uint8_t* __restrict__ etype, const uint8_t* __restrict__ visited, uint32_t* __restrict__ own_n,
                                                     uint32_t* n_circle_members)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (etype[i] != T_INTERIOR || visited[i]) return;
    atomicAdd(n_circle_members, 1u);
    // the walker with the smallest table index completes the loop; the circle is then owned by
    // its smallest K-MER, where canonicalizeCircle (BuildReadQGraph48.cc:375-397) starts it
    uint32_t m = i; Kmer mk = entry_kmer(tab[i]);
    uint32_t nk = walk_circle_links(links, i, true, [&](uint32_t, uint32_t j, uint32_t) { Kmer q = entry_kmer(tab[j]); if (q < mk) { mk = q; m = j; } });
    if (nk) { own_n[m] = nk; etype[m] = T_CIRCLE; }
}
__global__ void __launch_bounds__(256) k_edge_sizes(const uint32_t* __restrict__ own_n, uint32_t n, uint32_t* __restrict__ ebases, uint32_t* __restrict__ eflag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = own_n[i];
    ebases[i] = k ? k + (SN_K - 1) : 0u;
    eflag[i] = k ? 1u : 0u;
}
// owner walks its edge again: bases (one per byte, walk orientation) into `tmp`, and
// (edge id, step) into every entry on the edge; then the whole-edge canonical form
// (EdgeBuilder::addEdge :480-485 / extend :457-464) decides whether the edge is
// stored reverse-complemented.
__global__ void __launch_bounds__(128) k_walk_emit(DictEntry* tab, const Link2* __restrict__ links,
                                                   const uint32_t* __restrict__ owners, uint32_t n_owners, const uint8_t* __restrict__ etype,
                                                   const uint64_t* __restrict__ base_off, uint8_t* __restrict__ tmp,
                                                   uint32_t* __restrict__ elen, uint8_t* __restrict__ eflip, uint64_t* __restrict__ etmp_off)
{
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_owners) return;
    uint32_t i = owners[e];
    int t = etype[i];
    uint8_t* s = tmp + base_off[i];
    Kmer k = entry_kmer(tab[i]);
    if (t == T_END_UP) k = kmer_rc(k);
    for (int b = 0; b < SN_K; ++b) s[b] = (uint8_t)kmer_base(k, b);
    tab[i].edge = e; tab[i].off = 0;
    uint32_t nk = 1;
    auto visit = [&](uint32_t step, uint32_t j, uint32_t o) { s[SN_K - 1 + step] = (uint8_t)step_base(tab[j], o); tab[j].edge = e; tab[j].off = step; };
    if (t == T_END_DOWN || t == T_END_UP) nk = walk_links(links, i, t == T_END_UP ? 1u : 0u, visit);
    else if (t == T_CIRCLE) nk = walk_circle_links(links, i, false, visit);
    uint32_t len = nk + SN_K - 1;
    elen[e] = len;
    etmp_off[e] = base_off[i];
    eflip[e] = seq_form_u8(s, len) == REV ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_fix_offsets(DictEntry* tab, uint32_t n, const uint32_t* __restrict__ elen, const uint8_t* __restrict__ eflip)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t e = tab[i].edge;
    if (e != SN_NULL_EDGE && eflip[e]) tab[i].off = (elen[e] - (SN_K - 1)) - 1 - tab[i].off;
}
__global__ void __launch_bounds__(256) k_edge_bytes(const uint32_t* __restrict__ elen, uint32_t n_edges, uint32_t* __restrict__ ebytes)
{
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) ebytes[e] = (elen[e] + 3) >> 2;
}
// one thread per output byte of the packed edge store (fastb layout)
__global__ void __launch_bounds__(256) k_pack_edges(const uint8_t* __restrict__ tmp, const uint64_t* __restrict__ etmp_off,
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
__global__ void __launch_bounds__(128) k_hbv_keys(const uint8_t* __restrict__ ebases, const uint64_t* __restrict__ eoff, const uint32_t* __restrict__ elen,
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
__global__ void __launch_bounds__(256) k_hbv_mark(const uint4* __restrict__ rec, uint32_t n, uint32_t* __restrict__ flag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = rec[i];
    bool valid = r.w != 0xFFFFFFFFu;
    flag[i] = (valid && (i == 0 || !same_kmer(rec[i - 1], r))) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_hbv_assign(const uint4* __restrict__ rec, uint32_t n, const uint32_t* __restrict__ flag, const uint64_t* __restrict__ pos,
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
__device__ __forceinline__ void thread_path(const PathInputs& in, uint64_t r, const DictView& d, const EdgeStore& es, const HbvView& h,
                                            Part* parts, RPath& path, uint8_t* qbuf)
{
    const uint8_t* q;
    if (in.pq) { pqvec_decode(in.pq + in.pq_off[r], in.pq + in.pq_off[r + 1], qbuf, SN_MAX_READ_LEN); q = qbuf; }
    else q = in.quals + in.qoff[r];
    path_one_read(d, es, h, in.bases + in.boff[r], q, in.len[r], parts, path);
}
__global__ void __launch_bounds__(128) k_path_reads(PathInputs in, DictView d, EdgeStore es, HbvView h,
                                                    uint32_t* __restrict__ plen, int32_t* __restrict__ poffset, int32_t* __restrict__ scratch, uint32_t* overflow)
{
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= in.n_reads) return;
    Part parts[SN_MAX_PARTS];
    RPath path;
    uint8_t qbuf[SN_MAX_READ_LEN];
    thread_path(in, r, d, es, h, parts, path, qbuf);
    if (path.overflow) atomicAdd(overflow, 1u);
    plen[r] = path.n; poffset[r] = path.offset;
    int4 v = make_int4(path.n > 0 ? path.e[0] : 0, path.n > 1 ? path.e[1] : 0, path.n > 2 ? path.e[2] : 0, path.n > 3 ? path.e[3] : 0);
    reinterpret_cast<int4*>(scratch)[r] = v;
}
__global__ void __launch_bounds__(128) k_path_finish(PathInputs in, DictView d, EdgeStore es, HbvView h,
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

}  // namespace sn
