// sn_msp.cuh -- minimizer sharding ("MSP") of the k-mer stream and the per-bucket k-mer count.
//
// What it replaces in the reference:
//   * tada MSP: `simple_scan` / `compute_pvals` / `Bsp::new` (lib/tada/src/msp/mod.rs:17-134,
//     lib/tada/src/cmd_msp.rs:129-146): reads are cut into super-k-mers, maximal runs of
//     consecutive k-mers sharing their minimizer, and every super-k-mer goes to the shard its
//     minimizer names.  Every occurrence of a canonical k-mer carries the same minimizer
//     (`check_consistent_shard`, lib/tada/src/kmer/mod.rs:1102-1150), so a shard can be counted
//     on its own (`process_kmer_shard_opt`, lib/tada/src/utils.rs:322-408).
//   * C++ createDict: Kmerizer::map / MapReduceEngine::run / Kmerizer::reduce
//     (paths/long/BuildReadQGraph48.cc:91-137,155-181; MapReduceEngine.h:417-584): the k-mer
//     occurrences the super-k-mers expand to are exactly the records Kmerizer::map emits
//     (bases [0,goodLen) only, reads with goodLen < K+1 dropped, contexts never look past
//     goodLen-1), and a bucket is reduced with Kmerizer::reduce's rule.
//
// The on-device format is private (".msp ... are private and may change").
// Differences from tada's: P = 16 (not 8) with the order of the p-mers given by a 32-bit mix of
// the canonical p-mer instead of a frequency permutation -- 2^31 possible minimizers spread
// evenly over up to 2^24 buckets, where 4^8 p-mers could not -- and fixed 32-byte records.
//
//   SkRec (32 B):  w0 = bc24 | (nk-1) << 24 | hasL << 30 | hasR << 31
//                  w1 = bucket hash (mix of the minimizer; bucket = w1 >> (32 - bits))
//                  w2..w7 = bases, fastb packing (base j at bits 2*(j%16) of word j/16):
//                           [left neighbour if hasL] the nk+K-1 bases of the run [right neighbour if hasR]
//   A run is cut after 47 k-mers, so nk+K-1+2 <= 96 bases always fit.
//
// HBM traffic: the reads are scanned twice (histogram, then scatter: 0.25 B/base each), the
// super-k-mers are written once and read once (~0.38 B per k-mer occurrence instead of the
// 16 B records x 4 sort passes of a key sort); the k-mers themselves only ever exist in
// shared memory, in the per-bucket hash table of k_bucket_count.
#pragma once
#include "sn_kmer.cuh"

namespace sn {

#define SN_SK_WORDS 8

// Cuts the good part [0,gl) of one read into super-k-mers: maximal runs of consecutive k-mers
// with the same minimizer VALUE (pmer_order is a bijection, so that is the same canonical
// p-mer), cut after SN_SK_MAXK k-mers.  emit(start, nk, minval) is called for every run, in
// read order; k-mers [start, start+nk).
// The minimum over the W p-mers of each k-mer window is a sliding-window minimum computed
// without data-dependent branches (van Herk / Gil-Werman): p-mers are taken in blocks of W; a
// window is the suffix of one block plus the prefix of the next.  `ring` (W words, element t at
// ring[t * ring_stride]) holds the suffix minima of the last complete block in the slots not
// yet overwritten by the order values of the current block; when a block completes one
// backward sweep turns it into suffix minima in place.  All threads of a warp run the same
// iterations (thread per read), so the kernel does not diverge on the minimizer logic.
#define SN_SK_MAXK 47
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class F>
SN_HD void msp_scan(const uint8_t* rp, uint32_t gl, uint32_t* ring, uint32_t ring_stride, F&& emit, uint32_t min_gl = SN_K + 1)
{
    if (gl < min_gl) return;                         // (min_gl = K + 1: BuildReadQGraph48.cc:160; K in the tada variant, cmd_msp.rs:109-110)
    uint32_t fwd = 0, rc = 0, byte = 0;
    uint32_t r = 0;                                  // index of the current p-mer inside its block
    uint32_t pm = 0;                                 // prefix minimum of the current block
    uint32_t sk_start = 0, sk_min = 0;
    for (uint32_t j = 0; j < gl; ++j) {
        if ((j & 3u) == 0) byte = rp[j >> 2];
        const uint32_t b = (byte >> (2 * (j & 3u))) & 3u;
        fwd = (fwd << 2) | b;
        rc = (rc >> 2) | ((3u - b) << 30);
        if (j < SN_P - 1) continue;
        const uint32_t m = pmer_order(fwd < rc ? fwd : rc);
        pm = (r == 0 || m < pm) ? m : pm;
        uint32_t minval = pm;
        if (r != SN_W - 1) { const uint32_t sfx = ring[(r + 1) * ring_stride]; minval = sfx < pm ? sfx : pm; }
        ring[r * ring_stride] = m;
        if (r == SN_W - 1) {                         // block complete: suffix minima in place
            uint32_t sm = m;
            for (int t = SN_W - 2; t >= 0; --t) { const uint32_t v = ring[t * ring_stride]; sm = v < sm ? v : sm; ring[t * ring_stride] = sm; }
            r = 0;
        } else ++r;
        if (j < SN_K - 1) continue;                   // the first window completes with p-mer W-1 (base K-1)
        const uint32_t i = j - (SN_K - 1);
        if (i == 0) { sk_start = 0; sk_min = minval; }
        else if (minval != sk_min || i - sk_start >= SN_SK_MAXK) {
            emit(sk_start, i - sk_start, sk_min);
            sk_start = i; sk_min = minval;
        }
    }
    emit(sk_start, gl - SN_K + 1 - sk_start, sk_min);
}

// Builds the record of the run [start, start+nk) of a read with good length gl.
// `rp` may have any byte alignment; up to 28 bytes past the last needed base may be read.
SN_HD void sk_build(const uint8_t* rp, uint32_t gl, uint32_t start, uint32_t nk, uint32_t bc24, uint32_t bhash, uint32_t* w /*8*/)
{
    const uint32_t hasL = start > 0 ? 1u : 0u;
    const uint32_t hasR = start + nk + SN_K - 1 < gl ? 1u : 0u;
    const uint32_t s0 = start - hasL, nb = nk + SN_K - 1 + hasL + hasR;
    w[0] = bc24 | ((nk - 1) << 24) | (hasL << 30) | (hasR << 31);
    w[1] = bhash;
    const uint8_t* q = rp + (s0 >> 2);
    const uint32_t sh = 2 * (s0 & 3);
    uint32_t x[7];
    for (int i = 0; i < 7; ++i) {
        const uint8_t* p = q + 4 * i;
        x[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    for (int i = 0; i < 6; ++i) {
        uint32_t v = sh ? (x[i] >> sh) | (x[i + 1] << (32 - sh)) : x[i];
        const uint32_t first = 16u * i;                // first base of this word
        if (first >= nb) v = 0;
        else if (nb - first < 16) v &= (1u << (2 * (nb - first))) - 1u;
        w[2 + i] = v;
    }
}

SN_HD uint32_t sk_nk(uint32_t w0) { return ((w0 >> 24) & 0x3Fu) + 1u; }
#if defined(__CUDACC__)
// sk_build over a staged block of reads: `sb` is the 16-byte aligned staging buffer, `byte_off` the read's
// first byte in it.  Seven aligned 32-bit shared-memory loads and six funnel shifts instead of 28 byte loads
// (the byte loads made the scatter pass wait on the memory-instruction queue).
__device__ __forceinline__ void sk_build_staged(const uint8_t* sb, uint32_t byte_off, uint32_t gl, uint32_t start, uint32_t nk, uint32_t bc24, uint32_t bhash, uint32_t* w /*8*/)
{
    const uint32_t hasL = start > 0 ? 1u : 0u;
    const uint32_t hasR = start + nk + SN_K - 1 < gl ? 1u : 0u;
    const uint32_t s0 = start - hasL, nb = nk + SN_K - 1 + hasL + hasR;
    w[0] = bc24 | ((nk - 1) << 24) | (hasL << 30) | (hasR << 31);
    w[1] = bhash;
    const uint32_t a = byte_off + (s0 >> 2);
    const uint32_t* W = reinterpret_cast<const uint32_t*>(sb) + (a >> 2);
    const uint32_t S = 8u * (a & 3u) + 2u * (s0 & 3u);          // < 32
    uint32_t x[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = W[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        uint32_t v = __funnelshift_r(x[i], x[i + 1], S);
        const uint32_t first = 16u * i;
        if (first >= nb) v = 0;
        else if (nb - first < 16) v &= (1u << (2 * (nb - first))) - 1u;
        w[2 + i] = v;
    }
}
#endif

// Occurrence i (0 <= i < nk) of a record: canonical k-mer + context as Kmerizer::map emits them
// (BuildReadQGraph48.cc:155-172).  `w` = the 8 words of the record.
SN_HD void sk_occurrence(const uint32_t* w, uint32_t i, Kmer* out, uint32_t* ctx_out)
{
    const uint32_t w0 = w[0];
    const uint32_t hasL = (w0 >> 30) & 1u, hasR = w0 >> 31, nk = sk_nk(w0);
    const uint32_t p = hasL + i;                       // first base of the k-mer
    const uint32_t q = p >> 4, sh = 2 * (p & 15);
    const uint32_t* b = w + 2;
    uint32_t l0, l1, l2;
    if (sh) { l0 = (b[q] >> sh) | (b[q + 1] << (32 - sh)); l1 = (b[q + 1] >> sh) | (b[q + 2] << (32 - sh)); l2 = (b[q + 2] >> sh) | (b[q + 3] << (32 - sh)); }
    else { l0 = b[q]; l1 = b[q + 1]; l2 = b[q + 2]; }
    Kmer k; k.w0 = rev2(l0); k.w1 = rev2(l1); k.w2 = rev2(l2);
    uint32_t ctx = 0;
    if (i > 0 || hasL) { uint32_t j = p - 1; ctx |= 16u << ((b[j >> 4] >> (2 * (j & 15))) & 3u); }
    if (i + 1 < nk || hasR) { uint32_t j = p + SN_K; ctx |= 1u << ((b[j >> 4] >> (2 * (j & 15))) & 3u); }
    Kmer r;
    if (kmer_form(k, &r) == REV) { k = r; ctx = ctx_rc(ctx); }
    *out = k; *ctx_out = ctx;
}

// number of bucket bits for n_occ k-mer occurrences: 768..1536 occurrences per bucket.  The
// shared-memory table of k_bucket_count2 has 1024 slots; sequencing data has 0.15-0.3 distinct
// k-mers per occurrence (C2: 0.25), so a bucket normally needs one pass at <= 40 % load.
#define SN_MSP_TARGET_OCC 1536
inline int msp_bucket_bits(uint64_t n_occ)
{
    int b = 4;
    while (b < 24 && (n_occ >> b) > SN_MSP_TARGET_OCC) ++b;
    return b;
}

}  // namespace sn

// The bucket of the current pass's window a super-k-mer goes to, or 0xFFFFFFFF when it belongs to another pass.
//   plain window   : bucket - b_lo, inside [0, b_n)  (one GPU counting in several passes over bucket ranges)
//   interleaved    : pcfg = wb | lp << 8 | pass << 16 (sharded count in passes, sn_multi.cu): the global bucket is
//                    owner (top bits) | within (wb bits); the top lp bits of `within` name the pass.  The buckets of one pass
//                    are renumbered owner | rest, so that every owner's share of the pass is one contiguous range.
SN_HD uint32_t msp_window_bucket(uint32_t b, uint32_t b_lo, uint32_t b_n, uint32_t pcfg)
{
    if (pcfg) {
        const uint32_t wb = pcfg & 0xFFu, lp = (pcfg >> 8) & 0xFFu, ps = pcfg >> 16, rest = wb - lp;
        const uint32_t within = b & ((1u << wb) - 1u);
        if ((within >> rest) != ps) return 0xFFFFFFFFu;
        b = ((b >> wb) << rest) | (within & ((1u << rest) - 1u));
    }
    b -= b_lo;
    return b < b_n ? b : 0xFFFFFFFFu;
}

#if defined(__CUDACC__)
#include "sn_prims.cuh"
namespace sn {

// ---------------------------------------------------------------------------
// k_msp_scan: one read per thread; a CTA stages the packed bases of its 128 reads in shared
// memory with 16-byte loads (as k_extract does).  EMIT = false: histogram of the super-k-mers
// per bucket.  EMIT = true: every super-k-mer record goes straight to its place in the
// bucket-ordered record array (bucket start + a per-bucket cursor).
// ---------------------------------------------------------------------------
#define SN_MS_READS 128
#define SN_MS_BYTES (SN_MS_READS * (256 / 4) + 64)
#define SN_MS_QUEUE 12

template <bool EMIT>
static __global__ void __launch_bounds__(SN_MS_READS) k_msp_scan(uint64_t n_reads, const uint8_t* __restrict__ bases, const uint64_t* __restrict__ boff,
                                                            const uint32_t* __restrict__ goodlen, const int32_t* __restrict__ bc, int64_t ign_bc_below, uint32_t min_gl,
                                                            int bits, uint32_t* __restrict__ counter /* hist or cursor, one per bucket of the window */,
                                                            const uint64_t* __restrict__ bucket_off, uint4* __restrict__ recs,
                                                            uint32_t b_lo = 0u, uint32_t b_n = 0xFFFFFFFFu /* bucket window [b_lo, b_lo + b_n): a count in several passes */,
                                                            uint2* __restrict__ dsc = nullptr, uint8_t* __restrict__ nruns = nullptr /* out: the runs of every read (see k_msp_place) */,
                                                            const uint8_t* __restrict__ only_overflow = nullptr /* in: handle only the reads whose runs did not fit the descriptors */,
                                                            uint32_t* __restrict__ n_overflow = nullptr /* out: how many reads those are */, uint32_t pcfg = 0u)
{
    __shared__ __align__(16) uint8_t sb[SN_MS_BYTES];
    __shared__ uint32_t ring[SN_W * SN_MS_READS];
    __shared__ uint32_t qv[SN_MS_QUEUE * SN_MS_READS], qs[SN_MS_QUEUE * SN_MS_READS];
    const uint32_t tid = threadIdx.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * SN_MS_READS;
    const uint32_t nr = (uint32_t)min((uint64_t)SN_MS_READS, n_reads - r0);
    const uint64_t lo = boff[r0], hi = boff[r0 + nr];
    const uint64_t lo_al = lo & ~15ull;
    const uint32_t shift = (uint32_t)(lo - lo_al);
    {
        const uint4* src = reinterpret_cast<const uint4*>(bases + lo_al);
        uint4* dst = reinterpret_cast<uint4*>(sb);
        uint32_t nv = (uint32_t)((hi - lo_al + 32 + 15) >> 4);       // +32: sk_build reads ahead (the allocation is padded)
        if (nv > SN_MS_BYTES / 16) nv = SN_MS_BYTES / 16;
        for (uint32_t i = tid; i < nv; i += SN_MS_READS) dst[i] = src[i];
    }
    __syncthreads();
    if (tid >= nr) return;
    const uint64_t r = r0 + tid;
    const uint32_t gl = goodlen[r];
    if (only_overflow && only_overflow[r] != 255u) return;
    if (gl < min_gl) { if (nruns) nruns[r] = 0; return; }
    const uint8_t* rp = sb + (uint32_t)(boff[r] - lo) + shift;
    uint32_t bc24 = 0xFFFFFFu;
    if (EMIT) {
        int32_t b = -1;
        if (bc && (int64_t)r >= ign_bc_below) b = bc[r];
        bc24 = b < 0 ? 0xFFFFFFu : (uint32_t)b;
    }
    const int sh = 32 - bits;
    auto process = [&](uint32_t start, uint32_t nk, uint32_t minval) {
        const uint32_t bh = bucket_hash(minval);
        const uint32_t bkt = msp_window_bucket(bh >> sh, b_lo, b_n, pcfg);
        if (bkt == 0xFFFFFFFFu) return;                    // another pass's bucket
        if (!EMIT) atomicAdd(&counter[bkt], 1u);
        else {
            const uint64_t pos = bucket_off[bkt] + atomicAdd(&counter[bkt], 1u);
            uint32_t w[SN_SK_WORDS];
            sk_build_staged(sb, (uint32_t)(rp - sb), gl, start, nk, bc24, bh, w);
            recs[2 * pos] = make_uint4(w[0], w[1], w[2], w[3]);
            recs[2 * pos + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    };
    // the runs are queued while the read is scanned and handled afterwards, so that the scan loop
    // itself stays free of divergent work (a run ends at a different base in every lane)
    uint32_t nq = 0, nrun = 0;
    msp_scan(rp, gl, ring + tid, SN_MS_READS, [&](uint32_t start, uint32_t nk, uint32_t minval) {
        ++nrun;
        if (nq < SN_MS_QUEUE) { qv[nq * SN_MS_READS + tid] = minval; qs[nq * SN_MS_READS + tid] = start | (nk << 16); ++nq; }
        else process(start, nk, minval);
    }, min_gl);
    // the runs of the read, kept for the passes that follow (scatter, further bucket windows): [block][slot][thread],
    // so that a warp reads and writes them coalesced; a read with more runs than slots is marked and scanned again
    if (nruns) {
        nruns[r] = nrun <= SN_MS_QUEUE ? (uint8_t)nrun : (uint8_t)255;
        if (nrun > SN_MS_QUEUE && n_overflow) atomicAdd(n_overflow, 1u);
        uint2* d = dsc + (uint64_t)blockIdx.x * SN_MS_QUEUE * SN_MS_READS + tid;
        if (nrun <= SN_MS_QUEUE) for (uint32_t e = 0; e < nq; ++e) d[e * SN_MS_READS] = make_uint2(qv[e * SN_MS_READS + tid], qs[e * SN_MS_READS + tid]);
    }
    for (uint32_t e = 0; e < nq; ++e) { const uint32_t x = qs[e * SN_MS_READS + tid]; process(x & 0xFFFFu, x >> 16, qv[e * SN_MS_READS + tid]); }
}

// k_msp_place: the histogram (EMIT = false) or the scatter (EMIT = true) of the super-k-mers from the run
// descriptors the first scan left behind -- no minimizer is computed twice.  Same thread mapping as
// k_msp_scan (the descriptors are indexed by its blocks); reads marked 255 are left to k_msp_scan.
template <bool EMIT>
static __global__ void __launch_bounds__(SN_MS_READS) k_msp_place(uint64_t n_reads, const uint8_t* __restrict__ bases, const uint64_t* __restrict__ boff,
                                                             const uint32_t* __restrict__ goodlen, const int32_t* __restrict__ bc, int64_t ign_bc_below,
                                                             int bits, uint32_t* __restrict__ counter, const uint64_t* __restrict__ bucket_off, uint4* __restrict__ recs,
                                                             uint32_t b_lo, uint32_t b_n, const uint2* __restrict__ dsc, const uint8_t* __restrict__ nruns, uint32_t pcfg = 0u)
{
    __shared__ __align__(16) uint8_t sb[EMIT ? SN_MS_BYTES : 16];
    const uint32_t tid = threadIdx.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * SN_MS_READS;
    const uint32_t nr = (uint32_t)min((uint64_t)SN_MS_READS, n_reads - r0);
    uint64_t lo = 0; uint32_t shift = 0;
    if (EMIT) {
        lo = boff[r0]; const uint64_t hi = boff[r0 + nr];
        const uint64_t lo_al = lo & ~15ull;
        shift = (uint32_t)(lo - lo_al);
        const uint4* src = reinterpret_cast<const uint4*>(bases + lo_al);
        uint4* dst = reinterpret_cast<uint4*>(sb);
        uint32_t nv = (uint32_t)((hi - lo_al + 32 + 15) >> 4);
        if (nv > SN_MS_BYTES / 16) nv = SN_MS_BYTES / 16;
        for (uint32_t i = tid; i < nv; i += SN_MS_READS) dst[i] = src[i];
        __syncthreads();
    }
    if (tid >= nr) return;
    const uint64_t r = r0 + tid;
    const uint32_t n = nruns[r];
    if (n == 0u || n == 255u) return;
    const uint32_t gl = EMIT ? goodlen[r] : 0u;
    const uint8_t* rp = sb + (uint32_t)(EMIT ? boff[r] - lo : 0) + shift;
    uint32_t bc24 = 0xFFFFFFu;
    if (EMIT) {
        int32_t b = -1;
        if (bc && (int64_t)r >= ign_bc_below) b = bc[r];
        bc24 = b < 0 ? 0xFFFFFFu : (uint32_t)b;
    }
    const int sh = 32 - bits;
    const uint2* d = dsc + (uint64_t)blockIdx.x * SN_MS_QUEUE * SN_MS_READS + tid;
    for (uint32_t e = 0; e < n; ++e) {
        const uint2 x = d[e * SN_MS_READS];
        const uint32_t bh = bucket_hash(x.x);
        const uint32_t bkt = msp_window_bucket(bh >> sh, b_lo, b_n, pcfg);
        if (bkt == 0xFFFFFFFFu) continue;
        if (!EMIT) atomicAdd(&counter[bkt], 1u);
        else {
            const uint64_t pos = bucket_off[bkt] + atomicAdd(&counter[bkt], 1u);
            uint32_t w[SN_SK_WORDS];
            sk_build_staged(sb, (uint32_t)(rp - sb), gl, x.y & 0xFFFFu, x.y >> 16, bc24, bh, w);
            recs[2 * pos] = make_uint4(w[0], w[1], w[2], w[3]);
            recs[2 * pos + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

// ---------------------------------------------------------------------------
// k_bucket_count: Kmerizer::reduce / summarizeEntries / areIgnoredBarcodes / areEnoughBarcodes
// (BuildReadQGraph48.cc:91-137,174-181) for one bucket per CTA.  The bucket's super-k-mer
// records are staged in shared memory by the TMA (one cp.async.bulk per chunk, completion on an
// mbarrier), expanded to their k-mer occurrences, and aggregated in a shared-memory hash table:
//   tag[s]  : 0 = empty, else hash | 1 of the k-mer that claimed slot s (atomicCAS)
//   k0..k2  : the k-mer;  cnt: occurrences;  flg: ctx | ign << 8 | (>= 2 barcodes) << 9;  bc0: first barcode > 0
//   list[]  : the claimed slots in claim order (emission and clean-up touch only these)
// Every thread expands a CONTIGUOUS range of the chunk's occurrences: one search for its first
// occurrence, then the k-mer and its reverse complement are rolled base by base
// (kmer_succ / kmer_pred) along the record.
// Insertion runs in two phases per batch so that no thread ever compares against a key that is
// still being written: phase A claims or finds a slot by tag, the barrier publishes the keys,
// phase B verifies the key (a tag collision between different k-mers sends the item back to
// phase A one slot further) and accumulates.  The claimer initialises count/ctx/barcode with
// plain stores, so a k-mer seen once costs one CAS; every further occurrence costs one
// atomicAdd plus an atomicOr/CAS only while it still changes the slot.
// If the table fills beyond 3/4 the pass is split in two by one more bit of the hash (see the
// pass loop).  Survivors leave as 16-byte records
// {w0,w1,w2, count:24 | ctx << 24} appended to `out` (one global atomic per bucket and round).
// ---------------------------------------------------------------------------
template <int SN_BC_THREADS, int SN_BC_SLOTS>
struct BcSmem {
    static constexpr int SN_BC_CHUNK = SN_BC_THREADS;    // records staged per TMA copy: one per thread in the prefix scan
    uint4 rec[2 * SN_BC_CHUNK];                 // staging
    uint32_t pref[SN_BC_CHUNK + 1];
    uint32_t tag[SN_BC_SLOTS], k0[SN_BC_SLOTS], k1[SN_BC_SLOTS], k2[SN_BC_SLOTS], cnt[SN_BC_SLOTS], flg[SN_BC_SLOTS], bc0[SN_BC_SLOTS];
    uint16_t list[SN_BC_SLOTS];
    unsigned long long mbar;
    uint32_t fill, over, wsum[SN_BC_THREADS / 32 + 1], ndist;
    unsigned long long out_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// rolling expansion of one record: the k-mer at occurrence i in read orientation, its reverse
// complement, and what is needed to step to occurrence i+1
struct SkCursor {
    const uint32_t* w;      // the record (8 words, shared memory)
    uint32_t i, nk, p;      // occurrence, occurrences in the record, first base of the k-mer
    uint32_t hasL, hasR, bcv, ign;
    Kmer f, r;              // k-mer and reverse complement (MSB-first words)
};
__device__ __forceinline__ uint32_t sk_base(const uint32_t* w, uint32_t j) { return (w[2 + (j >> 4)] >> (2 * (j & 15))) & 3u; }
__device__ __forceinline__ void skc_open(SkCursor& c, const uint32_t* w, uint32_t i)
{
    const uint32_t w0 = w[0];
    c.w = w; c.i = i; c.nk = sk_nk(w0); c.hasL = (w0 >> 30) & 1u; c.hasR = w0 >> 31;
    const uint32_t bv = w0 & 0xFFFFFFu;
    c.ign = bv == 0xFFFFFFu ? 0x100u : 0u; c.bcv = bv == 0xFFFFFFu ? 0u : bv;
    c.p = c.hasL + i;
    const uint32_t q = c.p >> 4, sh = 2 * (c.p & 15);
    const uint32_t* b = w + 2;
    uint32_t l0 = __funnelshift_r(b[q], b[q + 1], sh), l1 = __funnelshift_r(b[q + 1], b[q + 2], sh), l2 = __funnelshift_r(b[q + 2], q + 3 < 6 ? b[q + 3] : 0u, sh);
    c.f.w0 = rev2(l0); c.f.w1 = rev2(l1); c.f.w2 = rev2(l2);
    c.r = kmer_rc(c.f);
}
// canonical k-mer + context of the current occurrence (Kmerizer::map, BuildReadQGraph48.cc:155-172)
__device__ __forceinline__ void skc_get(const SkCursor& c, Kmer* key, uint32_t* ctx_out, uint32_t* next_base)
{
    uint32_t ctx = 0;
    if (c.i > 0 || c.hasL) ctx |= 16u << sk_base(c.w, c.p - 1);
    const uint32_t nb = sk_base(c.w, c.p + SN_K);      // always inside the 96-base field
    if (c.i + 1 < c.nk || c.hasR) ctx |= 1u << nb;
    *next_base = nb;
    if (c.r < c.f) { *key = c.r; ctx = ctx_rc(ctx); } else *key = c.f;
    *ctx_out = ctx;
}
__device__ __forceinline__ void skc_step(SkCursor& c, uint32_t next_base)
{
    c.f = kmer_succ(c.f, next_base); c.r = kmer_pred(c.r, 3u - next_base);
    ++c.i; ++c.p;
}

template <int SN_BC_THREADS, int SN_BC_SLOTS, int SN_BC_ITEMS, int MINB>
static __global__ void __launch_bounds__(SN_BC_THREADS, MINB)
k_bucket_count(const uint4* __restrict__ recs, const uint64_t* __restrict__ bucket_off, uint32_t n_buckets, uint32_t n_seg,
               uint32_t min_freq, uint32_t min_bc, int has_bc,
               uint4* __restrict__ out, uint64_t out_cap, unsigned long long* out_cursor,
               uint64_t* __restrict__ seg_base, uint32_t* __restrict__ seg_cnt /* per bucket: where its survivors went; zeroed by the caller */,
               unsigned long long* n_distinct, uint32_t* err)
{
    constexpr int SN_BC_CHUNK = SN_BC_THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BcSmem<SN_BC_THREADS, SN_BC_SLOTS>& S = *reinterpret_cast<BcSmem<SN_BC_THREADS, SN_BC_SLOTS>*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t bkt = blockIdx.x;
    if (bkt >= n_buckets) return;
    // the bucket's records: one range per segment (one segment on a single GPU; one per source
    // rank after the multi-GPU exchange).  bucket_off is the exclusive scan of the counts laid out
    // segment-major, so segment s holds records [off[s*n_buckets + bkt], off[s*n_buckets + bkt + 1]).
    {
        bool any = false;
        for (uint32_t sg = 0; sg < n_seg; ++sg) any = any || bucket_off[(uint64_t)sg * n_buckets + bkt] != bucket_off[(uint64_t)sg * n_buckets + bkt + 1];
        if (!any) return;
    }
    if (tid == 0) { mbar_init(&S.mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); S.fill = 0; S.over = 0; S.ndist = 0; }
    for (uint32_t s = tid; s < SN_BC_SLOTS; s += SN_BC_THREADS) S.tag[s] = 0;
    __syncthreads();
    uint32_t phase = 0;
    const uint32_t* recw = reinterpret_cast<const uint32_t*>(S.rec);
    // A pass takes the k-mers whose hash starts with the `depth` bits `sub` (depth 0: all of them).
    // If the table overflows the pass is abandoned and split into its two halves (depth+1: 2*sub,
    // 2*sub+1).  The halves are walked depth first, i.e. in increasing hash order, and every pass
    // leaves its survivors sorted by (hash, k-mer), so the bucket comes out sorted.  A bucket that
    // fits one pass (the normal case) reserves its output range when the pass is done; a split
    // bucket walks its passes twice: mode 1 only counts the survivors, then the range is
    // reserved, mode 2 writes them.
    uint32_t depth = 0, sub = 0, mode = 0;
    uint32_t total = 0, run = 0;                      // survivors of the bucket / written so far (mode 2)
    uint64_t base = 0;                                // the bucket's range in `out`
    for (;;) {
        bool overflowed = false;
        // The bucket's records are the concatenation of its ranges in the n_seg segments; they are staged
        // SN_BC_CHUNK at a time, every contiguous piece by its own bulk copy on the same mbarrier.
        uint32_t sg = 0;
        uint64_t rpos = bucket_off[bkt], rend = bucket_off[bkt + 1];
        for (;;) {
            while (rpos == rend && ++sg < n_seg) { rpos = bucket_off[(uint64_t)sg * n_buckets + bkt]; rend = bucket_off[(uint64_t)sg * n_buckets + bkt + 1]; }
            if (sg >= n_seg || overflowed) break;
            uint32_t nc = 0;
            {
                // what the chunk takes (every thread, so that all agree on the new position) ...
                uint32_t s2 = sg; uint64_t p2 = rpos, e2 = rend;
                while (nc < (uint32_t)SN_BC_CHUNK && s2 < n_seg) {
                    const uint32_t take = (uint32_t)min((uint64_t)(SN_BC_CHUNK - nc), e2 - p2);
                    nc += take; p2 += take;
                    if (p2 == e2 && ++s2 < n_seg) { p2 = bucket_off[(uint64_t)s2 * n_buckets + bkt]; e2 = bucket_off[(uint64_t)s2 * n_buckets + bkt + 1]; }
                }
                if (tid == 0) {                                           // ... and the copies themselves
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the staging buffer finished at the last barrier
                    mbar_expect_tx(&S.mbar, nc * 32u);
                    uint32_t s3 = sg, done = 0; uint64_t p3 = rpos, e3 = rend;
                    while (done < nc) {
                        const uint32_t take = (uint32_t)min((uint64_t)(nc - done), e3 - p3);
                        if (take) tma_load_1d(S.rec + 2 * done, recs + 2 * p3, take * 32u, &S.mbar);
                        done += take; p3 += take;
                        if (p3 == e3 && ++s3 < n_seg) { p3 = bucket_off[(uint64_t)s3 * n_buckets + bkt]; e3 = bucket_off[(uint64_t)s3 * n_buckets + bkt + 1]; }
                    }
                }
                sg = s2; rpos = p2; rend = e2;
            }
            {
                mbar_wait(&S.mbar, phase); phase ^= 1u;
                // exclusive prefix of the k-mers per record
                {
                    uint32_t v = tid < nc ? sk_nk(S.rec[2 * tid].x) : 0u, x = v;
                    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, x, o); if (lane >= (uint32_t)o) x += y; }
                    if (lane == 31) S.wsum[warp] = x;
                    __syncthreads();
                    uint32_t wb = 0;
#pragma unroll
                    for (uint32_t k = 0; k < SN_BC_THREADS / 32; ++k) wb += k < warp ? S.wsum[k] : 0u;
                    S.pref[tid] = wb + x - v;
                    if (tid == SN_BC_THREADS - 1) S.pref[SN_BC_CHUNK] = wb + x;
                    __syncthreads();
                }
                const uint32_t T = S.pref[SN_BC_CHUNK];
                const uint32_t C = (T + SN_BC_THREADS - 1) / SN_BC_THREADS;      // consecutive occurrences per thread
                uint32_t x = tid * C;
                const uint32_t xe = min(x + C, T);
                SkCursor cur; uint32_t rec_i = 0;
                if (x < xe) {
                    uint32_t a = 0, b = nc;                                   // record of occurrence x: largest a with pref[a] <= x
                    while (b - a > 1) { uint32_t m = (a + b) >> 1; if (S.pref[m] <= x) a = m; else b = m; }
                    rec_i = a;
                    skc_open(cur, recw + 8 * a, x - S.pref[a]);
                }
                for (uint32_t x0 = 0; x0 < C && !overflowed; x0 += SN_BC_ITEMS) {
                    // st: 0 = done / nothing to do, 1 = probing from slot, 2 = found by tag (verify the key in phase B)
                    Kmer key[SN_BC_ITEMS]; uint32_t tg[SN_BC_ITEMS], slot[SN_BC_ITEMS], fl[SN_BC_ITEMS], bcv[SN_BC_ITEMS], st[SN_BC_ITEMS];
#pragma unroll
                    for (int j = 0; j < SN_BC_ITEMS; ++j) {
                        st[j] = 0;
                        if (x < xe) {
                            uint32_t ctx, nb;
                            skc_get(cur, &key[j], &ctx, &nb);
                            const uint32_t h = kmer_hash(key[j]);
                            if (depth == 0u || (h >> (32u - depth)) == sub) {
                                tg[j] = h | 1u; slot[j] = h & (SN_BC_SLOTS - 1u);
                                fl[j] = ctx | cur.ign; bcv[j] = cur.bcv;
                                st[j] = 1;
                            }
                            ++x;
                            if (x < xe) {
                                if (cur.i + 1 < cur.nk) skc_step(cur, nb);
                                else { ++rec_i; skc_open(cur, recw + 8 * rec_i, 0); }
                            }
                        }
                    }
                    for (;;) {
                        // ---- phase A: claim an empty slot (and initialise it) or find the k-mer's slot by tag ----
#pragma unroll
                        for (int j = 0; j < SN_BC_ITEMS; ++j) {
                            if (st[j] != 1) continue;
                            uint32_t s = slot[j], probes = 0;
                            for (;;) {
                                uint32_t t = S.tag[s];
                                if (t == 0u) {
                                    t = atomicCAS(&S.tag[s], 0u, tg[j]);
                                    if (t == 0u) {
                                        S.k0[s] = key[j].w0; S.k1[s] = key[j].w1; S.k2[s] = key[j].w2;
                                        S.cnt[s] = 1u; S.flg[s] = fl[j]; S.bc0[s] = bcv[j];
                                        S.list[atomicAdd(&S.fill, 1u)] = (uint16_t)s;
                                        st[j] = 0; break;
                                    }
                                }
                                if (t == tg[j]) { st[j] = 2; break; }
                                s = (s + 1u) & (SN_BC_SLOTS - 1u);
                                if (++probes >= SN_BC_SLOTS) { S.over = 1u; st[j] = 0; break; }
                            }
                            slot[j] = s;
                        }
                        __syncthreads();
                        // ---- phase B: verify the key, accumulate ----
                        bool pending = false;
#pragma unroll
                        for (int j = 0; j < SN_BC_ITEMS; ++j) {
                            if (st[j] != 2) continue;
                            const uint32_t s = slot[j];
                            if (S.k0[s] != key[j].w0 || S.k1[s] != key[j].w1 || S.k2[s] != key[j].w2) {
                                st[j] = 1; slot[j] = (s + 1u) & (SN_BC_SLOTS - 1u); pending = true;      // tag collision: keep probing
                                continue;
                            }
                            atomicAdd(&S.cnt[s], 1u);
                            const uint32_t f = S.flg[s];
                            if ((f & fl[j]) != fl[j]) atomicOr(&S.flg[s], fl[j]);
                            if (bcv[j] != 0u && !(f & 0x200u)) {
                                uint32_t b0 = S.bc0[s];
                                if (b0 == 0u) b0 = atomicCAS(&S.bc0[s], 0u, bcv[j]);
                                if (b0 != 0u && b0 != bcv[j]) atomicOr(&S.flg[s], 0x200u);
                            }
                            st[j] = 0;
                        }
                        if (!__syncthreads_or(pending ? 1 : 0)) break;
                    }
                    if (S.over || S.fill > (SN_BC_SLOTS * 3) / 4) overflowed = true;     // uniform: every thread reads before the barrier below
                    __syncthreads();
                }
            }
        }
        // ---- a pass is over: which of the claimed slots hold valid k-mers? ----
        const uint32_t nfill = S.fill;
        uint32_t vmask = 0, nvalid = 0;
        if (!overflowed) {
#pragma unroll
            for (int j = 0; j < SN_BC_SLOTS / SN_BC_THREADS; ++j) {
                const uint32_t e = j * SN_BC_THREADS + tid;
                if (e < nfill) {
                    const uint32_t s = S.list[e];
                    const uint32_t c = S.cnt[s], f = S.flg[s];
                    const bool enough = min_bc == 0 || (min_bc == 1 ? S.bc0[s] != 0u : (f & 0x200u) != 0u);
                    if (c >= min_freq && (!has_bc || (f & 0x100u) || enough)) { vmask |= 1u << j; ++nvalid; }
                }
            }
        }
        uint32_t xs = nvalid;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, xs, o); if (lane >= (uint32_t)o) xs += y; }
        if (lane == 31) S.wsum[warp] = xs;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
#pragma unroll
        for (uint32_t k = 0; k < SN_BC_THREADS / 32; ++k) { const uint32_t v = S.wsum[k]; wb += k < warp ? v : 0u; tot += v; }
        if (!overflowed) {
            if (mode != 2u && tid == 0) S.ndist += nfill;
            if (mode == 0u) {                                               // the whole bucket in one pass: reserve its range now
                if (tid == 0) S.out_base = tot ? atomicAdd(out_cursor, (unsigned long long)tot) : 0ull;
                __syncthreads();
                base = S.out_base; total = tot;
            }
            if (mode == 1u) total += tot;
            else if (tot) {
                // ---- order the pass's survivors by (hash, k-mer) in shared memory and write them ----
                // The hashes are uniform, so a counting sort on their top bits (SN_BC_THREADS bins) leaves
                // about one survivor per bin; inside a bin the rank is settled by comparing.
                // Parked in arrays that are free now: slot list + bin members in the staging buffer,
                // hashes in the tag array, bin counters in the prefix array.
                constexpr int NB = SN_BC_THREADS, NB_SHIFT = 32 - (SN_BC_THREADS == 128 ? 7 : (SN_BC_THREADS == 256 ? 8 : 9));
                constexpr int PER = (SN_BC_SLOTS * 3 / 4 + SN_BC_THREADS - 1) / SN_BC_THREADS;
                uint16_t* vs = reinterpret_cast<uint16_t*>(S.rec);
                uint16_t* members = vs + (SN_BC_SLOTS * 3) / 4 + 2;
                {
                    uint32_t p = wb + xs - nvalid;
#pragma unroll
                    for (int j = 0; j < SN_BC_SLOTS / SN_BC_THREADS; ++j)
                        if (vmask & (1u << j)) {
                            const uint32_t s = S.list[j * SN_BC_THREADS + tid];
                            Kmer k; k.w0 = S.k0[s]; k.w1 = S.k1[s]; k.w2 = S.k2[s];
                            vs[p] = (uint16_t)s; S.tag[p] = kmer_hash(k); ++p;
                        }
                }
                S.pref[tid] = 0;
                __syncthreads();
                uint32_t rin[PER];                                           // arrival order inside the bin
#pragma unroll
                for (int j = 0; j < PER; ++j) { const uint32_t e = j * SN_BC_THREADS + tid; if (e < tot) rin[j] = atomicAdd(&S.pref[S.tag[e] >> NB_SHIFT], 1u); }
                __syncthreads();
                {                                                           // exclusive scan of the NB bin counts (one per thread)
                    const uint32_t v = S.pref[tid]; uint32_t x = v;
                    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, x, o); if (lane >= (uint32_t)o) x += y; }
                    if (lane == 31) S.wsum[warp] = x;
                    __syncthreads();
                    uint32_t wbase = 0;
#pragma unroll
                    for (uint32_t k = 0; k < SN_BC_THREADS / 32; ++k) wbase += k < warp ? S.wsum[k] : 0u;
                    S.pref[tid] = wbase + x - v;
                    if (tid == NB - 1) S.pref[NB] = wbase + x;
                    __syncthreads();
                }
#pragma unroll
                for (int j = 0; j < PER; ++j) { const uint32_t e = j * SN_BC_THREADS + tid; if (e < tot) members[S.pref[S.tag[e] >> NB_SHIFT] + rin[j]] = (uint16_t)e; }
                __syncthreads();
                const uint64_t obase = base + run;
                if (obase + tot > out_cap) { if (tid == 0) atomicOr(err, 1u); }
                else {
#pragma unroll
                    for (int j = 0; j < PER; ++j) {
                        const uint32_t e = j * SN_BC_THREADS + tid;
                        if (e < tot) {
                            const uint32_t h = S.tag[e], s = vs[e], bin = h >> NB_SHIFT;
                            const uint32_t m0 = S.pref[bin], m1 = S.pref[bin + 1];
                            uint32_t rank = m0;
                            for (uint32_t m = m0; m < m1; ++m) {
                                const uint32_t u = members[m];
                                if (u == e) continue;
                                const uint32_t hu = S.tag[u];
                                if (hu < h) ++rank;
                                else if (hu == h) {                         // same 32-bit hash: the k-mer decides
                                    const uint32_t su = vs[u];
                                    const uint32_t a0 = S.k0[su], b0 = S.k0[s], a1 = S.k1[su], b1 = S.k1[s];
                                    if (a0 != b0 ? a0 < b0 : (a1 != b1 ? a1 < b1 : S.k2[su] < S.k2[s])) ++rank;
                                }
                            }
                            out[obase + rank] = make_uint4(S.k0[s], S.k1[s], S.k2[s], min(S.cnt[s], 0xFFFFFFu) | ((S.flg[s] & 0xFFu) << 24));
                        }
                    }
                }
                run += tot;
            }
            if (mode == 0u) break;
        }
        // clean the table through the claim list for the next pass
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SN_BC_SLOTS / SN_BC_THREADS; ++j) { const uint32_t e = j * SN_BC_THREADS + tid; if (e < nfill) S.tag[S.list[e]] = 0; }
        if (!overflowed) for (uint32_t e = tid; e < tot; e += SN_BC_THREADS) S.tag[e] = 0;      // the survivors' hashes were parked there
        __syncthreads();
        if (tid == 0) { S.fill = 0; S.over = 0; }
        __syncthreads();
        if (overflowed) {
            if (depth >= 20u) { if (tid == 0) atomicOr(err, 2u); total = 0; break; }
            if (mode == 0u) mode = 1u;
            ++depth; sub <<= 1;                                             // lower half first
            continue;
        }
        while (depth > 0u && (sub & 1u)) { --depth; sub >>= 1; }            // upper halves done: back up
        if (depth == 0u) {
            if (mode == 2u) break;
            // every pass counted: reserve the bucket's range, then walk the passes again and write
            if (tid == 0) S.out_base = total ? atomicAdd(out_cursor, (unsigned long long)total) : 0ull;
            __syncthreads();
            base = S.out_base;
            mode = 2u; run = 0; depth = 1; sub = 0;                          // (the root pass is known to overflow)
            continue;
        }
        sub |= 1u;                                                          // sibling half
    }
    if (tid == 0) {
        seg_base[bkt] = base; seg_cnt[bkt] = total;
        if (S.ndist) atomicAdd(n_distinct, (unsigned long long)S.ndist);
    }
}

// ---------------------------------------------------------------------------
// k_bucket_count2: the same reduction (one CTA per bucket, same passes / split rule / output
// order) with an insert loop that needs no block-wide barrier and a third of the shared-memory
// wavefronts per occurrence:
//   slot[s] (16 B) = {k0, k1, k2, tag};  tag = hash[31:12] << 12 | barcode rule settled << 10 | ctx << 2 | state
//                    (state 0 = empty, 2 = claimed and being written, 1 = published)
//   acc[s]  (8 B)  = {occurrences, first barcode > 0 | IGN << 24 | MULTI << 25}
// An occurrence reads its slot with ONE 16-byte load (tag and key together).  Empty: claim the
// tag by CAS, write key + accumulators, fence, publish the tag.  Published and same key: one
// atomicAdd on the count, an atomicOr on the tag only while the context still adds a bit, a
// barcode update only while the barcode state can still change.  Claimed by someone else with
// the same hash bits: look again in the next round.  The rounds of a warp are convergent
// (`__any_sync` loop + `__syncwarp`), so a lane never spins on a slot a lane of its own warp is
// still writing; a claim itself never waits, so the loop cannot deadlock.
// The cursor keeps the upcoming bases of its record in a register (one shared-memory load per 16
// occurrences) and takes the preceding base from the k-mer it just left.
// There is no claim list: validity, the count of distinct k-mers and the clean-up scan the table
// (SLOTS / THREADS slots per thread, 16-byte loads).  A pass overflows when an occurrence probes
// SN_BC2_MAXPROBE slots without finding its k-mer or a free slot.
// ---------------------------------------------------------------------------
#define SN_BC2_MAXPROBE 64u
#define SN_BC2_IGN (1u << 24)
#define SN_BC2_MULTI (1u << 25)
template <int T, int SLOTS>
struct BcSmem2 {
    uint4 rec[2 * T];                 // staging; vs[] + members[] of the survivor ordering afterwards
    uint32_t pref[T + 1];
    uint4 slot[SLOTS];
    uint2 acc[SLOTS];
    uint32_t dtab[2 * T];             // record de-duplication: hash table of record indices (+1) over the chunk
    uint32_t mult[T], rsum[T];        // per representative record: copies in the chunk, their barcode summary
    uint16_t rid[T];                  // the representative records, in chunk order
    unsigned long long mbar;
    uint32_t over, wsum[T / 32 + 1], ndist;
    unsigned long long out_base;
};
__device__ __forceinline__ uint4 lds128v(const uint4* p)
{ uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds32v(const uint32_t* p)
{ uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ void sts32v(uint32_t* p, uint32_t v)
{ asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory"); }

struct SkCur2 {
    const uint32_t* w;
    uint32_t i, nk, p, hasR, fbv, mult;
    uint32_t nextw, nleft;      // upcoming bases (base p+K in the low 2 bits), bases left in nextw
    uint32_t prev;              // context bit of the base before the k-mer, 0 if there is none
    Kmer f, r;
};
__device__ __forceinline__ void skc2_open(SkCur2& c, const uint32_t* w, uint32_t i, uint32_t mult, uint32_t bsum)
{
    const uint32_t w0 = w[0];
    const uint32_t hasL = (w0 >> 30) & 1u;
    c.w = w; c.i = i; c.nk = sk_nk(w0); c.hasR = w0 >> 31;
    c.fbv = bsum; c.mult = mult;
    c.p = hasL + i;
    const uint32_t q = c.p >> 4, sh = 2 * (c.p & 15);
    const uint32_t* b = w + 2;
    const uint32_t l0 = __funnelshift_r(b[q], b[q + 1], sh), l1 = __funnelshift_r(b[q + 1], b[q + 2], sh), l2 = __funnelshift_r(b[q + 2], b[q + 3], sh);   // p <= 47: q + 3 <= 5
    c.f.w0 = rev2(l0); c.f.w1 = rev2(l1); c.f.w2 = rev2(l2);
    c.r.w0 = ~l2; c.r.w1 = ~l1; c.r.w2 = ~l0;                // kmer_rc(f): rev2 is an involution, so ~rev2(f.w2) = ~l2
    c.prev = c.p > 0 ? 16u << sk_base(w, c.p - 1) : 0u;      // p > 0 <=> i > 0 || hasL
    const uint32_t idx = c.p + SN_K;                          // <= 95
    c.nextw = b[idx >> 4] >> (2 * (idx & 15)); c.nleft = 16u - (idx & 15u);
}
__device__ __forceinline__ void skc2_get(const SkCur2& c, Kmer* key, uint32_t* ctx_out)
{
    uint32_t ctx = c.prev;
    if (c.i + 1 < c.nk || c.hasR) ctx |= 1u << (c.nextw & 3u);
    if (c.r < c.f) { *key = c.r; ctx = ctx_rc(ctx); } else *key = c.f;
    *ctx_out = ctx;
}
__device__ __forceinline__ void skc2_step(SkCur2& c)
{
    const uint32_t nb = c.nextw & 3u;
    c.prev = 16u << (c.f.w0 >> 30);
    c.f = kmer_succ(c.f, nb); c.r = kmer_pred(c.r, 3u - nb);
    ++c.i; ++c.p;
    c.nextw >>= 2;
    if (--c.nleft == 0u) { const uint32_t idx = c.p + SN_K; c.nextw = idx < 96u ? c.w[2 + (idx >> 4)] : 0u; c.nleft = 16u; }
}

template <int T, int SLOTS, int ITEMS, int MINB>
static __global__ void __launch_bounds__(T, MINB)
k_bucket_count2(const uint4* __restrict__ recs, const uint64_t* __restrict__ bucket_off, uint32_t n_buckets, uint32_t n_seg,
                uint32_t min_freq, uint32_t min_bc, int has_bc,
                uint4* __restrict__ out, uint64_t out_cap, unsigned long long* out_cursor,
                uint64_t* __restrict__ seg_base, uint32_t* __restrict__ seg_cnt,
                unsigned long long* n_distinct, uint32_t* err,
                const uint2* __restrict__ vbucket /* {bucket, depth0 | prefix0 << 8} per CTA: a heavy bucket is shared out by hash prefix */, uint32_t n_virtual)
{
    static_assert(4 * SLOTS <= 32 * T, "the survivor ordering parks 2 x SLOTS u16 in the staging buffer");
    static_assert(SLOTS % T == 0 && SLOTS / T <= 32, "slots per thread");
    constexpr int CHUNK = T;
    constexpr int PER = SLOTS / T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BcSmem2<T, SLOTS>& S = *reinterpret_cast<BcSmem2<T, SLOTS>*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // A bucket with many records (low-complexity sequence, a repeat family: 10^5..10^7 k-mers under one minimizer) is shared
    // out over 2^d0 CTAs: CTA j counts the k-mers whose hash starts with the d0-bit prefix j (every CTA streams all of the
    // bucket's records, but its table only takes its own share), so the passes a big bucket needs run side by side
    // instead of one after the other on one CTA.  Prefix order is hash order: the CTAs' outputs concatenate.
    if (blockIdx.x >= n_virtual) return;
    const uint2 vb = vbucket[blockIdx.x];
    const uint32_t bkt = vb.x, d0 = vb.y & 0xFFu, s0 = vb.y >> 8;
    const uint32_t vslot = blockIdx.x;
    if (bkt >= n_buckets) return;
    {
        bool any = false;
        for (uint32_t sg = 0; sg < n_seg; ++sg) any = any || bucket_off[(uint64_t)sg * n_buckets + bkt] != bucket_off[(uint64_t)sg * n_buckets + bkt + 1];
        if (!any) return;
    }
    if (tid == 0) { mbar_init(&S.mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); S.over = 0; S.ndist = 0; }
#pragma unroll
    for (int j = 0; j < PER; ++j) S.slot[j * T + tid].w = 0u;
    __syncthreads();
    uint32_t phase = 0;
    const uint32_t* recw = reinterpret_cast<const uint32_t*>(S.rec);
    uint32_t depth = d0, sub = s0, mode = 0;          // passes: see k_bucket_count (this CTA's root is the prefix s0 of d0 bits)
    uint32_t total = 0, run = 0;
    uint64_t base = 0;
    for (;;) {
        bool overflowed = false;
        uint32_t sg = 0;
        uint64_t rpos = bucket_off[bkt], rend = bucket_off[bkt + 1];
        for (;;) {
            while (rpos == rend && ++sg < n_seg) { rpos = bucket_off[(uint64_t)sg * n_buckets + bkt]; rend = bucket_off[(uint64_t)sg * n_buckets + bkt + 1]; }
            if (sg >= n_seg || overflowed) break;
            uint32_t nc = 0;
            if (n_seg == 1u) {                                                  // one contiguous range: one bulk copy per chunk
                nc = (uint32_t)min((uint64_t)CHUNK, rend - rpos);
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&S.mbar, nc * 32u);
                    tma_load_1d(S.rec, recs + 2 * rpos, nc * 32u, &S.mbar);
                }
                rpos += nc;
            } else {
                uint32_t s2 = sg; uint64_t p2 = rpos, e2 = rend;
                while (nc < (uint32_t)CHUNK && s2 < n_seg) {
                    const uint32_t take = (uint32_t)min((uint64_t)(CHUNK - nc), e2 - p2);
                    nc += take; p2 += take;
                    if (p2 == e2 && ++s2 < n_seg) { p2 = bucket_off[(uint64_t)s2 * n_buckets + bkt]; e2 = bucket_off[(uint64_t)s2 * n_buckets + bkt + 1]; }
                }
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&S.mbar, nc * 32u);
                    uint32_t s3 = sg, done = 0; uint64_t p3 = rpos, e3 = rend;
                    while (done < nc) {
                        const uint32_t take = (uint32_t)min((uint64_t)(nc - done), e3 - p3);
                        if (take) tma_load_1d(S.rec + 2 * done, recs + 2 * p3, take * 32u, &S.mbar);
                        done += take; p3 += take;
                        if (p3 == e3 && ++s3 < n_seg) { p3 = bucket_off[(uint64_t)s3 * n_buckets + bkt]; e3 = bucket_off[(uint64_t)s3 * n_buckets + bkt + 1]; }
                    }
                }
                sg = s2; rpos = p2; rend = e2;
            }
            // (while the copy is in flight) reset the record de-duplication state of the chunk
            S.dtab[tid] = 0u; S.dtab[tid + T] = 0u; S.mult[tid] = 1u; S.rsum[tid] = 0u;
            mbar_wait(&S.mbar, phase); phase ^= 1u;
            __syncthreads();
            // ---- identical super-k-mers (same bases, neighbours and flags: reads covering the same stretch
            // of the genome) are expanded ONCE: the first copy in the table represents the others and
            // carries their number and the summary of their barcodes (none / one / several / ignored),
            // which is all Kmerizer::reduce needs from them (BuildReadQGraph48.cc:91-137).
            uint32_t my_nk = 0;
            if (tid < nc) {
                const uint4 ra = S.rec[2 * tid], rb = S.rec[2 * tid + 1];
                const uint32_t idw = ra.x >> 24;                                 // nk-1, hasL, hasR
                uint32_t hr = (ra.z * 0x9E3779B1u) ^ (ra.w * 0x85EBCA77u) ^ (rb.x * 0xC2B2AE3Du) ^ (rb.y * 0x27D4EB2Fu) ^ (rb.z * 0x165667B1u) ^ (rb.w * 0x9E3779B9u);
                hr += idw; hr ^= hr >> 15; hr *= 0x2C1B3C6Du; hr ^= hr >> 12;
                uint32_t rep = tid, sd = hr & (2u * T - 1u);
                for (;;) {
                    uint32_t v = lds32v(&S.dtab[sd]);
                    if (v == 0u) v = atomicCAS(&S.dtab[sd], 0u, tid + 1u);
                    if (v == 0u) break;                                          // first of its kind
                    const uint32_t u = v - 1u;
                    const uint4 ua = S.rec[2 * u], ub = S.rec[2 * u + 1];
                    if ((ua.x >> 24) == idw && ua.z == ra.z && ua.w == ra.w && ub.x == rb.x && ub.y == rb.y && ub.z == rb.z && ub.w == rb.w) { rep = u; break; }
                    sd = (sd + 1u) & (2u * T - 1u);
                }
                if (rep != tid) atomicAdd(&S.mult[rep], 1u); else my_nk = sk_nk(ra.x);
                const uint32_t bv = ra.x & 0xFFFFFFu;
                if (bv == 0xFFFFFFu) atomicOr(&S.rsum[rep], SN_BC2_IGN);
                else if (bv != 0u) {
                    const uint32_t old = atomicCAS(&S.rsum[rep], 0u, bv);
                    if (!(old & (SN_BC2_MULTI | SN_BC2_IGN)) && old != 0u && old != bv) atomicOr(&S.rsum[rep], SN_BC2_MULTI);   // (an ignored copy decides the rule by itself)
                }
            }
            uint32_t nrep;
            {                                                                   // representatives in chunk order + exclusive prefix of their k-mers
                const uint32_t v = my_nk ? (my_nk | 0x10000u) : 0u; uint32_t xsc = v;
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, xsc, o); if (lane >= (uint32_t)o) xsc += y; }
                if (lane == 31) S.wsum[warp] = xsc;
                __syncthreads();
                uint32_t wb = 0, all = 0;
#pragma unroll
                for (uint32_t k = 0; k < T / 32; ++k) { const uint32_t q = S.wsum[k]; wb += k < warp ? q : 0u; all += q; }
                const uint32_t incl = wb + xsc;
                if (my_nk) { const uint32_t pos = (incl >> 16) - 1u; S.pref[pos] = (incl & 0xFFFFu) - my_nk; S.rid[pos] = (uint16_t)tid; }
                nrep = all >> 16;
                if (tid == 0) S.pref[nrep] = all & 0xFFFFu;
                __syncthreads();
            }
            const uint32_t Tn = S.pref[nrep];
            const uint32_t C = (Tn + T - 1) / T;                                 // consecutive occurrences per thread
            uint32_t x = tid * C;
            const uint32_t xe = min(x + C, Tn);
            SkCur2 cur; uint32_t rec_i = 0;
            if (x < xe) {
                uint32_t a = 0, b = nrep;                                       // record of occurrence x: largest a with pref[a] <= x
                while (b - a > 1) { uint32_t m = (a + b) >> 1; if (S.pref[m] <= x) a = m; else b = m; }
                rec_i = a;
                const uint32_t r = S.rid[a];
                skc2_open(cur, recw + 8 * r, x - S.pref[a], S.mult[r], S.rsum[r]);
            }
            // every warp walks its threads' ranges on its own: no block barrier until the chunk is done.
            // ITEMS occurrences per thread are in flight together (independent hash chains and slot loads).
            for (uint32_t x0 = 0; x0 < C; x0 += ITEMS) {
                uint32_t act = 0;
                Kmer key[ITEMS]; uint32_t h[ITEMS], cx[ITEMS], fb[ITEMS], mu[ITEMS], sl[ITEMS];
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) {
                    h[j] = 0; cx[j] = 0; fb[j] = 0; mu[j] = 0; key[j].w0 = key[j].w1 = key[j].w2 = 0;
                    if (x < xe) {
                        skc2_get(cur, &key[j], &cx[j]);
                        h[j] = kmer_hash(key[j]);
                        fb[j] = cur.fbv; mu[j] = cur.mult;
                        if (depth == 0u || (h[j] >> (32u - depth)) == sub) act |= 1u << j;
                        ++x;
                        if (x < xe) {
                            if (cur.i + 1 < cur.nk) skc2_step(cur);
                            else { ++rec_i; const uint32_t r = S.rid[rec_i]; skc2_open(cur, recw + 8 * r, 0, S.mult[r], S.rsum[r]); }
                        }
                    }
                    sl[j] = h[j] & (SLOTS - 1u);                                // slot | probes << 16
                }
                if (lds32v(&S.over)) break;                                     // (the pass is abandoned anyway)
                uint32_t rounds = 0;
                while (__any_sync(SN_FULL, act != 0u)) {
                    // probe: a tight loop per item until it stands on (a) an empty slot, (b) the published slot
                    // of its k-mer, (c) a slot somebody is still writing under its hash bits
                    uint32_t tg[ITEMS];
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j) {
                        tg[j] = 0;
                        if (!(act & (1u << j))) continue;
                        const uint32_t hb = h[j] & ~4095u;
                        uint32_t s = sl[j] & 0xFFFFu, probes = sl[j] >> 16;
                        for (;;) {
                            const uint4 e = lds128v(&S.slot[s]);
                            tg[j] = e.w;
                            if (e.w == 0u) break;
                            if ((e.w & ~4095u) == hb && ((e.w & 3u) == 2u || (e.x == key[j].w0 && e.y == key[j].w1 && e.z == key[j].w2))) break;
                            s = (s + 1u) & (SLOTS - 1u);
                            if (++probes >= SN_BC2_MAXPROBE) { tg[j] = 3u; break; }         // state 3 = gave up
                        }
                        sl[j] = s | (probes << 16);
                    }
                    // act
#pragma unroll
                    for (int j = 0; j < ITEMS; ++j) {
                        if (!(act & (1u << j))) continue;
                        const uint32_t s = sl[j] & 0xFFFFu, hb = h[j] & ~4095u, ctx = cx[j], fbv = fb[j], t = tg[j];
                        if (t == 0u) {
                            if (atomicCAS(&S.slot[s].w, 0u, hb | 2u) == 0u) {                  // claimed: key + accumulators, then publish
                                S.acc[s] = make_uint2(mu[j], fbv);
                                S.slot[s] = make_uint4(key[j].w0, key[j].w1, key[j].w2, hb | 2u);
                                __threadfence_block();
                                sts32v(&S.slot[s].w, hb | ((fbv & (SN_BC2_IGN | SN_BC2_MULTI)) ? 1024u : 0u) | (ctx << 2) | 1u);
                                act &= ~(1u << j);
                            }                                                               // (lost the race: probe again from this slot)
                        } else if ((t & 3u) == 1u) {
                            atomicAdd(&S.acc[s].x, mu[j]);
                            if ((((t >> 2) & 0xFFu) & ctx) != ctx) atomicOr(&S.slot[s].w, ctx << 2);
                            if (fbv != 0u && !(t & 1024u)) {                                // bit 10: the barcode rule is already settled
                                if (fbv & (SN_BC2_IGN | SN_BC2_MULTI)) { atomicOr(&S.acc[s].y, fbv & (SN_BC2_IGN | SN_BC2_MULTI)); atomicOr(&S.slot[s].w, 1024u); }
                                else {
                                    uint32_t b0 = lds32v(&S.acc[s].y);
                                    if (b0 == 0u) b0 = atomicCAS(&S.acc[s].y, 0u, fbv);
                                    if (b0 & (SN_BC2_IGN | SN_BC2_MULTI)) atomicOr(&S.slot[s].w, 1024u);
                                    else if (b0 != 0u && b0 != fbv) { atomicOr(&S.acc[s].y, SN_BC2_MULTI); atomicOr(&S.slot[s].w, 1024u); }
                                }
                            }
                            act &= ~(1u << j);
                        } else if (t == 3u) { S.over = 1u; act &= ~(1u << j); }
                        else if (++rounds > (1u << 22)) { atomicOr(err, 4u); S.over = 1u; act = 0; }     // being written (cannot last: a claim never waits)
                    }
                    __syncwarp();
                }
            }
            overflowed = __syncthreads_or((int)lds32v(&S.over)) != 0;          // (whoever set it reads it back as set)
        }
        // ---- a pass is over: which slots hold valid k-mers? ----
        uint32_t vmask = 0, nvalid = 0, nfill = 0;
        if (!overflowed) {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const uint32_t s = j * T + tid;
                if (S.slot[s].w != 0u) {
                    ++nfill;
                    const uint2 a = S.acc[s];
                    const bool enough = min_bc == 0 || (min_bc == 1 ? (a.y & (0xFFFFFFu | SN_BC2_MULTI)) != 0u : (a.y & SN_BC2_MULTI) != 0u);
                    if (a.x >= min_freq && (!has_bc || (a.y & SN_BC2_IGN) || enough)) { vmask |= 1u << j; ++nvalid; }
                }
            }
        }
        uint32_t xs = nvalid | (nfill << 16);                                   // both counts in one scan (each <= SLOTS <= 2^15)
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, xs, o); if (lane >= (uint32_t)o) xs += y; }
        if (lane == 31) S.wsum[warp] = xs;
        __syncthreads();
        uint32_t wb = 0, tot = 0, totfill = 0;
#pragma unroll
        for (uint32_t k = 0; k < T / 32; ++k) { const uint32_t v = S.wsum[k]; wb += k < warp ? (v & 0xFFFFu) : 0u; tot += v & 0xFFFFu; totfill += v >> 16; }
        xs &= 0xFFFFu;
        if (!overflowed) {
            if (mode != 2u && tid == 0) S.ndist += totfill;
            if (mode == 0u) {
                if (tid == 0) S.out_base = tot ? atomicAdd(out_cursor, (unsigned long long)tot) : 0ull;
                __syncthreads();
                base = S.out_base; total = tot;
            }
            if (mode == 1u) total += tot;
            else if (tot) {
                // ---- order the pass's survivors by (hash, k-mer) and write them (see k_bucket_count) ----
                constexpr int NB = T, NB_SHIFT = 32 - (T == 64 ? 6 : (T == 128 ? 7 : (T == 256 ? 8 : 9)));
                uint16_t* vs = reinterpret_cast<uint16_t*>(S.rec);
                uint16_t* members = vs + SLOTS;
                {
                    uint32_t p = wb + xs - nvalid;
#pragma unroll
                    for (int j = 0; j < PER; ++j)
                        if (vmask & (1u << j)) {
                            const uint32_t s = j * T + tid;
                            const uint4 e = S.slot[s];
                            Kmer k; k.w0 = e.x; k.w1 = e.y; k.w2 = e.z;
                            vs[p] = (uint16_t)s; S.acc[s].y = kmer_hash(k); ++p;   // the barcode state is not needed any more
                        }
                }
                S.pref[tid] = 0;
                __syncthreads();
                uint32_t rin[PER];
#pragma unroll
                for (int j = 0; j < PER; ++j) { const uint32_t e = j * T + tid; if (e < tot) rin[j] = atomicAdd(&S.pref[S.acc[vs[e]].y >> NB_SHIFT], 1u); }
                __syncthreads();
                {
                    const uint32_t v = S.pref[tid]; uint32_t x = v;
                    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, x, o); if (lane >= (uint32_t)o) x += y; }
                    if (lane == 31) S.wsum[warp] = x;
                    __syncthreads();
                    uint32_t wbase = 0;
#pragma unroll
                    for (uint32_t k = 0; k < T / 32; ++k) wbase += k < warp ? S.wsum[k] : 0u;
                    S.pref[tid] = wbase + x - v;
                    if (tid == NB - 1) S.pref[NB] = wbase + x;
                    __syncthreads();
                }
#pragma unroll
                for (int j = 0; j < PER; ++j) { const uint32_t e = j * T + tid; if (e < tot) members[S.pref[S.acc[vs[e]].y >> NB_SHIFT] + rin[j]] = (uint16_t)e; }
                __syncthreads();
                const uint64_t obase = base + run;
                if (obase + tot > out_cap) { if (tid == 0) atomicOr(err, 1u); }
                else {
#pragma unroll
                    for (int j = 0; j < PER; ++j) {
                        const uint32_t e = j * T + tid;
                        if (e < tot) {
                            const uint32_t s = vs[e];
                            const uint4 me = S.slot[s];
                            const uint2 ma = S.acc[s];
                            const uint32_t h = ma.y, bin = h >> NB_SHIFT;
                            const uint32_t m0 = S.pref[bin], m1 = S.pref[bin + 1];
                            uint32_t rank = m0;
                            for (uint32_t m = m0; m < m1; ++m) {
                                const uint32_t u = members[m];
                                if (u == e) continue;
                                const uint32_t su = vs[u];
                                const uint32_t hu = S.acc[su].y;
                                if (hu < h) ++rank;
                                else if (hu == h) {                             // same 32-bit hash: the k-mer decides
                                    const uint4 o = S.slot[su];
                                    if (o.x != me.x ? o.x < me.x : (o.y != me.y ? o.y < me.y : o.z < me.z)) ++rank;
                                }
                            }
                            out[obase + rank] = make_uint4(me.x, me.y, me.z, min(ma.x, 0xFFFFFFu) | (((me.w >> 2) & 0xFFu) << 24));
                        }
                    }
                }
                run += tot;
            }
            if (mode == 0u) break;
        }
        // clean the table for the next pass
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) S.slot[j * T + tid].w = 0u;
        if (tid == 0) S.over = 0;
        __syncthreads();
        if (overflowed) {
            if (depth >= 30u) { if (tid == 0) atomicOr(err, 2u); total = 0; break; }
            if (mode == 0u) mode = 1u;
            ++depth; sub <<= 1;
            continue;
        }
        while (depth > d0 && (sub & 1u)) { --depth; sub >>= 1; }
        if (depth == d0) {
            if (mode == 2u) break;
            if (tid == 0) S.out_base = total ? atomicAdd(out_cursor, (unsigned long long)total) : 0ull;
            __syncthreads();
            base = S.out_base;
            mode = 2u; run = 0; depth = d0 + 1u; sub = s0 << 1;
            continue;
        }
        sub |= 1u;
    }
    if (tid == 0) {
        seg_base[vslot] = base; seg_cnt[vslot] = total;
        if (S.ndist) atomicAdd(n_distinct, (unsigned long long)S.ndist);
    }
}

// ---- virtual buckets: how many CTAs a bucket gets (1, or 2^d for a heavy one) and the CTA table ------------------------
#define SN_BC_HEAVY_RECORDS 192u         // a bucket holds ~100 records; above this it is shared out, one CTA per `heavy` records (repeat-family test set: 748 ms unshared, 31.6 ms at 1024, 7.0 ms at 64)
__device__ __forceinline__ uint32_t vb_depth(uint64_t n_rec, uint32_t heavy)
{
    uint32_t d = 0;
    while (d < 12u && (n_rec >> d) > heavy) ++d;
    return d;
}
static __global__ void __launch_bounds__(256) k_vb_count(const uint64_t* __restrict__ bucket_off, uint32_t n_buckets, uint32_t n_seg, uint32_t heavy, uint32_t* __restrict__ nv)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint64_t r = 0;
    for (uint32_t sg = 0; sg < n_seg; ++sg) r += bucket_off[(uint64_t)sg * n_buckets + b + 1] - bucket_off[(uint64_t)sg * n_buckets + b];
    nv[b] = 1u << vb_depth(r, heavy);
}
static __global__ void __launch_bounds__(256) k_vb_fill(const uint32_t* __restrict__ nv, const uint64_t* __restrict__ first, uint32_t n_buckets, uint2* __restrict__ vbucket)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    const uint32_t n = nv[b]; uint32_t d = 0;
    while ((1u << d) < n) ++d;
    uint2* o = vbucket + first[b];
    for (uint32_t j = 0; j < n; ++j) o[j] = make_uint2(b, d | (j << 8));
}
// survivors per real bucket = sum over its virtual buckets
static __global__ void __launch_bounds__(256) k_vb_sum(const uint32_t* __restrict__ vcnt, const uint64_t* __restrict__ first, uint32_t n_buckets, uint32_t* __restrict__ cnt)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint32_t s = 0;
    for (uint64_t v = first[b]; v < first[b + 1]; ++v) s += vcnt[v];
    cnt[b] = s;
}

// k-mer occurrences held by n records
static __global__ void __launch_bounds__(256) k_sum_nk(const uint4* __restrict__ recs, uint64_t n, unsigned long long* total)
{
    unsigned long long s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += sk_nk(recs[2 * i].x);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(SN_FULL, s, o);
    if (lane_id() == 0 && s) atomicAdd(total, s);
}

// every bucket's survivors from where k_bucket_count left them to their place in bucket order (a warp per bucket)
static __global__ void __launch_bounds__(256) k_gather_survivors(const uint4* __restrict__ scratch, const uint64_t* __restrict__ seg_base, const uint32_t* __restrict__ seg_cnt,
                                                          const uint64_t* __restrict__ off, uint32_t n_buckets, uint4* __restrict__ surv)
{
    const uint32_t lane = threadIdx.x & 31u;
    for (uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_buckets; b += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
        const uint32_t n = seg_cnt[b];
        const uint4* src = scratch + seg_base[b];
        uint4* dst = surv + off[b];
        for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    }
}
// surviving k-mers (bucket, hash, k-mer order) -> dictionary entries
static __global__ void __launch_bounds__(256) k_make_dict(const uint4* __restrict__ surv, uint32_t n, DictEntry* __restrict__ dict, uint32_t* __restrict__ hs)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = surv[i];
    DictEntry e;
    e.w0 = r.x; e.w1 = r.y; e.w2 = r.z; e.cc = r.w; e.edge = SN_NULL_EDGE; e.off = 0; e.ctx = r.w >> 24; e.h = rs_hash(r);
    dict[i] = e;
    hs[i] = e.h;
}
// the hashes of a finished table on their own (DictView::hs)
static __global__ void __launch_bounds__(256) k_dict_hs(const DictEntry* __restrict__ dict, uint32_t n, uint32_t* __restrict__ hs)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hs[i] = dict[i].h;
}
// offsets per (bucket, top sub_bits of the hash) from the bucket offsets: a search per cell
static __global__ void __launch_bounds__(256) k_dict_cells(const DictEntry* __restrict__ dict, const uint32_t* __restrict__ bucket_off, uint32_t n_buckets, int sub_bits,
                                                    uint32_t* __restrict__ cell_off)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_cells = (uint64_t)n_buckets << sub_bits;
    if (c > n_cells) return;
    if (c == n_cells) { cell_off[c] = bucket_off[n_buckets]; return; }
    const uint32_t b = (uint32_t)(c >> sub_bits), s = (uint32_t)(c & ((1u << sub_bits) - 1u));
    uint32_t lo = bucket_off[b], hi = bucket_off[b + 1];
    const uint32_t key = sub_bits ? s << (32 - sub_bits) : 0u;          // first entry of the bucket with hash >= key
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (dict[mid].h < key) lo = mid + 1; else hi = mid; }
    cell_off[c] = lo;
}
static __global__ void __launch_bounds__(256) k_narrow_u64(const uint64_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ out)
{ const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = (uint32_t)in[i]; }
static __global__ void __launch_bounds__(256) k_diff_u32(const uint32_t* __restrict__ off, uint32_t n, uint32_t* __restrict__ cnt)
{ const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) cnt[i] = off[i + 1] - off[i]; }

}  // namespace sn
#endif
