// sn_prims.cuh -- device-wide primitives written for this hot path (no CUB/Thrust):
//   * tile look-back (decoupled single-pass prefix over tiles)
//   * exclusive scan  (u32 -> u64)
//   * LSD radix sort of 128-bit records {w0,w1,w2,aux} by the 96-bit k-mer, one
//     read + one write of every record per 8-bit digit pass ("onesweep" shape:
//     all digit histograms in one upfront pass, per-tile offsets by look-back).
// Integer / byte work only; every kernel is HBM-bound by design and sized as
// persistent-ish grids of 256-thread CTAs (multiples of the 148 SMs are chosen by
// the launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "sn_kmer.cuh"

namespace sn {

#define SN_FULL 0xFFFFFFFFu

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// ---------------------------------------------------------------------------
// Tile look-back.  status word (u64) = flag(2 bits) << 62 | value(62 bits).
// A tile first publishes its AGGREGATE, then walks predecessors until it meets
// a PREFIX, then publishes its own inclusive PREFIX.  The word is self-contained
// (value and flag travel in one 64-bit store), so relaxed gpu-scope accesses suffice.
// Tiles take their ids from an atomic counter, so every predecessor is already
// resident and the spin cannot deadlock.
// ---------------------------------------------------------------------------
#define SN_ST_INVALID 0ull
#define SN_ST_AGG 1ull
#define SN_ST_PREFIX 2ull
#define SN_ST_VMASK 0x3FFFFFFFFFFFFFFFull

__device__ __forceinline__ void st_status(uint64_t* p, uint64_t flag, uint64_t v)
{ uint64_t x = (flag << 62) | v; asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(x) : "memory"); }
__device__ __forceinline__ uint64_t ld_status(const uint64_t* p)
{ uint64_t x; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory"); return x; }

// exclusive prefix of `aggregate` over tiles [0,tile); `status` has `stride` words per tile.
// The walk over the predecessors is windowed: SN_LB_WINDOW status words are requested at
// once, so the latency of the (L2) loads overlaps instead of adding up tile by tile.
#define SN_LB_WINDOW 8
__device__ __forceinline__ void lookback_publish(uint64_t* status, uint32_t stride, uint32_t slot, uint32_t tile, uint64_t aggregate)
{
    if (tile == 0) st_status(status + slot, SN_ST_PREFIX, aggregate);
    else st_status(status + (size_t)tile * stride + slot, SN_ST_AGG, aggregate);
}
__device__ __forceinline__ uint64_t lookback_walk(uint64_t* status, uint32_t stride, uint32_t slot, uint32_t tile, uint64_t aggregate)
{
    if (tile == 0) return 0;
    uint64_t excl = 0;
    int64_t t = (int64_t)tile - 1;
    bool done = false;
    while (!done && t >= 0) {
        uint64_t s[SN_LB_WINDOW];
#pragma unroll
        for (int i = 0; i < SN_LB_WINDOW; ++i) if (t - i >= 0) s[i] = ld_status(status + (size_t)(t - i) * stride + slot);
#pragma unroll
        for (int i = 0; i < SN_LB_WINDOW; ++i) {
            if (!done && t - i >= 0) {
                while ((s[i] >> 62) == SN_ST_INVALID) s[i] = ld_status(status + (size_t)(t - i) * stride + slot);
                excl += s[i] & SN_ST_VMASK;
                if ((s[i] >> 62) == SN_ST_PREFIX) done = true;
            }
        }
        t -= SN_LB_WINDOW;
    }
    st_status(status + (size_t)tile * stride + slot, SN_ST_PREFIX, excl + aggregate);
    return excl;
}
__device__ __forceinline__ uint64_t tile_lookback(uint64_t* status, uint32_t stride, uint32_t slot, uint32_t tile, uint64_t aggregate)
{
    lookback_publish(status, stride, slot, tile, aggregate);
    return lookback_walk(status, stride, slot, tile, aggregate);
}

// ---------------------------------------------------------------------------
// exclusive scan u32 -> u64 (three small kernels; inputs here are per-read or
// per-dictionary-entry counts, far smaller than the k-mer stream).
// ---------------------------------------------------------------------------
#define SN_SCAN_THREADS 256
#define SN_SCAN_ITEMS 8
#define SN_SCAN_TILE (SN_SCAN_THREADS * SN_SCAN_ITEMS)

__device__ __forceinline__ uint64_t block_excl_scan_u64(uint64_t v, uint64_t* total, uint64_t* sm /*>=8*/)
{
    // inclusive warp scan, then scan of warp totals
    uint32_t lane = lane_id(), w = threadIdx.x >> 5;
    uint64_t x = v;
    for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(SN_FULL, x, o); if (lane >= (uint32_t)o) x += y; }
    if (lane == 31) sm[w] = x;
    __syncthreads();
    if (w == 0) {
        uint64_t t = lane < (blockDim.x >> 5) ? sm[lane] : 0;
        uint64_t xs = t;
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(SN_FULL, xs, o); if (lane >= (uint32_t)o) xs += y; }
        if (lane < (blockDim.x >> 5)) sm[lane] = xs - t;
        if (lane == 31) sm[32] = xs;
    }
    __syncthreads();
    uint64_t r = sm[w] + x - v;
    if (total) *total = sm[32];
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(SN_SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ tile_sums)
{
    __shared__ uint64_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * SN_SCAN_TILE;
    uint64_t s = 0;
    for (int j = 0; j < SN_SCAN_ITEMS; ++j) { uint64_t i = base + (uint64_t)j * SN_SCAN_THREADS + threadIdx.x; if (i < n) s += in[i]; }
    uint64_t tot; block_excl_scan_u64(s, &tot, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
// single block: exclusive scan of tile sums in place; total written to tile_sums[ntiles]
static __global__ void __launch_bounds__(1024) k_scan_spine(uint64_t* tile_sums, uint64_t ntiles)
{
    __shared__ uint64_t sm[33];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t b = 0; b < ntiles; b += blockDim.x) {
        uint64_t i = b + threadIdx.x;
        uint64_t v = i < ntiles ? tile_sums[i] : 0;
        uint64_t tot; uint64_t e = block_excl_scan_u64(v, &tot, sm);
        uint64_t c = carry;
        if (i < ntiles) tile_sums[i] = c + e;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}
static __global__ void __launch_bounds__(SN_SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ tile_sums, uint64_t* __restrict__ out)
{
    __shared__ uint64_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * SN_SCAN_TILE + (uint64_t)threadIdx.x * SN_SCAN_ITEMS;   // blocked
    uint32_t v[SN_SCAN_ITEMS]; uint64_t s = 0;
    for (int j = 0; j < SN_SCAN_ITEMS; ++j) { uint64_t i = base + j; v[j] = i < n ? in[i] : 0; s += v[j]; }
    uint64_t e = block_excl_scan_u64(s, nullptr, sm) + tile_sums[blockIdx.x];
    for (int j = 0; j < SN_SCAN_ITEMS; ++j) { uint64_t i = base + j; if (i < n) out[i] = e; e += v[j]; }
}

// out[i] = sum_{j<i} in[j]; out[n] = total.  `tmp` needs (ceil(n/TILE)+1) u64.
static inline uint64_t scan_tmp_words(uint64_t n) { return (n + SN_SCAN_TILE - 1) / SN_SCAN_TILE + 1; }
static inline void exclusive_scan_u32_u64(const uint32_t* in, uint64_t n, uint64_t* out, uint64_t* tmp, cudaStream_t st)
{
    if (n == 0) { cudaMemsetAsync(out, 0, sizeof(uint64_t), st); return; }
    uint64_t nt = (n + SN_SCAN_TILE - 1) / SN_SCAN_TILE;
    k_scan_tile_sums<<<(unsigned)nt, SN_SCAN_THREADS, 0, st>>>(in, n, tmp);
    k_scan_spine<<<1, 1024, 0, st>>>(tmp, nt);
    k_scan_apply<<<(unsigned)nt, SN_SCAN_THREADS, 0, st>>>(in, n, tmp, out);
    cudaMemcpyAsync(out + n, tmp + nt, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
}

// ---------------------------------------------------------------------------
// LSD radix sort of uint4 records {x,y,z,w} = {w0,w1,w2,aux}, 8-bit digits, two key modes:
//   RS_KEY96  : the 96-bit (x,y,z) key, 12 passes            (HBV end keys, exports)
//   RS_HASH32 : kmer_hash(x,y,z), 4 passes                    (the k-mer stream)
// Every pass reads each record once and writes it once.
// ---------------------------------------------------------------------------
#define SN_RS_ITEMS 8
enum { RS_KEY96 = 0, RS_HASH32 = 1 };
template <int MODE> struct RsMode { static constexpr int PASSES = MODE == RS_KEY96 ? 12 : 4; };
#define SN_RS_MAX_PASSES 12
#define SN_RS_MIN_TILE (256 * SN_RS_ITEMS)

__device__ __forceinline__ uint32_t rs_hash(const uint4& k) { Kmer q; q.w0 = k.x; q.w1 = k.y; q.w2 = k.z; return kmer_hash(q); }
template <int MODE>
__device__ __forceinline__ uint32_t rs_digit(const uint4& k, int pass)
{
    if (MODE == RS_HASH32) return (rs_hash(k) >> (8 * pass)) & 0xFFu;
    uint32_t w = pass < 4 ? k.z : (pass < 8 ? k.y : k.x);
    return (w >> (8 * (pass & 3))) & 0xFFu;
}

// all digit histograms in one pass over the records (they are invariant under the
// permutations the later passes apply).  hist[pass*256 + digit], u32 counts.
template <int MODE>
static __global__ void __launch_bounds__(256) k_rs_histogram(const uint4* __restrict__ keys, uint32_t n, uint32_t* __restrict__ hist, int arg)
{
    constexpr int P = RsMode<MODE>::PASSES;
    __shared__ uint32_t sh[P * 256];
    for (int i = threadIdx.x; i < P * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 k = keys[i];
        if (MODE == RS_HASH32) {
            uint32_t h = rs_hash(k);
#pragma unroll
            for (int p = 0; p < P; ++p) atomicAdd(&sh[p * 256 + ((h >> (8 * p)) & 0xFFu)], 1u);
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) atomicAdd(&sh[p * 256 + rs_digit<MODE>(k, p)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * 256; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// exclusive scan of each pass's 256 bins, in place (one block of 256 threads per pass)
static __global__ void __launch_bounds__(256) k_rs_scan_hist(uint32_t* hist)
{
    __shared__ uint32_t sm[256];
    uint32_t* h = hist + blockIdx.x * 256;
    uint32_t v = h[threadIdx.x];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        uint32_t y = threadIdx.x >= (uint32_t)o ? sm[threadIdx.x - o] : 0;
        __syncthreads();
        sm[threadIdx.x] += y;
        __syncthreads();
    }
    h[threadIdx.x] = sm[threadIdx.x] - v;
}

template <int THREADS>
struct RsSmem {
    uint4 keys[THREADS * SN_RS_ITEMS];         // reorder buffer
    uint32_t warp_cnt[THREADS / 32][256];      // per-warp digit counters -> exclusive warp offsets
    uint32_t tile_start[256];                  // exclusive scan of the tile's digit totals
    uint32_t gdst[256];                        // global destination of smem slot s with digit d: gdst[d] + s
    uint32_t scan_tmp[8];
    uint32_t tile_id;
};

// One digit pass: read each record once, write it once.
template <int MODE, int THREADS>
static __global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_rs_scatter(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t n, int pass,
             const uint32_t* __restrict__ ghist /* exclusive starts, this pass */, uint64_t* status, uint32_t* tile_counter)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * SN_RS_ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem<THREADS>& S = *reinterpret_cast<RsSmem<THREADS>*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid == 0) S.tile_id = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < WARPS * 256; i += THREADS) (&S.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile_id;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const uint32_t valid = (uint32_t)min((uint64_t)TILE, (uint64_t)n - tile_base);

    // warp-striped load: warp w owns [w*256, (w+1)*256); item j of lane l is w*256 + j*32 + l.
    // Records past the end get digit 255 and the highest ranks of it: they are never written.
    uint4 key[SN_RS_ITEMS];
    uint32_t dr[SN_RS_ITEMS];                  // digit | rank << 8
#pragma unroll
    for (int j = 0; j < SN_RS_ITEMS; ++j) {
        uint32_t li = warp * (32 * SN_RS_ITEMS) + j * 32 + lane;
        if (li < valid) { key[j] = in[tile_base + li]; dr[j] = rs_digit<MODE>(key[j], pass); }
        else { key[j] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); dr[j] = 0xFFu; }
    }
    // stable in-warp ranking by digit ("warp-wide radix partitioning"): eight ballots, one per
    // digit bit, give every lane the mask of lanes holding the same digit; the lowest such
    // lane bumps the warp's private counter for that digit.
#pragma unroll
    for (int j = 0; j < SN_RS_ITEMS; ++j) {
        uint32_t d = dr[j];
        uint32_t peers = SN_FULL;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
            uint32_t bal = __ballot_sync(SN_FULL, (d >> bit) & 1u);
            peers &= ((d >> bit) & 1u) ? bal : ~bal;
        }
        uint32_t leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) { base = S.warp_cnt[warp][d]; S.warp_cnt[warp][d] = base + __popc(peers); }
        base = __shfl_sync(SN_FULL, base, leader);
        dr[j] = d | ((base + __popc(peers & lanemask_lt())) << 8);
        __syncwarp();                                // the next item's leader may be another lane reading the same counter
    }
    __syncthreads();
    // thread d (< 256): exclusive scan of digit d over the warps -> tile total, published at once
    // as this tile's AGGREGATE; then the exclusive scan of the totals over the 256 digits
    uint32_t tot = 0, texcl = 0;
    if (tid < 256) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { uint32_t c = S.warp_cnt[w][tid]; S.warp_cnt[w][tid] = tot; tot += c; }
        lookback_publish(status, 256, tid, tile, tot);
        uint32_t x = tot;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(SN_FULL, x, o); if (lane >= (uint32_t)o) x += y; }
        if (lane == 31) S.scan_tmp[warp] = x;
        texcl = x - tot;                           // exclusive within the warp
    }
    __syncthreads();
    if (tid < 256) {
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) if (w < (int)warp) wbase += S.scan_tmp[w];
        texcl += wbase;
        S.tile_start[tid] = texcl;
    }
    __syncthreads();
    // scatter into the smem reorder buffer (frees the key registers before the look-back)
#pragma unroll
    for (int j = 0; j < SN_RS_ITEMS; ++j) {
        uint32_t d = dr[j] & 0xFFu;
        uint32_t pos = S.tile_start[d] + S.warp_cnt[warp][d] + (dr[j] >> 8);
        S.keys[pos] = key[j];
    }
    // look-back over the predecessor tiles for this digit's global offset
    if (tid < 256) {
        uint32_t prev = (uint32_t)lookback_walk(status, 256, tid, tile, tot);
        S.gdst[tid] = ghist[tid] + prev - texcl;   // may wrap below zero; the later + s brings it back (mod 2^32)
    }
    __syncthreads();
    // coalesced write-out: consecutive slots of one digit are consecutive in global memory
#pragma unroll
    for (int j = 0; j < SN_RS_ITEMS; ++j) {
        uint32_t s = j * THREADS + tid;
        if (s < valid) {
            uint4 k = S.keys[s];
            uint32_t d = rs_digit<MODE>(k, pass);
            out[(uint32_t)(S.gdst[d] + s)] = k;
        }
    }
}

static inline size_t radix_sort_tmp_bytes(uint32_t n)
{
    uint32_t nt = (n + SN_RS_MIN_TILE - 1) / SN_RS_MIN_TILE;
    return (size_t)SN_RS_MAX_PASSES * 256 * 4 + 64 + (size_t)nt * 256 * 8;
}
// Sorts n (< 2^32) records in two steps so the caller can time them apart:
// radix_sort_histograms (one read of the records) then radix_sort_passes (PASSES digit passes,
// each one read + one write).  The result ends in `a` (even number of ping-pong passes).
template <int MODE>
static inline cudaError_t radix_sort_histograms(const uint4* a, uint32_t n, void* tmp, int num_sms, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    uint32_t* hist = (uint32_t*)tmp;
    cudaMemsetAsync(hist, 0, (SN_RS_MAX_PASSES * 256 + 16) * 4, st);
    k_rs_histogram<MODE><<<num_sms * 8, 256, 0, st>>>(a, n, hist, 0);
    k_rs_scan_hist<<<RsMode<MODE>::PASSES, 256, 0, st>>>(hist);
    return cudaGetLastError();
}
static inline int rs_threads()
{
    static int t = 0;
    if (!t) { const char* e = getenv("SN_RS_THREADS"); t = e ? atoi(e) : 512; if (t != 256 && t != 512) t = 512; }
    return t;
}
template <int MODE, int THREADS>
static inline cudaError_t radix_sort_passes_t(uint4* a, uint4* b, uint32_t n, void* tmp, int arg, cudaStream_t st)
{
    constexpr int TILE = THREADS * SN_RS_ITEMS;
    uint32_t nt = (n + TILE - 1) / TILE;
    uint32_t* hist = (uint32_t*)tmp;
    uint32_t* counters = hist + SN_RS_MAX_PASSES * 256;
    uint64_t* status = (uint64_t*)(counters + 16);
    {   // every call: the kernels have internal linkage (one copy per translation unit that sorts) and the attribute
        // belongs to the copy and to the current device; the call costs microseconds
        cudaError_t e = cudaFuncSetAttribute(k_rs_scatter<MODE, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem<THREADS>));
        if (e != cudaSuccess) return e;
    }
    uint4* src = a; uint4* dst = b;
    for (int p = 0; p < RsMode<MODE>::PASSES; ++p) {
        cudaMemsetAsync(status, 0, (size_t)nt * 256 * 8, st);
        k_rs_scatter<MODE, THREADS><<<nt, THREADS, sizeof(RsSmem<THREADS>), st>>>(src, dst, n, p, hist + p * 256, status, counters + p);
        uint4* t = src; src = dst; dst = t;
    }
    return cudaGetLastError();
}
template <int MODE>
static inline cudaError_t radix_sort_passes(uint4* a, uint4* b, uint32_t n, void* tmp, int arg, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    return rs_threads() == 512 ? radix_sort_passes_t<MODE, 512>(a, b, n, tmp, arg, st) : radix_sort_passes_t<MODE, 256>(a, b, n, tmp, arg, st);
}
template <int MODE>
static inline cudaError_t radix_sort(uint4* a, uint4* b, uint32_t n, void* tmp, int num_sms, cudaStream_t st)
{
    cudaError_t e = radix_sort_histograms<MODE>(a, n, tmp, num_sms, st);
    if (e != cudaSuccess) return e;
    return radix_sort_passes<MODE>(a, b, n, tmp, 0, st);
}
}  // namespace sn
