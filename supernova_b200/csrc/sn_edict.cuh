// sn_edict.cuh -- the k-mer dictionary of a given edge set: what buildGraphFromMSP does between reading the
// MSPEDGES file and pathReads (paths/long/BuildReadQGraph48.cc:1647-1664):
//     for every edge e, every position i:  dict[ canonical(k-mer at i) ].set(e, i)
// The reference fills a hopscotch hash set sequentially (a k-mer that occurs twice keeps the LAST (e, i)).
// Here: one record per edge k-mer, two radix sorts (by k-mer, then by (minimizer bucket, hash, k-mer rank)),
// and the product's bucket-ordered dictionary comes out with (edge, offset) filled in -- the same structure
// the count path builds, so the pathing kernels run on it unchanged.
#pragma once
#include "sn_prims.cuh"

namespace sn {

// k-mers per edge (0 for an edge shorter than K: flagged)
static __global__ void __launch_bounds__(256) k_ed_nk(const uint32_t* __restrict__ elen, uint32_t n_edges, uint32_t* __restrict__ nk, uint32_t* err)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const uint32_t l = elen[e];
    if (l < SN_K) { atomicOr(err, 1u); nk[e] = 0; } else nk[e] = l - (SN_K - 1);
}
// edge of global k-mer g: largest e with koff[e] <= g
__device__ __forceinline__ uint32_t ed_edge_of(const uint64_t* __restrict__ koff, uint32_t n_edges, uint64_t g)
{
    uint32_t a = 0, b = n_edges;
    while (b - a > 1) { const uint32_t m = (a + b) >> 1; if (koff[m] <= g) a = m; else b = m; }
    return a;
}
// record {canonical k-mer, g}
static __global__ void __launch_bounds__(256) k_ed_records(const uint8_t* __restrict__ ebases, const uint64_t* __restrict__ eoff, const uint64_t* __restrict__ koff,
                                                    uint32_t n_edges, uint32_t n_k, uint4* __restrict__ rec)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_k) return;
    const uint32_t e = ed_edge_of(koff, n_edges, g);
    Kmer k = kmer_from_packed(ebases + eoff[e], g - koff[e]), r;
    if (kmer_form(k, &r) == REV) k = r;
    rec[g] = make_uint4(k.w0, k.w1, k.w2, g);
}
// sorted by k-mer (stable: equal k-mers in increasing g): the last of every run stays
static __global__ void __launch_bounds__(256) k_ed_last_of_run(const uint4* __restrict__ rec, uint32_t n, uint32_t* __restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool last = i + 1 == n;
    if (!last) { const uint4 a = rec[i], b = rec[i + 1]; last = a.x != b.x || a.y != b.y || a.z != b.z; }
    flag[i] = last ? 1u : 0u;
}
// second key: {minimizer bucket, hash, rank by k-mer}; per-bucket counts
static __global__ void __launch_bounds__(256) k_ed_bucket_keys(const uint4* __restrict__ rec, const uint32_t* __restrict__ flag, const uint64_t* __restrict__ pos, uint32_t n, int bits,
                                                        uint4* __restrict__ key, uint32_t* __restrict__ gids, uint32_t* __restrict__ bucket_cnt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const uint4 r = rec[i];
    Kmer k; k.w0 = r.x; k.w1 = r.y; k.w2 = r.z;
    const uint32_t b = bucket_hash(kmer_minimizer(k)) >> (32 - bits);
    const uint32_t u = (uint32_t)pos[i];
    key[u] = make_uint4(b, kmer_hash(k), u, u);
    gids[u] = r.w;
    atomicAdd(&bucket_cnt[b], 1u);
}
// final order -> 16-byte dictionary seeds {w0,w1,w2,0} (k_make_dict turns them into entries) + (edge, offset) per entry
static __global__ void __launch_bounds__(256) k_ed_emit(const uint4* __restrict__ key, const uint32_t* __restrict__ gids, uint32_t n,
                                                 const uint8_t* __restrict__ ebases, const uint64_t* __restrict__ eoff, const uint64_t* __restrict__ koff, uint32_t n_edges,
                                                 uint4* __restrict__ surv, uint2* __restrict__ loc)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = gids[key[i].w];
    const uint32_t e = ed_edge_of(koff, n_edges, g), off = (uint32_t)(g - koff[e]);
    Kmer k = kmer_from_packed(ebases + eoff[e], off), r;
    if (kmer_form(k, &r) == REV) k = r;
    surv[i] = make_uint4(k.w0, k.w1, k.w2, 0u);
    loc[i] = make_uint2(e, off);
}
static __global__ void __launch_bounds__(256) k_ed_set_loc(DictEntry* __restrict__ dict, const uint2* __restrict__ loc, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dict[i].edge = loc[i].x; dict[i].off = loc[i].y;
}

}  // namespace sn
