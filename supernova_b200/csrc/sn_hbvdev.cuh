// sn_hbvdev.cuh -- device side of buildHBVFromEdges (paths/long/HBVFromEdges.cc:28-296) around
// the one step that is sequential by definition, the FIFO numbering of HBVBuilder::processQueue
// (:199-228):
//   before it  k_hbv_rank      exact BVComp order (:106-111) of the edges from the device pre-order
//              k_hbv_groups    per vertex: its edge ends in EEComp order (:113-121), one 64-byte record
//              k_hbv_erec      per oriented edge: its two vertices
//              k_hbv_union / k_hbv_flatten / k_hbv_compstats / k_hbv_roots / k_hbv_compgather
//                              connected components of the graph; for each the first item the reference's
//                              outer loop (:277-285) would reach, its vertex and edge counts -> id bases,
//                              so the host can number all components IN PARALLEL with final ids
//   after it   k_hbv_csr_rec + sort + k_hbv_csr_emit   the sorted adjacency lists digraphE::AddEdge
//                              (graph/DigraphTemplate.h:2572-2582) maintains; k_hbv_inv the involution
//                              (paths/HyperBasevector.cc:685-697)
#pragma once
#include "sn_prims.cuh"
#include "sn_hbv.h"

namespace sn {

// lexicographic comparison of two equally long fastb-packed sequences (byte aligned)
__device__ __forceinline__ int cmp_packed_dev(const uint8_t* x, const uint8_t* y, uint32_t len)
{
    const uint32_t nb = (len + 3) / 4;
    for (uint32_t i = 0; i < nb; ++i) {
        const uint32_t a = x[i], b = y[i];
        if (a != b) { const int sh = (__ffs(a ^ b) - 1) & ~1; return ((a >> sh) & 3u) < ((b >> sh) & 3u) ? -1 : 1; }
    }
    return 0;
}
// ord: records {~len, first 16 bases, next 16 bases, edge} sorted by key.  Ties (same length, same
// first 32 bases -- the two branches of a SNP bubble, for one) are ordered by the full sequence.
static __global__ void __launch_bounds__(128) k_hbv_rank(const uint4* __restrict__ ord, uint32_t n, const uint8_t* __restrict__ ebases, const uint64_t* __restrict__ eoff,
                                                  const uint32_t* __restrict__ elen, uint32_t* __restrict__ order, uint32_t* __restrict__ rank)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = ord[i];
    auto same = [&](uint32_t j) { const uint4 q = ord[j]; return q.x == r.x && q.y == r.y && q.z == r.z; };
    uint32_t pos = i;
    if ((i > 0 && same(i - 1)) || (i + 1 < n && same(i + 1))) {
        uint32_t a = i, b = i + 1;
        while (a > 0 && same(a - 1)) --a;
        while (b < n && same(b)) ++b;
        const uint32_t len = elen[r.w];
        const uint8_t* mine = ebases + eoff[r.w];
        uint32_t before = 0;
        for (uint32_t j = a; j < b; ++j) {
            if (j == i) continue;
            const uint32_t e = ord[j].w;
            const int c = cmp_packed_dev(ebases + eoff[e], mine, len);
            if (c < 0 || (c == 0 && e < r.w)) ++before;
        }
        pos = a + before;
    }
    order[pos] = r.w; rank[r.w] = pos;
}

// items: edge<<2 | rc<<1 | distal, grouped by vertex (gstart).  EEComp order = (edge rank, rc, pos).
static __global__ void __launch_bounds__(128) k_hbv_groups(const uint32_t* __restrict__ items, const uint32_t* __restrict__ gstart, uint32_t n_groups, uint32_t n_valid,
                                                    const uint32_t* __restrict__ rank, snh::GroupRec* __restrict__ groups, uint32_t* err)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t s = gstart[g], e = g + 1 < n_groups ? gstart[g + 1] : n_valid;
    uint32_t n = e - s;
    if (n > 8) { atomicOr(err, 1u); n = 8; }
    uint32_t it[8]; unsigned long long key[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < (int)n) { it[i] = items[s + i]; key[i] = ((unsigned long long)rank[it[i] >> 2] << 2) | (it[i] & 3u); }
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        if (i >= (int)n) break;
#pragma unroll
        for (int j = i; j > 0; --j)
            if (key[j - 1] > key[j]) { unsigned long long tk = key[j]; key[j] = key[j - 1]; key[j - 1] = tk; uint32_t ti = it[j]; it[j] = it[j - 1]; it[j - 1] = ti; }
    }
    uint4* o = reinterpret_cast<uint4*>(groups + g);
    o[0] = make_uint4(0xFFFFFFFFu /* vid = -1 */, n, n > 0 ? it[0] >> 1 : 0u, n > 1 ? it[1] >> 1 : 0u);
    o[1] = make_uint4(n > 2 ? it[2] >> 1 : 0u, n > 3 ? it[3] >> 1 : 0u, n > 4 ? it[4] >> 1 : 0u, n > 5 ? it[5] >> 1 : 0u);
    o[2] = make_uint4(n > 6 ? it[6] >> 1 : 0u, n > 7 ? it[7] >> 1 : 0u, 0u, 0u);
    o[3] = make_uint4(0u, 0u, 0u, 0u);
}
static __global__ void __launch_bounds__(256) k_hbv_erec(const int32_t* __restrict__ egrp, const uint8_t* __restrict__ pal, uint32_t n_items, snh::ERec* __restrict__ er)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const uint32_t e = t >> 1, rc = t & 1u;
    reinterpret_cast<uint4*>(er)[t] = make_uint4((uint32_t)egrp[4 * e + 2 * rc], (uint32_t)egrp[4 * e + 2 * rc + 1], 0xFFFFFFFFu, pal[e]);
}

// the numbering loop's record of an oriented unipath: both vertices with their items, one cache line
static __global__ void __launch_bounds__(128) k_hbv_itemrec(const int32_t* __restrict__ egrp, const uint8_t* __restrict__ pal, const snh::GroupRec* __restrict__ groups,
                                                     uint32_t n_items, snh::ItemRec* __restrict__ rec)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const uint32_t e = t >> 1, rc = t & 1u;
    const int32_t g1 = egrp[4 * e + 2 * rc], g2 = egrp[4 * e + 2 * rc + 1];
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = 0u;
    w[0] = (uint32_t)g1; w[1] = (uint32_t)g2;
    uint32_t n1 = 0, n2 = 0;
    if (g1 >= 0) { const snh::GroupRec& G = groups[g1]; n1 = G.n; if (n1 <= 6) { for (uint32_t i = 0; i < 6; ++i) w[3 + i] = i < n1 ? G.items[i] : 0u; } else n1 = 15; }
    if (g2 >= 0) { const snh::GroupRec& G = groups[g2]; n2 = G.n; if (n2 <= 6) { for (uint32_t i = 0; i < 6; ++i) w[9 + i] = i < n2 ? G.items[i] : 0u; } else n2 = 15; }
    w[2] = n1 | (n2 << 4) | ((uint32_t)pal[e] << 8);
    uint4* o = reinterpret_cast<uint4*>(rec + t);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]); o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    o[2] = make_uint4(w[8], w[9], w[10], w[11]); o[3] = make_uint4(w[12], w[13], w[14], w[15]);
}

// ---- connected components (lock-free union-find, smaller root wins) --------------------------
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x)
{
    for (;;) {
        const uint32_t p = *(volatile uint32_t*)(parent + x);
        if (p == x) return x;
        const uint32_t gp = *(volatile uint32_t*)(parent + p);
        if (gp != p) parent[x] = gp;                   // path halving (benign race: only ever moves towards the root)
        x = p;
    }
}
static __global__ void __launch_bounds__(256) k_hbv_uf_init(uint32_t* parent, uint32_t n, unsigned long long* ckey, uint32_t* cnt_v, uint32_t* cnt_e)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    parent[g] = g; ckey[g] = ~0ull; cnt_v[g] = 0; cnt_e[g] = 0;
}
static __global__ void __launch_bounds__(256) k_hbv_union(const snh::ERec* __restrict__ er, uint32_t n_items, uint32_t* parent)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const int32_t g1 = er[t].g1, g2 = er[t].g2;
    if (g1 < 0 || g2 < 0) return;                     // reverse item of a palindromic edge: it is the forward item
    uint32_t a = (uint32_t)g1, b = (uint32_t)g2;
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) break;
        if (a < b) { const uint32_t x = a; a = b; b = x; }
        if (atomicCAS(parent + a, a, b) == a) break;   // hang the larger root under the smaller
    }
}
static __global__ void __launch_bounds__(256) k_hbv_flatten(uint32_t* parent, uint32_t n, uint32_t* __restrict__ comp, uint32_t* cnt_v)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const uint32_t r = uf_find(parent, g);
    comp[g] = r;
    // one atomic per component and warp: a well-covered genome is one giant component per strand,
    // and every vertex adding to the same address serialises in the L2
    const unsigned m = __match_any_sync(__activemask(), r);
    if ((threadIdx.x & 31u) == (uint32_t)__ffs((int)m) - 1u) atomicAdd(cnt_v + r, (uint32_t)__popc(m));
}
// per component: number of HBV edges, and the first item of the reference's outer loop
// (pass 0 = forward items in edge order, pass 1 = reverse items): min of (rc, rank)
static __global__ void __launch_bounds__(256) k_hbv_compstats(const snh::ERec* __restrict__ er, uint32_t n_items, const uint32_t* __restrict__ comp,
                                                       const uint32_t* __restrict__ rank, unsigned long long* ckey, uint32_t* cnt_e)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const int32_t g1 = er[t].g1;
    if (g1 < 0 || er[t].g2 < 0) return;
    const uint32_t r = comp[g1];
    // aggregated per component inside the warp (see k_hbv_flatten): min of the 33-bit keys, count
    const unsigned m = __match_any_sync(__activemask(), r);
    const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs((int)m) - 1u;
    const uint32_t k32 = __reduce_min_sync(m, ((t & 1u) << 31) | rank[t >> 1]);       // rank < 2^31 unipaths
    const unsigned long long best = ((unsigned long long)(k32 >> 31) << 32) | (k32 & 0x7FFFFFFFu);
    if (lane == leader) { atomicMin(ckey + r, best); atomicAdd(cnt_e + r, (uint32_t)__popc(m)); }
}
static __global__ void __launch_bounds__(256) k_hbv_rootflag(const uint32_t* __restrict__ comp, uint32_t n, uint32_t* __restrict__ flag)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) flag[g] = comp[g] == g ? 1u : 0u;
}
static __global__ void __launch_bounds__(256) k_hbv_roots(const uint32_t* __restrict__ flag, const uint64_t* __restrict__ pos, uint32_t n,
                                                   const unsigned long long* __restrict__ ckey, uint4* __restrict__ rec)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n || !flag[g]) return;
    const unsigned long long k = ckey[g];
    rec[pos[g]] = make_uint4(0u, (uint32_t)(k >> 32), (uint32_t)k, g);
}
// rec sorted by (rc, rank): component i -> its start item (edge << 1 | rc) and sizes
static __global__ void __launch_bounds__(256) k_hbv_compgather(const uint4* __restrict__ rec, uint32_t n_comp, const uint32_t* __restrict__ order,
                                                        const uint32_t* __restrict__ cnt_v, const uint32_t* __restrict__ cnt_e,
                                                        uint32_t* __restrict__ start_item, uint32_t* __restrict__ cv, uint32_t* __restrict__ ce)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_comp) return;
    const uint4 r = rec[i];
    start_item[i] = (order[r.z] << 1) | r.y;
    cv[i] = cnt_v[r.w]; ce[i] = cnt_e[r.w];
}

// ---- the FIFO numbering (HBVBuilder::processQueue, HBVFromEdges.cc:199-228) of the SMALL components on the device --------
// A component is numbered by a sequential breadth-first traversal, and the components carry their final id bases, so
// they are independent: one thread per component runs the traversal over the item records (its queue in local memory).
// A sparsely covered genome is 10^5-10^6 components of a few edges each -- all of them take this path and nothing of the
// numbering touches the host.  Components above SN_HBV_GPU_MAX oriented edges (a well-covered genome: one giant
// component per strand) are left to the host loop (sn_hbv.cpp), which is then the faster of the two.
#define SN_HBV_GPU_MAX 512
static __global__ void __launch_bounds__(64) k_hbv_number_small(const snh::ItemRec* __restrict__ rec, const snh::GroupRec* __restrict__ groups, uint32_t n_comp,
                                                                const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ ce,
                                                                const uint64_t* __restrict__ basev, const uint64_t* __restrict__ basee, uint32_t max_items,
                                                                uint8_t* seen /* n_items, zeroed */, uint8_t* vbits /* n_vertices, zeroed */, int32_t* vid,
                                                                uint32_t* src, int32_t* to_left, int32_t* to_right, int32_t* fwd, int32_t* rev,
                                                                uint32_t* n_big, uint32_t* err)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_comp) return;
    const uint32_t size = ce[c];
    if (size > max_items) { atomicAdd(n_big, 1u); return; }
    uint32_t Q[SN_HBV_GPU_MAX];
    uint32_t qh = 0, qt = 0;
    int32_t nextV = (int32_t)basev[c]; uint32_t nh = (uint32_t)basee[c];
    const uint32_t start = cstart[c];
    Q[qt++] = start; seen[start] = 1;
    while (qh < qt) {
        const uint32_t it = Q[qh++];
        const uint4* rp = reinterpret_cast<const uint4*>(rec + it);
        const uint4 a = rp[0], b = rp[1], cc = rp[2], d = rp[3];
        const uint32_t w[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w, d.x, d.y, d.z, d.w};
        const int32_t g1 = (int32_t)w[0], g2 = (int32_t)w[1];
        const uint32_t info = w[2];
        if (g1 < 0 || g2 < 0) { atomicOr(err, 2u); return; }
        const bool new1 = vbits[g1] == 0;
        if (new1) { vbits[g1] = 1; vid[g1] = nextV++; }
        const bool new2 = vbits[g2] == 0;                          // (g1 == g2: already set)
        if (new2) { vbits[g2] = 1; vid[g2] = nextV++; }
        const uint32_t id = nh++;
        src[id] = it; to_left[id] = vid[g1]; to_right[id] = vid[g2];
        if (it & 1u) rev[it >> 1] = (int32_t)id; else fwd[it >> 1] = (int32_t)id;
        if ((info >> 8) & 1u) { rev[it >> 1] = (int32_t)id; seen[it ^ 1u] = 1; }       // a palindromic edge: one HBV edge, processed as its forward item
        for (int side = 0; side < 2; ++side) {
            if (!(side ? new2 : new1)) continue;                   // the vertex pushed its items when it was numbered
            const int32_t g = side ? g2 : g1;
            uint32_t n = (info >> (4 * side)) & 15u;
            if (n == 15u) {                                        // more than 6 edge ends on this vertex: the vertex record has them
                const snh::GroupRec& G = groups[g];
                n = G.n;
                for (uint32_t x = 0; x < n; ++x) {
                    const uint32_t t2 = G.items[x];
                    if (!seen[t2]) { seen[t2] = 1; if (qt < SN_HBV_GPU_MAX) Q[qt++] = t2; else { atomicOr(err, 4u); return; } }
                }
                continue;
            }
            for (uint32_t x = 0; x < n; ++x) {
                const uint32_t t2 = w[3 + 6 * side + x];
                if (!seen[t2]) { seen[t2] = 1; if (qt < SN_HBV_GPU_MAX) Q[qt++] = t2; else { atomicOr(err, 4u); return; } }
            }
        }
    }
    if ((uint64_t)nextV != basev[c + 1] || (uint64_t)nh != basee[c + 1]) atomicOr(err, 8u);
}
// the host's share (big components; zero elsewhere) joins what the device numbered
static __global__ void __launch_bounds__(256) k_add_u32(uint32_t* __restrict__ acc, const uint32_t* __restrict__ add, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += add[i];
}

// ---- memory layout for the numbering of a giant component -----------------------------------------
// The numbering of one component is a sequential FIFO traversal that touches one 64-byte record per item.
// With the records in unipath order (= dictionary order of the owners: random along the genome) every step
// is a cache miss.  A well-covered genome is one or two components, so before the records of such a graph
// travel to the host they are laid out along the graph: a multi-source breadth-first labelling from 1 item
// in 32 (k_lay_seed / k_lay_round: one launch per round over the frontier, no host round trip), then the
// records are sorted by (seed, signed distance from the seed: one arm, the seed, the other arm) and their item
// lists renamed.  The layout changes no result: the traversal order only depends on the lists.
#define SN_LAY_NONE 0xFFFFFFFFu
__device__ __forceinline__ bool lay_is_seed(uint32_t item) { return ((item * 0x9E3779B1u) >> 27) == 0u; }      // 1 in 32
static __global__ void __launch_bounds__(256) k_lay_seed(const snh::ItemRec* __restrict__ rec, uint32_t n_items, uint32_t* __restrict__ lab, uint32_t* __restrict__ lev,
                                                  uint32_t* __restrict__ frontier, uint32_t* __restrict__ cnt /* 3 counters */)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    uint32_t l = SN_LAY_NONE;
    if (rec[t].g1 >= 0 && rec[t].g2 >= 0 && lay_is_seed(t)) { l = t; frontier[atomicAdd(cnt, 1u)] = t; }
    lab[t] = l; lev[t] = 0u;
}
// round r: every item of frontier (r & 1) labels its unlabelled neighbours (the items on its two vertices)
static __global__ void __launch_bounds__(256) k_lay_round(const snh::ItemRec* __restrict__ rec, const snh::GroupRec* __restrict__ groups, uint32_t round,
                                                   uint32_t* __restrict__ lab, uint32_t* __restrict__ lev, uint32_t* __restrict__ fa, uint32_t* __restrict__ fb, uint32_t* __restrict__ cnt)
{
    const uint32_t* fin = (round & 1u) ? fb : fa;
    uint32_t* fout = (round & 1u) ? fa : fb;
    const uint32_t n = cnt[round % 3u];
    uint32_t* nout = cnt + (round + 1u) % 3u;
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[(round + 2u) % 3u] = 0u;          // the counter of the round after next
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t x = fin[i];
        const snh::ItemRec r = rec[x];
        const uint32_t l = lab[x];
        const uint32_t arm_x = lev[x] & 0x80000000u;        // which side of its seed the item lies on (set at the first hop)
        for (int side = 0; side < 2; ++side) {
            uint32_t m = (r.info >> (4 * side)) & 15u;
            const uint32_t* items = side ? r.it2 : r.it1;
            if (m == 15u) { const snh::GroupRec& G = groups[side ? r.g2 : r.g1]; m = G.n; items = G.items; }
            const uint32_t arm = round == 0u ? (side ? 0x80000000u : 0u) : arm_x;
            for (uint32_t k = 0; k < m; ++k) {
                const uint32_t y = items[k];
                if (lab[y] == SN_LAY_NONE && atomicCAS(lab + y, SN_LAY_NONE, l) == SN_LAY_NONE) { lev[y] = (round + 1u) | arm; fout[atomicAdd(nout, 1u)] = y; }
            }
        }
    }
}
// sort records {seed, signed distance, item}; what no seed reached is its own seed
static __global__ void __launch_bounds__(256) k_lay_keys(const uint32_t* __restrict__ lab, const uint32_t* __restrict__ lev, uint32_t n_items, uint4* __restrict__ key)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const uint32_t l = lab[t];
    const uint32_t v = lev[t], d = v & 0x7FFFFFFFu;
    key[t] = make_uint4(l == SN_LAY_NONE ? t : l, l == SN_LAY_NONE ? 0x40000000u : ((v >> 31) ? 0x40000000u + d : 0x40000000u - d), t, 0u);
}
static __global__ void __launch_bounds__(256) k_lay_pos(const uint4* __restrict__ key, uint32_t n_items, uint32_t* __restrict__ pos)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_items) pos[key[i].z] = i;
}
// record i of the new layout = the record of item key[i].z with its lists renamed; pad = the item it is
static __global__ void __launch_bounds__(128) k_lay_permute(const uint4* __restrict__ key, const uint32_t* __restrict__ pos, const snh::ItemRec* __restrict__ rec, uint32_t n_items,
                                                     snh::ItemRec* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const uint32_t x = key[i].z;
    snh::ItemRec r = rec[x];
    const uint32_t n1 = r.info & 15u, n2 = (r.info >> 4) & 15u;
    if (n1 != 15u) for (uint32_t k = 0; k < n1 && k < 6u; ++k) r.it1[k] = pos[r.it1[k]];
    if (n2 != 15u) for (uint32_t k = 0; k < n2 && k < 6u; ++k) r.it2[k] = pos[r.it2[k]];
    r.pad = x;
    out[i] = r;
}
// the vertex records the numbering loop reads (more than 6 items) with their lists renamed
static __global__ void __launch_bounds__(256) k_lay_groups(const snh::GroupRec* __restrict__ groups, uint32_t n_groups, const uint32_t* __restrict__ pos, snh::GroupRec* __restrict__ out)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    snh::GroupRec r = groups[g];
    if (r.n > 6u) for (uint32_t k = 0; k < r.n && k < 8u; ++k) r.items[k] = pos[r.items[k]];
    out[g] = r;
}

// ---- after the numbering: adjacency lists and involution ----------------------------------------
static __global__ void __launch_bounds__(256) k_hbv_csr_rec(const int32_t* __restrict__ to_left, const int32_t* __restrict__ to_right, uint32_t n_h,
                                                     uint4* __restrict__ rec_from, uint4* __restrict__ rec_to)
{
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_h) return;
    const uint32_t l = (uint32_t)to_left[h], r = (uint32_t)to_right[h];
    rec_from[h] = make_uint4(l, r, h, 0u);             // from_[l] gets (r, h), lists sorted by (neighbour, id)
    rec_to[h] = make_uint4(r, l, h, 0u);
}
static __global__ void __launch_bounds__(256) k_hbv_csr_emit(const uint4* __restrict__ rec, uint32_t n_h, uint32_t n_v, uint32_t* __restrict__ start,
                                                      int32_t* __restrict__ nb, int32_t* __restrict__ eo)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_h) { const uint4 r = rec[i]; nb[i] = (int32_t)r.y; eo[i] = (int32_t)r.z; }
    if (i <= n_v) {                                    // start[v] = first record with vertex >= v
        uint32_t lo = 0, hi = n_h;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (rec[mid].x < i) lo = mid + 1; else hi = mid; }
        start[i] = lo;
    }
}
static __global__ void __launch_bounds__(256) k_hbv_inv(const int32_t* __restrict__ fwd, const int32_t* __restrict__ rev, uint32_t n_e, int32_t* __restrict__ inv)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_e) return;
    const int32_t f = fwd[e], r = rev[e];
    inv[f] = r; inv[r] = f;
}

}  // namespace sn
