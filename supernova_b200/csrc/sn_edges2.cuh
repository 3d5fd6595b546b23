// sn_edges2.cuh -- unipath edges over a dictionary that is SHARDED by minimizer-bucket range over the ranks
// (recomputeAdjacencies + buildEdges, kmers/ReadPather.h:346-385, paths/long/BuildReadQGraph48.cc:327-541; the
// role of tada's per-shard sedges + global stitch, lib/tada/src/cmd_shard_asm.rs:37-94, debruijn.rs:733-826).
//
// A rank holds the k-mers of its buckets only.  What it needs from other ranks:
//   GHOSTS   the k-mers of other buckets that are neighbours of local ones (5-6 % of the neighbours: those whose
//            minimizer differs).  They are collected into an open-addressed table right behind the local entries
//            (tab[n + slot], see DictView), resolved by ONE query/answer exchange with their owners (present? index
//            there?) and, after everybody pruned, completed with their pruned context by a second answer.  From
//            then on prune_ctx / classify_links (sn_graph.cuh) work unchanged: a lookup that leaves the window
//            lands in the ghost table, a link may point at a ghost.
//   STOPS    Chains are cut at stops as on one GPU (sn_kernels.cuh) -- edge ends, 1 interior k-mer in 16 -- plus
//            every k-mer with a link to a ghost and every single-k-mer edge.  Segments between stops are local.
//            The stop table {type, next stop and steps on each side} of all ranks is gathered (~1/13 of the k-mers,
//            20 bytes each), and every rank runs the hop kernels on the whole of it: lengths, owners, edge ids and
//            the (edge, offset) of every stop come out identical everywhere.
//   BASES    every rank writes the bases of its own segments into the edge store; one all-reduce (sum) completes it.
// With one rank nothing is exchanged and there are no ghosts: the same kernels are the single-GPU edge stage.
#pragma once
#include "sn_kernels.cuh"

namespace sn {

#define SN_GHOST_REF 0x80000000u      // Seg.next of a segment that ends on another rank: SN_GHOST_REF | ghost slot (before k_seg_globalize)

__device__ __forceinline__ uint32_t ldv32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ uint32_t bucket_owner(uint32_t bucket, uint32_t n_ranks, int bits) { return (uint32_t)(((uint64_t)bucket * n_ranks) >> bits); }

// ---- ghosts: collect -----------------------------------------------------------------------------------------
// thread per local k-mer: every neighbour its (unpruned) context names whose bucket lies outside the window enters the
// ghost table (once: equal k-mers meet in the same slot).  Insertion is claim (CAS on the state) / write / fence /
// publish; a lane that finds a slot being written looks again in the next round, and the rounds of a warp are
// convergent, so no lane ever spins on a lane of its own warp.
static __global__ void __launch_bounds__(256) k_ghost_collect(DictEntry* tab, DictView d, uint32_t n_ranks, uint32_t* n_ghosts, uint32_t* overflow,
                                                              uint32_t* __restrict__ glist /* g_cap / 4 * 3 + 1: the slots taken, in no particular order */, uint32_t* __restrict__ per_owner)
{
    __shared__ uint32_t blk_owner[64];                        // per-owner counts of this block (a few owners: global atomics on them would serialise)
    if (threadIdx.x < 64) blk_owner[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < d.n;
    Kmer k; k.w0 = k.w1 = k.w2 = 0; KmerMin km; uint32_t todo = 0;
    if (live) { const DictEntry& e = d.tab[i]; k = entry_kmer(e); todo = e.cc >> 24; km = kmer_minimizer_nb(k); }
    const int sh = 32 - d.bits;
    DictEntry* gh = tab + d.n;
    const uint32_t mask = d.g_cap - 1u;
    bool have = false; Kmer q; uint32_t h = 0, slot = 0, owner = 0, probes = 0;
    while (__any_sync(SN_FULL, have || todo)) {
        if (!have && todo) {                                    // next neighbour of this k-mer that lives elsewhere
            while (todo && !have) {
                const uint32_t bit = low_bit_index(todo); todo &= todo - 1u;
                const uint32_t c = bit & 3u;
                const uint32_t mn = bit < 4 ? succ_minimizer(km, c) : pred_minimizer(km, c);
                const uint32_t b = bucket_hash(mn) >> sh;
                if (b - d.b_lo < d.b_n) continue;
                q = bit < 4 ? kmer_succ(k, c) : kmer_pred(k, c);
                Kmer r; if (kmer_form(q, &r) == REV) q = r;
                h = kmer_hash(q); slot = h & mask; owner = bucket_owner(b, n_ranks, d.bits); probes = 0; have = true;
            }
        }
        if (have) {
            DictEntry* e = gh + slot;
            const uint32_t st = ldv32(&e->off);
            if (st == GH_EMPTY) {
                if (atomicCAS(&e->off, (uint32_t)GH_EMPTY, (uint32_t)GH_WRITING) == GH_EMPTY) {
                    e->w0 = q.w0; e->w1 = q.w1; e->w2 = q.w2; e->h = h; e->cc = owner; e->edge = SN_NULL_EDGE; e->ctx = 0;
                    __threadfence();
                    *reinterpret_cast<volatile uint32_t*>(&e->off) = GH_PENDING;
                    // (one atomic on the shared counter per warp and round, not per ghost)
                    const unsigned am = __activemask();
                    const uint32_t lane = threadIdx.x & 31u, ldr = (uint32_t)__ffs((int)am) - 1u;
                    uint32_t at = 0;
                    if (lane == ldr) at = atomicAdd(n_ghosts, (uint32_t)__popc(am));
                    at = __shfl_sync(am, at, (int)ldr) + (uint32_t)__popc(am & ((1u << lane) - 1u));
                    if (at >= (d.g_cap / 4u) * 3u) *overflow = 1u; else { glist[at] = slot; if (owner < 64u) atomicAdd(&blk_owner[owner], 1u); else atomicAdd(&per_owner[owner], 1u); }
                    have = false;
                }                                               // (lost the race: the slot is looked at again)
            } else if (st != GH_WRITING) {
                __threadfence();
                if (ldv32(&e->h) == h && ldv32(&e->w0) == q.w0 && ldv32(&e->w1) == q.w1 && ldv32(&e->w2) == q.w2) have = false;     // already there
                else { slot = (slot + 1u) & mask; if (++probes >= d.g_cap) { *overflow = 1u; have = false; } }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 64 && threadIdx.x < n_ranks && blk_owner[threadIdx.x]) atomicAdd(&per_owner[threadIdx.x], blk_owner[threadIdx.x]);
}
// ---- ghosts: queries grouped by owner -----------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_ghost_fill(const DictEntry* __restrict__ gh, const uint32_t* __restrict__ glist, uint32_t n_ghosts,
                                                           const uint32_t* __restrict__ base, uint32_t* __restrict__ cursor,
                                                           uint32_t* __restrict__ qk /* 3 words per query */, uint32_t* __restrict__ qslot)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < n_ghosts;
    uint32_t s = 0; DictEntry e; e.cc = 0xFFFFFFFFu; e.w0 = e.w1 = e.w2 = 0;
    if (live) { s = glist[t]; e = gh[s]; }
    // a handful of owners: one atomic per owner and warp
    const unsigned m = __match_any_sync(SN_FULL, e.cc);
    const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs((int)m) - 1u;
    uint32_t first = 0;
    if (live && lane == leader) first = atomicAdd(&cursor[e.cc], (uint32_t)__popc(m));
    first = __shfl_sync(SN_FULL, first, (int)leader);
    if (!live) return;
    const uint32_t p = base[e.cc] + first + (uint32_t)__popc(m & ((1u << lane) - 1u));
    qk[3 * p] = e.w0; qk[3 * p + 1] = e.w1; qk[3 * p + 2] = e.w2; qslot[p] = s;
}
// owner side: index of every queried k-mer in the local table, or SN_NULL_EDGE
static __global__ void __launch_bounds__(256) k_ghost_answer(DictView d /* g_cap = 0 */, const uint32_t* __restrict__ qk, uint32_t nq, uint32_t* __restrict__ ans)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    Kmer k; k.w0 = qk[3 * t]; k.w1 = qk[3 * t + 1]; k.w2 = qk[3 * t + 2];
    ans[t] = dict_find_canonical(d, k);
}
static __global__ void __launch_bounds__(256) k_ghost_apply(DictEntry* __restrict__ gh, const uint32_t* __restrict__ qslot, const uint32_t* __restrict__ ans, uint32_t nq)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    DictEntry* e = gh + qslot[t];
    e->edge = ans[t]; e->off = ans[t] == SN_NULL_EDGE ? GH_ABSENT : GH_PRESENT;
}
// owner side, after k_prune: the pruned context of every k-mer it was asked about
static __global__ void __launch_bounds__(256) k_ghost_ctx_send(const DictEntry* __restrict__ tab, const uint32_t* __restrict__ ans, uint32_t nq, uint8_t* __restrict__ out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq) out[t] = ans[t] == SN_NULL_EDGE ? (uint8_t)0 : (uint8_t)tab[ans[t]].ctx;
}
static __global__ void __launch_bounds__(256) k_ghost_ctx_apply(DictEntry* __restrict__ gh, const uint32_t* __restrict__ qslot, const uint8_t* __restrict__ ctx, uint32_t nq)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq) gh[qslot[t]].ctx = ctx[t];
}

// ---- links, stops, local segments ---------------------------------------------------------------------------
__device__ __forceinline__ bool link_is_ghost(uint32_t l, uint32_t n) { return l != SN_NO_LINK && (l >> 1) >= n; }
__device__ __forceinline__ bool has_ghost_link(const Link2& l, uint32_t n) { return link_is_ghost(l.x, n) || link_is_ghost(l.y, n); }

static __global__ void __launch_bounds__(256) k_classify2(DictView d, Link2* links, uint8_t* __restrict__ etype, uint32_t* __restrict__ is_stop)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    Link2 l;
    const int t = classify_links(d, i, links[i], &l);            // links[] holds prune's candidates on entry
    links[i] = l;
    etype[i] = (uint8_t)t;
    is_stop[i] = (t != T_INTERIOR || stop_sampled(i) || has_ghost_link(l, d.n)) ? 1u : 0u;
}
// thread per (stop, side): side 0 leaves through the down link, side 1 through the up link
static __global__ void __launch_bounds__(128) k_seg_walk2(const Link2* __restrict__ links, const uint32_t* __restrict__ stops, uint32_t n_stops,
                                                          const uint64_t* __restrict__ stop_pos, uint32_t n, Seg* __restrict__ segs)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_stops) return;
    const uint32_t s = stops[t >> 1];
    uint32_t o = t & 1u, cur = s, steps = 0;
    Seg out; out.next = SN_NO_LINK; out.steps_o = 0;
    Link2 lk = links[cur];
    for (;;) {
        const uint32_t l = o ? lk.y : lk.x;
        if (l == SN_NO_LINK) break;                               // only at the start: this side of an edge end is closed
        cur = l >> 1; o = l & 1u; ++steps;
        if (cur >= n) { out.next = SN_GHOST_REF | (cur - n); out.steps_o = (steps << 1) | o; break; }     // the chain goes on on another rank
        lk = links[cur];
        const uint32_t cont = o ? lk.y : lk.x;
        if (cont == SN_NO_LINK || stop_sampled(cur) || cur == s || has_ghost_link(lk, n)) {
            out.next = (uint32_t)stop_pos[cur]; out.steps_o = (steps << 1) | o;
            break;
        }
    }
    segs[t] = out;
}
static __global__ void __launch_bounds__(256) k_stop_types(const uint32_t* __restrict__ stops, const uint8_t* __restrict__ etype, uint32_t n_stops, uint8_t* __restrict__ stype)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_stops) stype[s] = etype[stops[s]];
}
// local stop ids -> ids in the gathered stop table; a segment that ends at a ghost is looked up in the stop list of
// the ghost's owner (sorted by table index there)
static __global__ void __launch_bounds__(256) k_seg_globalize(const Seg* __restrict__ local, uint32_t n_stops, const DictEntry* __restrict__ gh,
                                                              const uint32_t* __restrict__ gstops, const uint64_t* __restrict__ stop_base /* n_ranks + 1 */,
                                                              uint32_t rank, Seg* __restrict__ out, uint32_t* err)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_stops) return;
    Seg s = local[t];
    if (s.next != SN_NO_LINK) {
        if (s.next & SN_GHOST_REF) {
            const DictEntry& g = gh[s.next & ~SN_GHOST_REF];
            const uint32_t b = g.cc, j = g.edge;
            const uint32_t* L = gstops + stop_base[b];
            uint32_t lo = 0, hi = (uint32_t)(stop_base[b + 1] - stop_base[b]);
            while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (L[m] < j) lo = m + 1; else hi = m; }
            if (lo >= (uint32_t)(stop_base[b + 1] - stop_base[b]) || L[lo] != j) { atomicOr(err, 1u); s.next = SN_NO_LINK; }
            else s.next = (uint32_t)stop_base[b] + lo;
        } else s.next += (uint32_t)stop_base[rank];
    }
    out[t] = s;
}

// ---- the gathered stop table: lengths, owners, edge ids, offsets (identical on every rank) --------------------
// (a rank does this for ITS stops [first, first + count) only: 1/N of the hops; `own_n` is indexed by local stop)
static __global__ void __launch_bounds__(128) k_gs_end_hop(const Seg* __restrict__ segs, const uint8_t* __restrict__ stype, uint32_t first, uint32_t count, uint32_t* __restrict__ own_n)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t sid = first + t;
    const int ty = stype[sid];
    uint32_t mine = ty == T_SINGLE ? 1u : 0u;
    if (ty == T_END_DOWN || ty == T_END_UP) {
        uint32_t cur = sid, o = ty == T_END_UP ? 1u : 0u, nk = 1;
        for (;;) {
            const Seg sg = segs[2 * cur + o];
            if (sg.next == SN_NO_LINK) break;
            nk += sg.steps_o >> 1; cur = sg.next; o = sg.steps_o & 1u;
        }
        if (sid <= cur) mine = nk;                               // the other end walks the same edge; the smaller stop owns it
    }
    own_n[t] = mine;
}
// phase 0: singles and edges with ends; phase 1: circles
static __global__ void __launch_bounds__(256) k_gs_sizes(const uint32_t* __restrict__ own_n, const uint8_t* __restrict__ stype /* of the same stops */, int circles_only, uint32_t n_stops,
                                                         uint32_t* __restrict__ ebases, uint32_t* __restrict__ eflag)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_stops) return;
    uint32_t k = own_n[s];
    if (circles_only ? stype[s] != T_CIRCLE : stype[s] == T_CIRCLE) k = 0;
    ebases[s] = k ? (k + (SN_K - 1) + 3u) >> 2 : 0u;          // bytes of the edge in the packed store (4 bases per byte, byte aligned)
    eflag[s] = k ? 1u : 0u;
}
// thread per edge OWNED BY A LOCAL STOP: the owner hops over the stops of its edge -- wherever they live -- and leaves
// {edge + 1, offset, walk orientation} for each in `sinfo` (indexed by global stop; zeroed: the ranks' arrays add up)
static __global__ void __launch_bounds__(128) k_gs_owner_hop(const Seg* __restrict__ segs, const uint32_t* __restrict__ owners /* local stop ids */, uint32_t n_owners, uint32_t edge0, uint32_t first,
                                                             const uint8_t* __restrict__ stype, const uint32_t* __restrict__ own_n, const uint64_t* __restrict__ base_off, uint64_t base_shift,
                                                             uint32_t* __restrict__ elen, uint64_t* __restrict__ etmp_off, StopInfo* __restrict__ sinfo,
                                                             unsigned long long* n_kmers_on_edges)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_owners) return;
    const uint32_t ls = owners[k], sid = first + ls, e = edge0 + k;    // e: the edge's global id; elen / etmp_off are this rank's slices (indexed by k)
    const int t = stype[sid];
    elen[k] = own_n[ls] + SN_K - 1;
    atomicAdd(n_kmers_on_edges, (unsigned long long)own_n[ls]);
    etmp_off[k] = base_off[ls] + base_shift;
    uint32_t cur = sid, o = t == T_END_UP ? 1u : 0u, off = 0;
    for (;;) {
        StopInfo si; si.edge = e + 1u; si.off_o = (off << 1) | o;
        sinfo[cur] = si;
        const Seg sg = segs[2 * cur + o];
        if (sg.next == SN_NO_LINK) break;
        off += sg.steps_o >> 1; cur = sg.next; o = sg.steps_o & 1u;
        if (cur == sid) break;                                   // once around a circle
    }
}
// interior stops no edge end reached lie on circles: the smallest stop of each becomes its owner
static __global__ void __launch_bounds__(128) k_gs_circle_elect(const Seg* __restrict__ segs, uint8_t* __restrict__ stype, const StopInfo* __restrict__ sinfo, uint32_t first, uint32_t count,
                                                                uint32_t* __restrict__ own_n /* local */, uint32_t* n_found)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t sid = first + t;
    if (stype[sid] != T_INTERIOR || sinfo[sid].edge != 0u) return;
    uint32_t cur = sid, o = 0, nk = 0;
    for (;;) {
        const Seg sg = segs[2 * cur + o];
        if (sg.next == SN_NO_LINK) return;                       // (cannot happen on a circle)
        nk += sg.steps_o >> 1; cur = sg.next; o = sg.steps_o & 1u;
        if (cur == sid) break;
        if (cur < sid) return;                                   // a smaller stop owns this circle
    }
    own_n[t] = nk; stype[sid] = T_CIRCLE;
    atomicAdd(n_found, 1u);
}
// sinfo += add over a range (the circle phase is reduced separately: the first phase's sums must not be added twice)
static __global__ void __launch_bounds__(256) k_sinfo_add(StopInfo* __restrict__ acc, const StopInfo* __restrict__ add, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { acc[i].edge += add[i].edge; acc[i].off_o += add[i].off_o; }
}

// ---- bases ------------------------------------------------------------------------------------------------------
// The edge store is PACKED from the start (fastb layout: 4 bases per byte, every edge byte aligned, walk orientation):
// a rank ORs the bases of its own segments into a zeroed store, 16 bases per atomic; the stores of all ranks add up
// (all-reduce) to the complete one.  `slot` = 4 * byte offset of the edge + base index.
struct PkWriter {
    uint32_t* W; uint64_t word; uint32_t acc; bool any;
    __device__ __forceinline__ void put(uint64_t slot, uint32_t code)
    {
        const uint64_t w = slot >> 4;
        if (any && w != word) { if (acc) atomicOr(W + word, acc); acc = 0; }
        word = w; any = true; acc |= code << (2u * (uint32_t)(slot & 15u));
    }
    __device__ __forceinline__ void flush() { if (any && acc) atomicOr(W + word, acc); acc = 0; }
};
// thread per LOCAL stop: its own k-mer, then the k-mers up to (not including) the next stop; `sinfo` is this rank's
// slice of the global stop information, `segs` its local segments
static __global__ void __launch_bounds__(128) k_seg_emit2(DictEntry* tab, const Link2* __restrict__ links, const uint32_t* __restrict__ stops, uint32_t n_stops,
                                                          const StopInfo* __restrict__ sinfo, const Seg* __restrict__ segs, const uint64_t* __restrict__ etmp_off, uint32_t* __restrict__ store)
{
    const uint32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
    if (sid >= n_stops) return;
    StopInfo si = sinfo[sid];
    if (si.edge == 0u) return;                                    // (edge + 1; 0 = on no edge yet)
    si.edge -= 1u;
    uint32_t cur = stops[sid], o = si.off_o & 1u, off = si.off_o >> 1;
    const uint64_t s0 = 4ull * etmp_off[si.edge];
    PkWriter pw; pw.W = store; pw.word = 0; pw.acc = 0; pw.any = false;
    if (off == 0) {                                               // the owner: all K bases of its k-mer, in walk orientation
        Kmer km = entry_kmer(tab[cur]);
        if (o) km = kmer_rc(km);
        for (int b = 0; b < SN_K; ++b) pw.put(s0 + b, kmer_base(km, b));
    } else pw.put(s0 + SN_K - 1 + off, step_base(tab[cur], o));
    tab[cur].edge = si.edge; tab[cur].off = off;
    const Seg sg = segs[2 * sid + o];
    if (sg.next != SN_NO_LINK) {
        const uint32_t steps = sg.steps_o >> 1;
        for (uint32_t k = 1; k < steps; ++k) {
            const Link2 lk = links[cur];
            const uint32_t l = o ? lk.y : lk.x;
            cur = l >> 1; o = l & 1u;
            pw.put(s0 + SN_K - 1 + off + k, step_base(tab[cur], o));
            tab[cur].edge = si.edge; tab[cur].off = off + k;
        }
    }
    pw.flush();
}
// ---- circles without any stop: a few k-mers, all on this rank (a link to another rank makes a stop) -----------------
// the walker with the smallest table index completes the loop; the circle is owned by its smallest K-MER, where
// canonicalizeCircle (BuildReadQGraph48.cc:375-397) starts it
static __global__ void __launch_bounds__(128) k_lc_count(const DictEntry* __restrict__ tab, const Link2* __restrict__ links, uint32_t n,
                                                         uint8_t* __restrict__ etype, uint32_t* __restrict__ own_n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (etype[i] != T_INTERIOR || tab[i].edge != SN_NULL_EDGE) return;
    uint32_t m = i, cur = i, o = 0, nk = 1; Kmer mk = entry_kmer(tab[i]);
    for (;;) {
        const uint32_t l = o ? links[cur].y : links[cur].x;
        if (l == SN_NO_LINK || (l >> 1) >= n) return;            // (not a local circle: cannot happen)
        cur = l >> 1; o = l & 1u;
        if (cur == i) break;
        if (cur < i) return;
        const Kmer q = entry_kmer(tab[cur]);
        if (q < mk) { mk = q; m = cur; }
        ++nk;
    }
    own_n[m] = nk; etype[m] = T_CIRCLE;
}
static __global__ void __launch_bounds__(256) k_lc_sizes(const uint32_t* __restrict__ own_n, const uint8_t* __restrict__ etype, uint32_t n, uint32_t* __restrict__ ebytes, uint32_t* __restrict__ eflag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = etype[i] == T_CIRCLE ? own_n[i] : 0u;
    ebytes[i] = k ? (k + (SN_K - 1) + 3u) >> 2 : 0u;
    eflag[i] = k ? 1u : 0u;
}
// thread per such circle: its owner walks it; packed bases into this rank's private store, final edge ids into the dictionary
static __global__ void __launch_bounds__(64) k_lc_emit(DictEntry* tab, const Link2* __restrict__ links, const uint32_t* __restrict__ owners, uint32_t n_owners, uint32_t edge0,
                                                       const uint32_t* __restrict__ own_n, const uint64_t* __restrict__ base_off, uint8_t* __restrict__ store /* zeroed */, uint32_t* __restrict__ len_out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_owners) return;
    const uint32_t i = owners[k], e = edge0 + k;
    len_out[k] = own_n[i] + SN_K - 1;
    uint8_t* s = store + base_off[i];
    const Kmer km = entry_kmer(tab[i]);
    for (int b = 0; b < SN_K; ++b) s[b >> 2] |= (uint8_t)(kmer_base(km, b) << (2 * (b & 3)));
    tab[i].edge = e; tab[i].off = 0;
    walk_circle_links(links, i, false, [&](uint32_t step, uint32_t j, uint32_t o) {
        const uint32_t p = SN_K - 1 + step;
        s[p >> 2] |= (uint8_t)(step_base(tab[j], o) << (2 * (p & 3)));
        tab[j].edge = e; tab[j].off = step; });
}
// offsets of edges [e0, e0 + m) laid out back to back from `at` (m is tiny)
static __global__ void k_lc_offsets(const uint32_t* __restrict__ elen, uint64_t* __restrict__ etmp_off, uint32_t e0, uint32_t m, uint64_t at)
{
    if (blockIdx.x || threadIdx.x) return;
    for (uint32_t k = 0; k < m; ++k) { etmp_off[e0 + k] = at; at += (elen[e0 + k] + 3u) >> 2; }
}

// canonicalizeCircle (BuildReadQGraph48.cc:375-397) on an assembled circle: the edge must start at the smallest
// canonical k-mer of the circle, read in that k-mer's canonical orientation.  Thread per circle (circles are rare).
// rot[c] = offset of that k-mer in the sequence as assembled | (it is seen reverse-complemented) << 31.
static __global__ void __launch_bounds__(64) k_circle_canon(const uint8_t* __restrict__ store, const uint64_t* __restrict__ etmp_off, const uint32_t* __restrict__ elen,
                                                            uint32_t edge0, uint32_t n_circles, uint8_t* __restrict__ out /* same layout, offsets relative to the first circle; zeroed */,
                                                            uint32_t* __restrict__ rot)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_circles) return;
    const uint32_t e = edge0 + c, n = elen[e] - (SN_K - 1);
    const uint8_t* S = store + etmp_off[e];
    Kmer f; f.w0 = f.w1 = f.w2 = 0;
    for (int b = 0; b < SN_K; ++b) f = kmer_succ(f, packed_base(S, b));
    Kmer best = f; uint32_t p = 0, rev = 0;
    { Kmer r; if (kmer_form(f, &r) == REV) { best = r; rev = 1; } }
    for (uint32_t j = 1; j < n; ++j) {
        f = kmer_succ(f, packed_base(S, SN_K - 1 + j));
        Kmer r; const bool isrev = kmer_form(f, &r) == REV;
        const Kmer& cand = isrev ? r : f;
        if (cand < best) { best = cand; p = j; rev = isrev ? 1u : 0u; }
    }
    rot[c] = p | (rev << 31);
    uint8_t* T = out + (etmp_off[e] - etmp_off[edge0]);
    const uint32_t len = n + SN_K - 1, q = n - 1 - p;
    for (uint32_t i = 0; i < len; ++i) {
        const uint32_t code = !rev ? packed_base(S, (p + i) % n) : 3u - packed_base(S, n + SN_K - 2 - (q + i) % n);
        T[i >> 2] |= (uint8_t)(code << (2 * (i & 3)));
    }
}
// whole-edge canonical form over the packed store
static __global__ void __launch_bounds__(128) k_edge_form_pk(const uint8_t* __restrict__ store, const uint64_t* __restrict__ etmp_off, const uint32_t* __restrict__ elen,
                                                             uint32_t n_edges, uint8_t* __restrict__ eflip)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) eflip[e] = seq_form_packed(store + etmp_off[e], elen[e]) == REV ? 1 : 0;
}
// the final edge store: same layout, every edge in its canonical orientation (one thread per output byte)
static __global__ void __launch_bounds__(256) k_orient_edges(const uint8_t* __restrict__ store, const uint64_t* __restrict__ eoff, const uint32_t* __restrict__ elen,
                                                             const uint8_t* __restrict__ eflip, uint32_t n_edges, uint64_t total_bytes, uint8_t* __restrict__ packed)
{
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= total_bytes) return;
    uint32_t lo = 0, hi = n_edges;                     // largest e with eoff[e] <= x
    while (hi - lo > 1) { const uint32_t m = (lo + hi) >> 1; if (eoff[m] <= x) lo = m; else hi = m; }
    const uint32_t e = lo;
    if (!eflip[e]) { packed[x] = store[x]; return; }
    const uint32_t len = elen[e], b0 = (uint32_t)(x - eoff[e]) * 4;
    const uint8_t* s = store + eoff[e];
    uint32_t v = 0;
    for (uint32_t j = 0; j < 4 && b0 + j < len; ++j) v |= (packed_base(s, len - 1 - (b0 + j)) ^ 3u) << (2 * j);
    packed[x] = (uint8_t)v;
}
// offsets of the members of rotated circles, then of the edges stored reverse-complemented
static __global__ void __launch_bounds__(256) k_fix_offsets2(DictEntry* tab, uint32_t n, const uint32_t* __restrict__ elen, const uint8_t* __restrict__ eflip,
                                                             uint32_t circle0, uint32_t n_circles, const uint32_t* __restrict__ rot)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = tab[i].edge;
    if (e == SN_NULL_EDGE) return;
    const uint32_t nk = elen[e] - (SN_K - 1);
    uint32_t off = tab[i].off;
    if (e - circle0 < n_circles) {
        const uint32_t r = rot[e - circle0], p = r & 0x7FFFFFFFu;
        off = (r >> 31) ? (p + nk - off) % nk : (off + nk - p) % nk;
    }
    if (eflip[e]) off = nk - 1 - off;
    tab[i].off = off;
}

}  // namespace sn
