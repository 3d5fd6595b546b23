// sn_multi.cu -- the multi-GPU path with the collectives issued from the C++ host (sn_comm.h).
#include "sn_kernels.cuh"
#include "sn_msp.cuh"
#include "sn_edges2.cuh"
#include "sn_ctx.h"

using namespace sn;

extern "C" {

int sn_nccl_unique_id(void* out128)
{
    std::string err;
    if (!out128) return SN_ERR_ARG;
    if (!snc::nccl_unique_id(out128, err)) { g_sn_create_error = err; return SN_ERR_CUDA; }
    return SN_OK;
}
int sn_comm_init_nccl(sn_ctx* c, int rank, int n_ranks, const void* unique_id128)
{
    if (!c || !unique_id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return SN_ERR_ARG;
    CU(cudaSetDevice(c->device));
    delete c->comm; c->comm = nullptr;
    std::string err;
    c->comm = snc::make_nccl_comm(rank, n_ranks, unique_id128, err);
    if (!c->comm) return fail(c, SN_ERR_CUDA, err);
    return SN_OK;
}
void* sn_local_group_create(int n_ranks) { return snc::local_group_create(n_ranks); }
void sn_local_group_destroy(void* g) { snc::local_group_destroy(static_cast<snc::LocalGroup*>(g)); }
void sn_local_group_abort(void* g) { snc::local_group_abort(static_cast<snc::LocalGroup*>(g)); }
int sn_comm_init_local(sn_ctx* c, void* group, int rank)
{
    if (!c || !group) return SN_ERR_ARG;
    delete c->comm;
    c->comm = snc::make_local_comm(static_cast<snc::LocalGroup*>(group), rank);
    if (!c->comm) return fail(c, SN_ERR_ARG, "sn_comm_init_local: rank out of range");
    return SN_OK;
}
void sn_comm_free(sn_ctx* c) { if (c) { delete c->comm; c->comm = nullptr; } }
int sn_mg_dict_is_sharded(const sn_ctx* c) { return c && c->dict_sharded ? 1 : 0; }

}  // extern "C"

// =====================================================================================================================
// small host values across the ranks (through a device staging buffer: the collectives move device memory)
namespace {

struct MgScratch {
    DevBuf& dev; HostBuf& host;
};
int comm_fail(sn_ctx* c, const char* what) { return fail(c, SN_ERR_CUDA, std::string(what) + ": " + (c->comm ? c->comm->err : std::string("no communicator"))); }

// every rank contributes `k` u64 values; all[r * k + i] = value i of rank r
int allgather_u64(sn_ctx* c, const uint64_t* mine, uint32_t k, std::vector<uint64_t>& all)
{
    const int n = c->comm ? c->comm->n : 1, rank = c->comm ? c->comm->rank : 0;
    all.assign((size_t)n * k, 0);
    if (n == 1) { for (uint32_t i = 0; i < k; ++i) all[i] = mine[i]; return SN_OK; }
    DevBuf& d = c->pool["mg_small"];
    CU(d.alloc(8ull * n * k));
    CU(cudaMemcpyAsync(d.as<uint64_t>() + (size_t)rank * k, mine, 8ull * k, cudaMemcpyHostToDevice, c->st));
    std::vector<size_t> bytes(n, 8ull * k), off(n);
    for (int r = 0; r < n; ++r) off[r] = 8ull * k * r;
    if (c->comm->allgatherv(d.as<uint64_t>() + (size_t)rank * k, d.p, bytes.data(), off.data(), c->st)) return comm_fail(c, "allgather");
    CU(cudaMemcpyAsync(all.data(), d.p, 8ull * n * k, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return SN_OK;
}
// uneven allgather of device arrays: `mine` (n_mine elements of elem bytes) -> `out` (all ranks back to back, rank order);
// counts[r] = elements of rank r (already known on every rank)
int allgatherv_dev(sn_ctx* c, const void* mine, void* out, const std::vector<uint64_t>& counts, size_t elem)
{
    const int n = c->comm ? c->comm->n : 1, rank = c->comm ? c->comm->rank : 0;
    std::vector<size_t> bytes(n), off(n);
    size_t at = 0;
    for (int r = 0; r < n; ++r) { bytes[r] = counts[r] * elem; off[r] = at; at += bytes[r]; }
    if (n == 1) { if (bytes[0] && mine != out) CU(cudaMemcpyAsync(out, mine, bytes[0], cudaMemcpyDeviceToDevice, c->st)); return SN_OK; }
    (void)rank;
    if (c->comm->allgatherv(mine, out, bytes.data(), off.data(), c->st)) return comm_fail(c, "allgatherv");
    return SN_OK;
}

// a bigger allocation that keeps the first `used` bytes
int grow_keep(sn_ctx* c, DevBuf& b, size_t used, size_t want)
{
    if (b.cap >= want) { b.bytes = want; return SN_OK; }
    DevBuf bigger;
    CU(bigger.alloc(want));
    CU(cudaMemsetAsync(bigger.p, 0, want, c->st));
    if (used) CU(cudaMemcpyAsync(bigger.p, b.p, used, cudaMemcpyDeviceToDevice, c->st));
    CU(cudaStreamSynchronize(c->st));
    std::swap(b.p, bigger.p); std::swap(b.bytes, bigger.bytes); std::swap(b.cap, bigger.cap);
    return SN_OK;
}

}  // namespace

// =====================================================================================================================
// recomputeAdjacencies + buildEdges over a dictionary that may be one rank's shard (sn_edges2.cuh).  Leaves, on EVERY
// rank: all edges (device + host, as sn_build_edges does) and, for the local k-mers, pruned context, edge and offset.
extern "C" int sn_i_build_edges2(sn_ctx* c)
{
    CU(cudaSetDevice(c->device));
    if (c->edges_copy_inflight) { cudaEventSynchronize(c->ev_edges); c->edges_copy_inflight = false; }     // (an earlier step's edges still on their way)
    const int NR = c->comm ? c->comm->n : 1, rank = c->comm ? c->comm->rank : 0;
    const uint32_t n = (uint32_t)c->cnt.n_kmers;
    c->cnt.n_edges = 0; c->cnt.n_edge_bases = 0;
    DictEntry* tab = c->dict.as<DictEntry>();         // (re-read after the ghost table has grown)
    uint32_t* u32c = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8);        // [4] circles found, [20..23] ghost counters / errors
    int r;
    // ---- ghosts: the neighbours that live on other ranks ----------------------------------------------------------
    t_begin(c, "ghosts");
    DevBuf &qk = c->pool["gh_qk"], &qslot = c->pool["gh_qslot"], &rk = c->pool["gh_rk"], &rans = c->pool["gh_rans"], &ans = c->pool["gh_ans"],
           &gcnt = c->pool["gh_cnt"], &gbase = c->pool["gh_base"], &gctx_s = c->pool["gh_ctx_s"], &gctx_r = c->pool["gh_ctx_r"];
    std::vector<size_t> q_sb(NR, 0), q_so(NR, 0), q_rb(NR, 0), q_ro(NR, 0);         // query exchange: send/recv counts (in queries) and offsets
    uint64_t nq_send = 0, nq_recv = 0;
    if (NR > 1) {
        if (!c->ghost_cap) return fail(c, SN_ERR_STATE, "sharded dictionary without a ghost region");
        CU(gcnt.alloc(8ull * NR)); CU(gbase.alloc(4ull * NR));
        DevBuf& glist = c->pool["gh_list"];
        std::vector<uint32_t> h_cnt(NR); uint32_t h_g[2] = {0, 0};
        for (int attempt = 0;; ++attempt) {
            DictView dv = dict_view(c);
            CU(glist.alloc(4ull * (c->ghost_cap / 4 * 3 + 1)));
            CU(cudaMemsetAsync(c->dict.as<DictEntry>() + n, 0, (size_t)c->ghost_cap * sizeof(DictEntry), c->st));
            CU(cudaMemsetAsync(u32c + 20, 0, 16, c->st));
            CU(cudaMemsetAsync(gcnt.p, 0, 8ull * NR, c->st));
            if (n) {
                k_ghost_collect<<<blocks_for(n, 256), 256, 0, c->st>>>(c->dict.as<DictEntry>(), dv, (uint32_t)NR, u32c + 20, u32c + 21, glist.as<uint32_t>(), gcnt.as<uint32_t>());
                KCHECK("k_ghost_collect");
            }
            CU(cudaMemcpyAsync(h_cnt.data(), gcnt.p, 4ull * NR, cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(h_g, u32c + 20, 8, cudaMemcpyDeviceToHost, c->st));
            CU(cudaStreamSynchronize(c->st));
            if (!h_g[1]) break;
            // more neighbours on other ranks than the table was sized for: a bigger one (the local entries move along)
            if (attempt >= 3 || getenv("SN_GHOST_CAP") || (uint64_t)n + 4ull * c->ghost_cap >= (1ull << 31))
                return fail(c, SN_ERR_DATA, "the ghost table overflowed (" + std::to_string(h_g[0]) + " remote neighbours for " + std::to_string(c->ghost_cap) + " slots)");
            if ((r = grow_keep(c, c->dict, (size_t)n * sizeof(DictEntry), ((size_t)n + 4ull * c->ghost_cap) * sizeof(DictEntry) + 64))) return r;
            c->ghost_cap *= 4;
        }
        tab = c->dict.as<DictEntry>();
        const uint32_t n_gh = h_g[0];
        std::vector<uint32_t> h_base(NR);
        for (int d = 0; d < NR; ++d) { h_base[d] = (uint32_t)nq_send; q_sb[d] = h_cnt[d]; q_so[d] = nq_send; nq_send += h_cnt[d]; }
        // who asks me how much: every rank's per-destination counts
        std::vector<uint64_t> mine(NR), all;
        for (int d = 0; d < NR; ++d) mine[d] = h_cnt[d];
        if ((r = allgather_u64(c, mine.data(), (uint32_t)NR, all))) return r;
        for (int s = 0; s < NR; ++s) { q_rb[s] = all[(size_t)s * NR + rank]; q_ro[s] = nq_recv; nq_recv += q_rb[s]; }
        CU(qk.alloc(12 * nq_send + 16)); CU(qslot.alloc(4 * nq_send + 16)); CU(rk.alloc(12 * nq_recv + 16)); CU(rans.alloc(4 * nq_recv + 16)); CU(ans.alloc(4 * nq_send + 16));
        CU(cudaMemcpyAsync(gbase.p, h_base.data(), 4ull * NR, cudaMemcpyHostToDevice, c->st));
        CU(cudaMemsetAsync(gcnt.as<uint32_t>() + NR, 0, 4ull * NR, c->st));
        if (n_gh) { k_ghost_fill<<<blocks_for(n_gh, 256), 256, 0, c->st>>>(tab + n, glist.as<uint32_t>(), n_gh, gbase.as<uint32_t>(), gcnt.as<uint32_t>() + NR, qk.as<uint32_t>(), qslot.as<uint32_t>()); KCHECK("k_ghost_fill"); }
        auto scaled = [&](const std::vector<size_t>& v, size_t f) { std::vector<size_t> o(v.size()); for (size_t i = 0; i < v.size(); ++i) o[i] = v[i] * f; return o; };
        {   // queries out (12 bytes each) ...
            auto sb = scaled(q_sb, 12), so = scaled(q_so, 12), rb = scaled(q_rb, 12), ro = scaled(q_ro, 12);
            if (c->comm->alltoallv(qk.p, sb.data(), so.data(), rk.p, rb.data(), ro.data(), c->st)) return comm_fail(c, "alltoallv (ghost queries)");
        }
        DictView own = dict_view(c); own.g_cap = 0;
        if (nq_recv) { k_ghost_answer<<<blocks_for(nq_recv, 256), 256, 0, c->st>>>(own, rk.as<uint32_t>(), (uint32_t)nq_recv, rans.as<uint32_t>()); KCHECK("k_ghost_answer"); }
        {   // ... answers back (4 bytes each, the same sizes reversed)
            auto sb = scaled(q_rb, 4), so = scaled(q_ro, 4), rb = scaled(q_sb, 4), ro = scaled(q_so, 4);
            if (c->comm->alltoallv(rans.p, sb.data(), so.data(), ans.p, rb.data(), ro.data(), c->st)) return comm_fail(c, "alltoallv (ghost answers)");
        }
        if (nq_send) { k_ghost_apply<<<blocks_for(nq_send, 256), 256, 0, c->st>>>(tab + n, qslot.as<uint32_t>(), ans.as<uint32_t>(), (uint32_t)nq_send); KCHECK("k_ghost_apply"); }
    }
    t_end(c, "ghosts");
    if (!n && NR == 1) { c->edges_host_stale = false; resize_pinned(c, c->hedges.len, 0); resize_pinned(c, c->hedges.off, 1); c->hedges.off[0] = 0; resize_pinned(c, c->hedges.packed, 16); c->stage = 3; return SN_OK; }
    const DictView dv = dict_view(c);
    // ---- recomputeAdjacencies ---------------------------------------------------------------------------------------
    t_begin(c, "prune");
    DevBuf& links = c->pool["links"];
    CU(links.alloc(8ull * n + 16));
    if (n) { k_prune<<<blocks_for(n, 256), 256, 0, c->st>>>(tab, dv, links.as<Link2>()); KCHECK("k_prune"); }
    if (NR > 1) {   // the pruned contexts of the ghosts, from their owners
        CU(gctx_s.alloc(nq_recv + 16)); CU(gctx_r.alloc(nq_send + 16));
        if (nq_recv) { k_ghost_ctx_send<<<blocks_for(nq_recv, 256), 256, 0, c->st>>>(tab, rans.as<uint32_t>(), (uint32_t)nq_recv, gctx_s.as<uint8_t>()); KCHECK("k_ghost_ctx_send"); }
        if (c->comm->alltoallv(gctx_s.p, q_rb.data(), q_ro.data(), gctx_r.p, q_sb.data(), q_so.data(), c->st)) return comm_fail(c, "alltoallv (ghost contexts)");
        if (nq_send) { k_ghost_ctx_apply<<<blocks_for(nq_send, 256), 256, 0, c->st>>>(tab + n, qslot.as<uint32_t>(), gctx_r.as<uint8_t>(), (uint32_t)nq_send); KCHECK("k_ghost_ctx_apply"); }
    }
    t_end(c, "prune");
    // ---- links, stops, local segments -----------------------------------------------------------------------------
    t_begin(c, "edges");
    DevBuf &etype = c->pool["etype"], &flag = c->pool["flag"], &stop_pos = c->pool["stop_pos"], &stops = c->pool["stops"], &segs = c->pool["segs"], &stype = c->pool["stype"];
    CU(etype.alloc(n + 16)); CU(flag.alloc(4ull * n + 16)); CU(stop_pos.alloc(8ull * (n + 1)));
    if (n) { k_classify2<<<blocks_for(n, 256), 256, 0, c->st>>>(dv, links.as<Link2>(), etype.as<uint8_t>(), flag.as<uint32_t>()); KCHECK("k_classify2"); }
    uint64_t n_stops = 0;
    if ((r = scan_u32(c, flag.as<uint32_t>(), n, stop_pos.as<uint64_t>(), &n_stops))) return r;
    if (n_stops >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 stops on one rank");
    CU(stops.alloc(4 * n_stops + 16)); CU(segs.alloc(16 * n_stops + 16)); CU(stype.alloc(n_stops + 16));
    if (n_stops) {
        k_scatter_flagged<<<blocks_for(n, 256), 256, 0, c->st>>>(flag.as<uint32_t>(), stop_pos.as<uint64_t>(), n, stops.as<uint32_t>());
        KCHECK("k_scatter_flagged");
        k_seg_walk2<<<blocks_for(2 * n_stops, 128), 128, 0, c->st>>>(links.as<Link2>(), stops.as<uint32_t>(), (uint32_t)n_stops, stop_pos.as<uint64_t>(), n, segs.as<Seg>());
        KCHECK("k_seg_walk2");
        k_stop_types<<<blocks_for(n_stops, 256), 256, 0, c->st>>>(stops.as<uint32_t>(), etype.as<uint8_t>(), (uint32_t)n_stops, stype.as<uint8_t>());
        KCHECK("k_stop_types");
    }
    // ---- the stop table of all ranks --------------------------------------------------------------------------------
    std::vector<uint64_t> cnt_all, mine1(1, n_stops);
    if ((r = allgather_u64(c, mine1.data(), 1, cnt_all))) return r;
    std::vector<uint64_t> sbase(NR + 1, 0);
    for (int q = 0; q < NR; ++q) sbase[q + 1] = sbase[q] + cnt_all[q];
    const uint64_t S = sbase[NR];
    if (S >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 stops over all ranks");
    DevBuf *g_segs = &segs, *g_stype = &stype;
    DevBuf &gstops = c->pool["g_stops"], &gsegs = c->pool["g_segs"], &gstype = c->pool["g_stype"], &sbase_d = c->pool["g_sbase"], &lsegs = c->pool["g_lsegs"];
    if (NR > 1) {
        CU(gstops.alloc(4 * S + 16)); CU(gsegs.alloc(16 * S + 16)); CU(gstype.alloc(S + 16)); CU(sbase_d.alloc(8ull * (NR + 1))); CU(lsegs.alloc(16 * n_stops + 16));
        CU(cudaMemcpyAsync(sbase_d.p, sbase.data(), 8ull * (NR + 1), cudaMemcpyHostToDevice, c->st));
        if ((r = allgatherv_dev(c, stops.p, gstops.p, cnt_all, 4))) return r;
        if ((r = allgatherv_dev(c, stype.p, gstype.p, cnt_all, 1))) return r;
        CU(cudaMemsetAsync(u32c + 22, 0, 4, c->st));
        if (n_stops) {
            k_seg_globalize<<<blocks_for(2 * n_stops, 256), 256, 0, c->st>>>(segs.as<Seg>(), (uint32_t)n_stops, tab + n, gstops.as<uint32_t>(), sbase_d.as<uint64_t>(), (uint32_t)rank,
                lsegs.as<Seg>(), u32c + 22);
            KCHECK("k_seg_globalize");
        }
        if ((r = allgatherv_dev(c, lsegs.p, gsegs.p, cnt_all, 16))) return r;
        uint32_t h_err = 0;
        CU(cudaMemcpyAsync(&h_err, u32c + 22, 4, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        if (h_err) return fail(c, SN_ERR_DATA, "a chain continues on a k-mer its owner does not list as a stop (internal error)");
        g_segs = &gsegs; g_stype = &gstype;
    }
    // ---- lengths, owners, edge ids and offsets -------------------------------------------------------------------------------
    // Every rank hops only for the edges whose OWNER (the smaller end stop) is one of its stops -- 1/N of the hops over the
    // gathered table.  Edge ids are global by construction (owners in stop order = rank order): a rank's edges are one
    // contiguous id range; lengths and store offsets are gathered, the {edge, offset} of the stops is summed (all-reduce of
    // arrays that are zero wherever another rank's edge passes).
    const uint32_t L = (uint32_t)n_stops, first = (uint32_t)sbase[rank];
    DevBuf &own_n = c->pool["g_own_n"], &eb32 = c->pool["g_eb32"], &eflag = c->pool["g_eflag"], &base_off = c->pool["g_base_off"], &pos = c->pool["g_pos"],
           &owners = c->pool["owners"], &sinfo = c->pool["sinfo"], &sinfo2 = c->pool["sinfo2"], &elen_l = c->pool["g_elen_l"], &eoff_l = c->pool["g_eoff_l"];
    CU(own_n.alloc(4ull * L + 16)); CU(eb32.alloc(4ull * L + 16)); CU(eflag.alloc(4ull * L + 16)); CU(base_off.alloc(8ull * (L + 1))); CU(pos.alloc(8ull * (L + 1)));
    CU(sinfo.alloc(8 * S + 16)); CU(owners.alloc(4ull * L + 16));
    const Seg* SG = g_segs->as<Seg>(); uint8_t* ST = g_stype->as<uint8_t>();
    // (sizes below are BYTES of the packed edge store: 4 bases per byte, every edge byte aligned)
    uint64_t total_bytes_main = 0, n_edges = 0, circle_bytes = 0, n_circles = 0;
    DevBuf &store = c->pool["tmpb"], &eflip = c->pool["eflip"], &etmp_off = c->pool["etmp_off"];
    unsigned long long* n_on_edges = c->counters.as<unsigned long long>() + 5;
    CU(cudaMemsetAsync(n_on_edges, 0, 8, c->st));
    CU(cudaMemsetAsync(sinfo.p, 0, 8 * S + 16, c->st));
    // one phase: sizes of the edges my stops own -> their global ids and store offsets -> hop -> gather / sum
    auto phase = [&](int circles, uint64_t edge_base, uint64_t byte_base, DevBuf& si, uint64_t* n_out, uint64_t* bytes_out) -> int {
        uint64_t my_bytes = 0, my_n = 0;
        if (L) { k_gs_sizes<<<blocks_for(L, 256), 256, 0, c->st>>>(own_n.as<uint32_t>(), ST + first, circles, L, eb32.as<uint32_t>(), eflag.as<uint32_t>()); KCHECK("k_gs_sizes"); }
        int rr;
        if ((rr = scan_u32(c, eb32.as<uint32_t>(), L, base_off.as<uint64_t>(), &my_bytes))) return rr;
        if ((rr = scan_u32(c, eflag.as<uint32_t>(), L, pos.as<uint64_t>(), &my_n))) return rr;
        std::vector<uint64_t> all, mine = {my_n, my_bytes};
        if ((rr = allgather_u64(c, mine.data(), 2, all))) return rr;
        uint64_t tot_n = 0, tot_b = 0, e0 = 0, b0 = 0;
        std::vector<uint64_t> cnt(NR);
        for (int q = 0; q < NR; ++q) { if (q == rank) { e0 = tot_n; b0 = tot_b; } cnt[q] = all[2 * q]; tot_n += all[2 * q]; tot_b += all[2 * q + 1]; }
        *n_out = tot_n; *bytes_out = tot_b;
        if (edge_base + tot_n + 1 >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 edges");
        // room for the global per-edge arrays (keeps what earlier phases wrote)
        if ((rr = grow_keep(c, c->elen, 4 * edge_base, 4 * (edge_base + tot_n + 1) + 16)) || (rr = grow_keep(c, etmp_off, 8 * edge_base, 8 * (edge_base + tot_n + 2) + 16))) return rr;
        CU(elen_l.alloc(4 * my_n + 16)); CU(eoff_l.alloc(8 * my_n + 16));
        if (my_n) {
            k_scatter_flagged<<<blocks_for(L, 256), 256, 0, c->st>>>(eflag.as<uint32_t>(), pos.as<uint64_t>(), L, owners.as<uint32_t>());
            KCHECK("k_scatter_flagged");
            k_gs_owner_hop<<<blocks_for(my_n, 128), 128, 0, c->st>>>(SG, owners.as<uint32_t>(), (uint32_t)my_n, (uint32_t)(edge_base + e0), first, ST, own_n.as<uint32_t>(), base_off.as<uint64_t>(),
                byte_base + b0, elen_l.as<uint32_t>(), eoff_l.as<uint64_t>(), si.as<StopInfo>(), n_on_edges);
            KCHECK("k_gs_owner_hop");
        }
        if ((rr = allgatherv_dev(c, elen_l.p, c->elen.as<uint32_t>() + edge_base, cnt, 4))) return rr;
        if ((rr = allgatherv_dev(c, eoff_l.p, etmp_off.as<uint64_t>() + edge_base, cnt, 8))) return rr;
        if (NR > 1 && tot_n && c->comm->allreduce_sum(si.p, 2 * S, 4, c->st)) return comm_fail(c, "allreduce (stop info)");
        return SN_OK;
    };
    if (L) { k_gs_end_hop<<<blocks_for(L, 128), 128, 0, c->st>>>(SG, ST, first, L, own_n.as<uint32_t>()); KCHECK("k_gs_end_hop"); }
    CU(c->elen.alloc(16)); CU(etmp_off.alloc(16));
    if ((r = phase(0, 0, 0, sinfo, &n_edges, &total_bytes_main))) return r;
    // circles that hold a stop (simpleCircle, BuildReadQGraph48.cc:348-372): numbered after the edges above
    {
        CU(cudaMemsetAsync(u32c + 4, 0, 4, c->st));
        if (L) { k_gs_circle_elect<<<blocks_for(L, 128), 128, 0, c->st>>>(SG, ST, sinfo.as<StopInfo>(), first, L, own_n.as<uint32_t>(), u32c + 4); KCHECK("k_gs_circle_elect"); }
        uint32_t h_found = 0;
        CU(cudaMemcpyAsync(&h_found, u32c + 4, 4, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        std::vector<uint64_t> all, mine(1, h_found);
        if ((r = allgather_u64(c, mine.data(), 1, all))) return r;
        uint64_t any = 0; for (uint64_t v : all) any += v;
        if (any) {
            CU(sinfo2.alloc(8 * S + 16));
            CU(cudaMemsetAsync(sinfo2.p, 0, 8 * S + 16, c->st));
            if ((r = phase(1, n_edges, total_bytes_main, sinfo2, &n_circles, &circle_bytes))) return r;
            k_sinfo_add<<<blocks_for(S, 256), 256, 0, c->st>>>(sinfo.as<StopInfo>(), sinfo2.as<StopInfo>(), (uint32_t)S);
            KCHECK("k_sinfo_add");
        }
    }
    const uint64_t edge_cap = n_edges + n_circles + 1;
    CU(eflip.alloc(edge_cap + 16)); CU(c->eoff.alloc(8 * (edge_cap + 1)));
    const uint64_t circle0 = n_edges;                        // the circles with stops are the edges [circle0, circle0 + n_circles)
    const uint64_t main_bytes = total_bytes_main + circle_bytes;
    const uint64_t main_words = (main_bytes + 3) / 4;
    // ---- bases of the local segments; (edge, offset) of the local k-mers ---------------------------------------------------
    uint32_t h_unreached = 0;
    CU(store.alloc(4 * main_words + 64));
    CU(cudaMemsetAsync(store.p, 0, 4 * main_words + 64, c->st));
    if (n_stops) {
        k_seg_emit2<<<blocks_for(n_stops, 128), 128, 0, c->st>>>(tab, links.as<Link2>(), stops.as<uint32_t>(), (uint32_t)n_stops, sinfo.as<StopInfo>() + sbase[rank], segs.as<Seg>(),
            etmp_off.as<uint64_t>(), store.as<uint32_t>());
        KCHECK("k_seg_emit2");
    }
    if (n) {
        CU(cudaMemsetAsync(u32c + 23, 0, 4, c->st));
        k_count_unreached<<<blocks_for(n, 256), 256, 0, c->st>>>(tab, etype.as<uint8_t>(), n, u32c + 23);
        KCHECK("k_count_unreached");
        CU(cudaMemcpyAsync(&h_unreached, u32c + 23, 4, cudaMemcpyDeviceToHost, c->st));
    }
    if (NR > 1 && c->comm->allreduce_sum(store.p, main_words, 4, c->st)) return comm_fail(c, "allreduce (edge bases)");     // disjoint bits: the sum is the union
    unsigned long long h_on_edges = 0;
    CU(cudaMemcpyAsync(&h_on_edges, n_on_edges, 8, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    // ---- circles without a stop: local by construction (a link to another rank makes a stop) --------------------------------
    std::vector<uint64_t> lc_all, lc_mine(3, 0);          // per rank: {circles, bytes, k-mers}
    DevBuf &lc_own = c->pool["lc_own_n"], &lc_eb = c->pool["lc_eb"], &lc_flag = c->pool["lc_flag"], &lc_boff = c->pool["lc_boff"], &lc_pos = c->pool["lc_pos"],
           &lc_owners = c->pool["lc_owners"], &lc_tmp = c->pool["lc_tmp"], &lc_len = c->pool["lc_len"];
    uint64_t lc_n = 0, lc_bytes = 0;
    if (h_unreached) {
        CU(lc_own.alloc(4ull * n)); CU(lc_eb.alloc(4ull * n)); CU(lc_flag.alloc(4ull * n)); CU(lc_boff.alloc(8ull * (n + 1))); CU(lc_pos.alloc(8ull * (n + 1)));
        CU(cudaMemsetAsync(lc_own.p, 0, 4ull * n, c->st));
        k_lc_count<<<blocks_for(n, 128), 128, 0, c->st>>>(tab, links.as<Link2>(), n, etype.as<uint8_t>(), lc_own.as<uint32_t>());
        KCHECK("k_lc_count");
        k_lc_sizes<<<blocks_for(n, 256), 256, 0, c->st>>>(lc_own.as<uint32_t>(), etype.as<uint8_t>(), n, lc_eb.as<uint32_t>(), lc_flag.as<uint32_t>());
        KCHECK("k_lc_sizes");
        if ((r = scan_u32(c, lc_eb.as<uint32_t>(), n, lc_boff.as<uint64_t>(), &lc_bytes))) return r;
        if ((r = scan_u32(c, lc_flag.as<uint32_t>(), n, lc_pos.as<uint64_t>(), &lc_n))) return r;
        if (!lc_n) return fail(c, SN_ERR_DATA, "k-mers on no edge and on no circle (internal error)");
    }
    lc_mine[0] = lc_n; lc_mine[1] = lc_bytes; lc_mine[2] = (uint64_t)h_unreached + (uint64_t)h_on_edges;      // k-mers this rank's edges and circles hold
    if ((r = allgather_u64(c, lc_mine.data(), 3, lc_all))) return r;
    uint64_t lc_tot_n = 0, lc_tot_b = 0, lc_tot_k = 0, lc_e0 = 0;
    std::vector<uint64_t> lc_cnt(NR), lc_bcnt(NR);
    for (int q = 0; q < NR; ++q) { if (q == rank) lc_e0 = lc_tot_n; lc_cnt[q] = lc_all[3 * q]; lc_bcnt[q] = lc_all[3 * q + 1]; lc_tot_n += lc_cnt[q]; lc_tot_b += lc_bcnt[q]; lc_tot_k += lc_all[3 * q + 2]; }
    const uint64_t E = n_edges + n_circles + lc_tot_n, all_bytes = main_bytes + lc_tot_b;
    if (E >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 edges");
    if (E + 1 > edge_cap) {        // the per-edge arrays were sized before the stop-less circles were known
        const uint64_t used = n_edges + n_circles;
        if ((r = grow_keep(c, c->elen, 4 * used, 4 * (E + 1) + 16)) || (r = grow_keep(c, etmp_off, 8 * used, 8 * (E + 2) + 16))) return r;
        CU(eflip.alloc(E + 17)); CU(c->eoff.alloc(8 * (E + 2)));
    }
    if (lc_tot_n) {
        // the local circles' bases and lengths are gathered behind the rest of the edge store
        if ((r = grow_keep(c, store, main_bytes, all_bytes + 64))) return r;
        CU(lc_tmp.alloc(lc_bytes + 64)); CU(lc_len.alloc(4 * (lc_n + 1))); CU(lc_owners.alloc(4 * (lc_n + 1)));
        CU(cudaMemsetAsync(lc_tmp.p, 0, lc_bytes + 64, c->st));
        if (lc_n) {
            k_scatter_flagged<<<blocks_for(n, 256), 256, 0, c->st>>>(lc_flag.as<uint32_t>(), lc_pos.as<uint64_t>(), n, lc_owners.as<uint32_t>());
            KCHECK("k_scatter_flagged");
            k_lc_emit<<<blocks_for(lc_n, 64), 64, 0, c->st>>>(tab, links.as<Link2>(), lc_owners.as<uint32_t>(), (uint32_t)lc_n, (uint32_t)(n_edges + n_circles + lc_e0),
                lc_own.as<uint32_t>(), lc_boff.as<uint64_t>(), lc_tmp.as<uint8_t>(), lc_len.as<uint32_t>());
            KCHECK("k_lc_emit");
        }
        if ((r = allgatherv_dev(c, lc_tmp.p, store.as<uint8_t>() + main_bytes, lc_bcnt, 1))) return r;
        if ((r = allgatherv_dev(c, lc_len.p, c->elen.as<uint32_t>() + n_edges + n_circles, lc_cnt, 4))) return r;
        k_lc_offsets<<<1, 1, 0, c->st>>>(c->elen.as<uint32_t>(), etmp_off.as<uint64_t>(), (uint32_t)(n_edges + n_circles), (uint32_t)lc_tot_n, main_bytes);
        KCHECK("k_lc_offsets");
    }
    // ---- circles with stops start where canonicalizeCircle starts them ------------------------------------------------------
    DevBuf &crot = c->pool["crot"], &ctmp = c->pool["ctmp"];
    CU(crot.alloc(4 * n_circles + 16));
    if (n_circles) {
        CU(ctmp.alloc(circle_bytes + 64));
        CU(cudaMemsetAsync(ctmp.p, 0, circle_bytes + 64, c->st));
        k_circle_canon<<<blocks_for(n_circles, 64), 64, 0, c->st>>>(store.as<uint8_t>(), etmp_off.as<uint64_t>(), c->elen.as<uint32_t>(), (uint32_t)circle0, (uint32_t)n_circles,
            ctmp.as<uint8_t>(), crot.as<uint32_t>());
        KCHECK("k_circle_canon");
        CU(cudaMemcpyAsync(store.as<uint8_t>() + total_bytes_main, ctmp.p, circle_bytes, cudaMemcpyDeviceToDevice, c->st));
    }
    // ---- canonical orientation of every edge, offsets, final store (same layout as the walk-orientation store) --------------------
    const uint64_t total_bytes = all_bytes;
    CU(cudaMemcpyAsync(etmp_off.as<uint64_t>() + E, &total_bytes, 8, cudaMemcpyHostToDevice, c->st));
    CU(c->ebases.alloc(total_bytes + 64));
    CU(cudaMemsetAsync((char*)c->ebases.p + total_bytes, 0, 64, c->st));
    if (E) {
        k_edge_form_pk<<<blocks_for(E, 128), 128, 0, c->st>>>(store.as<uint8_t>(), etmp_off.as<uint64_t>(), c->elen.as<uint32_t>(), (uint32_t)E, eflip.as<uint8_t>());
        KCHECK("k_edge_form_pk");
        if (n) { k_fix_offsets2<<<blocks_for(n, 256), 256, 0, c->st>>>(tab, n, c->elen.as<uint32_t>(), eflip.as<uint8_t>(), (uint32_t)circle0, (uint32_t)n_circles, crot.as<uint32_t>()); KCHECK("k_fix_offsets2"); }
        k_orient_edges<<<blocks_for(total_bytes, 256), 256, 0, c->st>>>(store.as<uint8_t>(), etmp_off.as<uint64_t>(), c->elen.as<uint32_t>(), eflip.as<uint8_t>(),
            (uint32_t)E, total_bytes, c->ebases.as<uint8_t>());
        KCHECK("k_orient_edges");
    }
    CU(cudaMemcpyAsync(c->eoff.p, etmp_off.p, 8 * (E + 1), cudaMemcpyDeviceToDevice, c->st));
    t_end(c, "edges");
    const uint64_t kmers_on_edges = lc_tot_k;
    c->cnt.n_edges = E; c->cnt.n_edge_bases = kmers_on_edges + (uint64_t)(SN_K - 1) * E;
    // the edges in host memory: on one rank of a multi-GPU job; the others keep them on the device until asked
    // (the copy runs on the second stream, under the HBV stage; sn_build_hbv / the getters complete it)
    c->edges_host_stale = true;
    if (NR == 1 || rank == 0) { if ((r = sn_i_start_edges_copy(c, total_bytes))) return r; }
    CU(cudaStreamSynchronize(c->st));
    // every dictionary k-mer of every rank sits on exactly one edge
    std::vector<uint64_t> nk_all, nk_mine(1, n);
    if ((r = allgather_u64(c, nk_mine.data(), 1, nk_all))) return r;
    uint64_t n_total = 0; for (uint64_t v : nk_all) n_total += v;
    c->mg_n_kmers_total = n_total;
    if (kmers_on_edges != n_total)
        return fail(c, SN_ERR_DATA, "edge stage covered " + std::to_string(kmers_on_edges) + " of " + std::to_string(n_total) + " dictionary k-mers");
    c->stage = 3;
    // a job counted in several passes is a big one: what this stage leaves behind per k-mer is not needed again
    if (c->cnt.n_kmer_occurrences >= 3600000000ull) pool_release(c, {"links", "stop_pos", "flag", "etype"});
    return SN_OK;
}

// =====================================================================================================================
// The whole hot path over the ranks of this context's communicator (sn_comm_init_*): reads sharded over the ranks,
// super-k-mers routed to the owner of their minimizer bucket by one alltoallv (owner(b) = b * N >> bits: a rank's
// buckets are one contiguous range; the role of `shard % total_chunks`, lib/tada/src/cmd_shard_asm.rs:40), count +
// filter per owner, and from there a dictionary that STAYS sharded: ghosts / stops / bases as in sn_edges2.cuh.
// Every rank ends up with all edges and the whole HyperBasevector; the k-mer table stays distributed unless ReadPaths
// are asked for (then the finished entries are gathered, 32 bytes per k-mer, and every rank paths its own reads).
namespace {

int mg_count_sharded(sn_ctx* c)
{
    const int NR = c->comm->n, rank = c->comm->rank;
    int r; uint64_t n_occ = 0, n_sk = 0;
    if ((r = sn_i_count_goodlen(c, &n_occ))) return r;
    std::vector<uint64_t> all, mine(1, n_occ);
    if ((r = allgather_u64(c, mine.data(), 1, all))) return r;
    uint64_t occ_total = 0; for (uint64_t v : all) occ_total += v;
    int bits = sn_i_pick_bucket_bits(occ_total);
    while ((1u << bits) < (uint32_t)NR) ++bits;
    // A rank counts what it RECEIVES: about occ_total / NR occurrences.  Above 2^32 of them the count runs in passes, like on one
    // GPU -- but every pass must take a slice of EVERY owner's bucket range, so the passes interleave: the global bucket is
    // owner | within, the top bits of `within` name the pass (msp_window_bucket).  That needs a power-of-two number of ranks.
    uint32_t lp = 0;
    while ((occ_total / (uint64_t)NR + 1) / (1ull << lp) >= 3400000000ull) ++lp;
    if (const char* e = getenv("SN_MG_PASSES")) { int v = atoi(e); lp = 0; while ((1 << lp) < v) ++lp; }     // tests
    uint32_t lnr = 0; while ((1u << lnr) < (uint32_t)NR) ++lnr;
    if (lp && (1u << lnr) != (uint32_t)NR) return fail(c, SN_ERR_ARG, "a sharded count in passes (more than 2^32 k-mer occurrences per rank) needs a power-of-two number of ranks");
    if (lp) while ((uint32_t)bits < lnr + lp + 1) ++bits;
    const uint32_t P = 1u << lp, wb = (uint32_t)bits - lnr;
    // this pass's buckets, renumbered so that owner o's share is [fb[o], fb[o + 1])
    std::vector<uint32_t> fb(NR + 1);
    for (int o = 0; o <= NR; ++o) fb[o] = lp ? (uint32_t)o << (wb - lp) : sn_i_first_bucket((uint32_t)o, (uint32_t)NR, bits);
    const uint32_t nbl = fb[rank + 1] - fb[rank];                       // buckets this rank receives per pass
    DevBuf &recs = c->pool["mg_recs"], &cnts = c->pool["mg_counts"], &roff = c->pool["mg_off"];
    DevBuf &surv = c->pool["surv_a"], &surv_off = c->pool["surv_off"], &acc = c->pool["surv_all"], &acc_cnt = c->pool["surv_all_cnt"];
    if (lp) CU(acc_cnt.alloc(4ull * P * nbl + 16));
    uint64_t n_acc = 0, n_dist = 0, n_recv_total = 0, n_surv = 0;
    unsigned long long* occ = c->counters.as<unsigned long long>();
    for (uint32_t ps = 0; ps < P; ++ps) {
        // super-k-mers of the local reads for this pass, bucket order: an owner's buckets are one contiguous range of the record array
        if ((r = sn_i_msp_partition(c, bits, &n_sk, 0, 0, lp ? (wb | lp << 8 | ps << 16) : 0u))) return r;
        const uint64_t* off = c->pool["sk_off"].as<uint64_t>();
        std::vector<uint64_t> cut(NR + 1);
        for (int o = 0; o <= NR; ++o) CU(cudaMemcpyAsync(&cut[o], off + fb[o], 8, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        std::vector<uint64_t> send_n(NR);
        for (int o = 0; o < NR; ++o) send_n[o] = cut[o + 1] - cut[o];
        if ((r = allgather_u64(c, send_n.data(), (uint32_t)NR, all))) return r;
        std::vector<size_t> sb(NR), so(NR), rb(NR), ro(NR);
        uint64_t n_recv = 0;
        for (int s2 = 0; s2 < NR; ++s2) { sb[s2] = 32 * send_n[s2]; so[s2] = 32 * cut[s2]; rb[s2] = 32 * all[(size_t)s2 * NR + rank]; ro[s2] = 32 * n_recv; n_recv += all[(size_t)s2 * NR + rank]; }
        t_begin(c, "exchange");
        CU(recs.alloc(std::max<uint64_t>(n_recv, 1) * 32 + 64)); CU(cnts.alloc(4ull * NR * nbl + 16)); CU(roff.alloc(8ull * ((uint64_t)NR * nbl + 1)));
        if (c->comm->alltoallv(c->pool["sk_recs"].p, sb.data(), so.data(), recs.p, rb.data(), ro.data(), c->st)) return comm_fail(c, "alltoallv (super-k-mer records)");
        // the per-bucket record counts of the same ranges (after the scatter the per-bucket cursors equal the counts)
        for (int s2 = 0; s2 < NR; ++s2) { sb[s2] = 4ull * (fb[s2 + 1] - fb[s2]); so[s2] = 4ull * fb[s2]; rb[s2] = 4ull * nbl; ro[s2] = 4ull * nbl * s2; }
        if (c->comm->alltoallv(c->pool["sk_hist"].p, sb.data(), so.data(), cnts.p, rb.data(), ro.data(), c->st)) return comm_fail(c, "alltoallv (bucket counts)");
        t_end(c, "exchange");
        const uint64_t n_cnt = (uint64_t)NR * nbl;
        uint64_t total = 0;
        if ((r = scan_u32(c, cnts.as<uint32_t>(), n_cnt, roff.as<uint64_t>(), &total))) return r;
        if (total != n_recv) return fail(c, SN_ERR_DATA, "received per-bucket counts do not add up to the received records");
        CU(cudaMemsetAsync(occ + 4, 0, 8, c->st));
        if (n_recv) { k_sum_nk<<<std::min(blocks_for(n_recv, 256), 8u * (unsigned)c->num_sms), 256, 0, c->st>>>(recs.as<uint4>(), n_recv, occ + 4); KCHECK("k_sum_nk"); }
        unsigned long long h_occ = 0;
        CU(cudaMemcpyAsync(&h_occ, occ + 4, 8, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        if (h_occ >= (1ull << 32)) return fail(c, SN_ERR_ARG, "a rank received more than 2^32-1 k-mer occurrences in one pass (skewed buckets): raise SN_MG_PASSES");
        if ((r = sn_i_msp_bucket_count(c, recs.as<uint4>(), roff.as<uint64_t>(), nbl, (uint32_t)NR, h_occ, surv, surv_off, &n_surv))) return r;
        n_recv_total += n_recv; n_dist += c->cnt.n_kmers_distinct;
        if (!lp) break;
        if (16 * (n_acc + n_surv) + 64 > acc.cap) {                    // grow towards the expected final size, keeping what is there
            const uint64_t guess = (uint64_t)((double)(n_acc + n_surv) * (double)P / (double)(ps + 1) * 1.03) + 1024;
            if ((r = grow_keep(c, acc, 16 * n_acc, 16 * std::max<uint64_t>(n_acc + n_surv, guess) + 64))) return r;
        }
        if (n_surv) CU(cudaMemcpyAsync(acc.as<uint4>() + n_acc, surv.p, 16 * n_surv, cudaMemcpyDeviceToDevice, c->st));
        k_diff_u32<<<blocks_for(nbl, 256), 256, 0, c->st>>>(surv_off.as<uint32_t>(), nbl, acc_cnt.as<uint32_t>() + (uint64_t)ps * nbl);
        KCHECK("k_diff_u32");
        n_acc += n_surv;
    }
    c->cnt.n_superkmers = n_recv_total;
    const uint64_t n_mine = lp ? n_acc : n_surv;
    const uint32_t nb_mine = P * nbl;                                    // the rank's whole bucket window (the passes' slices in order ARE bucket order)
    // the rank's shard of the dictionary, with room for its ghosts: ~0.1 neighbours per k-mer live in other buckets
    // (those whose minimizer differs), (N-1)/N of them on other ranks; the table starts at n/4 slots and grows if it must
    uint32_t cap = 1024;
    while (cap < n_mine / 4 + 1024) cap <<= 1;
    if (const char* e = getenv("SN_GHOST_CAP")) { uint32_t v = (uint32_t)atoll(e); if (v >= 64 && !(v & (v - 1))) cap = v; }     // tests: force overflow handling
    if (lp) {
        pool_release(c, {"sk_recs", "sk_off", "sk_dsc", "sk_nruns", "sk_hist", "surv_a", "surv_off", "surv_scratch", "mg_recs", "mg_counts", "mg_off"});
        if ((r = sn_i_msp_install_dict(c, acc.as<uint4>(), n_mine, bits, acc_cnt.as<uint32_t>(), false, nb_mine, cap))) return r;
        pool_release(c, {"surv_all", "surv_all_cnt"});
        c->cnt.n_kmers_distinct = n_dist;
    } else if ((r = sn_i_msp_install_dict(c, surv.as<uint4>(), n_mine, bits, surv_off.as<uint32_t>(), true, nbl, cap))) return r;
    c->dict_b_lo = lp ? (uint32_t)rank << wb : fb[rank]; c->dict_b_n = nb_mine; c->ghost_cap = cap; c->dict_sharded = true;
    return SN_OK;
}

// the finished dictionary entries of all ranks, in rank = bucket order: what a single GPU would hold
int mg_replicate_dict(sn_ctx* c)
{
    const int NR = c->comm->n, rank = c->comm->rank;
    int r;
    const uint64_t n = c->cnt.n_kmers;
    std::vector<uint64_t> all, mine(1, n);
    if ((r = allgather_u64(c, mine.data(), 1, all))) return r;
    uint64_t total = 0, my_base = 0;
    for (int q = 0; q < NR; ++q) { if (q == rank) my_base = total; total += all[q]; }
    if (total >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 dictionary k-mers: ReadPaths over a replicated dictionary do not fit one context");
    const int bits = c->dict_bits;
    const uint64_t nb = 1ull << bits;
    DevBuf &full = c->pool["dict_full"], &bcnt_l = c->pool["mg_bcnt_l"], &bcnt = c->pool["surv_gcnt"];
    CU(full.alloc(total * sizeof(DictEntry) + 64));
    if ((r = allgatherv_dev(c, c->dict.p, full.p, all, sizeof(DictEntry)))) return r;
    // bucket counts of every rank's window -> global bucket offsets
    std::vector<uint64_t> wn(NR);
    for (int q = 0; q < NR; ++q) wn[q] = sn_i_first_bucket((uint32_t)q + 1, (uint32_t)NR, bits) - sn_i_first_bucket((uint32_t)q, (uint32_t)NR, bits);
    CU(bcnt_l.alloc(4ull * c->dict_b_n + 16)); CU(bcnt.alloc(4 * nb + 16));
    k_diff_u32<<<blocks_for(c->dict_b_n, 256), 256, 0, c->st>>>(c->dboff.as<uint32_t>(), c->dict_b_n, bcnt_l.as<uint32_t>());
    KCHECK("k_diff_u32");
    if ((r = allgatherv_dev(c, bcnt_l.p, bcnt.p, wn, 4))) return r;
    CU(cudaStreamSynchronize(c->st));
    // swap the full table in (the shard is not needed any more: its entries are in the full table)
    std::swap(c->dict.p, full.p); std::swap(c->dict.bytes, full.bytes); std::swap(c->dict.cap, full.cap);
    c->cnt.n_kmers = total; c->dict_b_lo = 0; c->dict_b_n = 0; c->ghost_cap = 0; c->dict_sharded = false;
    CU(c->dict_hs.alloc(4 * total + 64));
    if (total) { k_dict_hs<<<blocks_for(total, 256), 256, 0, c->st>>>(c->dict.as<DictEntry>(), (uint32_t)total, c->dict_hs.as<uint32_t>()); KCHECK("k_dict_hs"); }
    CU(c->dboff.alloc(4 * (nb + 1)));
    DevBuf& o64 = c->pool["boff64"];
    CU(o64.alloc(8 * (nb + 1)));
    if ((r = scan_u32(c, bcnt.as<uint32_t>(), nb, o64.as<uint64_t>(), nullptr))) return r;
    k_narrow_u64<<<blocks_for(nb + 1, 256), 256, 0, c->st>>>(o64.as<uint64_t>(), nb + 1, c->dboff.as<uint32_t>());
    KCHECK("k_narrow_u64");
    int sub = 0;
    while (sub < 6 && (nb << sub) < (1ull << 26) && total / (nb << sub) > 32) ++sub;
    c->dict_sub_bits = sub;
    if (sub) {
        DevBuf& cells = c->pool["dict_cells"];
        CU(cells.alloc(4 * ((nb << sub) + 1)));
        k_dict_cells<<<blocks_for((nb << sub) + 1, 256), 256, 0, c->st>>>(c->dict.as<DictEntry>(), c->dboff.as<uint32_t>(), (uint32_t)nb, sub, cells.as<uint32_t>());
        KCHECK("k_dict_cells");
    }
    CU(cudaStreamSynchronize(c->st));
    (void)my_base;
    return SN_OK;
}

}  // namespace

extern "C" int sn_mg_build_graph(sn_ctx* c, const sn_params* params, int with_paths)
{
    if (!c) return SN_ERR_ARG;
    if (!c->comm) return fail(c, SN_ERR_STATE, "sn_mg_build_graph: no communicator (sn_comm_init_nccl / sn_comm_init_local)");
    if (c->stage < 1) return fail(c, SN_ERR_STATE, "sn_mg_build_graph: no reads loaded");
    CU(cudaSetDevice(c->device));
    int r;
    if ((r = sn_i_count_set_params(c, params))) return r;
    if ((r = mg_count_sharded(c))) return r;
    if ((r = sn_i_build_edges2(c))) return r;
    if ((r = sn_build_hbv(c))) return r;
    if (with_paths) {
        if ((r = mg_replicate_dict(c))) return r;
        if ((r = sn_path_reads(c))) return r;
    }
    return SN_OK;
}
