// sn_dfside.cu -- what DF does with the graph and the paths right after the hot path (10X/DF.cc:573-600,
// 10X/WriteFiles.cc:16-60): ReadPathVecX, MarkDups and the files next to a.hbv.  Kernels: sn_dfside.cuh.
#include "sn_dfside.cuh"
#include "sn_ctx.h"

using namespace sn;

namespace {

HbvView hbv_view(sn_ctx* c)
{
    HbvView h;
    h.fwd_xlat = c->d_fwd.as<int32_t>(); h.rev_xlat = c->d_rev.as<int32_t>();
    h.to_left = c->d_toleft.as<int32_t>(); h.to_right = c->d_toright.as<int32_t>(); h.src = c->d_src.as<uint32_t>();
    h.from_start = c->d_from_start.as<uint32_t>(); h.from_v = c->d_from_v.as<int32_t>(); h.from_e = c->d_from_e.as<int32_t>();
    h.to_start = c->d_to_start.as<uint32_t>(); h.to_v = c->d_to_v.as<int32_t>(); h.to_e = c->d_to_e.as<int32_t>();
    return h;
}
ReadsView reads_view(sn_ctx* c)
{
    ReadsView rv;
    rv.n_reads = c->cnt.n_reads; rv.bases = c->bases.as<uint8_t>(); rv.boff = c->boff.as<uint64_t>(); rv.len = c->len.as<uint32_t>();
    rv.quals = c->have_pq ? nullptr : c->quals.as<uint8_t>(); rv.qoff = c->qoff.as<uint64_t>();
    rv.pq = c->have_pq ? c->pq.as<uint8_t>() : nullptr; rv.pq_off = c->pqoff.as<uint64_t>();
    return rv;
}

}  // namespace

extern "C" {

int sn_build_pathsx(sn_ctx* c)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 5) return fail(c, SN_ERR_STATE, "sn_build_pathsx: run sn_path_reads first");
    CU(cudaSetDevice(c->device));
    const uint64_t n = c->cnt.n_reads, n_index = (n + 9) / 10;
    DevBuf &sz = c->pool["px_sz"], &off = c->pool["px_off"], &data = c->pool["px_data"], &index = c->pool["px_index"];
    CU(sz.alloc(4 * n + 16)); CU(off.alloc(8 * (n + 1))); CU(index.alloc(8 * n_index + 16));
    t_begin(c, "pathsx");
    uint64_t total = 0;
    if (n) {
        k_rpx_sizes<<<blocks_for(n, 256), 256, 0, c->st>>>(c->path_off.as<uint64_t>(), n, sz.as<uint32_t>());
        KCHECK("k_rpx_sizes");
    }
    { int r = scan_u32(c, sz.as<uint32_t>(), n, off.as<uint64_t>(), &total); if (r) return r; }
    CU(data.alloc(total + 16));
    if (n) {
        k_rpx_encode<<<blocks_for(n, 256), 256, 0, c->st>>>(c->pedges.as<int32_t>(), c->path_off.as<uint64_t>(), c->poffset.as<int32_t>(), n, hbv_view(c),
                                                           off.as<uint64_t>(), data.as<uint8_t>(), index.as<long long>());
        KCHECK("k_rpx_encode");
    }
    t_end(c, "pathsx");
    resize_pinned(c, c->px_data, total); resize_pinned(c, c->px_index, n_index);
    if (total) CU(cudaMemcpyAsync(c->px_data.data(), data.p, total, cudaMemcpyDeviceToHost, c->st));
    if (n_index) CU(cudaMemcpyAsync(c->px_index.data(), index.p, 8 * n_index, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    c->px_ready = true;
    return SN_OK;
}
int sn_get_pathsx(sn_ctx* c, uint64_t* n_index, const int64_t** zip_index, uint64_t* n_bytes, const uint8_t** zipped_data)
{
    if (!c) return SN_ERR_ARG;
    if (!c->px_ready) return fail(c, SN_ERR_STATE, "run sn_build_pathsx first");
    if (n_index) *n_index = c->px_index.size();
    if (zip_index) *zip_index = c->px_index.data();
    if (n_bytes) *n_bytes = c->px_data.size();
    if (zipped_data) *zipped_data = c->px_data.data();
    return SN_OK;
}
int sn_write_pathsx(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (!c->px_ready) return fail(c, SN_ERR_STATE, "run sn_build_pathsx first");
    std::string err;
    if (!snf::write_pathsx(path, c->cnt.n_reads, c->px_index.data(), c->px_index.size(), c->px_data.data(), c->px_data.size(), err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}

int sn_mark_dups(sn_ctx* c, sn_dup_stats* stats)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 5) return fail(c, SN_ERR_STATE, "sn_mark_dups: run sn_path_reads first");
    const uint64_t n64 = c->cnt.n_reads;
    if (n64 & 1) return fail(c, SN_ERR_ARG, "sn_mark_dups: the reads are not pairs (odd number of reads)");
    if (n64 >= (1ull << 32)) return fail(c, SN_ERR_ARG, "sn_mark_dups: more than 2^32-1 reads in one context");
    CU(cudaSetDevice(c->device));
    const uint32_t n = (uint32_t)n64; const uint64_t np = n64 / 2;
    DevBuf &ra = c->pool["pi_a"], &rb = c->pool["pi_b"], &tmp = c->pool["rs_tmp"], &qs = c->pool["md_qsum"], &fl = c->pool["md_flags"], &tg = c->pool["md_tie"],
           &dup = c->pool["md_dup"], &art = c->pool["md_art"];
    CU(ra.alloc(16ull * n + 16)); CU(rb.alloc(16ull * n + 16)); CU(tmp.alloc(radix_sort_tmp_bytes(n))); CU(qs.alloc(4ull * n + 16)); CU(fl.alloc(n + 16ull));
    CU(tg.alloc(4ull * n + 16)); CU(dup.alloc(np + 16)); CU(art.alloc(np + 16));
    unsigned long long* ctr = c->counters.as<unsigned long long>() + 24;      // [24] ndups, [25] interdups
    const ReadsView rv = reads_view(c);
    t_begin(c, "mark_dups");
    CU(cudaMemsetAsync(dup.p, 0, np + 16, c->st)); CU(cudaMemsetAsync(art.p, 0, np + 16, c->st)); CU(cudaMemsetAsync(tg.p, 0, 4ull * n + 16, c->st));
    CU(cudaMemsetAsync(ctr, 0, 16, c->st));
    if (n) {
        k_md_records<<<blocks_for(n, 256), 256, 0, c->st>>>(rv, c->pedges.as<int32_t>(), c->path_off.as<uint64_t>(), c->poffset.as<int32_t>(), ra.as<uint4>());
        KCHECK("k_md_records");
        cudaError_t e = radix_sort<RS_KEY96>(ra.as<uint4>(), rb.as<uint4>(), n, tmp.p, c->num_sms, c->st);       // 12 passes: the result is back in ra
        if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, cudaGetErrorString(e));
        c->launches += 13;
        k_md_members<<<blocks_for(n, 128), 128, 0, c->st>>>(rv, ra.as<uint4>(), n, qs.as<uint32_t>(), fl.as<uint8_t>());
        KCHECK("k_md_members");
        k_md_groups<<<blocks_for(n, 128), 128, 0, c->st>>>(ra.as<uint4>(), n, qs.as<uint32_t>(), fl.as<uint8_t>(), c->have_bc ? c->bc.as<int32_t>() : nullptr,
                                                          dup.as<uint8_t>(), tg.as<uint32_t>(), ctr);
        KCHECK("k_md_groups");
        // the artifactual duplicates of the groups that saw a tie: the records are rebuilt in rb, sorted there (the result of
        // an even number of passes ends where it started)
        k_md_art_records<<<blocks_for(n, 128), 128, 0, c->st>>>(rv, ra.as<uint4>(), n, tg.as<uint32_t>(), rb.as<uint4>());
        KCHECK("k_md_art_records");
        e = radix_sort<RS_KEY96>(rb.as<uint4>(), ra.as<uint4>(), n, tmp.p, c->num_sms, c->st);
        if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, cudaGetErrorString(e));
        c->launches += 13;
        k_md_art<<<blocks_for(n, 128), 128, 0, c->st>>>(rv, rb.as<uint4>(), n, art.as<uint8_t>());
        KCHECK("k_md_art");
    }
    t_end(c, "mark_dups");
    resize_pinned(c, c->md_dup, np); resize_pinned(c, c->md_art, np);
    unsigned long long hc[2] = {0, 0};
    if (np) { CU(cudaMemcpyAsync(c->md_dup.data(), dup.p, np, cudaMemcpyDeviceToHost, c->st)); CU(cudaMemcpyAsync(c->md_art.data(), art.p, np, cudaMemcpyDeviceToHost, c->st)); }
    CU(cudaMemcpyAsync(hc, ctr, 16, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    sn_dup_stats& s = c->md_stats;
    s.n_pairs = np; s.n_dups = hc[0]; s.n_interdups = hc[1]; s.n_dup_pairs = 0; s.n_art_pairs = 0;
    for (uint64_t p = 0; p < np; ++p) { s.n_dup_pairs += c->md_dup[p]; s.n_art_pairs += c->md_art[p]; }
    c->md_ready = true;
    if (stats) *stats = s;
    return SN_OK;
}
int sn_get_dups(sn_ctx* c, uint8_t* dup, uint8_t* art)
{
    if (!c) return SN_ERR_ARG;
    if (!c->md_ready) return fail(c, SN_ERR_STATE, "run sn_mark_dups first");
    if (dup && !c->md_dup.empty()) memcpy(dup, c->md_dup.data(), c->md_dup.size());
    if (art && !c->md_art.empty()) memcpy(art, c->md_art.data(), c->md_art.size());
    return SN_OK;
}
int sn_write_dup(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (!c->md_ready) return fail(c, SN_ERR_STATE, "run sn_mark_dups first");
    std::string err;
    if (!snf::write_vec_u8(path, c->md_dup.data(), c->md_dup.size(), err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}

// ---- the files WriteAssemblyFiles leaves next to a.hbv (10X/WriteFiles.cc:33-51) --------------------------------------
int sn_write_hbx(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    { int r = sn_i_fetch_edges_host(c); if (r) return r; if ((r = sn_i_fetch_hbv_host(c))) return r; }
    std::string err; const snh::Hbv& H = c->hbv;
    std::vector<uint8_t> ep; std::vector<uint64_t> eo; std::vector<uint32_t> el;
    snh::hbv_edge_sequences(c->hedges, H, ep, eo, el);
    if (!snf::write_hbx(path, H.K, (uint64_t)H.n_vert, H.from_start.data(), H.from_v.data(), H.from_e.data(), H.to_start.data(), H.to_v.data(), H.to_e.data(),
                        ep.data(), eo.data(), el.data(), el.size(), H.to_left.data(), H.to_right.data(), err))
        return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
int sn_write_edges_fastb(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    { int r = sn_i_fetch_edges_host(c); if (r) return r; if ((r = sn_i_fetch_hbv_host(c))) return r; }
    std::string err;
    snf::Fastb fb;
    snh::hbv_edge_sequences(c->hedges, c->hbv, fb.var, fb.off, fb.len);
    fb.var.resize(fb.off.back());                      // (hbv_edge_sequences pads its buffer)
    if (!snf::write_fastb(path, fb, err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
int sn_write_kmers(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    { int r = sn_i_fetch_edges_host(c); if (r) return r; if ((r = sn_i_fetch_hbv_host(c))) return r; }
    const snh::Hbv& H = c->hbv;
    std::vector<int32_t> km(H.src.size());
    for (size_t e = 0; e < km.size(); ++e) km[e] = (int32_t)c->hedges.len[H.src[e] >> 1] - H.K + 1;      // HyperBasevector::Kmers(e)
    std::string err;
    if (!snf::write_vec_int(path, km, err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
int sn_write_k(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    std::string err;
    if (!snf::write_text(path, std::to_string(c->hbv.K) + "\n", err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}

}  // extern "C"
