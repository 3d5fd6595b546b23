// sn_ctx.h -- internal: the context behind the opaque sn_ctx of include/supernova_b200.h and the small host helpers
// every translation unit of the library shares (buffers that keep their allocation, timers, error plumbing).
#pragma once
#include "../../include/supernova_b200.h"
#include "sn_kmer.cuh"
#include "sn_formats.h"
#include "sn_hbv.h"
#include "sn_comm.h"
#include "sn_prims.cuh"

#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <map>
#include <string>
#include <vector>

extern std::string g_sn_create_error;

// Device buffer that keeps its allocation across steps: alloc() only goes to cudaMalloc when
// the request outgrows the capacity (cudaMalloc/cudaFree of multi-GB buffers cost
// milliseconds each and serialise the device).
struct DevBuf {
    void* p = nullptr; size_t bytes = 0, cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; cap = 0; }
    cudaError_t alloc(size_t n)
    {
        if (!n) n = 16;
        if (p && cap >= n) { bytes = n; return cudaSuccess; }
        release();
        size_t want = n + n / 16;                      // a little headroom so step-to-step jitter does not reallocate
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); want = n; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) { bytes = n; cap = want; }
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// pinned host buffer, kept across steps like DevBuf
struct HostBuf {
    void* p = nullptr; size_t cap = 0;
    HostBuf() = default;
    HostBuf(const HostBuf&) = delete; HostBuf& operator=(const HostBuf&) = delete;
    ~HostBuf() { if (p) cudaFreeHost(p); }
    cudaError_t alloc(size_t n)
    {
        if (!n) n = 64;
        if (p && cap >= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, n + n / 8);
        if (e == cudaSuccess) cap = n + n / 8;
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Timer { cudaEvent_t a = nullptr, b = nullptr; bool used = false; };

struct sn_ctx {
    int device = 0, num_sms = 148;
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;                 // copy stream of sn_load_reads_streamed
    cudaEvent_t ev_copy[16] = {};
    // work sn_load_reads_streamed already did under the copies; consumed by the next count call only
    bool gl_ready = false; uint32_t gl_min_qual = 0; uint64_t gl_occ = 0; int hist_ready_bits = -1;
    bool dsc_ready = false; uint32_t dsc_overflow = 0;      // run descriptors of the loaded reads under the current good lengths (k_msp_place)
    std::string err;
    uint64_t launches = 0;
    std::map<std::string, Timer> timers;
    std::map<std::string, double> host_ms;
    sn_params params{7, 3, 2, 0};
    uint32_t min_gl = SN_K + 1;      // reads trimmed below this give no k-mer: K + 1 (BuildReadQGraph48.cc:160), K with sn_set_semantics(SN_SEM_TADA)
    sn_counts cnt{};
    int stage = 0;       // 0 none, 1 reads, 2 counted, 3 edges, 4 hbv, 5 paths
    bool reads_ok = false;   // reads are resident (a graph can also come from an edge file without any)

    // reads
    DevBuf bases, boff, len, quals, qoff, bc, pq, pqoff, goodlen;
    bool have_bc = false, have_pq = false;
    // dictionary
    DevBuf dict, dboff;      // dictionary (bucket, hash, k-mer order) and its bucket offsets (2^dict_bits + 1)
    DevBuf dict_hs;          // the entries' hashes on their own (DictView::hs)
    int dict_bits = 4, dict_sub_bits = 0;
    // edges (device) + host copy
    DevBuf ebases, eoff, elen;
    snh::Edges hedges;
    // hbv
    snh::Hbv hbv;
    DevBuf d_fwd, d_rev, d_toleft, d_toright, d_src, d_from_start, d_from_v, d_from_e, d_to_start, d_to_v, d_to_e;
    // paths
    DevBuf plen, poffset, path_off, pedges;
    std::vector<int32_t> h_poffset, h_pedges; std::vector<uint64_t> h_path_off; bool paths_on_host = false;
    std::vector<uint64_t> pi_off, pi_ids; std::vector<int32_t> pi_countsb; bool pi_ready = false;     // paths index (writePathsIndex)
    // DF side (sn_dfside.cu): the ReadPathVecX of the paths and MarkDups' flags, in host memory once built
    std::vector<uint8_t> px_data; std::vector<int64_t> px_index; bool px_ready = false;
    std::vector<uint8_t> md_dup, md_art; sn_dup_stats md_stats{}; bool md_ready = false;
    DevBuf counters;     // small scratch of u64 counters
    std::map<std::string, DevBuf> pool;      // stage temporaries, kept across steps
    std::map<std::string, HostBuf> hpool;    // pinned staging, kept across steps
    std::map<const void*, size_t> pinned;    // host vectors whose storage is page-locked (result arrays reused across steps)
    // multi-GPU (sn_multi.cu): the communicator this context is a rank of, and its shard of the dictionary
    snc::Comm* comm = nullptr;
    uint32_t dict_b_lo = 0, dict_b_n = 0;    // minimizer buckets [b_lo, b_lo + b_n) of 2^dict_bits held by `dict` (b_n = 0: all of them)
    uint32_t ghost_cap = 0;                  // ghost region behind the dictionary: remote neighbours of the local k-mers (open addressing, power of two)
    uint64_t mg_n_kmers_total = 0;           // dictionary size over all ranks
    bool dict_sharded = false;
    // Results a rank other than 0 leaves on the device until somebody asks for them (every rank computes the whole edge
    // set and HBV; the job needs them in host memory once): sn_i_fetch_edges_host / sn_i_fetch_hbv_host bring them over.
    bool edges_host_stale = false, hbv_host_stale = false;
    bool edges_copy_inflight = false;        // the edges are on their way to the host on st2 (under the HBV stage); ev_edges marks the end
    cudaEvent_t ev_edges = nullptr, ev_edges_go = nullptr;
};

namespace {

// Result arrays live in std::vectors that keep their storage from step to step; their storage is
// page-locked once (cudaHostRegister) so that the copies to and from them run at full PCIe speed
// and asynchronously.  resize_pinned never lets a registered block be freed behind CUDA's back.
template <class T> void resize_pinned(sn_ctx* c, std::vector<T>& v, size_t n)
{
    if (v.capacity() < n) {
        auto it = c->pinned.find(v.data());
        if (it != c->pinned.end()) { cudaHostUnregister(const_cast<void*>(it->first)); c->pinned.erase(it); }
        std::vector<T>().swap(v);
        v.reserve(n + n / 8 + 16);
    }
    v.resize(n);
    if (v.capacity() && !c->pinned.count(v.data())) {
        if (cudaHostRegister(v.data(), v.capacity() * sizeof(T), cudaHostRegisterDefault) == cudaSuccess) c->pinned[v.data()] = v.capacity() * sizeof(T);
        else cudaGetLastError();                       // not fatal: the copies fall back to pageable memory
    }
}
void unpin_all(sn_ctx* c) { for (auto& kv : c->pinned) cudaHostUnregister(const_cast<void*>(kv.first)); c->pinned.clear(); }

int fail(sn_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; else g_sn_create_error = msg; return code; }

#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return fail(c, SN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
#define KCHECK(name) do { ++c->launches; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) \
    return fail(c, SN_ERR_CUDA, std::string("launch ") + name + ": " + cudaGetErrorString(e__)); } while (0)

void t_begin(sn_ctx* c, const char* name)
{
    Timer& t = c->timers[name];
    if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
    cudaEventRecord(t.a, c->st); t.used = false;
}
void t_end(sn_ctx* c, const char* name) { Timer& t = c->timers[name]; if (!t.a) return; cudaEventRecord(t.b, c->st); t.used = true; }

// the dictionary of this context as the kernels see it (its bucket window and ghost region included)
inline sn::DictView dict_view(sn_ctx* c)
{
    sn::DictView d;
    d.tab = c->dict.as<sn::DictEntry>();
    d.boff = c->dict_sub_bits ? c->pool["dict_cells"].as<uint32_t>() : c->dboff.as<uint32_t>();
    d.n = (uint32_t)c->cnt.n_kmers; d.bits = c->dict_bits; d.sub_bits = c->dict_sub_bits;
    d.b_lo = c->dict_b_lo; d.b_n = c->dict_b_n ? c->dict_b_n : (1u << c->dict_bits);
    d.g_cap = c->ghost_cap;
    d.hs = c->dict_hs.as<uint32_t>();
    return d;
}

// SN_POOL_REPORT=1: the device buffers above 64 MB at a stage boundary, to stderr (sizing big jobs)
inline void pool_report(sn_ctx* c, const char* when)
{
    static const bool on = getenv("SN_POOL_REPORT") != nullptr;
    if (!on) return;
    size_t total = 0; std::string line;
    auto add = [&](const char* name, const DevBuf& b) { total += b.cap; if (b.cap >= (64u << 20)) line += std::string(" ") + name + "=" + std::to_string(b.cap >> 20); };
    add("bases", c->bases); add("boff", c->boff); add("len", c->len); add("quals", c->quals); add("bc", c->bc); add("pq", c->pq); add("pqoff", c->pqoff); add("goodlen", c->goodlen);
    add("dict", c->dict); add("dboff", c->dboff); add("dict_hs", c->dict_hs); add("ebases", c->ebases); add("eoff", c->eoff); add("elen", c->elen);
    add("plen", c->plen); add("poffset", c->poffset); add("path_off", c->path_off); add("pedges", c->pedges);
    for (auto& kv : c->pool) add(kv.first.c_str(), kv.second);
    fprintf(stderr, "[pools %s] total %zu MB:%s\n", when, total >> 20, line.c_str());
}
// frees stage temporaries by name (big jobs: what a finished stage leaves behind is not needed by the next)
inline void pool_release(sn_ctx* c, std::initializer_list<const char*> names)
{ for (const char* n : names) { auto it = c->pool.find(n); if (it != c->pool.end()) it->second.release(); } }

inline unsigned blocks_for(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

// exclusive scan helper that owns its temporaries; out has n+1 entries; total returned through *total (host, after sync)
int scan_u32(sn_ctx* c, const uint32_t* in, uint64_t n, uint64_t* out, uint64_t* total)
{
    DevBuf& tmp = c->pool["scan_tmp"];
    CU(tmp.alloc(sn::scan_tmp_words(n) * 8 + 16));
    sn::exclusive_scan_u32_u64(in, n, out, tmp.as<uint64_t>(), c->st);
    c->launches += n ? 3 : 0;
    CU(cudaGetLastError());
    if (total) { CU(cudaMemcpyAsync(total, out + n, 8, cudaMemcpyDeviceToHost, c->st)); }
    CU(cudaStreamSynchronize(c->st));
    return SN_OK;
}

int upload(sn_ctx* c, DevBuf& b, const void* src, size_t bytes, size_t pad = 0)
{
    CU(b.alloc(bytes + pad));
    if (pad) CU(cudaMemsetAsync((char*)b.p + bytes, 0, pad, c->st));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->st));
    return SN_OK;
}


}  // namespace

// ---- pieces of the count / HBV stages defined in sn_pipeline.cu and reused by the sharded path (sn_multi.cu) ----
extern "C" {
int sn_i_pick_bucket_bits(uint64_t n_occ);
uint32_t sn_i_first_bucket(uint32_t owner, uint32_t nparts, int bits);
int sn_i_count_set_params(sn_ctx* c, const sn_params* p);
int sn_i_count_goodlen(sn_ctx* c, uint64_t* n_occ_out);
int sn_i_msp_partition(sn_ctx* c, int bits, uint64_t* n_sk_out, uint32_t b_lo = 0, uint32_t b_n = 0, uint32_t pcfg = 0 /* interleaved pass: msp_window_bucket */);
int sn_i_msp_bucket_count(sn_ctx* c, const uint4* recs, const uint64_t* off, uint32_t n_buckets, uint32_t n_seg, uint64_t occ_bound,
                          DevBuf& surv, DevBuf& surv_off, uint64_t* n_surv_out);
// nb_window != 0: the table holds that many buckets only (a rank's shard); extra_entries: room behind the table (ghosts)
int sn_i_msp_install_dict(sn_ctx* c, const uint4* surv, uint64_t n_surv, int bits, const uint32_t* counts_or_off, bool is_offsets,
                          uint32_t nb_window = 0, uint64_t extra_entries = 0);
// recomputeAdjacencies + buildEdges over this context's (possibly sharded) dictionary: sn_multi.cu
int sn_i_build_edges2(sn_ctx* c);
int sn_i_fetch_edges_host(sn_ctx* c);
int sn_i_start_edges_copy(sn_ctx* c, uint64_t total_bytes);      // the same, asynchronously on the copy stream (sn_i_fetch_edges_host completes it)
int sn_i_fetch_hbv_host(sn_ctx* c);
}
