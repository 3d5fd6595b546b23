// sn_pipeline.cu -- host orchestration + extern "C" ABI (include/supernova_b200.h) of the
// B200-native hot path.  Mirrors the call sequence of buildReadQGraph48
// (paths/long/BuildReadQGraph48.cc:1688-1774): createDict -> buildEdges ->
// buildHBVFromEdges -> pathReads.  Everything heavy runs in the kernels of
// sn_kernels.cuh; the host keeps sizes, allocations and the sequential HBV numbering.
#include "../../include/supernova_b200.h"
#include "sn_kernels.cuh"
#include "sn_msp.cuh"
#include "sn_ingest.cuh"
#include "sn_synth.cuh"
#include "sn_hbvdev.cuh"
#include "sn_edict.cuh"
#include "sn_ctx.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <sys/mman.h>

using namespace sn;

std::string g_sn_create_error;

namespace {

int load_common(sn_ctx* c, uint64_t n_reads, const uint8_t* bases, const uint64_t* base_off, const uint32_t* len, const int32_t* bc)
{
    if (!n_reads || !bases || !base_off || !len) return fail(c, SN_ERR_ARG, "sn_load_reads: empty or NULL input");
    if (n_reads >= (1ull << 32)) return fail(c, SN_ERR_ARG, "sn_load_reads: more than 2^32-1 reads per context");
    c->cnt = sn_counts{}; c->stage = 0; c->reads_ok = false; c->paths_on_host = false;
    c->gl_ready = false; c->hist_ready_bits = -1; c->dsc_ready = false;
    c->cnt.n_reads = n_reads;
    t_begin(c, "h2d");
    int r;
    if ((r = upload(c, c->bases, bases, base_off[n_reads], 64))) return r;
    if ((r = upload(c, c->boff, base_off, 8 * (n_reads + 1)))) return r;
    if ((r = upload(c, c->len, len, 4 * n_reads))) return r;
    c->have_bc = bc != nullptr;
    if (bc && (r = upload(c, c->bc, bc, 4 * n_reads))) return r;
    return SN_OK;
}

int finish_load(sn_ctx* c)
{
    t_end(c, "h2d");
    const uint64_t n = c->cnt.n_reads;
    // limits are checked on the device (no pass over the caller's arrays on the host):
    // total bases, longest read, largest barcode ordinal
    unsigned long long* cnt64 = c->counters.as<unsigned long long>() + 16;     // [16] bases, [17] max len | max bc
    CU(cudaMemsetAsync(cnt64, 0, 16, c->st));
    k_read_stats<<<std::min(blocks_for(n, 256), 8u * (unsigned)c->num_sms), 256, 0, c->st>>>(n, c->len.as<uint32_t>(), c->have_bc ? c->bc.as<int32_t>() : nullptr,
        cnt64, reinterpret_cast<uint32_t*>(cnt64 + 1), reinterpret_cast<int32_t*>(cnt64 + 1) + 1);
    KCHECK("k_read_stats");
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, cnt64, 16, cudaMemcpyDeviceToHost, c->st));
    if (!c->have_pq) {
        // element offsets of the unpacked quals = exclusive scan of the read lengths
        CU(c->qoff.alloc(8 * (n + 1)));
        int r = scan_u32(c, c->len.as<uint32_t>(), n, c->qoff.as<uint64_t>(), nullptr);
        if (r) return r;
    }
    CU(cudaStreamSynchronize(c->st));                  // the caller's buffers are free again when this returns
    c->cnt.n_bases = h[0];
    const uint32_t max_len = (uint32_t)(h[1] & 0xFFFFFFFFu); const int32_t max_bc = (int32_t)(h[1] >> 32);
    // longest read must fit the per-thread buffers of the pathing kernel
    if (max_len > SN_MAX_READ_LEN) return fail(c, SN_ERR_ARG, "reads longer than " + std::to_string(SN_MAX_READ_LEN) + " bases are not supported");
    // barcode ordinals travel in 24 bits of a super-k-mer record (0xFFFFFF is reserved for "-1")
    if (max_bc >= 0xFFFFFF) return fail(c, SN_ERR_ARG, "more than 2^24-2 distinct barcodes in one context");
    c->stage = 1; c->reads_ok = true;
    return SN_OK;
}

}  // namespace

// =============================================================================
extern "C" {

int sn_msp_bucket_bits(uint64_t n_occ_total) { return msp_bucket_bits(n_occ_total); }
int sn_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

int sn_ctx_create(sn_ctx** out, int device)
{
    if (!out) return SN_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, SN_ERR_CUDA, std::string("no CUDA device: the hot path has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= n) return fail(nullptr, SN_ERR_ARG, "device index out of range");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, SN_ERR_CUDA, cudaGetErrorString(e));
    sn_ctx* c = new sn_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return fail(nullptr, SN_ERR_CUDA, cudaGetErrorString(e)); }
    if ((e = c->counters.alloc(256)) != cudaSuccess) { delete c; return fail(nullptr, SN_ERR_CUDA, cudaGetErrorString(e)); }
    *out = c;
    return SN_OK;
}
void sn_ctx_destroy(sn_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    for (auto& kv : c->timers) { if (kv.second.a) cudaEventDestroy(kv.second.a); if (kv.second.b) cudaEventDestroy(kv.second.b); }
    unpin_all(c);
    for (auto& e : c->ev_copy) if (e) cudaEventDestroy(e);
    if (c->ev_edges) { cudaEventDestroy(c->ev_edges); cudaEventDestroy(c->ev_edges_go); }
    if (c->st2) cudaStreamDestroy(c->st2);
    cudaStreamDestroy(c->st);
    delete c->comm;
    delete c;
}
const char* sn_last_error(const sn_ctx* c) { return c ? c->err.c_str() : g_sn_create_error.c_str(); }
uint64_t sn_kernel_launches(const sn_ctx* c) { return c ? c->launches : 0; }
void* sn_stream(const sn_ctx* c) { return c ? (void*)c->st : nullptr; }
double sn_stage_ms(const sn_ctx* c, const char* name)
{
    if (!c || !name) return -1.0;
    auto h = c->host_ms.find(name);
    if (h != c->host_ms.end()) return h->second;
    auto it = c->timers.find(name);
    if (it == c->timers.end() || !it->second.used) return -1.0;
    cudaEventSynchronize(it->second.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, it->second.a, it->second.b) != cudaSuccess) return -1.0;
    return (double)ms;
}

int sn_load_reads(sn_ctx* c, uint64_t n_reads, const uint8_t* bases, const uint64_t* base_off, const uint32_t* len,
                  const uint8_t* pq, const uint64_t* pq_off, const int32_t* bc)
{
    if (!c) return SN_ERR_ARG;
    CU(cudaSetDevice(c->device));
    int r = load_common(c, n_reads, bases, base_off, len, bc);
    if (r) return r;
    if (!pq || !pq_off) return fail(c, SN_ERR_ARG, "sn_load_reads: NULL quals");
    if ((r = upload(c, c->pq, pq, pq_off[n_reads], 16))) return r;
    if ((r = upload(c, c->pqoff, pq_off, 8 * (n_reads + 1)))) return r;
    c->have_pq = true;
    c->quals.release();
    return finish_load(c);
}
int sn_load_reads_q8(sn_ctx* c, uint64_t n_reads, const uint8_t* bases, const uint64_t* base_off, const uint32_t* len,
                     const uint8_t* quals, const uint64_t* qual_off, const int32_t* bc)
{
    if (!c) return SN_ERR_ARG;
    CU(cudaSetDevice(c->device));
    int r = load_common(c, n_reads, bases, base_off, len, bc);
    if (r) return r;
    if (!quals || !qual_off) return fail(c, SN_ERR_ARG, "sn_load_reads_q8: NULL quals");
    if ((r = upload(c, c->quals, quals, qual_off[n_reads], 16))) return r;
    c->have_pq = false;
    c->pq.release(); c->pqoff.release();
    return finish_load(c);
}
// ---- ingest on the device: barcoded pseudo-FASTQ text -> the context's reads (sn_ingest.cuh) ----
// What ParseBarcodedFastqs does (10X/ParseBarcodedFastqs.cc:56-146,284-303); `text` is the decompressed
// file content in host memory.  The reads land in the context exactly as sn_load_reads would leave them
// (unbarcoded reads first, barcode ordinals from 1), so sn_build_read_qgraph48 can follow directly and
// sn_save_read_files writes the reference's .fastb/.qualp/.bci byte for byte.
static int load_fasth_impl(sn_ctx* c, const char* text, uint64_t n_bytes, const uint64_t* file_first_byte, uint32_t n_files);
int sn_load_fasth_text(sn_ctx* c, const char* text, uint64_t n_bytes)
{
    const uint64_t zero = 0;
    return load_fasth_impl(c, text, n_bytes, &zero, 1);
}
// several barcode-sorted files, as ParseBarcodedFastqs FASTQS={a,b,...} takes them (:258-264): the unbarcoded reads
// of all files first, then the barcoded ones in file order; a file never continues the barcode of the one before
int sn_load_fasth_files(sn_ctx* c, const char* const* paths, uint32_t n_files)
{
    if (!c || !paths || !n_files) return SN_ERR_ARG;
    if (n_files > 4096) return fail(c, SN_ERR_ARG, "sn_load_fasth_files: more than 4096 files");
    std::vector<char> text, one; std::vector<uint64_t> first; std::string err;
    for (uint32_t f = 0; f < n_files; ++f) {
        if (!paths[f] || !snf::read_text_maybe_gz(paths[f], one, err)) return fail(c, SN_ERR_IO, paths[f] ? err : "NULL path");
        if (!one.empty() && one.back() != '\n') return fail(c, SN_ERR_DATA, std::string("fasth: out of sync reading line (") + paths[f] + " does not end with a newline)");
        first.push_back(text.size());
        text.insert(text.end(), one.begin(), one.end());
    }
    return load_fasth_impl(c, text.data(), text.size(), first.data(), n_files);
}
static int load_fasth_impl(sn_ctx* c, const char* text, uint64_t n_bytes, const uint64_t* file_first_byte, uint32_t n_files)
{
    if (!c || (!text && n_bytes)) return SN_ERR_ARG;
    CU(cudaSetDevice(c->device));
    c->cnt = sn_counts{}; c->stage = 0; c->reads_ok = false; c->paths_on_host = false;
    c->gl_ready = false; c->hist_ready_bits = -1; c->dsc_ready = false;
    if (!n_bytes) return fail(c, SN_ERR_ARG, "sn_load_fasth_text: empty input");
    if (text[n_bytes - 1] != '\n') return fail(c, SN_ERR_DATA, "fasth: out of sync reading line (the text does not end with a newline)");
    DevBuf &dtext = c->pool["ing_text"], &segc = c->pool["ing_segc"], &sego = c->pool["ing_sego"], &ls = c->pool["ing_lines"];
    t_begin(c, "ingest_h2d");
    CU(dtext.alloc(n_bytes + 64));
    CU(cudaMemcpyAsync(dtext.p, text, n_bytes, cudaMemcpyHostToDevice, c->st));
    t_end(c, "ingest_h2d");
    t_begin(c, "ingest_parse");
    const uint8_t* T = dtext.as<uint8_t>();
    const uint64_t n_seg = (n_bytes + SN_ING_SEG - 1) / SN_ING_SEG;
    CU(segc.alloc(4 * n_seg)); CU(sego.alloc(8 * (n_seg + 1)));
    k_nl_count<<<blocks_for(n_seg, 256), 256, 0, c->st>>>(T, n_bytes, segc.as<uint32_t>(), n_seg);
    KCHECK("k_nl_count");
    uint64_t n_lines = 0;
    int r = scan_u32(c, segc.as<uint32_t>(), n_seg, sego.as<uint64_t>(), &n_lines);
    if (r) return r;
    if (n_lines % 9) return fail(c, SN_ERR_DATA, "fasth: out of sync reading line (" + std::to_string(n_lines) + " lines: not 9 per record)");
    const uint64_t n_rec = n_lines / 9, n = 2 * n_rec;
    if (!n_rec) return fail(c, SN_ERR_ARG, "sn_load_fasth_text: no records");
    if (n >= (1ull << 32)) return fail(c, SN_ERR_ARG, "sn_load_reads: more than 2^32-1 reads per context");
    CU(ls.alloc(8 * (n_lines + 1)));
    k_nl_fill<<<blocks_for(n_seg, 256), 256, 0, c->st>>>(T, n_bytes, sego.as<uint64_t>(), n_seg, ls.as<uint64_t>());
    KCHECK("k_nl_fill");
    // file boundaries -> first record of every file (each file must hold whole records)
    DevBuf &ffb = c->pool["ing_ffb"], &ffl = c->pool["ing_ffl"], &ffr = c->pool["ing_ffr"];
    CU(ffb.alloc(8ull * n_files)); CU(ffl.alloc(8ull * n_files)); CU(ffr.alloc(4ull * n_files));
    {
        CU(cudaMemcpyAsync(ffb.p, file_first_byte, 8ull * n_files, cudaMemcpyHostToDevice, c->st));
        k_fasth_file_lines<<<blocks_for(n_files, 64), 64, 0, c->st>>>(ls.as<uint64_t>(), n_lines, ffb.as<uint64_t>(), n_files, ffl.as<uint64_t>());
        KCHECK("k_fasth_file_lines");
        std::vector<uint64_t> h_l(n_files); std::vector<uint32_t> h_r(n_files);
        CU(cudaMemcpyAsync(h_l.data(), ffl.p, 8ull * n_files, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        for (uint32_t f = 0; f < n_files; ++f) {
            if (h_l[f] == ~0ull || h_l[f] % 9) return fail(c, SN_ERR_DATA, "fasth: out of sync reading line (input file " + std::to_string(f) + " does not start on a record boundary: the file before it is not 9 lines per record)");
            h_r[f] = (uint32_t)(h_l[f] / 9);
        }
        CU(cudaMemcpyAsync(ffr.p, h_r.data(), 4ull * n_files, cudaMemcpyHostToDevice, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    uint32_t* err = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8) + 10;
    CU(cudaMemsetAsync(err, 0, 4, c->st));
    DevBuf &bflag = c->pool["ing_bflag"], &bbefore = c->pool["ing_bbefore"], &blist = c->pool["ing_blist"], &isnew = c->pool["ing_isnew"], &nbefore = c->pool["ing_nbefore"];
    CU(bflag.alloc(4 * n_rec)); CU(bbefore.alloc(8 * (n_rec + 1)));
    k_fasth_flags<<<blocks_for(n_rec, 256), 256, 0, c->st>>>(T, ls.as<uint64_t>(), n_rec, bflag.as<uint32_t>(), err);
    KCHECK("k_fasth_flags");
    uint64_t n_bc = 0;
    if ((r = scan_u32(c, bflag.as<uint32_t>(), n_rec, bbefore.as<uint64_t>(), &n_bc))) return r;
    CU(blist.alloc(4 * n_bc + 16)); CU(isnew.alloc(4 * n_bc + 16)); CU(nbefore.alloc(8 * (n_bc + 1) + 16));
    uint64_t n_barcodes = 0;
    if (n_bc) {
        k_fasth_blist<<<blocks_for(n_rec, 256), 256, 0, c->st>>>(bflag.as<uint32_t>(), bbefore.as<uint64_t>(), n_rec, blist.as<uint32_t>());
        KCHECK("k_fasth_blist");
        k_fasth_newbc<<<blocks_for(n_bc, 256), 256, 0, c->st>>>(T, ls.as<uint64_t>(), blist.as<uint32_t>(), n_bc, isnew.as<uint32_t>(), ffr.as<uint32_t>(), n_files);
        KCHECK("k_fasth_newbc");
        if ((r = scan_u32(c, isnew.as<uint32_t>(), n_bc, nbefore.as<uint64_t>(), &n_barcodes))) return r;
    }
    if (n_barcodes >= 0xFFFFFFull) return fail(c, SN_ERR_ARG, "more than 2^24-2 distinct barcodes in one context");
    c->cnt.n_reads = n;
    DevBuf &bpos = c->pool["ing_bpos"], &qpos = c->pool["ing_qpos"], &nby = c->pool["ing_nbytes"], &cap = c->pool["ing_pqcap"], &slot = c->pool["ing_slot"],
           &psz = c->pool["ing_pqsize"], &scratch = c->pool["ing_scratch"];
    CU(c->len.alloc(4 * n)); CU(c->bc.alloc(4 * n)); CU(bpos.alloc(8 * n)); CU(qpos.alloc(8 * n)); CU(nby.alloc(4 * n)); CU(cap.alloc(4 * n));
    CU(c->boff.alloc(8 * (n + 1))); CU(slot.alloc(8 * (n + 1))); CU(psz.alloc(4 * n)); CU(c->pqoff.alloc(8 * (n + 1)));
    k_fasth_layout<<<blocks_for(n_rec, 256), 256, 0, c->st>>>(T, ls.as<uint64_t>(), n_rec, bflag.as<uint32_t>(), bbefore.as<uint64_t>(), n_rec - n_bc,
        nbefore.as<uint64_t>(), isnew.as<uint32_t>(), c->len.as<uint32_t>(), c->bc.as<int32_t>(), bpos.as<uint64_t>(), qpos.as<uint64_t>(),
        nby.as<uint32_t>(), cap.as<uint32_t>(), err);
    KCHECK("k_fasth_layout");
    uint64_t n_base_bytes = 0, n_scratch = 0, n_pq = 0;
    if ((r = scan_u32(c, nby.as<uint32_t>(), n, c->boff.as<uint64_t>(), &n_base_bytes))) return r;
    if ((r = scan_u32(c, cap.as<uint32_t>(), n, slot.as<uint64_t>(), &n_scratch))) return r;
    CU(c->bases.alloc(n_base_bytes + 64)); CU(scratch.alloc(n_scratch + 16));
    CU(cudaMemsetAsync((char*)c->bases.p + n_base_bytes, 0, 64, c->st));
    k_fasth_pack<<<blocks_for(n, 128), 128, 0, c->st>>>(T, bpos.as<uint64_t>(), c->len.as<uint32_t>(), c->boff.as<uint64_t>(), n, c->bases.as<uint8_t>(), err);
    KCHECK("k_fasth_pack");
    k_fasth_pqvec<<<blocks_for(n, 128), 128, 0, c->st>>>(T, qpos.as<uint64_t>(), c->len.as<uint32_t>(), slot.as<uint64_t>(), n, scratch.as<uint8_t>(), psz.as<uint32_t>(), err);
    KCHECK("k_fasth_pqvec");
    if ((r = scan_u32(c, psz.as<uint32_t>(), n, c->pqoff.as<uint64_t>(), &n_pq))) return r;
    CU(c->pq.alloc(n_pq + 16));
    CU(cudaMemsetAsync((char*)c->pq.p + n_pq, 0, 16, c->st));
    k_fasth_pq_compact<<<blocks_for(n * 32, 256), 256, 0, c->st>>>(scratch.as<uint8_t>(), slot.as<uint64_t>(), psz.as<uint32_t>(), c->pqoff.as<uint64_t>(), n, c->pq.as<uint8_t>());
    KCHECK("k_fasth_pq_compact");
    t_end(c, "ingest_parse");
    uint32_t h_err = 0;
    CU(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (h_err & SN_ING_E_NAME) return fail(c, SN_ERR_DATA, "fasth: out of sync reading line (a record does not start with '@')");
    if (h_err & SN_ING_E_QLEN) return fail(c, SN_ERR_DATA, "fasth: a quality line and its base line differ in length");
    if (h_err & SN_ING_E_LONG) return fail(c, SN_ERR_ARG, "reads longer than " + std::to_string(SN_MAX_READ_LEN) + " bases are not supported");
    if (h_err & SN_ING_E_QUAL) return fail(c, SN_ERR_DATA, "fasth: a quality character outside '!'..'`' (Q0..Q63): PQVec cannot hold it (the reference's PQVecEncoder gives up on it too)");
    if (h_err & SN_ING_E_BASE) return fail(c, SN_ERR_DATA, "fasth: a base other than ACGTN (the reference draws other ambiguity codes at random: not reproducible)");
    c->have_bc = true; c->have_pq = true; c->quals.release();
    return finish_load(c);
}
// the synthetic linked reads of pairs [first_pair, first_pair + n_pairs) of a job of spec->total_pairs pairs,
// generated in place (sn_synth.cuh); quals go through the ingest path's PQVec encoder chunk by chunk
int sn_generate_reads(sn_ctx* c, const sn_synth* spec, uint64_t first_pair, uint64_t n_pairs, const uint32_t* err_thresholds /* 150 */)
{
    if (!c || !spec || !err_thresholds) return SN_ERR_ARG;
    if (!n_pairs || first_pair + n_pairs > spec->total_pairs || spec->genome_bases < 1000 || !spec->n_barcodes) return fail(c, SN_ERR_ARG, "sn_generate_reads: empty slice, slice beyond total_pairs, genome below 1000 bases or no barcodes");
    if (spec->n_barcodes >= 0xFFFFFEu) return fail(c, SN_ERR_ARG, "more than 2^24-2 distinct barcodes in one context");
    const uint64_t n = 2 * n_pairs;
    if (n >= (1ull << 32)) return fail(c, SN_ERR_ARG, "sn_generate_reads: more than 2^32-1 reads per context");
    CU(cudaSetDevice(c->device));
    c->cnt = sn_counts{}; c->stage = 0; c->reads_ok = false; c->paths_on_host = false;
    c->gl_ready = false; c->hist_ready_bits = -1; c->dsc_ready = false;
    c->cnt.n_reads = n;
    SynSpec sp; sp.genome_bases = spec->genome_bases; sp.total_pairs = spec->total_pairs; sp.seed = spec->seed; sp.n_barcodes = spec->n_barcodes;
    constexpr uint64_t NB = (SN_SYN_L + 3) / 4, CAP = SN_SYN_L + 8;
    uint64_t CH = std::min<uint64_t>(n, 4ull << 20);                            // reads per PQVec chunk
    if (const char* e = getenv("SN_SYN_CHUNK")) { const long long v = atoll(e); if (v >= 2) CH = std::min<uint64_t>(n, (uint64_t)v); }   // tests: many chunks on a small set
    DevBuf &T = c->pool["syn_T"], &qtext = c->pool["syn_qtext"], &qpos = c->pool["ing_qpos"], &slot = c->pool["ing_slot"], &psz = c->pool["ing_pqsize"],
           &scratch = c->pool["ing_scratch"], &coff = c->pool["syn_coff"];
    uint32_t* err = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8) + 10;
    CU(cudaMemsetAsync(err, 0, 4, c->st));
    CU(T.alloc(4 * SN_SYN_L)); CU(cudaMemcpyAsync(T.p, err_thresholds, 4 * SN_SYN_L, cudaMemcpyHostToDevice, c->st));
    CU(c->bases.alloc(n * NB + 64)); CU(c->boff.alloc(8 * (n + 1))); CU(c->len.alloc(4 * n)); CU(c->bc.alloc(4 * n)); CU(c->pqoff.alloc(8 * (n + 1)));
    CU(qtext.alloc(CH * SN_SYN_L)); CU(qpos.alloc(8 * CH)); CU(slot.alloc(8 * (CH + 1))); CU(psz.alloc(4 * CH)); CU(scratch.alloc(CH * CAP + 16)); CU(coff.alloc(8 * (CH + 1)));
    CU(cudaMemsetAsync((char*)c->bases.p + n * NB, 0, 64, c->st));
    // the PQVec stream grows chunk by chunk: ~0.31 bytes per qual on these reads, a quarter more to start with
    uint64_t pq_cap = n * 60 + 1024, pq_used = 0;
    CU(c->pq.alloc(pq_cap));
    c->have_bc = true; c->have_pq = true; c->quals.release();
    t_begin(c, "h2d");
    for (uint64_t r0 = 0; r0 < n; r0 += CH) {
        const uint64_t m = std::min(CH, n - r0);
        k_synth_reads<<<blocks_for(m, 128), 128, 0, c->st>>>(sp, T.as<uint32_t>(), first_pair, r0, m, c->bases.as<uint8_t>(), c->boff.as<uint64_t>(), c->len.as<uint32_t>(),
            c->bc.as<int32_t>(), qtext.as<uint8_t>(), qpos.as<uint64_t>(), slot.as<uint64_t>());
        KCHECK("k_synth_reads");
        k_fasth_pqvec<<<blocks_for(m, 128), 128, 0, c->st>>>(qtext.as<uint8_t>(), qpos.as<uint64_t>(), c->len.as<uint32_t>() + r0, slot.as<uint64_t>(), m, scratch.as<uint8_t>(), psz.as<uint32_t>(), err);
        KCHECK("k_fasth_pqvec");
        uint64_t bytes = 0;
        { int r = scan_u32(c, psz.as<uint32_t>(), m, coff.as<uint64_t>(), &bytes); if (r) return r; }
        if (pq_used + bytes + 16 > pq_cap) {                                  // grow, keeping what is there
            DevBuf bigger;
            pq_cap = std::max<uint64_t>(pq_used + bytes + 16, (uint64_t)((double)(pq_used + bytes) * (double)n / (double)(r0 + m)) + (64ull << 20));
            CU(bigger.alloc(pq_cap));
            if (pq_used) CU(cudaMemcpyAsync(bigger.p, c->pq.p, pq_used, cudaMemcpyDeviceToDevice, c->st));
            CU(cudaStreamSynchronize(c->st));
            std::swap(c->pq.p, bigger.p); std::swap(c->pq.bytes, bigger.bytes); std::swap(c->pq.cap, bigger.cap);
        }
        k_add_u64<<<blocks_for(m + 1, 256), 256, 0, c->st>>>(coff.as<uint64_t>(), m + 1, pq_used, c->pqoff.as<uint64_t>() + r0);
        KCHECK("k_add_u64");
        k_fasth_pq_compact<<<blocks_for(m * 32, 256), 256, 0, c->st>>>(scratch.as<uint8_t>(), slot.as<uint64_t>(), psz.as<uint32_t>(), c->pqoff.as<uint64_t>() + r0, m, c->pq.as<uint8_t>());
        KCHECK("k_fasth_pq_compact");
        pq_used += bytes;
    }
    CU(cudaMemsetAsync((char*)c->pq.p + pq_used, 0, 16, c->st));
    {   // boff[n]
        const uint64_t endb = n * NB;
        CU(cudaMemcpyAsync(c->boff.as<uint64_t>() + n, &endb, 8, cudaMemcpyHostToDevice, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    pool_release(c, {"syn_qtext", "syn_coff", "ing_scratch", "ing_qpos", "ing_slot", "ing_pqsize"});
    return finish_load(c);
}
int sn_load_fasth_file(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    std::vector<char> text; std::string err;
    if (!snf::read_text_maybe_gz(path, text, err)) return fail(c, SN_ERR_IO, err);
    return sn_load_fasth_text(c, text.data(), text.size());
}
// the loaded reads as the three files ParseBarcodedFastqs writes
int sn_save_read_files(sn_ctx* c, const char* fastb, const char* qualp, const char* bci)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 1 || !c->have_pq) return fail(c, SN_ERR_STATE, "sn_save_read_files: no PQVec reads loaded");
    CU(cudaSetDevice(c->device));
    const uint64_t n = c->cnt.n_reads;
    std::vector<uint64_t> boff(n + 1), pqoff(n + 1); std::vector<uint32_t> len(n); std::vector<int32_t> bc(n, 0);
    CU(cudaMemcpy(boff.data(), c->boff.p, 8 * (n + 1), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(pqoff.data(), c->pqoff.p, 8 * (n + 1), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(len.data(), c->len.p, 4 * n, cudaMemcpyDeviceToHost));
    if (c->have_bc) CU(cudaMemcpy(bc.data(), c->bc.p, 4 * n, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> bases(boff[n] + 1), pq(pqoff[n] + 1);
    if (boff[n]) CU(cudaMemcpy(bases.data(), c->bases.p, boff[n], cudaMemcpyDeviceToHost));
    if (pqoff[n]) CU(cudaMemcpy(pq.data(), c->pq.p, pqoff[n], cudaMemcpyDeviceToHost));
    int r = sn_write_read_files(fastb, qualp, bci, n, bases.data(), boff.data(), len.data(), pq.data(), pqoff.data(), bc.data());
    if (r) return fail(c, r, g_sn_create_error);
    return SN_OK;
}

int sn_load_read_files(sn_ctx* c, const char* fastb, const char* qualp, const char* bci)
{
    if (!c || !fastb || !qualp) return SN_ERR_ARG;
    snf::Fastb fb; snf::Qualp qp; std::vector<int64_t> bi; std::vector<int32_t> bc; std::string err;
    if (!snf::read_fastb(fastb, fb, err) || !snf::read_qualp(qualp, qp, err)) return fail(c, SN_ERR_IO, err);
    if (fb.len.size() + 1 != qp.off.size()) return fail(c, SN_ERR_DATA, "fastb and qualp hold different numbers of reads");
    if (bci) {
        if (!snf::read_bci(bci, bi, err)) return fail(c, SN_ERR_IO, err);
        if (!snf::expand_bci(bi, bc)) return fail(c, SN_ERR_DATA, std::string(bci) + ": barcode index is not sorted");
        if (bc.size() != fb.len.size()) return fail(c, SN_ERR_DATA, "bci.back() != number of reads");
    }
    return sn_load_reads(c, fb.len.size(), fb.var.data(), fb.off.data(), fb.len.data(), qp.var.data(), qp.off.data(), bci ? bc.data() : nullptr);
}
// a slice of the read files: reads [first_read, first_read + n_reads) (n_reads = 0: to the end) -- the shard of one rank
int sn_load_read_files_range(sn_ctx* c, const char* fastb, const char* qualp, const char* bci, uint64_t first_read, uint64_t n_reads)
{
    if (!c || !fastb || !qualp) return SN_ERR_ARG;
    snf::Fastb fb; snf::Qualp qp; std::vector<int64_t> bi; std::vector<int32_t> bc; std::string err;
    if (!snf::read_fastb(fastb, fb, err) || !snf::read_qualp(qualp, qp, err)) return fail(c, SN_ERR_IO, err);
    if (fb.len.size() + 1 != qp.off.size()) return fail(c, SN_ERR_DATA, "fastb and qualp hold different numbers of reads");
    if (bci) {
        if (!snf::read_bci(bci, bi, err)) return fail(c, SN_ERR_IO, err);
        if (!snf::expand_bci(bi, bc)) return fail(c, SN_ERR_DATA, std::string(bci) + ": barcode index is not sorted");
        if (bc.size() != fb.len.size()) return fail(c, SN_ERR_DATA, "bci.back() != number of reads");
    }
    const uint64_t N = fb.len.size();
    if (first_read > N) return fail(c, SN_ERR_ARG, "sn_load_read_files_range: first_read beyond the file");
    if (!n_reads || first_read + n_reads > N) n_reads = N - first_read;
    if (!n_reads) return fail(c, SN_ERR_ARG, "sn_load_read_files_range: empty range");
    std::vector<uint64_t> bo(n_reads + 1), qo(n_reads + 1);
    for (uint64_t i = 0; i <= n_reads; ++i) { bo[i] = fb.off[first_read + i] - fb.off[first_read]; qo[i] = qp.off[first_read + i] - qp.off[first_read]; }
    return sn_load_reads(c, n_reads, fb.var.data() + fb.off[first_read], bo.data(), fb.len.data() + first_read, qp.var.data() + qp.off[first_read], qo.data(),
                         bci ? bc.data() + first_read : nullptr);
}
// the same with the per-read barcode ordinals the caller already holds in memory (the vec<int32_t> bc of
// buildReadQGraph48's caller, 10X/DF.cc:464-469); bc may be NULL
int sn_load_read_files_bc(sn_ctx* c, const char* fastb, const char* qualp, const int32_t* bc, uint64_t n_bc)
{
    if (!c || !fastb || !qualp) return SN_ERR_ARG;
    snf::Fastb fb; snf::Qualp qp; std::string err;
    if (!snf::read_fastb(fastb, fb, err) || !snf::read_qualp(qualp, qp, err)) return fail(c, SN_ERR_IO, err);
    if (fb.len.size() + 1 != qp.off.size()) return fail(c, SN_ERR_DATA, "fastb and qualp hold different numbers of reads");
    if (bc && n_bc != fb.len.size()) return fail(c, SN_ERR_DATA, "the barcode array and the fastb hold different numbers of reads");
    return sn_load_reads(c, fb.len.size(), fb.var.data(), fb.off.data(), fb.len.data(), qp.var.data(), qp.off.data(), bc);
}

// ---------------------------------------------------------------------------
// ---- pieces of the count stage (shared by the single-GPU call and the multi-GPU calls) ----
uint32_t sn_i_first_bucket(uint32_t owner, uint32_t nparts, int bits) { return (uint32_t)((((uint64_t)owner << bits) + nparts - 1) / nparts); }
int sn_i_pick_bucket_bits(uint64_t n_occ)
{
    int bits = msp_bucket_bits(n_occ);
    if (const char* e = getenv("SN_MSP_OCC")) { int t = atoi(e); if (t >= 64) { bits = 4; while (bits < 24 && (n_occ >> bits) > (uint64_t)t) ++bits; } }
    if (const char* e = getenv("SN_MSP_BITS")) { int b = atoi(e); if (b >= 1 && b <= 24) bits = b; }     // tests: few huge buckets force the split passes
    return bits;
}
int sn_i_count_set_params(sn_ctx* c, const sn_params* p)
{
    if (p) c->params = *p;
    if (c->params.min_bc > 2) return fail(c, SN_ERR_ARG, "min_bc > 2 is not supported (the reference pipeline fixes MIN_BC=2, 10X/DF.cc:140)");
    if (c->params.min_freq == 0) c->params.min_freq = 1;
    return SN_OK;
}
// a1: good lengths and the number of k-mer occurrences Kmerizer::map will emit
int sn_i_count_goodlen(sn_ctx* c, uint64_t* n_occ_out)
{
    const uint64_t n = c->cnt.n_reads;
    if (c->gl_ready) {                           // computed under the copies of sn_load_reads_streamed (once: a repeated count recomputes)
        c->gl_ready = false;
        if (c->gl_min_qual == c->params.min_qual) { c->cnt.n_kmer_occurrences = c->gl_occ; *n_occ_out = c->gl_occ; return SN_OK; }
    }
    c->hist_ready_bits = -1; c->dsc_ready = false;
    unsigned long long* occ = c->counters.as<unsigned long long>();        // [0] occurrences, [1] cursor, [2] distinct
    uint32_t* u32c = reinterpret_cast<uint32_t*>(occ + 8);                  // [0] bad reads, [3] reduce overflow
    CU(cudaMemsetAsync(c->counters.p, 0, 256, c->st));
    CU(c->goodlen.alloc(4 * n));
    t_begin(c, "goodlen");
    if (c->have_pq) {
        k_pqvec_goodlen<<<blocks_for(n, 256), 256, 0, c->st>>>(n, c->pq.as<uint8_t>(), c->pqoff.as<uint64_t>(), c->len.as<uint32_t>(),
            c->params.min_qual, c->min_gl, c->goodlen.as<uint32_t>(), occ, u32c);
        KCHECK("k_pqvec_goodlen");
    } else {
        k_q8_goodlen<<<blocks_for(n, 256), 256, 0, c->st>>>(n, c->quals.as<uint8_t>(), c->qoff.as<uint64_t>(), c->len.as<uint32_t>(),
            c->params.min_qual, c->min_gl, c->goodlen.as<uint32_t>(), occ);
        KCHECK("k_q8_goodlen");
    }
    t_end(c, "goodlen");
    unsigned long long h_occ = 0; uint32_t h_bad = 0;
    CU(cudaMemcpyAsync(&h_occ, occ, 8, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(&h_bad, u32c, 4, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (h_bad) return fail(c, SN_ERR_DATA, std::to_string(h_bad) + " reads whose PQVec length differs from their base count");
    c->cnt.n_kmer_occurrences = h_occ;
    *n_occ_out = h_occ;
    return SN_OK;
}
// a14 (MSP): cuts this context's reads into super-k-mers and groups them by minimizer bucket:
// pool["sk_recs"] (32-byte records, bucket order) and pool["sk_off"] (2^bits + 1 record offsets).
int sn_i_msp_partition(sn_ctx* c, int bits, uint64_t* n_sk_out, uint32_t b_lo, uint32_t b_n, uint32_t pcfg)
{
    const uint64_t n = c->cnt.n_reads;
    if (pcfg) { b_lo = 0; b_n = (1u << bits) >> ((pcfg >> 8) & 0xFFu); }      // an interleaved pass (msp_window_bucket): all of its renumbered buckets
    const bool window = b_n != 0;                 // a count in several passes: only the buckets [b_lo, b_lo + b_n)
    const uint64_t nb = window ? b_n : 1ull << bits;
    const uint32_t w_lo = window ? b_lo : 0u, w_n = window ? b_n : 0xFFFFFFFFu;
    if (window) c->hist_ready_bits = -1;
    DevBuf &hist = c->pool["sk_hist"], &off = c->pool["sk_off"], &recs = c->pool["sk_recs"], &dsc = c->pool["sk_dsc"], &nruns = c->pool["sk_nruns"];
    CU(hist.alloc(4 * nb)); CU(off.alloc(8 * (nb + 1)));
    const int32_t* bc = c->have_bc ? c->bc.as<int32_t>() : nullptr;
    const unsigned grid = blocks_for(n, SN_MS_READS);
    uint32_t* ovf = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8) + 9;     // reads whose runs did not fit the descriptors
    // The first scan of the reads under the current good lengths leaves the runs of every read behind
    // (12 slots of 8 bytes per read); the scatter pass, and every later bucket window, places the
    // super-k-mers from those instead of computing the minimizers again.
    t_begin(c, "msp_hist");
    if (c->hist_ready_bits != bits || window) {      // (else: the histogram was built under the copies of sn_load_reads_streamed)
        CU(cudaMemsetAsync(hist.p, 0, 4 * nb, c->st));
        if (c->dsc_ready) {
            k_msp_place<false><<<grid, SN_MS_READS, 0, c->st>>>(n, c->bases.as<uint8_t>(), c->boff.as<uint64_t>(), c->goodlen.as<uint32_t>(),
                bc, c->params.ign_bc_below, bits, hist.as<uint32_t>(), nullptr, nullptr, w_lo, w_n, dsc.as<uint2>(), nruns.as<uint8_t>(), pcfg);
            KCHECK("k_msp_place<hist>");
            if (c->dsc_overflow) {
                k_msp_scan<false><<<grid, SN_MS_READS, 0, c->st>>>(n, c->bases.as<uint8_t>(), c->boff.as<uint64_t>(), c->goodlen.as<uint32_t>(),
                    bc, c->params.ign_bc_below, c->min_gl, bits, hist.as<uint32_t>(), nullptr, nullptr, w_lo, w_n, nullptr, nullptr, nruns.as<uint8_t>(), nullptr, pcfg);
                KCHECK("k_msp_scan<hist, overflow>");
            }
        } else {
            CU(dsc.alloc(8ull * SN_MS_QUEUE * SN_MS_READS * grid)); CU(nruns.alloc(n));
            CU(cudaMemsetAsync(ovf, 0, 4, c->st));
            k_msp_scan<false><<<grid, SN_MS_READS, 0, c->st>>>(n, c->bases.as<uint8_t>(), c->boff.as<uint64_t>(), c->goodlen.as<uint32_t>(),
                bc, c->params.ign_bc_below, c->min_gl, bits, hist.as<uint32_t>(), nullptr, nullptr, w_lo, w_n, dsc.as<uint2>(), nruns.as<uint8_t>(), nullptr, ovf, pcfg);
            KCHECK("k_msp_scan<hist>");
            CU(cudaMemcpyAsync(&c->dsc_overflow, ovf, 4, cudaMemcpyDeviceToHost, c->st));      // (host value valid after the scan's sync below)
            c->dsc_ready = true;
        }
    }
    c->hist_ready_bits = -1;
    uint64_t n_sk = 0;
    int r = scan_u32(c, hist.as<uint32_t>(), nb, off.as<uint64_t>(), &n_sk);
    if (r) return r;
    t_end(c, "msp_hist");
    *n_sk_out = n_sk;
    CU(recs.alloc(32 * n_sk + 64));
    t_begin(c, "msp_scatter");
    CU(cudaMemsetAsync(hist.p, 0, 4 * nb, c->st));                 // now the per-bucket cursors
    if (c->dsc_ready) {
        k_msp_place<true><<<grid, SN_MS_READS, 0, c->st>>>(n, c->bases.as<uint8_t>(), c->boff.as<uint64_t>(), c->goodlen.as<uint32_t>(),
            bc, c->params.ign_bc_below, bits, hist.as<uint32_t>(), off.as<uint64_t>(), recs.as<uint4>(), w_lo, w_n, dsc.as<uint2>(), nruns.as<uint8_t>(), pcfg);
        KCHECK("k_msp_place<scatter>");
    }
    if (!c->dsc_ready || c->dsc_overflow) {
        k_msp_scan<true><<<grid, SN_MS_READS, 0, c->st>>>(n, c->bases.as<uint8_t>(), c->boff.as<uint64_t>(), c->goodlen.as<uint32_t>(),
            bc, c->params.ign_bc_below, c->min_gl, bits, hist.as<uint32_t>(), off.as<uint64_t>(), recs.as<uint4>(), w_lo, w_n, nullptr, nullptr, c->dsc_ready ? nruns.as<uint8_t>() : nullptr, nullptr, pcfg);
        KCHECK("k_msp_scan<scatter>");
    }
    t_end(c, "msp_scatter");
    return SN_OK;
}
// a15 / a4 + a5: per-bucket count + filter of bucket-ordered super-k-mer records (n_seg ranges per
// bucket, see k_bucket_count).  The surviving k-mers land in `surv` ordered by (bucket, hash,
// k-mer); surv_off[n_buckets + 1] = first survivor of every bucket.
// `occ_bound` bounds the k-mer occurrences the records hold.
int sn_i_msp_bucket_count(sn_ctx* c, const uint4* recs, const uint64_t* off, uint32_t n_buckets, uint32_t n_seg, uint64_t occ_bound,
                            DevBuf& surv, DevBuf& surv_off, uint64_t* n_surv_out)
{
    unsigned long long* occ = c->counters.as<unsigned long long>();        // [2] distinct, [3] scratch cursor
    uint32_t* u32c = reinterpret_cast<uint32_t*>(occ + 8);                  // [3] error flags, [8] bucket ticket
    c->cnt.n_kmers = 0; c->cnt.n_kmers_distinct = 0;
    *n_surv_out = 0;
    CU(surv_off.alloc(4ull * (n_buckets + 1)));
    if (!occ_bound) { CU(cudaMemsetAsync(surv_off.p, 0, 4ull * (n_buckets + 1), c->st)); CU(surv.alloc(64)); return SN_OK; }
    const uint64_t cap = occ_bound / c->params.min_freq + 16;
    if (cap >= (1ull << 32)) return fail(c, SN_ERR_ARG, "more than 2^32-1 candidate k-mers in one context");
    DevBuf &scratch = c->pool["surv_scratch"], &seg_base = c->pool["bc_seg_base"], &seg_cnt = c->pool["bc_seg_cnt"], &off64 = c->pool["bc_off64"],
           &vnv = c->pool["bc_vnv"], &vfirst = c->pool["bc_vfirst"], &vbucket = c->pool["bc_vbucket"], &bcnt = c->pool["bc_cnt"], &boff64 = c->pool["bc_boff64"];
    CU(surv.alloc(16 * cap)); CU(scratch.alloc(16 * cap));
    t_begin(c, "bucket_count");
    // the CTA table: one CTA per bucket, 2^d CTAs for a heavy one (k_bucket_count2)
    uint32_t heavy = SN_BC_HEAVY_RECORDS;
    if (const char* e = getenv("SN_BC_HEAVY")) { const int v = atoi(e); if (v >= 1) heavy = (uint32_t)v; }      // tests
    CU(vnv.alloc(4ull * n_buckets + 16)); CU(vfirst.alloc(8ull * (n_buckets + 1)));
    k_vb_count<<<blocks_for(n_buckets, 256), 256, 0, c->st>>>(off, n_buckets, n_seg, heavy, vnv.as<uint32_t>());
    KCHECK("k_vb_count");
    uint64_t n_virtual = 0;
    { int rv = scan_u32(c, vnv.as<uint32_t>(), n_buckets, vfirst.as<uint64_t>(), &n_virtual); if (rv) return rv; }
    if (n_virtual >= (1ull << 31)) return fail(c, SN_ERR_ARG, "too many virtual buckets");
    CU(vbucket.alloc(8 * n_virtual + 16)); CU(seg_base.alloc(8 * n_virtual + 16)); CU(seg_cnt.alloc(4 * n_virtual + 16)); CU(off64.alloc(8 * (n_virtual + 1)));
    CU(bcnt.alloc(4ull * n_buckets + 16)); CU(boff64.alloc(8ull * (n_buckets + 1)));
    k_vb_fill<<<blocks_for(n_buckets, 256), 256, 0, c->st>>>(vnv.as<uint32_t>(), vfirst.as<uint64_t>(), n_buckets, vbucket.as<uint2>());
    KCHECK("k_vb_fill");
    CU(cudaMemsetAsync(occ + 2, 0, 16, c->st)); CU(cudaMemsetAsync(u32c + 3, 0, 4, c->st));
    CU(cudaMemsetAsync(seg_cnt.p, 0, 4 * n_virtual + 16, c->st));
    const int variant = getenv("SN_BC_VARIANT") ? atoi(getenv("SN_BC_VARIANT")) : 0;
#define SN_BC_LAUNCH(T, S, I, M) do { \
        CU(cudaFuncSetAttribute(k_bucket_count<T, S, I, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BcSmem<T, S>)));   /* (per device: every call) */ \
        k_bucket_count<T, S, I, M><<<n_buckets, T, sizeof(BcSmem<T, S>), c->st>>>(recs, off, n_buckets, n_seg, c->params.min_freq, c->params.min_bc, c->have_bc ? 1 : 0, \
            scratch.as<uint4>(), cap, occ + 3, seg_base.as<uint64_t>(), seg_cnt.as<uint32_t>(), occ + 2, u32c + 3); } while (0)
#define SN_BC2_LAUNCH(T, S, I, M) do { \
        CU(cudaFuncSetAttribute(k_bucket_count2<T, S, I, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BcSmem2<T, S>))); \
        k_bucket_count2<T, S, I, M><<<(unsigned)n_virtual, T, sizeof(BcSmem2<T, S>), c->st>>>(recs, off, n_buckets, n_seg, c->params.min_freq, c->params.min_bc, c->have_bc ? 1 : 0, \
            scratch.as<uint4>(), cap, occ + 3, seg_base.as<uint64_t>(), seg_cnt.as<uint32_t>(), occ + 2, u32c + 3, vbucket.as<uint2>(), (uint32_t)n_virtual); } while (0)
    if ((variant == 1 || variant == 4) && n_virtual != n_buckets) return fail(c, SN_ERR_ARG, "SN_BC_VARIANT 1/4 (first count kernel) cannot share out heavy buckets: set SN_BC_HEAVY high");
    switch (variant) {
        case 1: SN_BC_LAUNCH(256, 2048, 4, 3); break;        // the two-phase insert (kept for A/B timing)
        case 4: SN_BC_LAUNCH(128, 1024, 4, 6); break;
        case 12: SN_BC2_LAUNCH(256, 2048, 1, 3); break;      // with SN_MSP_OCC=3072
        case 31: SN_BC2_LAUNCH(64, 512, 1, 12); break;       // with SN_MSP_OCC=768
        // 4 warps per bucket, 6 buckets per SM: the kernel is bound by the shared-memory LSU (atomics: 2 cycles/lane) and
        // by the barriers around its per-bucket phases, not by latency -- more occurrences in flight per thread do not help,
        // smaller CTAs do (C2: 256 threads x 2048 slots 13.7 ms, 128 x 1024 12.6 ms, 64 x 512 16.3 ms)
        default: SN_BC2_LAUNCH(128, 1024, 1, 6); break;
    }
#undef SN_BC_LAUNCH
#undef SN_BC2_LAUNCH
    KCHECK("k_bucket_count");
    t_end(c, "bucket_count");
    // the buckets' survivor ranges, in bucket order
    uint64_t n_surv = 0;
    int r = scan_u32(c, seg_cnt.as<uint32_t>(), n_virtual, off64.as<uint64_t>(), &n_surv);
    if (r) return r;
    unsigned long long h_dist = 0; uint32_t h_err = 0;
    CU(cudaMemcpyAsync(&h_dist, occ + 2, 8, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(&h_err, u32c + 3, 4, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (h_err & 1u) return fail(c, SN_ERR_DATA, "k_bucket_count: more surviving k-mers than occurrences / min_freq (internal error)");
    if (h_err & 2u) return fail(c, SN_ERR_DATA, "k_bucket_count: a bucket does not fit the shared-memory table after 20 splits (pathological hash collisions)");
    if (h_err & 4u) return fail(c, SN_ERR_DATA, "k_bucket_count: a claimed slot was never published (internal error)");
    t_begin(c, "make_dict");
    // (virtual buckets are in bucket order, a heavy bucket's in hash-prefix order: the gathered array is in (bucket, hash, k-mer) order)
    k_gather_survivors<<<std::min(blocks_for(n_virtual * 32, 256), 16u * (unsigned)c->num_sms), 256, 0, c->st>>>(scratch.as<uint4>(), seg_base.as<uint64_t>(), seg_cnt.as<uint32_t>(),
        off64.as<uint64_t>(), (uint32_t)n_virtual, surv.as<uint4>());
    KCHECK("k_gather_survivors");
    if (n_virtual == n_buckets) {
        k_narrow_u64<<<blocks_for((uint64_t)n_buckets + 1, 256), 256, 0, c->st>>>(off64.as<uint64_t>(), (uint64_t)n_buckets + 1, surv_off.as<uint32_t>());
        KCHECK("k_narrow_u64");
    } else {       // survivors per real bucket, then their offsets
        k_vb_sum<<<blocks_for(n_buckets, 256), 256, 0, c->st>>>(seg_cnt.as<uint32_t>(), vfirst.as<uint64_t>(), n_buckets, bcnt.as<uint32_t>());
        KCHECK("k_vb_sum");
        if ((r = scan_u32(c, bcnt.as<uint32_t>(), n_buckets, boff64.as<uint64_t>(), nullptr))) return r;
        k_narrow_u64<<<blocks_for((uint64_t)n_buckets + 1, 256), 256, 0, c->st>>>(boff64.as<uint64_t>(), (uint64_t)n_buckets + 1, surv_off.as<uint32_t>());
        KCHECK("k_narrow_u64");
    }
    c->cnt.n_kmers_distinct = h_dist;
    *n_surv_out = n_surv;
    return SN_OK;
}
// the surviving k-mers (bucket, hash, k-mer order) become the dictionary; `counts_or_off`: the bucket
// offsets (is_offsets) or the per-bucket survivor counts of all 2^bits buckets
int sn_i_msp_install_dict(sn_ctx* c, const uint4* surv, uint64_t n_surv, int bits, const uint32_t* counts_or_off, bool is_offsets, uint32_t nb_window, uint64_t extra_entries)
{
    if (n_surv + extra_entries >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 dictionary k-mers in one context");
    const uint64_t nb = nb_window ? nb_window : 1ull << bits;      // buckets this table holds
    c->cnt.n_kmers = n_surv; c->dict_bits = bits;
    CU(c->dict.alloc((size_t)(n_surv + extra_entries) * sizeof(DictEntry) + 64));
    CU(c->dboff.alloc(4 * (nb + 1)));
    if (is_offsets) CU(cudaMemcpyAsync(c->dboff.p, counts_or_off, 4 * (nb + 1), cudaMemcpyDeviceToDevice, c->st));
    else {
        DevBuf& o64 = c->pool["boff64"];
        CU(o64.alloc(8 * (nb + 1)));
        int r = scan_u32(c, counts_or_off, nb, o64.as<uint64_t>(), nullptr);
        if (r) return r;
        k_narrow_u64<<<blocks_for(nb + 1, 256), 256, 0, c->st>>>(o64.as<uint64_t>(), nb + 1, c->dboff.as<uint32_t>());
        KCHECK("k_narrow_u64");
    }
    CU(c->dict_hs.alloc(4 * n_surv + 64));
    if (n_surv) {
        k_make_dict<<<blocks_for(n_surv, 256), 256, 0, c->st>>>(surv, (uint32_t)n_surv, c->dict.as<DictEntry>(), c->dict_hs.as<uint32_t>());
        KCHECK("k_make_dict");
    }
    // lookup cells of ~32 entries: (bucket, top sub_bits of the hash)
    int sub = 0;
    while (sub < 6 && (nb << sub) < (1ull << 26) && n_surv / (nb << sub) > 32) ++sub;
    c->dict_sub_bits = sub;
    if (sub) {
        DevBuf& cells = c->pool["dict_cells"];
        CU(cells.alloc(4 * ((nb << sub) + 1)));
        k_dict_cells<<<blocks_for((nb << sub) + 1, 256), 256, 0, c->st>>>(c->dict.as<DictEntry>(), c->dboff.as<uint32_t>(), (uint32_t)nb, sub, cells.as<uint32_t>());
        KCHECK("k_dict_cells");
    }
    t_end(c, "make_dict");
    CU(cudaStreamSynchronize(c->st));
    c->dict_b_lo = 0; c->dict_b_n = 0; c->ghost_cap = 0; c->dict_sharded = false;      // (the sharded path sets its window after this call)
    c->stage = 2;
    return SN_OK;
}

// sn_load_reads with the first two stages of the count running UNDER the copies: the reads travel
// in chunks on a copy stream -- quals first, then bases -- and as soon as a chunk has landed its
// good lengths (a1) and, once those fix the bucket count, its share of the super-k-mer histogram
// (a14, first pass) are computed on the compute stream.  The PCIe link never waits for a kernel; the
// next sn_count_kmers (same min_qual) starts at the scatter pass.  with_hist = 0 (multi-GPU: the
// bucket count comes from an allreduce) overlaps the good lengths only.
int sn_load_reads_streamed(sn_ctx* c, uint64_t n_reads, const uint8_t* bases, const uint64_t* base_off, const uint32_t* len,
                           const uint8_t* pq, const uint64_t* pq_off, const int32_t* bc, const sn_params* p, int with_hist)
{
    if (!c) return SN_ERR_ARG;
    CU(cudaSetDevice(c->device));
    if (!n_reads || !bases || !base_off || !len || !pq || !pq_off) return fail(c, SN_ERR_ARG, "sn_load_reads_streamed: empty or NULL input");
    if (n_reads >= (1ull << 32)) return fail(c, SN_ERR_ARG, "sn_load_reads: more than 2^32-1 reads per context");
    int r;
    if ((r = sn_i_count_set_params(c, p))) return r;
    c->cnt = sn_counts{}; c->stage = 0; c->reads_ok = false; c->paths_on_host = false;
    c->gl_ready = false; c->hist_ready_bits = -1; c->dsc_ready = false;
    c->cnt.n_reads = n_reads;
    const uint64_t n = n_reads;
    constexpr int MAXCH = 8;
    const int nch = n >= (1u << 16) ? 4 : 1;
    if (!c->st2) CU(cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking));
    if (!c->ev_copy[0]) for (int i = 0; i < 2 * MAXCH; ++i) CU(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
    c->have_bc = bc != nullptr; c->have_pq = true; c->quals.release();
    CU(c->bases.alloc(base_off[n] + 64)); CU(c->boff.alloc(8 * (n + 1))); CU(c->len.alloc(4 * n));
    if (bc) CU(c->bc.alloc(4 * n));
    CU(c->pq.alloc(pq_off[n] + 16)); CU(c->pqoff.alloc(8 * (n + 1))); CU(c->goodlen.alloc(4 * n));
    unsigned long long* occ = c->counters.as<unsigned long long>();
    uint32_t* u32c = reinterpret_cast<uint32_t*>(occ + 8);
    unsigned long long* cnt64 = occ + 16;
    CU(cudaMemsetAsync(c->counters.p, 0, 256, c->st));
    CU(cudaMemsetAsync((char*)c->bases.p + base_off[n], 0, 64, c->st));
    CU(cudaMemsetAsync((char*)c->pq.p + pq_off[n], 0, 16, c->st));
    CU(cudaEventRecord(c->ev_copy[0], c->st));
    CU(cudaStreamWaitEvent(c->st2, c->ev_copy[0], 0));            // nothing of an earlier step still reads the buffers
    auto lo = [&](int ch) { return ch >= nch ? n : (n * (uint64_t)ch / (uint64_t)nch) & ~(uint64_t)(SN_MS_READS - 1); };   // whole k_msp_scan blocks per chunk
    auto bail = [&](int rc) { cudaStreamSynchronize(c->st2); cudaStreamSynchronize(c->st); return rc; };
#define CUB_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return bail(fail(c, SN_ERR_CUDA, cudaGetErrorString(e_))); } while (0)
    t_begin(c, "goodlen");
    for (int ch = 0; ch < nch; ++ch) {                             // quals -> good lengths
        const uint64_t r0 = lo(ch), r1 = lo(ch + 1);
        CUB_(cudaMemcpyAsync(c->pq.as<uint8_t>() + pq_off[r0], pq + pq_off[r0], pq_off[r1] - pq_off[r0], cudaMemcpyHostToDevice, c->st2));
        CUB_(cudaMemcpyAsync(c->pqoff.as<uint64_t>() + r0, pq_off + r0, 8 * (r1 - r0 + 1), cudaMemcpyHostToDevice, c->st2));
        CUB_(cudaMemcpyAsync(c->len.as<uint32_t>() + r0, len + r0, 4 * (r1 - r0), cudaMemcpyHostToDevice, c->st2));
        CUB_(cudaEventRecord(c->ev_copy[ch], c->st2));
        CUB_(cudaStreamWaitEvent(c->st, c->ev_copy[ch], 0));
        k_pqvec_goodlen<<<blocks_for(r1 - r0, 256), 256, 0, c->st>>>(r1 - r0, c->pq.as<uint8_t>(), c->pqoff.as<uint64_t>() + r0, c->len.as<uint32_t>() + r0,
            c->params.min_qual, c->min_gl, c->goodlen.as<uint32_t>() + r0, occ, u32c);
        ++c->launches;
    }
    t_end(c, "goodlen");
    for (int ch = 0; ch < nch; ++ch) {                             // bases (the copies queue up behind the quals)
        const uint64_t r0 = lo(ch), r1 = lo(ch + 1);
        CUB_(cudaMemcpyAsync(c->bases.as<uint8_t>() + base_off[r0], bases + base_off[r0], base_off[r1] - base_off[r0], cudaMemcpyHostToDevice, c->st2));
        CUB_(cudaMemcpyAsync(c->boff.as<uint64_t>() + r0, base_off + r0, 8 * (r1 - r0 + 1), cudaMemcpyHostToDevice, c->st2));
        if (bc) CUB_(cudaMemcpyAsync(c->bc.as<int32_t>() + r0, bc + r0, 4 * (r1 - r0), cudaMemcpyHostToDevice, c->st2));
        CUB_(cudaEventRecord(c->ev_copy[MAXCH + ch], c->st2));
    }
    unsigned long long h_occ = 0; uint32_t h_bad = 0;
    CUB_(cudaMemcpyAsync(&h_occ, occ, 8, cudaMemcpyDeviceToHost, c->st));
    CUB_(cudaMemcpyAsync(&h_bad, u32c, 4, cudaMemcpyDeviceToHost, c->st));
    CUB_(cudaStreamSynchronize(c->st));                             // the good lengths are done; the bases are still arriving
    if (h_bad) return bail(fail(c, SN_ERR_DATA, std::to_string(h_bad) + " reads whose PQVec length differs from their base count"));
    const int bits = sn_i_pick_bucket_bits(h_occ);
    if (h_occ >= 3600000000ull) with_hist = 0;                     // counted in several passes: each pass has its own histogram
    {   // the MSP scan stages 128 reads of at most SN_MAX_READ_LEN bases in shared memory: check the lengths (all on the device by now:
        // they travelled with the quals) before it runs, not after
        unsigned long long* c64 = occ + 16;
        k_read_stats<<<std::min(blocks_for(n, 256), 8u * (unsigned)c->num_sms), 256, 0, c->st>>>(n, c->len.as<uint32_t>(), nullptr, c64, reinterpret_cast<uint32_t*>(c64 + 1), reinterpret_cast<int32_t*>(c64 + 1) + 1);
        ++c->launches;
        unsigned long long hh[2] = {0, 0};
        CUB_(cudaMemcpyAsync(hh, c64, 16, cudaMemcpyDeviceToHost, c->st));
        CUB_(cudaStreamSynchronize(c->st));
        if ((uint32_t)(hh[1] & 0xFFFFFFFFu) > SN_MAX_READ_LEN) return bail(fail(c, SN_ERR_ARG, "reads longer than " + std::to_string(SN_MAX_READ_LEN) + " bases are not supported"));
        CUB_(cudaMemsetAsync(c64, 0, 16, c->st));
    }
    const uint64_t nb = 1ull << bits;
    DevBuf &hist = c->pool["sk_hist"], &dsc = c->pool["sk_dsc"], &nruns = c->pool["sk_nruns"];
    uint32_t* ovf = u32c + 9;
    if (with_hist && h_occ) {
        CUB_(hist.alloc(4 * nb));
        CUB_(cudaMemsetAsync(hist.p, 0, 4 * nb, c->st));
        CUB_(dsc.alloc(8ull * SN_MS_QUEUE * SN_MS_READS * blocks_for(n, SN_MS_READS))); CUB_(nruns.alloc(n));
        CUB_(cudaMemsetAsync(ovf, 0, 4, c->st));
    }
    for (int ch = 0; ch < nch; ++ch) {
        const uint64_t r0 = lo(ch), r1 = lo(ch + 1);
        CUB_(cudaStreamWaitEvent(c->st, c->ev_copy[MAXCH + ch], 0));
        if (with_hist && h_occ) {
            k_msp_scan<false><<<blocks_for(r1 - r0, SN_MS_READS), SN_MS_READS, 0, c->st>>>(r1 - r0, c->bases.as<uint8_t>(), c->boff.as<uint64_t>() + r0,
                c->goodlen.as<uint32_t>() + r0, nullptr, c->params.ign_bc_below, c->min_gl, bits, hist.as<uint32_t>(), nullptr, nullptr, 0u, 0xFFFFFFFFu,
                dsc.as<uint2>() + (r0 / SN_MS_READS) * (uint64_t)(SN_MS_QUEUE * SN_MS_READS), nruns.as<uint8_t>() + r0, nullptr, ovf);
            ++c->launches;
        }
    }
    k_read_stats<<<std::min(blocks_for(n, 256), 8u * (unsigned)c->num_sms), 256, 0, c->st>>>(n, c->len.as<uint32_t>(), c->have_bc ? c->bc.as<int32_t>() : nullptr,
        cnt64, reinterpret_cast<uint32_t*>(cnt64 + 1), reinterpret_cast<int32_t*>(cnt64 + 1) + 1);
    ++c->launches;
    unsigned long long h[2] = {0, 0};
    uint32_t h_ovf = 0;
    CUB_(cudaMemcpyAsync(h, cnt64, 16, cudaMemcpyDeviceToHost, c->st));
    CUB_(cudaMemcpyAsync(&h_ovf, ovf, 4, cudaMemcpyDeviceToHost, c->st));
    CUB_(cudaStreamSynchronize(c->st));                             // (st waited for every copy event: the caller's buffers are free again)
    CUB_(cudaGetLastError());
#undef CUB_
    c->cnt.n_bases = h[0];
    const uint32_t max_len = (uint32_t)(h[1] & 0xFFFFFFFFu); const int32_t max_bc = (int32_t)(h[1] >> 32);
    if (max_len > SN_MAX_READ_LEN) return fail(c, SN_ERR_ARG, "reads longer than " + std::to_string(SN_MAX_READ_LEN) + " bases are not supported");
    if (max_bc >= 0xFFFFFF) return fail(c, SN_ERR_ARG, "more than 2^24-2 distinct barcodes in one context");
    c->cnt.n_kmer_occurrences = h_occ;
    c->gl_ready = true; c->gl_min_qual = c->params.min_qual; c->gl_occ = h_occ;
    c->hist_ready_bits = (with_hist && h_occ) ? bits : -1;
    c->dsc_ready = with_hist && h_occ; c->dsc_overflow = h_ovf;
    c->stage = 1; c->reads_ok = true;
    return SN_OK;
}

int sn_set_semantics(sn_ctx* c, int semantics)
{
    if (!c || (semantics != SN_SEM_CXX && semantics != SN_SEM_TADA)) return SN_ERR_ARG;
    const uint32_t want = semantics == SN_SEM_TADA ? SN_K : SN_K + 1;
    if (want != c->min_gl) { c->min_gl = want; c->gl_ready = false; c->hist_ready_bits = -1; c->dsc_ready = false; }    // (work a streamed load did under the other rule is void)
    return SN_OK;
}

int sn_count_kmers(sn_ctx* c, const sn_params* p)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 1) return fail(c, SN_ERR_STATE, "sn_count_kmers: no reads loaded");
    CU(cudaSetDevice(c->device));
    int r; uint64_t n_occ = 0, n_sk = 0;
    if ((r = sn_i_count_set_params(c, p))) return r;
    if ((r = sn_i_count_goodlen(c, &n_occ))) return r;
    const int bits = sn_i_pick_bucket_bits(n_occ);
    // One pass holds < 2^32 k-mer occurrences (32-bit positions inside k_bucket_count's output) and its
    // super-k-mer records in HBM.  More occurrences than 3.6e9 are counted in several passes over
    // consecutive bucket ranges: every pass scans the resident reads again (0.25 B/base, cheap next to the
    // count itself) and keeps only its buckets; the passes' survivors, concatenated, are in bucket order.
    uint32_t passes = (uint32_t)(n_occ / 3600000000ull + 1);      // (bucket ranges are even to a fraction of a percent: a pass stays below 2^32)
    if (const char* e = getenv("SN_COUNT_PASSES")) { int v = atoi(e); if (v >= 1 && v <= 4096) passes = (uint32_t)v; }     // tests
    if (passes > (1u << bits)) passes = 1u << bits;
    DevBuf &surv = c->pool["surv_a"], &surv_off = c->pool["surv_off"];
    if (passes == 1) {
        if (n_occ && (r = sn_i_msp_partition(c, bits, &n_sk))) return r;
        c->cnt.n_superkmers = n_sk;
        uint64_t n_surv = 0;
        if ((r = sn_i_msp_bucket_count(c, c->pool["sk_recs"].as<uint4>(), c->pool["sk_off"].as<uint64_t>(), 1u << bits, 1, n_occ, surv, surv_off, &n_surv))) return r;
        return sn_i_msp_install_dict(c, surv.as<uint4>(), n_surv, bits, surv_off.as<uint32_t>(), true);
    }
    const uint64_t nb = 1ull << bits;
    DevBuf &all = c->pool["surv_all"], &all_cnt = c->pool["surv_all_cnt"];
    CU(all_cnt.alloc(4 * nb));
    unsigned long long* occ = c->counters.as<unsigned long long>();
    uint64_t n_total = 0, n_dist = 0, n_sk_total = 0;
    for (uint32_t ps = 0; ps < passes; ++ps) {
        const uint32_t b0 = sn_i_first_bucket(ps, passes, bits), b1 = sn_i_first_bucket(ps + 1, passes, bits);
        if (b1 == b0) continue;
        if ((r = sn_i_msp_partition(c, bits, &n_sk, b0, b1 - b0))) return r;
        n_sk_total += n_sk;
        DevBuf& recs = c->pool["sk_recs"];
        unsigned long long h_occ = 0;
        CU(cudaMemsetAsync(occ + 4, 0, 8, c->st));
        if (n_sk) { k_sum_nk<<<std::min(blocks_for(n_sk, 256), 8u * (unsigned)c->num_sms), 256, 0, c->st>>>(recs.as<uint4>(), n_sk, occ + 4); KCHECK("k_sum_nk"); }
        CU(cudaMemcpyAsync(&h_occ, occ + 4, 8, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        if (h_occ >= (1ull << 32)) return fail(c, SN_ERR_ARG, "a count pass holds more than 2^32-1 k-mer occurrences (skewed buckets): raise SN_COUNT_PASSES");
        uint64_t n_surv = 0;
        if ((r = sn_i_msp_bucket_count(c, recs.as<uint4>(), c->pool["sk_off"].as<uint64_t>(), b1 - b0, 1, h_occ, surv, surv_off, &n_surv))) return r;
        n_dist += c->cnt.n_kmers_distinct;
        if (16 * (n_total + n_surv) > all.cap) {                   // grow the concatenation (keeps what it holds)
            DevBuf bigger;
            const uint64_t want = std::max<uint64_t>(n_total + n_surv, (uint64_t)((n_total + n_surv) * (double)passes / (ps + 1)) + 1024);
            CU(bigger.alloc(16 * want));
            if (n_total) CU(cudaMemcpyAsync(bigger.p, all.p, 16 * n_total, cudaMemcpyDeviceToDevice, c->st));
            CU(cudaStreamSynchronize(c->st));
            std::swap(all.p, bigger.p); std::swap(all.bytes, bigger.bytes); std::swap(all.cap, bigger.cap);
        }
        if (n_surv) CU(cudaMemcpyAsync(all.as<uint4>() + n_total, surv.p, 16 * n_surv, cudaMemcpyDeviceToDevice, c->st));
        k_diff_u32<<<blocks_for(b1 - b0, 256), 256, 0, c->st>>>(surv_off.as<uint32_t>(), b1 - b0, all_cnt.as<uint32_t>() + b0);
        KCHECK("k_diff_u32");
        n_total += n_surv;
    }
    c->cnt.n_superkmers = n_sk_total;
    // (a job this size needs the room: the passes' temporaries go before the dictionary is laid out, the survivors after)
    pool_release(c, {"sk_recs", "sk_off", "sk_dsc", "sk_nruns", "sk_hist", "surv_a", "surv_off", "surv_scratch"});
    r = sn_i_msp_install_dict(c, all.as<uint4>(), n_total, bits, all_cnt.as<uint32_t>(), false);
    pool_release(c, {"surv_all", "surv_all_cnt"});
    c->cnt.n_kmers_distinct = n_dist;
    return r;
}

// (the multi-GPU path -- reads sharded, super-k-mers routed by bucket range, the dictionary kept sharded -- is sn_multi.cu)

// ---------------------------------------------------------------------------
int sn_build_edges(sn_ctx* c)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 2) return fail(c, SN_ERR_STATE, "sn_build_edges: run sn_count_kmers first");
    CU(cudaSetDevice(c->device));
    {   // the stop-indexed stage of the sharded path (sn_multi.cu) also runs on one rank; it is the default; SN_EDGES2=0 runs the first implementation below for A/B timing
        const char* e2 = getenv("SN_EDGES2");
        pool_report(c, "before edges");
        if (!(e2 && atoi(e2) == 0) || c->dict_sharded) return sn_i_build_edges2(c);
    }
    const uint32_t n = (uint32_t)c->cnt.n_kmers;
    c->cnt.n_edges = 0; c->cnt.n_edge_bases = 0;
    if (c->edges_copy_inflight) { cudaEventSynchronize(c->ev_edges); c->edges_copy_inflight = false; }
    c->edges_host_stale = false;
    if (!n) { resize_pinned(c, c->hedges.len, 0); resize_pinned(c, c->hedges.off, 1); c->hedges.off[0] = 0; resize_pinned(c, c->hedges.packed, 16); c->stage = 3; return SN_OK; }
    DictEntry* tab = c->dict.as<DictEntry>();
    const DictView dv = dict_view(c);
    t_begin(c, "prune");
    DevBuf& links = c->pool["links"];
    CU(links.alloc(8ull * n));
    k_prune<<<blocks_for(n, 256), 256, 0, c->st>>>(tab, dv, links.as<Link2>());
    KCHECK("k_prune");
    t_end(c, "prune");

    t_begin(c, "edges");
    DevBuf &etype = c->pool["etype"], &own_n = c->pool["own_n"], &flag = c->pool["flag"], &stop_pos = c->pool["stop_pos"], &pos = c->pool["pos"],
           &stops = c->pool["stops"], &owners = c->pool["owners"], &segs = c->pool["segs"], &sinfo = c->pool["sinfo"];
    CU(etype.alloc(n)); CU(own_n.alloc(4ull * n)); CU(flag.alloc(4ull * n)); CU(stop_pos.alloc(8ull * (n + 1))); CU(pos.alloc(8ull * (n + 1)));
    k_classify<<<blocks_for(n, 256), 256, 0, c->st>>>(dv, links.as<Link2>(), etype.as<uint8_t>(), own_n.as<uint32_t>(), flag.as<uint32_t>());
    KCHECK("k_classify");
    // stops = edge ends + sampled interiors; segments between neighbouring stops; ends hop to the far end
    uint64_t n_stops = 0;
    int r = scan_u32(c, flag.as<uint32_t>(), n, stop_pos.as<uint64_t>(), &n_stops);
    if (r) return r;
    CU(stops.alloc(4 * n_stops + 16)); CU(segs.alloc(16 * n_stops + 16)); CU(sinfo.alloc(8 * n_stops + 16));
    if (n_stops) {
        k_scatter_flagged<<<blocks_for(n, 256), 256, 0, c->st>>>(flag.as<uint32_t>(), stop_pos.as<uint64_t>(), n, stops.as<uint32_t>());
        KCHECK("k_scatter_flagged");
        CU(cudaMemsetAsync(sinfo.p, 0xFF, 8 * n_stops, c->st));
        k_seg_walk<<<blocks_for(2 * n_stops, 128), 128, 0, c->st>>>(links.as<Link2>(), stops.as<uint32_t>(), (uint32_t)n_stops, stop_pos.as<uint64_t>(), segs.as<Seg>());
        KCHECK("k_seg_walk");
        k_end_hop<<<blocks_for(n_stops, 128), 128, 0, c->st>>>(segs.as<Seg>(), stops.as<uint32_t>(), (uint32_t)n_stops, etype.as<uint8_t>(), own_n.as<uint32_t>());
        KCHECK("k_end_hop");
    }
    // allocation, first the singles and the edges with ends: bases per owner -> offsets in the
    // unpacked scratch; owner rank -> edge id
    DevBuf &ebases_u32 = c->pool["ebases_u32"], &base_off = c->pool["base_off"];
    CU(ebases_u32.alloc(4ull * n)); CU(base_off.alloc(8ull * (n + 1)));
    k_edge_sizes<<<blocks_for(n, 256), 256, 0, c->st>>>(own_n.as<uint32_t>(), etype.as<uint8_t>(), 0, n, ebases_u32.as<uint32_t>(), flag.as<uint32_t>());
    KCHECK("k_edge_sizes");
    uint64_t total_bases = 0, n_edges = 0;
    if ((r = scan_u32(c, ebases_u32.as<uint32_t>(), n, base_off.as<uint64_t>(), &total_bases))) return r;
    if ((r = scan_u32(c, flag.as<uint32_t>(), n, pos.as<uint64_t>(), &n_edges))) return r;
    // every k-mer sits on exactly one edge: what the edges so far do not hold lies on circles
    const uint64_t n_unreached = (uint64_t)n - (total_bases - (uint64_t)(SN_K - 1) * n_edges);
    const uint64_t edge_cap = n_edges + n_unreached;
    if (edge_cap >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 edges");
    CU(owners.alloc(4 * edge_cap + 16));
    DevBuf &tmpb = c->pool["tmpb"], &eflip = c->pool["eflip"], &etmp_off = c->pool["etmp_off"], &ebytes = c->pool["ebytes"];
    CU(tmpb.alloc(total_bases + n_unreached * SN_K + 16)); CU(eflip.alloc(edge_cap + 16)); CU(etmp_off.alloc(8 * edge_cap + 16)); CU(ebytes.alloc(4 * edge_cap + 16));
    CU(c->elen.alloc(4 * edge_cap + 16)); CU(c->eoff.alloc(8 * (edge_cap + 1)));
    if (n_edges) {
        k_scatter_flagged<<<blocks_for(n, 256), 256, 0, c->st>>>(flag.as<uint32_t>(), pos.as<uint64_t>(), n, owners.as<uint32_t>());
        KCHECK("k_scatter_flagged");
        k_owner_hop<<<blocks_for(n_edges, 128), 128, 0, c->st>>>(tab, links.as<Link2>(), segs.as<Seg>(), stop_pos.as<uint64_t>(), owners.as<uint32_t>(), (uint32_t)n_edges, 0u,
            etype.as<uint8_t>(), own_n.as<uint32_t>(), base_off.as<uint64_t>(), 0ull, tmpb.as<uint8_t>(), c->elen.as<uint32_t>(), etmp_off.as<uint64_t>(), sinfo.as<StopInfo>());
        KCHECK("k_owner_hop");
        if (n_stops) {
            k_seg_emit<<<blocks_for(n_stops, 128), 128, 0, c->st>>>(tab, links.as<Link2>(), stops.as<uint32_t>(), (uint32_t)n_stops, sinfo.as<StopInfo>(), segs.as<Seg>(),
                etmp_off.as<uint64_t>(), tmpb.as<uint8_t>());
            KCHECK("k_seg_emit");
        }
    }
    if (n_unreached) {
        // circles (simpleCircle / canonicalizeCircle, BuildReadQGraph48.cc:348-397): rare; numbered after the edges above
        uint32_t* u32c = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8);
        CU(cudaMemsetAsync(u32c + 4, 0, 4, c->st));
        k_circle_count<<<blocks_for(n, 128), 128, 0, c->st>>>(tab, links.as<Link2>(), n, etype.as<uint8_t>(), own_n.as<uint32_t>(), u32c + 4);
        KCHECK("k_circle_count");
        k_edge_sizes<<<blocks_for(n, 256), 256, 0, c->st>>>(own_n.as<uint32_t>(), etype.as<uint8_t>(), 1, n, ebases_u32.as<uint32_t>(), flag.as<uint32_t>());
        KCHECK("k_edge_sizes");
        uint64_t circle_bases = 0, n_circles = 0;
        if ((r = scan_u32(c, ebases_u32.as<uint32_t>(), n, base_off.as<uint64_t>(), &circle_bases))) return r;
        if ((r = scan_u32(c, flag.as<uint32_t>(), n, pos.as<uint64_t>(), &n_circles))) return r;
        if (n_circles > n_unreached || circle_bases > n_unreached * SN_K) return fail(c, SN_ERR_DATA, "circle stage: inconsistent sizes");
        if (n_circles) {
            k_scatter_flagged<<<blocks_for(n, 256), 256, 0, c->st>>>(flag.as<uint32_t>(), pos.as<uint64_t>(), n, owners.as<uint32_t>() + n_edges);
            KCHECK("k_scatter_flagged");
            k_owner_hop<<<blocks_for(n_circles, 128), 128, 0, c->st>>>(tab, links.as<Link2>(), segs.as<Seg>(), stop_pos.as<uint64_t>(), owners.as<uint32_t>() + n_edges,
                (uint32_t)n_circles, (uint32_t)n_edges, etype.as<uint8_t>(), own_n.as<uint32_t>(), base_off.as<uint64_t>(), total_bases, tmpb.as<uint8_t>(),
                c->elen.as<uint32_t>(), etmp_off.as<uint64_t>(), sinfo.as<StopInfo>());
            KCHECK("k_owner_hop");
        }
        total_bases += circle_bases; n_edges += n_circles;
    }
    k_edge_form<<<blocks_for(n_edges, 128), 128, 0, c->st>>>(tmpb.as<uint8_t>(), etmp_off.as<uint64_t>(), c->elen.as<uint32_t>(), (uint32_t)n_edges, eflip.as<uint8_t>());
    KCHECK("k_edge_form");
    k_fix_offsets<<<blocks_for(n, 256), 256, 0, c->st>>>(tab, n, c->elen.as<uint32_t>(), eflip.as<uint8_t>());
    KCHECK("k_fix_offsets");
    k_edge_bytes<<<blocks_for(n_edges, 256), 256, 0, c->st>>>(c->elen.as<uint32_t>(), (uint32_t)n_edges, ebytes.as<uint32_t>());
    KCHECK("k_edge_bytes");
    uint64_t total_bytes = 0;
    if ((r = scan_u32(c, ebytes.as<uint32_t>(), n_edges, c->eoff.as<uint64_t>(), &total_bytes))) return r;
    CU(c->ebases.alloc(total_bytes + 64));
    CU(cudaMemsetAsync((char*)c->ebases.p + total_bytes, 0, 64, c->st));
    k_pack_edges<<<blocks_for(total_bytes, 256), 256, 0, c->st>>>(tmpb.as<uint8_t>(), etmp_off.as<uint64_t>(), c->elen.as<uint32_t>(), eflip.as<uint8_t>(),
        c->eoff.as<uint64_t>(), (uint32_t)n_edges, total_bytes, c->ebases.as<uint8_t>());
    KCHECK("k_pack_edges");
    t_end(c, "edges");
    // every dictionary entry must now sit on exactly one edge
    c->cnt.n_edges = n_edges; c->cnt.n_edge_bases = total_bases;
    resize_pinned(c, c->hedges.len, n_edges); resize_pinned(c, c->hedges.off, n_edges + 1); resize_pinned(c, c->hedges.packed, total_bytes + 16);
    memset(c->hedges.packed.data() + total_bytes, 0, 16);
    CU(cudaMemcpyAsync(c->hedges.len.data(), c->elen.p, 4 * n_edges, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(c->hedges.off.data(), c->eoff.p, 8 * (n_edges + 1), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(c->hedges.packed.data(), c->ebases.p, total_bytes, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    uint64_t kmers_on_edges = total_bases - (uint64_t)(SN_K - 1) * n_edges;
    if (kmers_on_edges != n)
        return fail(c, SN_ERR_DATA, "edge stage covered " + std::to_string(kmers_on_edges) + " of " + std::to_string(n) + " dictionary k-mers");
    c->stage = 3;
    return SN_OK;
}

// ---------------------------------------------------------------------------
int sn_build_hbv(sn_ctx* c)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 3) return fail(c, SN_ERR_STATE, "sn_build_hbv: run sn_build_edges first");
    CU(cudaSetDevice(c->device));
    pool_report(c, "before hbv");
    const uint32_t nE = (uint32_t)c->cnt.n_edges;
    c->cnt.n_hbv_vertices = 0; c->cnt.n_hbv_edges = 0;
    if (!nE) {
        snh::Hbv& H0 = c->hbv;
        H0.n_vert = 0; H0.fwd.clear(); H0.rev.clear(); H0.src.clear(); H0.to_left.clear(); H0.to_right.clear(); H0.inv.clear();
        H0.from_v.clear(); H0.from_e.clear(); H0.to_v.clear(); H0.to_e.clear();
        H0.from_start.assign(1, 0); H0.to_start.assign(1, 0);
        c->hbv_host_stale = false;
        c->stage = 4; return SN_OK;
    }
    if (nE >= (1u << 29)) return fail(c, SN_ERR_ARG, "more than 2^29 unipath edges");
    const uint32_t n4 = 4 * nE, n_items = 2 * nE;
    uint32_t* u32c = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8);        // [6] npal, [7] error flags
    DevBuf &ord_a = c->pool["ord_a"], &ord_b = c->pool["ord_b"], &end_a = c->pool["end_a"], &end_b = c->pool["end_b"], &tmp = c->pool["hbv_tmp"],
           &pal = c->pool["pal"], &flag = c->pool["hflag"], &pos = c->pool["hpos"], &egrp = c->pool["egrp"], &items = c->pool["items"], &gstart = c->pool["gstart"],
           &order = c->pool["hbv_order"], &rank = c->pool["hbv_rank"], &groups = c->pool["hbv_groups"], &er = c->pool["hbv_er"],
           &parent = c->pool["hbv_parent"], &comp = c->pool["hbv_comp"], &ckey = c->pool["hbv_ckey"], &cntv = c->pool["hbv_cntv"], &cnte = c->pool["hbv_cnte"],
           &crec_a = c->pool["hbv_crec_a"], &crec_b = c->pool["hbv_crec_b"], &cstart = c->pool["hbv_cstart"], &cv = c->pool["hbv_cv"], &ce = c->pool["hbv_ce"],
           &basev = c->pool["hbv_basev"], &basee = c->pool["hbv_basee"];
    // ---- a8 on the device: end keys, exact edge order, vertex groups --------------------------------
    t_begin(c, "hbv_dev");
    CU(ord_a.alloc(16ull * nE)); CU(ord_b.alloc(16ull * nE)); CU(end_a.alloc(16ull * n4)); CU(end_b.alloc(16ull * n4));
    CU(tmp.alloc(radix_sort_tmp_bytes(n4))); CU(pal.alloc(nE)); CU(flag.alloc(4ull * n4)); CU(pos.alloc(8ull * (n4 + 1)));
    CU(egrp.alloc(4ull * n4)); CU(items.alloc(4ull * n4)); CU(gstart.alloc(4ull * (n4 + 1)));
    CU(order.alloc(4ull * nE)); CU(rank.alloc(4ull * nE)); CU(er.alloc(16ull * n_items));
    CU(cudaMemsetAsync(egrp.p, 0xFF, 4ull * n4, c->st));
    CU(cudaMemsetAsync(u32c + 6, 0, 8, c->st));
    k_hbv_keys<<<blocks_for(nE, 128), 128, 0, c->st>>>(c->ebases.as<uint8_t>(), c->eoff.as<uint64_t>(), c->elen.as<uint32_t>(), nE,
        ord_a.as<uint4>(), end_a.as<uint4>(), pal.as<uint8_t>(), u32c + 6);
    KCHECK("k_hbv_keys");
    cudaError_t e = radix_sort<RS_KEY96>(ord_a.as<uint4>(), ord_b.as<uint4>(), nE, tmp.p, c->num_sms, c->st);
    if (e == cudaSuccess) e = radix_sort<RS_KEY96>(end_a.as<uint4>(), end_b.as<uint4>(), n4, tmp.p, c->num_sms, c->st);
    c->launches += 2 * (2 + RsMode<RS_KEY96>::PASSES);
    if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, std::string("hbv sort: ") + cudaGetErrorString(e));
    k_hbv_rank<<<blocks_for(nE, 128), 128, 0, c->st>>>(ord_a.as<uint4>(), nE, c->ebases.as<uint8_t>(), c->eoff.as<uint64_t>(), c->elen.as<uint32_t>(),
        order.as<uint32_t>(), rank.as<uint32_t>());
    KCHECK("k_hbv_rank");
    k_hbv_mark<<<blocks_for(n4, 256), 256, 0, c->st>>>(end_a.as<uint4>(), n4, flag.as<uint32_t>());
    KCHECK("k_hbv_mark");
    uint64_t n_groups = 0;
    int r0 = scan_u32(c, flag.as<uint32_t>(), n4, pos.as<uint64_t>(), &n_groups);
    if (r0) return r0;
    uint32_t h_npal = 0;
    CU(cudaMemcpy(&h_npal, u32c + 6, 4, cudaMemcpyDeviceToHost));
    const uint32_t n_valid = n4 - 2 * h_npal;                  // the invalid (palindrome rc) end records sort last
    const uint32_t nV = (uint32_t)n_groups;
    k_hbv_assign<<<blocks_for(n4, 256), 256, 0, c->st>>>(end_a.as<uint4>(), n4, flag.as<uint32_t>(), pos.as<uint64_t>(),
        egrp.as<int32_t>(), items.as<uint32_t>(), gstart.as<uint32_t>());
    KCHECK("k_hbv_assign");
    CU(groups.alloc(64ull * nV));
    k_hbv_groups<<<blocks_for(nV, 128), 128, 0, c->st>>>(items.as<uint32_t>(), gstart.as<uint32_t>(), nV, n_valid, rank.as<uint32_t>(),
        groups.as<snh::GroupRec>(), u32c + 7);
    KCHECK("k_hbv_groups");
    k_hbv_erec<<<blocks_for(n_items, 256), 256, 0, c->st>>>(egrp.as<int32_t>(), pal.as<uint8_t>(), n_items, er.as<snh::ERec>());
    KCHECK("k_hbv_erec");
    // the numbering loop's records travel to pinned host memory while the components are analysed
    DevBuf& irec = c->pool["hbv_irec"];
    CU(irec.alloc(64ull * n_items));
    k_hbv_itemrec<<<blocks_for(n_items, 128), 128, 0, c->st>>>(egrp.as<int32_t>(), pal.as<uint8_t>(), groups.as<snh::GroupRec>(), n_items, irec.as<snh::ItemRec>());
    KCHECK("k_hbv_itemrec");
    HostBuf &h_groups = c->hpool["hbv_groups"], &h_irec = c->hpool["hbv_irec"], &h_comp = c->hpool["hbv_comp"];
    // ---- connected components, their discovery order and id bases ----------------------------------------
    CU(parent.alloc(4ull * nV)); CU(comp.alloc(4ull * nV)); CU(ckey.alloc(8ull * nV)); CU(cntv.alloc(4ull * nV)); CU(cnte.alloc(4ull * nV));
    k_hbv_uf_init<<<blocks_for(nV, 256), 256, 0, c->st>>>(parent.as<uint32_t>(), nV, ckey.as<unsigned long long>(), cntv.as<uint32_t>(), cnte.as<uint32_t>());
    KCHECK("k_hbv_uf_init");
    k_hbv_union<<<blocks_for(n_items, 256), 256, 0, c->st>>>(er.as<snh::ERec>(), n_items, parent.as<uint32_t>());
    KCHECK("k_hbv_union");
    k_hbv_flatten<<<blocks_for(nV, 256), 256, 0, c->st>>>(parent.as<uint32_t>(), nV, comp.as<uint32_t>(), cntv.as<uint32_t>());
    KCHECK("k_hbv_flatten");
    k_hbv_compstats<<<blocks_for(n_items, 256), 256, 0, c->st>>>(er.as<snh::ERec>(), n_items, comp.as<uint32_t>(), rank.as<uint32_t>(),
        ckey.as<unsigned long long>(), cnte.as<uint32_t>());
    KCHECK("k_hbv_compstats");
    k_hbv_rootflag<<<blocks_for(nV, 256), 256, 0, c->st>>>(comp.as<uint32_t>(), nV, flag.as<uint32_t>());
    KCHECK("k_hbv_rootflag");
    uint64_t n_comp = 0;
    if ((r0 = scan_u32(c, flag.as<uint32_t>(), nV, pos.as<uint64_t>(), &n_comp))) return r0;
    CU(crec_a.alloc(16 * n_comp + 16)); CU(crec_b.alloc(16 * n_comp + 16)); CU(cstart.alloc(4 * n_comp + 16)); CU(cv.alloc(4 * n_comp + 16)); CU(ce.alloc(4 * n_comp + 16));
    CU(basev.alloc(8 * (n_comp + 1))); CU(basee.alloc(8 * (n_comp + 1)));
    k_hbv_roots<<<blocks_for(nV, 256), 256, 0, c->st>>>(flag.as<uint32_t>(), pos.as<uint64_t>(), nV, ckey.as<unsigned long long>(), crec_a.as<uint4>());
    KCHECK("k_hbv_roots");
    e = radix_sort<RS_KEY96>(crec_a.as<uint4>(), crec_b.as<uint4>(), (uint32_t)n_comp, tmp.p, c->num_sms, c->st);
    c->launches += 2 + RsMode<RS_KEY96>::PASSES;
    if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, std::string("hbv component sort: ") + cudaGetErrorString(e));
    k_hbv_compgather<<<blocks_for(n_comp, 256), 256, 0, c->st>>>(crec_a.as<uint4>(), (uint32_t)n_comp, order.as<uint32_t>(), cntv.as<uint32_t>(), cnte.as<uint32_t>(),
        cstart.as<uint32_t>(), cv.as<uint32_t>(), ce.as<uint32_t>());
    KCHECK("k_hbv_compgather");
    uint64_t tot_v = 0, tot_h = 0;
    if ((r0 = scan_u32(c, cv.as<uint32_t>(), n_comp, basev.as<uint64_t>(), &tot_v))) return r0;
    if ((r0 = scan_u32(c, ce.as<uint32_t>(), n_comp, basee.as<uint64_t>(), &tot_h))) return r0;
    // ---- the FIFO numbering: small components on the device, one thread each (k_hbv_number_small) -------------------
    const uint64_t nH = tot_h;
    if (tot_h >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 HBV edges");
    uint32_t max_items = SN_HBV_GPU_MAX;
    if (getenv("SN_HBV_LAYOUT_MIN")) max_items = 0;                 // tests of the host path: every component goes there
    if (const char* e = getenv("SN_HBV_GPU_MAX")) { const int v = atoi(e); if (v >= 0 && v <= SN_HBV_GPU_MAX) max_items = (uint32_t)v; }
    DevBuf &seen = c->pool["hbv_seen"], &vbits = c->pool["hbv_vbits"], &vid = c->pool["hbv_vid"];
    CU(seen.alloc(n_items + 16)); CU(vbits.alloc(nV + 16)); CU(vid.alloc(4ull * nV + 16));
    CU(c->d_fwd.alloc(4ull * nE + 16)); CU(c->d_rev.alloc(4ull * nE + 16)); CU(c->d_toleft.alloc(4 * nH + 16)); CU(c->d_toright.alloc(4 * nH + 16)); CU(c->d_src.alloc(4 * nH + 16));
    CU(cudaMemsetAsync(seen.p, 0, n_items + 16, c->st)); CU(cudaMemsetAsync(vbits.p, 0, nV + 16, c->st));
    CU(cudaMemsetAsync(c->d_fwd.p, 0, 4ull * nE + 16, c->st)); CU(cudaMemsetAsync(c->d_rev.p, 0, 4ull * nE + 16, c->st));
    CU(cudaMemsetAsync(c->d_toleft.p, 0, 4 * nH + 16, c->st)); CU(cudaMemsetAsync(c->d_toright.p, 0, 4 * nH + 16, c->st)); CU(cudaMemsetAsync(c->d_src.p, 0, 4 * nH + 16, c->st));
    CU(cudaMemsetAsync(u32c + 16, 0, 8, c->st));
    if (n_comp) {
        k_hbv_number_small<<<blocks_for(n_comp, 64), 64, 0, c->st>>>(irec.as<snh::ItemRec>(), groups.as<snh::GroupRec>(), (uint32_t)n_comp, cstart.as<uint32_t>(), ce.as<uint32_t>(),
            basev.as<uint64_t>(), basee.as<uint64_t>(), max_items, seen.as<uint8_t>(), vbits.as<uint8_t>(), vid.as<int32_t>(),
            c->d_src.as<uint32_t>(), c->d_toleft.as<int32_t>(), c->d_toright.as<int32_t>(), c->d_fwd.as<int32_t>(), c->d_rev.as<int32_t>(), u32c + 16, u32c + 17);
        KCHECK("k_hbv_number_small");
    }
    t_end(c, "hbv_dev");
    uint32_t h_err = 0, h_big[2] = {0, 0};
    CU(cudaMemcpyAsync(&h_err, u32c + 7, 4, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(h_big, u32c + 16, 8, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (h_err) return fail(c, SN_ERR_DATA, "HBV: a vertex has more than 8 edge ends (HBVFromEdges.cc:83)");
    if (h_big[1]) return fail(c, SN_ERR_DATA, "HBV: the device numbering of a component failed (flags " + std::to_string(h_big[1]) + ")");
    if (tot_v != nV) return fail(c, SN_ERR_DATA, "HBV: component vertex counts do not add up");
    c->host_ms["hbv_layout"] = 0.0; c->host_ms["hbv_host"] = 0.0;
    snh::Hbv& H = c->hbv;
    H.n_vert = (int32_t)nV;
    c->cnt.n_hbv_vertices = nV; c->cnt.n_hbv_edges = nH;

    if (h_big[0]) {
    // ---- big components: the host loop (sn_hbv.cpp) over the records, which travel to pinned host memory ---------------
    CU(h_groups.alloc(64ull * nV)); CU(h_irec.alloc(64ull * n_items));
    CU(cudaMemcpyAsync(h_irec.p, irec.p, 64ull * n_items, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(h_groups.p, groups.p, 64ull * nV, cudaMemcpyDeviceToHost, c->st));
    CU(h_comp.alloc(4 * n_comp + 16 * (n_comp + 1) + 64));
    uint64_t* h_basev = h_comp.as<uint64_t>(); uint64_t* h_basee = h_basev + (n_comp + 1); uint32_t* h_cstart = reinterpret_cast<uint32_t*>(h_basee + (n_comp + 1));
    CU(cudaMemcpyAsync(h_basev, basev.p, 8 * (n_comp + 1), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(h_basee, basee.p, 8 * (n_comp + 1), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(h_cstart, cstart.p, 4 * n_comp, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    // ---- a giant component: lay the records out along the graph first (sn_hbvdev.cuh, k_lay_*) -----------
    const uint32_t* layout = nullptr;
    {
        uint64_t max_comp = 0;
        for (uint64_t k = 0; k < n_comp; ++k) max_comp = std::max(max_comp, h_basee[k + 1] - h_basee[k]);
        const bool no_layout = getenv("SN_HBV_NO_LAYOUT") != nullptr;
        const uint64_t lay_min = getenv("SN_HBV_LAYOUT_MIN") ? (uint64_t)atoll(getenv("SN_HBV_LAYOUT_MIN")) : 262144;      // tests lower it
        c->host_ms["hbv_layout"] = 0.0;
        if (max_comp >= lay_min && !no_layout) {
            auto tl0 = std::chrono::steady_clock::now();
            DevBuf &lab = c->pool["lay_lab"], &lev = c->pool["lay_lev"], &fa = c->pool["lay_fa"], &fb = c->pool["lay_fb"], &ka = c->pool["lay_ka"], &kb = c->pool["lay_kb"],
                   &posd = c->pool["lay_pos"], &irec2 = c->pool["lay_irec"], &groups2 = c->pool["lay_groups"];
            CU(lab.alloc(4ull * n_items)); CU(lev.alloc(4ull * n_items)); CU(fa.alloc(4ull * n_items)); CU(fb.alloc(4ull * n_items));
            CU(ka.alloc(16ull * n_items)); CU(kb.alloc(16ull * n_items)); CU(posd.alloc(4ull * n_items)); CU(irec2.alloc(64ull * n_items)); CU(groups2.alloc(64ull * nV));
            uint32_t* cnt3 = u32c + 12;
            CU(cudaMemsetAsync(cnt3, 0, 12, c->st));
            k_lay_seed<<<blocks_for(n_items, 256), 256, 0, c->st>>>(irec.as<snh::ItemRec>(), n_items, lab.as<uint32_t>(), lev.as<uint32_t>(), fa.as<uint32_t>(), cnt3);
            KCHECK("k_lay_seed");
            const unsigned lay_grid = (unsigned)std::min<uint64_t>(2 * c->num_sms, blocks_for(n_items / 32 + 1, 256));
            const uint32_t lay_rounds = 64;              // seeds are ~32 items apart: what 64 rounds do not reach becomes its own cluster
            for (uint32_t round = 0; round < lay_rounds; ++round)
                k_lay_round<<<lay_grid, 256, 0, c->st>>>(irec.as<snh::ItemRec>(), groups.as<snh::GroupRec>(), round, lab.as<uint32_t>(), lev.as<uint32_t>(),
                    fa.as<uint32_t>(), fb.as<uint32_t>(), cnt3);
            c->launches += lay_rounds;
            k_lay_keys<<<blocks_for(n_items, 256), 256, 0, c->st>>>(lab.as<uint32_t>(), lev.as<uint32_t>(), n_items, ka.as<uint4>());
            KCHECK("k_lay_keys");
            e = radix_sort<RS_KEY96>(ka.as<uint4>(), kb.as<uint4>(), n_items, tmp.p, c->num_sms, c->st);
            c->launches += 2 + RsMode<RS_KEY96>::PASSES;
            if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, std::string("hbv layout sort: ") + cudaGetErrorString(e));
            k_lay_pos<<<blocks_for(n_items, 256), 256, 0, c->st>>>(ka.as<uint4>(), n_items, posd.as<uint32_t>());
            KCHECK("k_lay_pos");
            k_lay_permute<<<blocks_for(n_items, 128), 128, 0, c->st>>>(ka.as<uint4>(), posd.as<uint32_t>(), irec.as<snh::ItemRec>(), n_items, irec2.as<snh::ItemRec>());
            KCHECK("k_lay_permute");
            k_lay_groups<<<blocks_for(nV, 256), 256, 0, c->st>>>(groups.as<snh::GroupRec>(), nV, posd.as<uint32_t>(), groups2.as<snh::GroupRec>());
            KCHECK("k_lay_groups");
            HostBuf& h_lay = c->hpool["hbv_layout"];
            CU(h_lay.alloc(4ull * n_items));
            CU(cudaMemcpyAsync(h_irec.p, irec2.p, 64ull * n_items, cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(h_groups.p, groups2.p, 64ull * nV, cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(h_lay.p, posd.p, 4ull * n_items, cudaMemcpyDeviceToHost, c->st));
            CU(cudaStreamSynchronize(c->st));
            layout = h_lay.as<uint32_t>();
            c->host_ms["hbv_layout"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl0).count();
        }
    }
    auto t0 = std::chrono::steady_clock::now();
    snh::HbvComponents comps; comps.n_comp = n_comp; comps.start_item = h_cstart; comps.base_v = h_basev; comps.base_e = h_basee;
    static const unsigned hbv_threads = [] { const char* e = getenv("SN_HBV_THREADS"); unsigned n = e ? (unsigned)atoi(e) : std::min(32u, std::thread::hardware_concurrency()); return n ? n : 1u; }();
    snh::Hbv Hh;                                                   // the host's share: zero where the device (or another rank) numbers
    Hh.src.assign(tot_h, 0); Hh.to_left.assign(tot_h, 0); Hh.to_right.assign(tot_h, 0); Hh.fwd.assign(nE, 0); Hh.rev.assign(nE, 0);
    // several ranks (sn_multi.cu): the graph is the same on all of them, so every rank numbers 1/N of the big components
    // (components carry their final id bases) and the arrays are summed on the device by one all-reduce each
    const unsigned n_parts = c->comm && c->comm->n > 1 ? (unsigned)c->comm->n : 1u, part = n_parts > 1 ? (unsigned)c->comm->rank : 0u;
    const unsigned nthreads = n_parts > 1 ? std::max(2u, hbv_threads / n_parts) : hbv_threads;      // (the ranks of one box share its cores)
    try { snh::number_hbv(comps, h_irec.as<snh::ItemRec>(), h_groups.as<snh::GroupRec>(), nV, nE, Hh, nthreads, layout, part, n_parts, (uint64_t)max_items + 1); }
    catch (const std::exception& ex) { return fail(c, SN_ERR_DATA, ex.what()); }
    c->host_ms["hbv_host"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    DevBuf& hs = c->pool["hbv_hostshare"];
    CU(hs.alloc(4 * std::max<uint64_t>(nH, nE) + 16));
    struct { DevBuf* d; const void* h; size_t n; } parts[5] = {{&c->d_fwd, Hh.fwd.data(), nE}, {&c->d_rev, Hh.rev.data(), nE}, {&c->d_toleft, Hh.to_left.data(), nH},
                                                                {&c->d_toright, Hh.to_right.data(), nH}, {&c->d_src, Hh.src.data(), nH}};
    for (auto& q : parts) {
        CU(cudaMemcpyAsync(hs.p, q.h, 4 * q.n, cudaMemcpyHostToDevice, c->st));
        if (n_parts > 1 && c->comm->allreduce_sum(hs.p, q.n, 4, c->st)) return fail(c, SN_ERR_CUDA, "allreduce (HBV numbering): " + c->comm->err);
        k_add_u32<<<blocks_for(q.n, 256), 256, 0, c->st>>>(q.d->as<uint32_t>(), hs.as<uint32_t>(), q.n);
        KCHECK("k_add_u32");
        CU(cudaStreamSynchronize(c->st));                          // (hs and the pageable host vectors are reused)
    }
    }
    // ---- adjacency lists and involution on the device; the graph stays there for the pathing stage ----
    t_begin(c, "hbv_csr");
    CU(c->d_from_start.alloc(4ull * (nV + 1) + 16)); CU(c->d_to_start.alloc(4ull * (nV + 1) + 16));
    CU(c->d_from_v.alloc(4 * nH + 16)); CU(c->d_from_e.alloc(4 * nH + 16)); CU(c->d_to_v.alloc(4 * nH + 16)); CU(c->d_to_e.alloc(4 * nH + 16));
    DevBuf &ra = c->pool["hbv_csr_a"], &rb = c->pool["hbv_csr_b"], &rs = c->pool["hbv_csr_s"], &dinv = c->pool["hbv_inv"];
    CU(ra.alloc(16 * nH + 16)); CU(rb.alloc(16 * nH + 16)); CU(rs.alloc(16 * nH + 16)); CU(dinv.alloc(4 * nH + 16));
    CU(tmp.alloc(radix_sort_tmp_bytes((uint32_t)nH)));
    k_hbv_csr_rec<<<blocks_for(nH, 256), 256, 0, c->st>>>(c->d_toleft.as<int32_t>(), c->d_toright.as<int32_t>(), (uint32_t)nH, ra.as<uint4>(), rb.as<uint4>());
    KCHECK("k_hbv_csr_rec");
    const unsigned csr_blocks = blocks_for(std::max<uint64_t>(nH, (uint64_t)nV + 1), 256);
    e = radix_sort<RS_KEY96>(ra.as<uint4>(), rs.as<uint4>(), (uint32_t)nH, tmp.p, c->num_sms, c->st);
    if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, std::string("hbv adjacency sort: ") + cudaGetErrorString(e));
    k_hbv_csr_emit<<<csr_blocks, 256, 0, c->st>>>(ra.as<uint4>(), (uint32_t)nH, nV, c->d_from_start.as<uint32_t>(), c->d_from_v.as<int32_t>(), c->d_from_e.as<int32_t>());
    KCHECK("k_hbv_csr_emit");
    e = radix_sort<RS_KEY96>(rb.as<uint4>(), rs.as<uint4>(), (uint32_t)nH, tmp.p, c->num_sms, c->st);
    if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, std::string("hbv adjacency sort: ") + cudaGetErrorString(e));
    k_hbv_csr_emit<<<csr_blocks, 256, 0, c->st>>>(rb.as<uint4>(), (uint32_t)nH, nV, c->d_to_start.as<uint32_t>(), c->d_to_v.as<int32_t>(), c->d_to_e.as<int32_t>());
    KCHECK("k_hbv_csr_emit");
    c->launches += 2 * (2 + RsMode<RS_KEY96>::PASSES);
    k_hbv_inv<<<blocks_for(nE, 256), 256, 0, c->st>>>(c->d_fwd.as<int32_t>(), c->d_rev.as<int32_t>(), nE, dinv.as<int32_t>());
    KCHECK("k_hbv_inv");
    t_end(c, "hbv_csr");
    // the graph in host memory: on one rank of a multi-GPU job; the others keep it on the device until asked (sn_get_hbv, sn_write_hbv ...)
    c->hbv_host_stale = true;
    c->stage = 4;
    if (!(c->comm && c->comm->n > 1 && c->comm->rank != 0)) { int rf = sn_i_fetch_hbv_host(c); if (!rf) rf = sn_i_fetch_edges_host(c); if (rf) { c->stage = 3; return rf; } }
    else CU(cudaStreamSynchronize(c->st));
    return SN_OK;
}
int sn_i_fetch_hbv_host(sn_ctx* c)
{
    if (!c->hbv_host_stale) return SN_OK;
    CU(cudaSetDevice(c->device));
    snh::Hbv& H = c->hbv;
    const uint64_t nH = c->cnt.n_hbv_edges, nV = c->cnt.n_hbv_vertices, nE = c->cnt.n_edges;
    DevBuf& dinv = c->pool["hbv_inv"];
    resize_pinned(c, H.src, nH); resize_pinned(c, H.to_left, nH); resize_pinned(c, H.to_right, nH); resize_pinned(c, H.fwd, nE); resize_pinned(c, H.rev, nE);
    CU(cudaMemcpyAsync(H.src.data(), c->d_src.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.to_left.data(), c->d_toleft.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.to_right.data(), c->d_toright.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.fwd.data(), c->d_fwd.p, 4ull * nE, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.rev.data(), c->d_rev.p, 4ull * nE, cudaMemcpyDeviceToHost, c->st));
    resize_pinned(c, H.from_start, nV + 1); resize_pinned(c, H.to_start, nV + 1);
    resize_pinned(c, H.from_v, nH); resize_pinned(c, H.from_e, nH); resize_pinned(c, H.to_v, nH); resize_pinned(c, H.to_e, nH); resize_pinned(c, H.inv, nH);
    CU(cudaMemcpyAsync(H.from_start.data(), c->d_from_start.p, 4ull * (nV + 1), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.to_start.data(), c->d_to_start.p, 4ull * (nV + 1), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.from_v.data(), c->d_from_v.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.from_e.data(), c->d_from_e.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.to_v.data(), c->d_to_v.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.to_e.data(), c->d_to_e.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(H.inv.data(), dinv.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    c->hbv_host_stale = false;
    // every side of a vertex holds at most 4 edges (one per base)
    for (uint64_t v = 0; v < nV; ++v)
        if (H.from_start[v + 1] - H.from_start[v] > 4 || H.to_start[v + 1] - H.to_start[v] > 4)
            return fail(c, SN_ERR_DATA, "HBV: more than 4 edges on one side of a vertex");
    return SN_OK;
}
int sn_i_start_edges_copy(sn_ctx* c, uint64_t total_bytes)
{
    CU(cudaSetDevice(c->device));
    if (!c->st2) CU(cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking));
    if (!c->ev_edges) { CU(cudaEventCreateWithFlags(&c->ev_edges, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->ev_edges_go, cudaEventDisableTiming)); }
    const uint64_t E = c->cnt.n_edges;
    resize_pinned(c, c->hedges.len, E); resize_pinned(c, c->hedges.off, E + 1); resize_pinned(c, c->hedges.packed, total_bytes + 16);
    memset(c->hedges.packed.data() + total_bytes, 0, 16);
    CU(cudaEventRecord(c->ev_edges_go, c->st));
    CU(cudaStreamWaitEvent(c->st2, c->ev_edges_go, 0));
    if (E) CU(cudaMemcpyAsync(c->hedges.len.data(), c->elen.p, 4 * E, cudaMemcpyDeviceToHost, c->st2));
    CU(cudaMemcpyAsync(c->hedges.off.data(), c->eoff.p, 8 * (E + 1), cudaMemcpyDeviceToHost, c->st2));
    if (total_bytes) CU(cudaMemcpyAsync(c->hedges.packed.data(), c->ebases.p, total_bytes, cudaMemcpyDeviceToHost, c->st2));
    CU(cudaEventRecord(c->ev_edges, c->st2));
    c->edges_copy_inflight = true; c->edges_host_stale = true;
    return SN_OK;
}
int sn_i_fetch_edges_host(sn_ctx* c)
{
    if (!c->edges_host_stale) return SN_OK;
    CU(cudaSetDevice(c->device));
    if (c->edges_copy_inflight) { CU(cudaEventSynchronize(c->ev_edges)); c->edges_copy_inflight = false; c->edges_host_stale = false; return SN_OK; }
    const uint64_t E = c->cnt.n_edges;
    uint64_t total_bytes = 0;
    CU(cudaMemcpyAsync(&total_bytes, c->eoff.as<uint64_t>() + E, 8, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    resize_pinned(c, c->hedges.len, E); resize_pinned(c, c->hedges.off, E + 1); resize_pinned(c, c->hedges.packed, total_bytes + 16);
    memset(c->hedges.packed.data() + total_bytes, 0, 16);
    if (E) CU(cudaMemcpyAsync(c->hedges.len.data(), c->elen.p, 4 * E, cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(c->hedges.off.data(), c->eoff.p, 8 * (E + 1), cudaMemcpyDeviceToHost, c->st));
    if (total_bytes) CU(cudaMemcpyAsync(c->hedges.packed.data(), c->ebases.p, total_bytes, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    c->edges_host_stale = false;
    return SN_OK;
}

// ---------------------------------------------------------------------------
// buildGraphFromMSP (paths/long/BuildReadQGraph48.h:24-26, .cc:1631-1684) up to its pathReads call: the edges of
// an MSPEDGES file (vec<basevector>: tada's asm_graph, or sn_write_edges_bv) become the context's edge set,
// the HyperBasevector is built from them (mspEdgesToHBV = buildHBVFromEdges) and every edge k-mer enters the
// dictionary with its (edge, offset) (:1656-1664).  sn_path_reads / sn_write_paths / sn_write_hbv follow as usual.
static int graph_from_edges(sn_ctx* c, const snf::Fastb& E)
{
    CU(cudaSetDevice(c->device));
    const uint64_t nE = E.len.size();
    if (nE >= (1ull << 29)) return fail(c, SN_ERR_ARG, "more than 2^29 unipath edges");
    uint64_t n_bases = 0, n_k64 = 0;
    for (uint32_t l : E.len) { n_bases += l; n_k64 += l >= SN_K ? l - (SN_K - 1) : 0; }
    if (n_k64 >= (1ull << 31)) return fail(c, SN_ERR_ARG, "more than 2^31 dictionary k-mers in one context");
    const uint32_t n_k = (uint32_t)n_k64;
    // the edge set, on the device and on the host (what sn_build_edges leaves behind)
    const uint64_t total_bytes = E.off[nE];
    CU(c->ebases.alloc(total_bytes + 64)); CU(c->eoff.alloc(8 * (nE + 1))); CU(c->elen.alloc(4 * nE + 16));
    CU(cudaMemsetAsync((char*)c->ebases.p + total_bytes, 0, 64, c->st));
    if (total_bytes) CU(cudaMemcpyAsync(c->ebases.p, E.var.data(), total_bytes, cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->eoff.p, E.off.data(), 8 * (nE + 1), cudaMemcpyHostToDevice, c->st));
    if (nE) CU(cudaMemcpyAsync(c->elen.p, E.len.data(), 4 * nE, cudaMemcpyHostToDevice, c->st));
    resize_pinned(c, c->hedges.len, nE); resize_pinned(c, c->hedges.off, nE + 1); resize_pinned(c, c->hedges.packed, total_bytes + 16);
    if (nE) memcpy(c->hedges.len.data(), E.len.data(), 4 * nE);
    memcpy(c->hedges.off.data(), E.off.data(), 8 * (nE + 1));
    if (total_bytes) memcpy(c->hedges.packed.data(), E.var.data(), total_bytes);
    memset(c->hedges.packed.data() + total_bytes, 0, 16);
    if (c->edges_copy_inflight) { cudaEventSynchronize(c->ev_edges); c->edges_copy_inflight = false; }
    c->cnt.n_edges = nE; c->cnt.n_edge_bases = n_bases; c->edges_host_stale = false;
    c->cnt.n_kmer_occurrences = 0; c->cnt.n_kmers_distinct = 0; c->cnt.n_superkmers = 0;
    // ---- dictionary of the edge k-mers ----
    t_begin(c, "edge_dict");
    uint32_t* err = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8) + 11;
    DevBuf &nk = c->pool["ed_nk"], &koff = c->pool["ed_koff"], &ra = c->pool["ed_a"], &rb = c->pool["ed_b"], &tmp = c->pool["rs_tmp"], &flag = c->pool["ed_flag"],
           &pos = c->pool["ed_pos"], &gids = c->pool["ed_gids"], &bcnt = c->pool["ed_bcnt"], &surv = c->pool["surv_a"], &loc = c->pool["ed_loc"];
    CU(nk.alloc(4 * nE + 16)); CU(koff.alloc(8 * (nE + 1)));
    CU(cudaMemsetAsync(err, 0, 4, c->st));
    if (nE) { k_ed_nk<<<blocks_for(nE, 256), 256, 0, c->st>>>(c->elen.as<uint32_t>(), (uint32_t)nE, nk.as<uint32_t>(), err); KCHECK("k_ed_nk"); }
    uint64_t tot = 0;
    int r = scan_u32(c, nk.as<uint32_t>(), nE, koff.as<uint64_t>(), &tot);
    if (r) return r;
    uint32_t h_err = 0;
    CU(cudaMemcpy(&h_err, err, 4, cudaMemcpyDeviceToHost));
    if (h_err) return fail(c, SN_ERR_DATA, "an edge is shorter than K = 48 bases");
    uint64_t n_unique = 0;
    int bits = 4;
    if (n_k) {
        CU(ra.alloc(16ull * n_k)); CU(rb.alloc(16ull * n_k)); CU(tmp.alloc(radix_sort_tmp_bytes(n_k))); CU(flag.alloc(4ull * n_k)); CU(pos.alloc(8ull * (n_k + 1)));
        k_ed_records<<<blocks_for(n_k, 256), 256, 0, c->st>>>(c->ebases.as<uint8_t>(), c->eoff.as<uint64_t>(), koff.as<uint64_t>(), (uint32_t)nE, n_k, ra.as<uint4>());
        KCHECK("k_ed_records");
        cudaError_t e = radix_sort<RS_KEY96>(ra.as<uint4>(), rb.as<uint4>(), n_k, tmp.p, c->num_sms, c->st);
        if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, std::string("edge dictionary sort: ") + cudaGetErrorString(e));
        c->launches += 2 + RsMode<RS_KEY96>::PASSES;
        k_ed_last_of_run<<<blocks_for(n_k, 256), 256, 0, c->st>>>(ra.as<uint4>(), n_k, flag.as<uint32_t>());
        KCHECK("k_ed_last_of_run");
        if ((r = scan_u32(c, flag.as<uint32_t>(), n_k, pos.as<uint64_t>(), &n_unique))) return r;
        while (bits < 24 && (n_unique >> bits) > 128) ++bits;           // ~128 entries per bucket, as the count path leaves them
        CU(gids.alloc(4 * n_unique + 16)); CU(bcnt.alloc(4ull << bits)); CU(surv.alloc(16 * n_unique + 16)); CU(loc.alloc(8 * n_unique + 16));
        CU(cudaMemsetAsync(bcnt.p, 0, 4ull << bits, c->st));
        k_ed_bucket_keys<<<blocks_for(n_k, 256), 256, 0, c->st>>>(ra.as<uint4>(), flag.as<uint32_t>(), pos.as<uint64_t>(), n_k, bits, rb.as<uint4>(), gids.as<uint32_t>(), bcnt.as<uint32_t>());
        KCHECK("k_ed_bucket_keys");
        e = radix_sort<RS_KEY96>(rb.as<uint4>(), ra.as<uint4>(), (uint32_t)n_unique, tmp.p, c->num_sms, c->st);
        if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, std::string("edge dictionary sort: ") + cudaGetErrorString(e));
        c->launches += 2 + RsMode<RS_KEY96>::PASSES;
        k_ed_emit<<<blocks_for(n_unique, 256), 256, 0, c->st>>>(rb.as<uint4>(), gids.as<uint32_t>(), (uint32_t)n_unique, c->ebases.as<uint8_t>(), c->eoff.as<uint64_t>(),
            koff.as<uint64_t>(), (uint32_t)nE, surv.as<uint4>(), loc.as<uint2>());
        KCHECK("k_ed_emit");
    } else { CU(bcnt.alloc(4ull << bits)); CU(cudaMemsetAsync(bcnt.p, 0, 4ull << bits, c->st)); CU(surv.alloc(64)); }
    t_begin(c, "make_dict");
    if ((r = sn_i_msp_install_dict(c, surv.as<uint4>(), n_unique, bits, bcnt.as<uint32_t>(), false))) return r;
    if (n_unique) {
        k_ed_set_loc<<<blocks_for(n_unique, 256), 256, 0, c->st>>>(c->dict.as<DictEntry>(), loc.as<uint2>(), (uint32_t)n_unique);
        KCHECK("k_ed_set_loc");
    }
    t_end(c, "edge_dict");
    CU(cudaStreamSynchronize(c->st));
    c->stage = 3;
    return sn_build_hbv(c);
}
int sn_build_graph_from_edges(sn_ctx* c, const char* bv_path)
{
    if (!c || !bv_path) return SN_ERR_ARG;
    snf::Fastb E; std::string err;
    if (!snf::read_bv(bv_path, E, err)) return fail(c, SN_ERR_IO, err);
    const uint64_t keep_reads = c->cnt.n_reads, keep_bases = c->cnt.n_bases;       // the reads (if loaded) stay for sn_path_reads
    int r = graph_from_edges(c, E);
    c->cnt.n_reads = keep_reads; c->cnt.n_bases = keep_bases;
    return r;
}

// ---------------------------------------------------------------------------
int sn_path_reads(sn_ctx* c)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "sn_path_reads: run sn_build_hbv first");
    if (!c->reads_ok) return fail(c, SN_ERR_STATE, "sn_path_reads: no reads loaded");
    CU(cudaSetDevice(c->device));
    const uint64_t n = c->cnt.n_reads;
    const DictView d = dict_view(c);
    EdgeStore es; es.bases = c->ebases.as<uint8_t>(); es.off = c->eoff.as<uint64_t>(); es.len = c->elen.as<uint32_t>();
    HbvView h;
    h.fwd_xlat = c->d_fwd.as<int32_t>(); h.rev_xlat = c->d_rev.as<int32_t>();
    h.to_left = c->d_toleft.as<int32_t>(); h.to_right = c->d_toright.as<int32_t>(); h.src = c->d_src.as<uint32_t>();
    h.from_start = c->d_from_start.as<uint32_t>(); h.from_v = c->d_from_v.as<int32_t>(); h.from_e = c->d_from_e.as<int32_t>();
    h.to_start = c->d_to_start.as<uint32_t>(); h.to_v = c->d_to_v.as<int32_t>(); h.to_e = c->d_to_e.as<int32_t>();
    CU(c->plen.alloc(4 * n)); CU(c->poffset.alloc(4 * n)); CU(c->path_off.alloc(8 * (n + 1)));
    DevBuf& scratch = c->pool["path_scratch"];
    CU(scratch.alloc(16 * n));
    uint32_t* u32c = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + 8);
    CU(cudaMemsetAsync(u32c + 5, 0, 4, c->st));
    c->paths_on_host = false; c->pi_ready = false; c->px_ready = false; c->md_ready = false;
    PathInputs in;
    in.n_reads = n; in.bases = c->bases.as<uint8_t>(); in.boff = c->boff.as<uint64_t>(); in.len = c->len.as<uint32_t>();
    in.quals = c->have_pq ? nullptr : c->quals.as<uint8_t>(); in.qoff = c->qoff.as<uint64_t>();
    in.pq = c->have_pq ? c->pq.as<uint8_t>() : nullptr; in.pq_off = c->pqoff.as<uint64_t>();
    t_begin(c, "path");
    if (c->cnt.n_kmers == 0) {
        CU(cudaMemsetAsync(c->plen.p, 0, 4 * n, c->st)); CU(cudaMemsetAsync(c->poffset.p, 0, 4 * n, c->st));
    } else {
        k_path_reads<<<blocks_for(n, 128), 128, 0, c->st>>>(in, d, es, h, c->plen.as<uint32_t>(), c->poffset.as<int32_t>(), scratch.as<int32_t>(), u32c + 5);
        KCHECK("k_path_reads");
    }
    uint64_t total = 0;
    int r = scan_u32(c, c->plen.as<uint32_t>(), n, c->path_off.as<uint64_t>(), &total);
    if (r) return r;
    CU(c->pedges.alloc(4 * total + 16));
    if (total) {
        k_path_finish<<<blocks_for(n, 128), 128, 0, c->st>>>(in, d, es, h, c->plen.as<uint32_t>(), scratch.as<int32_t>(), c->path_off.as<uint64_t>(), c->pedges.as<int32_t>());
        KCHECK("k_path_finish");
    }
    t_end(c, "path");
    uint32_t h_over = 0;
    CU(cudaMemcpyAsync(&h_over, u32c + 5, 4, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (h_over) return fail(c, SN_ERR_DATA, "a ReadPath exceeded " + std::to_string(SN_MAX_PATH) + " edges");
    c->cnt.n_path_edges = total;
    c->stage = 5;
    return SN_OK;
}

// ---------------------------------------------------------------------------
int sn_get_counts(const sn_ctx* c, sn_counts* out) { if (!c || !out) return SN_ERR_ARG; *out = c->cnt; return SN_OK; }

int sn_get_good_lengths(sn_ctx* c, uint32_t* out)
{
    if (!c || !out) return SN_ERR_ARG;
    if (c->stage < 2) return fail(c, SN_ERR_STATE, "good lengths are available after sn_count_kmers");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpy(out, c->goodlen.p, 4 * c->cnt.n_reads, cudaMemcpyDeviceToHost));
    return SN_OK;
}
// The dictionary lives in (hash, k-mer) order; the two getters below hand it out sorted by
// k-mer (the order of a sorted kmers.kvec), which is an export/debug path, not a hot one.
static int fetch_dict_sorted(sn_ctx* c, std::vector<DictEntry>& h)
{
    size_t n = c->cnt.n_kmers;
    h.resize(n);
    if (n) CU(cudaMemcpy(h.data(), c->dict.p, n * sizeof(DictEntry), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end(), [](const DictEntry& a, const DictEntry& b) {
        return a.w0 != b.w0 ? a.w0 < b.w0 : (a.w1 != b.w1 ? a.w1 < b.w1 : a.w2 < b.w2); });
    return SN_OK;
}
int sn_get_kmers(sn_ctx* c, sn_kmer_rec* out)
{
    if (!c || !out) return SN_ERR_ARG;
    if (c->stage < 2) return fail(c, SN_ERR_STATE, "no dictionary yet");
    CU(cudaSetDevice(c->device));
    std::vector<DictEntry> h;
    int r = fetch_dict_sorted(c, h); if (r) return r;
    for (size_t i = 0; i < h.size(); ++i) { out[i].w[0] = h[i].w0; out[i].w[1] = h[i].w1; out[i].w[2] = h[i].w2; out[i].count_ctx = h[i].cc; }
    return SN_OK;
}
int sn_get_kmer_graph_info(sn_ctx* c, uint8_t* ctx_pruned, uint32_t* edge, uint32_t* offset)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 3) return fail(c, SN_ERR_STATE, "run sn_build_edges first");
    CU(cudaSetDevice(c->device));
    std::vector<DictEntry> h;
    int r = fetch_dict_sorted(c, h); if (r) return r;
    for (size_t i = 0; i < h.size(); ++i) {
        if (ctx_pruned) ctx_pruned[i] = (uint8_t)h[i].ctx;
        if (edge) edge[i] = h[i].edge;
        if (offset) offset[i] = h[i].off;
    }
    return SN_OK;
}
int sn_get_edges_bytes(const sn_ctx* c, uint64_t* packed_bytes)
{
    if (!c || !packed_bytes) return SN_ERR_ARG;
    if (c->edges_host_stale) { int r = sn_i_fetch_edges_host(const_cast<sn_ctx*>(c)); if (r) return r; }
    *packed_bytes = c->hedges.off.empty() ? 0 : c->hedges.off.back();
    return SN_OK;
}
int sn_get_edges(sn_ctx* c, uint32_t* len, uint64_t* off, uint8_t* packed)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 3) return fail(c, SN_ERR_STATE, "run sn_build_edges first");
    { int r = sn_i_fetch_edges_host(c); if (r) return r; }
    const snh::Edges& E = c->hedges;
    if (len && E.n()) memcpy(len, E.len.data(), 4 * E.n());
    if (off) memcpy(off, E.off.data(), 8 * E.off.size());
    if (packed && !E.off.empty()) memcpy(packed, E.packed.data(), E.off.back());
    return SN_OK;
}
int sn_get_hbv(sn_ctx* c, uint32_t* from_start, int32_t* from_v, int32_t* from_e, uint32_t* to_start, int32_t* to_v, int32_t* to_e,
               int32_t* fwd_xlat, int32_t* rev_xlat, int32_t* inv)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    { int r = sn_i_fetch_hbv_host(c); if (r) return r; }
    const snh::Hbv& H = c->hbv;
    size_t nV = (size_t)H.n_vert, nH = H.src.size();
    if (from_start) memcpy(from_start, H.from_start.data(), 4 * (nV + 1));
    if (to_start) memcpy(to_start, H.to_start.data(), 4 * (nV + 1));
    if (nH) {
        if (from_v) memcpy(from_v, H.from_v.data(), 4 * nH);
        if (from_e) memcpy(from_e, H.from_e.data(), 4 * nH);
        if (to_v) memcpy(to_v, H.to_v.data(), 4 * nH);
        if (to_e) memcpy(to_e, H.to_e.data(), 4 * nH);
    }
    if (fwd_xlat && !H.fwd.empty()) memcpy(fwd_xlat, H.fwd.data(), 4 * H.fwd.size());
    if (rev_xlat && !H.rev.empty()) memcpy(rev_xlat, H.rev.data(), 4 * H.rev.size());
    if (inv && !H.inv.empty()) memcpy(inv, H.inv.data(), 4 * H.inv.size());
    return SN_OK;
}
static int fetch_paths(sn_ctx* c)
{
    if (c->stage < 5) return fail(c, SN_ERR_STATE, "run sn_path_reads first");
    if (c->paths_on_host) return SN_OK;
    CU(cudaSetDevice(c->device));
    uint64_t n = c->cnt.n_reads, m = c->cnt.n_path_edges;
    c->h_poffset.resize(n); c->h_path_off.resize(n + 1); c->h_pedges.resize(m);
    CU(cudaMemcpy(c->h_poffset.data(), c->poffset.p, 4 * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(c->h_path_off.data(), c->path_off.p, 8 * (n + 1), cudaMemcpyDeviceToHost));
    if (m) CU(cudaMemcpy(c->h_pedges.data(), c->pedges.p, 4 * m, cudaMemcpyDeviceToHost));
    c->paths_on_host = true;
    return SN_OK;
}
int sn_get_paths(sn_ctx* c, int32_t* offset, uint64_t* path_off, int32_t* edges)
{
    if (!c) return SN_ERR_ARG;
    int r = fetch_paths(c); if (r) return r;
    if (offset) memcpy(offset, c->h_poffset.data(), 4 * c->h_poffset.size());
    if (path_off) memcpy(path_off, c->h_path_off.data(), 8 * c->h_path_off.size());
    if (edges && !c->h_pedges.empty()) memcpy(edges, c->h_pedges.data(), 4 * c->h_pedges.size());
    return SN_OK;
}

int sn_write_hbv(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    { int r = sn_i_fetch_edges_host(c); if (r) return r; if ((r = sn_i_fetch_hbv_host(c))) return r; }
    std::string err; const snh::Hbv& H = c->hbv;
    std::vector<uint8_t> ep; std::vector<uint64_t> eo; std::vector<uint32_t> el;
    snh::hbv_edge_sequences(c->hedges, H, ep, eo, el);
    if (!snf::write_hbv(path, H.K, (uint64_t)H.n_vert, H.from_start.data(), H.from_v.data(), H.from_e.data(), H.to_start.data(), H.to_e.data(),
                        ep.data(), eo.data(), el.data(), el.size(), err))
        return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
int sn_write_paths(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    int r = fetch_paths(c); if (r) return r;
    std::string err;
    if (!snf::write_paths(path, c->cnt.n_reads, c->h_poffset.data(), c->h_path_off.data(), c->h_pedges.data(), err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
// tmp.paths from arrays (host helper: the ranks' paths of a multi-GPU job, concatenated by the caller)
int sn_write_paths_arrays(const char* path, uint64_t n_reads, const int32_t* offset, const uint64_t* path_off, const int32_t* edges)
{
    if (!path || !offset || !path_off) return SN_ERR_ARG;
    std::string err;
    if (!snf::write_paths(path, n_reads, offset, path_off, edges, err)) { g_sn_create_error = err; return SN_ERR_IO; }
    return SN_OK;
}
int sn_write_edges_bv(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (c->stage < 3) return fail(c, SN_ERR_STATE, "run sn_build_edges first");
    { int r = sn_i_fetch_edges_host(c); if (r) return r; }
    std::string err; const snh::Edges& E = c->hedges;
    if (!snf::write_bv(path, E.packed.data(), E.off.data(), E.len.data(), E.n(), err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
// writePathsIndex (10X/PathsIndex.cc:23-143, called from 10X/DF.cc:588): a.paths.inv + a.countsb
int sn_build_paths_index(sn_ctx* c)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 5) return fail(c, SN_ERR_STATE, "sn_build_paths_index: run sn_path_reads first");
    CU(cudaSetDevice(c->device));
    const uint64_t n = c->cnt.n_reads, m = c->cnt.n_path_edges, nH = c->cnt.n_hbv_edges;
    if (m >= (1ull << 32)) return fail(c, SN_ERR_ARG, "more than 2^32-1 path entries in one context");
    DevBuf &ra = c->pool["pi_a"], &rb = c->pool["pi_b"], &tmp = c->pool["rs_tmp"], &cnt = c->pool["pi_cnt"], &off = c->pool["pi_off"], &ids = c->pool["pi_ids"], &cb = c->pool["pi_cb"];
    CU(ra.alloc(16 * m + 16)); CU(rb.alloc(16 * m + 16)); CU(tmp.alloc(radix_sort_tmp_bytes((uint32_t)m))); CU(cnt.alloc(4 * nH + 16)); CU(off.alloc(8 * (nH + 1)));
    CU(ids.alloc(8 * m + 16)); CU(cb.alloc(4 * nH + 16));
    t_begin(c, "paths_index");
    CU(cudaMemsetAsync(cnt.p, 0, 4 * nH + 16, c->st));
    k_pi_records<<<blocks_for(n, 256), 256, 0, c->st>>>(c->pedges.as<int32_t>(), c->path_off.as<uint64_t>(), n, ra.as<uint4>(), cnt.as<uint32_t>());
    KCHECK("k_pi_records");
    if (m) {
        cudaError_t e = radix_sort<RS_KEY96>(ra.as<uint4>(), rb.as<uint4>(), (uint32_t)m, tmp.p, c->num_sms, c->st);      // 12 passes: the result is back in ra
        if (e != cudaSuccess) return fail(c, SN_ERR_CUDA, cudaGetErrorString(e));
        c->launches += 13;
        k_pi_ids<<<blocks_for(m, 256), 256, 0, c->st>>>(ra.as<uint4>(), m, ids.as<unsigned long long>());
        KCHECK("k_pi_ids");
    }
    int r = scan_u32(c, cnt.as<uint32_t>(), nH, off.as<uint64_t>(), nullptr);
    if (r) return r;
    if (nH) {
        k_pi_countsb<<<blocks_for(nH, 256), 256, 0, c->st>>>(cnt.as<uint32_t>(), c->pool["hbv_inv"].as<int32_t>(), (uint32_t)nH, cb.as<int32_t>());
        KCHECK("k_pi_countsb");
    }
    t_end(c, "paths_index");
    resize_pinned(c, c->pi_off, nH + 1); resize_pinned(c, c->pi_ids, m); resize_pinned(c, c->pi_countsb, nH);
    CU(cudaMemcpyAsync(c->pi_off.data(), off.p, 8 * (nH + 1), cudaMemcpyDeviceToHost, c->st));
    if (m) CU(cudaMemcpyAsync(c->pi_ids.data(), ids.p, 8 * m, cudaMemcpyDeviceToHost, c->st));
    if (nH) CU(cudaMemcpyAsync(c->pi_countsb.data(), cb.p, 4 * nH, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    c->pi_ready = true;
    return SN_OK;
}
int sn_get_paths_index(sn_ctx* c, uint64_t* off, uint64_t* ids, int32_t* countsb)
{
    if (!c) return SN_ERR_ARG;
    if (!c->pi_ready) return fail(c, SN_ERR_STATE, "run sn_build_paths_index first");
    if (off) memcpy(off, c->pi_off.data(), 8 * c->pi_off.size());
    if (ids && !c->pi_ids.empty()) memcpy(ids, c->pi_ids.data(), 8 * c->pi_ids.size());
    if (countsb && !c->pi_countsb.empty()) memcpy(countsb, c->pi_countsb.data(), 4 * c->pi_countsb.size());
    return SN_OK;
}
int sn_write_paths_index(sn_ctx* c, const char* paths_inv, const char* countsb)
{
    if (!c) return SN_ERR_ARG;
    if (!c->pi_ready) return fail(c, SN_ERR_STATE, "run sn_build_paths_index first");
    std::string err;
    if (paths_inv && !snf::write_ulongvecs(paths_inv, c->pi_off.size() - 1, c->pi_ids.data(), c->pi_off.data(), err)) return fail(c, SN_ERR_IO, err);
    if (countsb && !snf::write_vec_vec_int(countsb, c->pi_countsb, err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
// a.to_left / a.to_right (HyperBasevector::ToLeft/ToRight, written by 10X/WriteFiles.cc:16-60): source and
// target vertex of every HBV edge
int sn_write_to_left_right(sn_ctx* c, const char* to_left, const char* to_right)
{
    if (!c) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    { int r = sn_i_fetch_hbv_host(c); if (r) return r; }
    std::string err;
    if (to_left && !snf::write_vec_int(to_left, c->hbv.to_left, err)) return fail(c, SN_ERR_IO, err);
    if (to_right && !snf::write_vec_int(to_right, c->hbv.to_right, err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
int sn_write_inv(sn_ctx* c, const char* path)
{
    if (!c || !path) return SN_ERR_ARG;
    if (c->stage < 4) return fail(c, SN_ERR_STATE, "run sn_build_hbv first");
    { int r = sn_i_fetch_hbv_host(c); if (r) return r; }
    std::string err;
    if (!snf::write_vec_int(path, c->hbv.inv, err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}
// WriteKmerSpectrum (BuildReadQGraph48.cc:199-216) + WriteHistToJson (10X/MakeHist.cc:68-92)
int sn_write_kmer_spectrum(sn_ctx* c, const char* json)
{
    if (!c || !json) return SN_ERR_ARG;
    if (c->stage < 2) return fail(c, SN_ERR_STATE, "run sn_count_kmers first");
    CU(cudaSetDevice(c->device));
    size_t n = c->cnt.n_kmers;
    std::vector<uint32_t> cc(n);
    if (n) CU(cudaMemcpy2D(cc.data(), 4, (const char*)c->dict.p + 12, sizeof(DictEntry), 4, n, cudaMemcpyDeviceToHost));
    std::vector<int64_t> spec;
    for (uint32_t v : cc) { uint32_t k = v & 0xFFFFFFu; if (spec.size() <= k) spec.resize(k + 1, 0); spec[k]++; }
    int64_t maxc = (int64_t)spec.size() - 1;
    // (through the checked writer: a failed write or close is SN_ERR_IO, and the file appears only when complete)
    std::string text = "{\n\t\"description\": \"kmer_count\",\n\t\"stage\": \"DF\",\n\t\"binsize\": 1,\n\t\"min\": 0,\n\t\"max\": " + std::to_string((long long)maxc) +
                       ",\n\t\"numbins\": " + std::to_string(spec.size()) + ",\n\t\"vals\": [";
    for (size_t i = 0; i < spec.size(); ++i) { text += std::to_string((long long)spec[i]); if (i + 1 != spec.size()) text += ","; }
    text += "]\n}\n";
    std::string err;
    if (!snf::write_text(json, text, err)) return fail(c, SN_ERR_IO, err);
    return SN_OK;
}

int sn_build_read_qgraph48(sn_ctx* c, const char* work_dir, const sn_params* params, int with_paths, int write_files)
{
    if (!c) return SN_ERR_ARG;
    int r;
    if ((r = sn_count_kmers(c, params))) return r;
    if ((r = sn_build_edges(c))) return r;
    if ((r = sn_build_hbv(c))) return r;
    if (with_paths && (r = sn_path_reads(c))) return r;
    if (write_files) {
        if (!work_dir) return fail(c, SN_ERR_ARG, "work_dir is NULL");
        std::string wd(work_dir);
        mkdir((wd + "/stats").c_str(), 0777);
        if ((r = sn_write_kmer_spectrum(c, (wd + "/stats/histogram_kmer_count.json").c_str()))) return r;
        if ((r = sn_write_hbv(c, (wd + "/a.hbv").c_str()))) return r;
        if (with_paths && (r = sn_write_paths(c, (wd + "/tmp.paths").c_str()))) return r;
    }
    return SN_OK;
}

// ---- host-only helpers -------------------------------------------------------------------
uint64_t sn_pqvec_encode(const uint8_t* quals, uint32_t n, uint8_t* out)
{
    std::vector<uint8_t> v; snf::pqvec_encode(quals, n, v);
    memcpy(out, v.data(), v.size());
    return v.size();
}
uint32_t sn_pqvec_decode(const uint8_t* pq, uint64_t pq_bytes, uint8_t* out, uint32_t cap) { return snf::pqvec_decode(pq, pq + pq_bytes, out, cap); }
void sn_free(void* p) { free(p); }

int sn_pack_reads(uint64_t n, const uint8_t* codes, const uint8_t* quals, const uint64_t* off, int threads,
                  uint8_t** bases, uint64_t** base_off, uint32_t** len, uint8_t** pq, uint64_t** pq_off)
{
    if (!codes || !quals || !off || !bases || !base_off || !len || !pq || !pq_off) return SN_ERR_ARG;
    if (threads < 1) threads = 1;
    uint64_t* bo = (uint64_t*)malloc(8 * (n + 1)); uint32_t* ln = (uint32_t*)malloc(4 * (n ? n : 1)); uint64_t* po = (uint64_t*)malloc(8 * (n + 1));
    bo[0] = 0;
    for (uint64_t r = 0; r < n; ++r) { ln[r] = (uint32_t)(off[r + 1] - off[r]); bo[r + 1] = bo[r] + (ln[r] + 3) / 4; }
    uint8_t* bs = (uint8_t*)calloc(bo[n] + 64, 1);
    std::vector<std::vector<uint8_t>> chunks(threads);
    std::vector<std::vector<uint64_t>> sizes(threads);
    auto work = [&](int t) {
        uint64_t a = n * t / threads, b = n * (t + 1) / threads;
        sizes[t].reserve(b - a);
        for (uint64_t r = a; r < b; ++r) {
            const uint8_t* s = codes + off[r]; uint8_t* d = bs + bo[r];
            for (uint32_t i = 0; i < ln[r]; ++i) d[i >> 2] |= (uint8_t)((s[i] & 3u) << (2 * (i & 3)));
            size_t before = chunks[t].size();
            snf::pqvec_encode(quals + off[r], ln[r], chunks[t]);
            sizes[t].push_back(chunks[t].size() - before);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
    uint64_t tot = 0; for (auto& v : chunks) tot += v.size();
    uint8_t* pqb = (uint8_t*)malloc(tot + 64);
    uint64_t at = 0, r = 0; po[0] = 0;
    for (int t = 0; t < threads; ++t) {
        memcpy(pqb + at, chunks[t].data(), chunks[t].size()); at += chunks[t].size();
        for (uint64_t s : sizes[t]) { po[r + 1] = po[r] + s; ++r; }
    }
    memset(pqb + tot, 0, 64);
    *bases = bs; *base_off = bo; *len = ln; *pq = pqb; *pq_off = po;
    return SN_OK;
}
int sn_write_read_files(const char* fastb, const char* qualp, const char* bci, uint64_t n,
                        const uint8_t* bases, const uint64_t* base_off, const uint32_t* len,
                        const uint8_t* pq, const uint64_t* pq_off, const int32_t* bc)
{
    std::string err;
    if (fastb) {
        snf::Fastb fb; fb.var.assign(bases, bases + base_off[n]); fb.off.assign(base_off, base_off + n + 1); fb.len.assign(len, len + n);
        if (!snf::write_fastb(fastb, fb, err)) { g_sn_create_error = err; return SN_ERR_IO; }
    }
    if (qualp) {
        snf::Qualp qp; qp.var.assign(pq, pq + pq_off[n]); qp.off.assign(pq_off, pq_off + n + 1);
        if (!snf::write_qualp(qualp, qp, err)) { g_sn_create_error = err; return SN_ERR_IO; }
    }
    if (bci) {
        // inverse of the expansion in 10X/DF.cc:464-469: bci[b] = first read of barcode ordinal b
        std::vector<int64_t> bi; bi.push_back(0);
        int32_t cur = 0;
        for (uint64_t r = 0; r < n; ++r) {
            int32_t b = bc ? bc[r] : 0;
            if (b < cur) { g_sn_create_error = "barcode ordinals are not sorted"; return SN_ERR_DATA; }
            while (cur < b) { bi.push_back((int64_t)r); ++cur; }
        }
        bi.push_back((int64_t)n);
        if (!snf::write_bci(bci, bi, err)) { g_sn_create_error = err; return SN_ERR_IO; }
    }
    return SN_OK;
}

}  // extern "C"
