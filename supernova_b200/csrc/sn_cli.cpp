// sn_build_graph -- the C++ host side of the hot path as one command, over the C ABI only
// (include/supernova_b200.h): what StageBuildGraph + the start of DF do around buildReadQGraph48
// (10X/runstages/RunStages.cc:404-413, 10X/DF.cc:573-600), for a maintainer who wants to try the
// library without touching DF, and for the parity tests (tests/test_cli.py).
//
//   sn_build_graph HEAD=<dir>/reads OUT=<dir> [FASTH=<file[.gz]>] [MIN_QUAL=7] [MIN_FREQ=3] [MIN_BC=2]
//                  [PATHS=True] [INDEX=False] [DEVICE=0] [NGPU=1]
//
// HEAD   : reads.fastb / reads.qualp / reads.bci as ParseBarcodedFastqs writes them; with FASTH= they are
//          produced first, on the device, from the barcoded pseudo-FASTQ (and written to HEAD.*).
// OUT    : a.hbv, tmp.paths, stats/histogram_kmer_count.json as the reference leaves them; with INDEX=True
//          also what DF leaves in a.48/ before it goes on (10X/WriteFiles.cc:16-60, 10X/DF.cc:573-600): a.k, a.inv, a.to_left,
//          a.to_right, a.hbx, a.fastb, a.kmers and, from the paths, a.paths.inv, a.countsb, a.pathsX, a.dup (MarkDups; its
//          three percentages go to stdout as the reference prints them).
// NGPU   : > 1 = the sharded multi-GPU path (sn_mg_build_graph): one host thread and one context per GPU of this
//          box, reads split evenly (whole pairs), every collective issued by the library on NCCL -- no Python, no
//          launcher.  Rank 0 writes a.hbv; tmp.paths is the ranks' ReadPaths in read order.  (INDEX=True: one GPU.)
// Exit status 0, or 1 with the library's message on stderr (the reference: FatalErr -> exit(1)).
#include "../../include/supernova_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

static bool truth(const std::string& s) { return s == "True" || s == "true" || s == "1"; }

static int run_multi(std::map<std::string, std::string>& a, int ngpu)
{
    const std::string head = a["HEAD"], out = a["OUT"];
    const bool paths = truth(a["PATHS"]);
    sn_params prm;
    prm.min_qual = (uint32_t)atoi(a["MIN_QUAL"].c_str()); prm.min_freq = (uint32_t)atoi(a["MIN_FREQ"].c_str());
    prm.min_bc = (uint32_t)atoi(a["MIN_BC"].c_str()); prm.ign_bc_below = 0;
    if (sn_device_count() < ngpu) { fprintf(stderr, "sn_build_graph: NGPU=%d but %d CUDA device(s) visible\n", ngpu, sn_device_count()); return 1; }
    unsigned char uid[128];
    if (sn_nccl_unique_id(uid)) { fprintf(stderr, "sn_build_graph: %s\n", sn_last_error(nullptr)); return 1; }
    // the number of reads: from the fastb header (u32 n at offset 0 of a feudal file)
    uint64_t n_reads = 0;
    { FILE* f = fopen((head + ".fastb").c_str(), "rb"); uint32_t n = 0; if (!f || fread(&n, 4, 1, f) != 1) { fprintf(stderr, "sn_build_graph: cannot read %s.fastb\n", head.c_str()); if (f) fclose(f); return 1; } fclose(f); n_reads = n; }
    std::vector<uint64_t> cut(ngpu + 1);
    for (int r = 0; r <= ngpu; ++r) cut[r] = r == ngpu ? n_reads : (n_reads * (uint64_t)r / (uint64_t)ngpu) & ~1ull;
    std::vector<std::vector<int32_t>> p_off(ngpu), p_edges(ngpu); std::vector<std::vector<uint64_t>> p_poff(ngpu);
    std::vector<std::string> errs(ngpu);
    std::vector<sn_counts> cnt(ngpu);
    auto work = [&](int r) {
        sn_ctx* ctx = nullptr;
        auto die = [&](const char* what) { errs[r] = std::string(what) + ": " + sn_last_error(ctx); if (ctx) sn_ctx_destroy(ctx); };
        if (sn_ctx_create(&ctx, r)) { errs[r] = sn_last_error(nullptr); return; }
        if (sn_comm_init_nccl(ctx, r, ngpu, uid)) return die("NCCL");
        if (sn_load_read_files_range(ctx, (head + ".fastb").c_str(), (head + ".qualp").c_str(), (head + ".bci").c_str(), cut[r], cut[r + 1] - cut[r])) return die("loading the read files");
        if (sn_mg_build_graph(ctx, &prm, paths ? 1 : 0)) return die("sn_mg_build_graph");
        sn_get_counts(ctx, &cnt[r]);
        if (paths) {
            p_off[r].resize(cnt[r].n_reads); p_poff[r].resize(cnt[r].n_reads + 1); p_edges[r].resize(cnt[r].n_path_edges + 1);
            if (sn_get_paths(ctx, p_off[r].data(), p_poff[r].data(), p_edges[r].data())) return die("sn_get_paths");
        }
        if (r == 0) {
            mkdir((out + "/stats").c_str(), 0777);
            if (sn_write_hbv(ctx, (out + "/a.hbv").c_str())) return die("a.hbv");
            if (!sn_mg_dict_is_sharded(ctx) && sn_write_kmer_spectrum(ctx, (out + "/stats/histogram_kmer_count.json").c_str())) return die("k-mer spectrum");
        }
        sn_ctx_destroy(ctx);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < ngpu; ++r) th.emplace_back(work, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < ngpu; ++r) if (!errs[r].empty()) { fprintf(stderr, "sn_build_graph: rank %d: %s\n", r, errs[r].c_str()); return 1; }
    uint64_t n_path_edges = 0, n_kmers = 0;
    if (paths) {
        std::vector<int32_t> off, edges; std::vector<uint64_t> poff(1, 0);
        for (int r = 0; r < ngpu; ++r) {
            off.insert(off.end(), p_off[r].begin(), p_off[r].end());
            for (uint64_t i = 0; i < cnt[r].n_reads; ++i) poff.push_back(poff.back() + (p_poff[r][i + 1] - p_poff[r][i]));
            edges.insert(edges.end(), p_edges[r].begin(), p_edges[r].begin() + cnt[r].n_path_edges);
            n_path_edges += cnt[r].n_path_edges;
        }
        edges.push_back(0);
        if (sn_write_paths_arrays((out + "/tmp.paths").c_str(), n_reads, off.data(), poff.data(), edges.data())) { fprintf(stderr, "sn_build_graph: tmp.paths: %s\n", sn_last_error(nullptr)); return 1; }
        n_kmers = cnt[0].n_kmers;
    } else for (int r = 0; r < ngpu; ++r) n_kmers += cnt[r].n_kmers;
    printf("%d GPUs: reads %llu k-mers %llu unipaths %llu hbv %llu vertices / %llu edges path entries %llu\n", ngpu, (unsigned long long)n_reads, (unsigned long long)n_kmers,
           (unsigned long long)cnt[0].n_edges, (unsigned long long)cnt[0].n_hbv_vertices, (unsigned long long)cnt[0].n_hbv_edges, (unsigned long long)n_path_edges);
    return 0;
}

int main(int argc, char** argv)
{
    std::map<std::string, std::string> a = {{"MIN_QUAL", "7"}, {"MIN_FREQ", "3"}, {"MIN_BC", "2"}, {"PATHS", "True"}, {"INDEX", "False"}, {"DEVICE", "0"}, {"NGPU", "1"}};
    for (int i = 1; i < argc; ++i) {
        const char* eq = strchr(argv[i], '=');
        if (!eq) { fprintf(stderr, "sn_build_graph: argument without '=': %s\n", argv[i]); return 2; }
        a[std::string(argv[i], eq - argv[i])] = eq + 1;
    }
    if (!a.count("HEAD") || !a.count("OUT")) {
        fprintf(stderr, "usage: sn_build_graph HEAD=<dir>/reads OUT=<dir> [FASTH=<file[.gz]>] [MIN_QUAL=7] [MIN_FREQ=3] [MIN_BC=2] [PATHS=True] [INDEX=False] [DEVICE=0] [NGPU=1]\n");
        return 2;
    }
    const std::string head = a["HEAD"], out = a["OUT"];
    const int ngpu = atoi(a["NGPU"].c_str());
    sn_ctx* ctx = nullptr;
    if (sn_ctx_create(&ctx, atoi(a["DEVICE"].c_str()))) { fprintf(stderr, "sn_build_graph: %s\n", sn_last_error(nullptr)); return 1; }
    auto die = [&](const char* what) { fprintf(stderr, "sn_build_graph: %s: %s\n", what, sn_last_error(ctx)); sn_ctx_destroy(ctx); return 1; };
    if (a.count("FASTH")) {
        if (sn_load_fasth_file(ctx, a["FASTH"].c_str())) return die("ingest");
        if (sn_save_read_files(ctx, (head + ".fastb").c_str(), (head + ".qualp").c_str(), (head + ".bci").c_str())) return die("writing the read files");
    }
    if (ngpu > 1) {
        if (truth(a["INDEX"])) { fprintf(stderr, "sn_build_graph: INDEX=True needs NGPU=1\n"); sn_ctx_destroy(ctx); return 2; }
        sn_ctx_destroy(ctx);
        return run_multi(a, ngpu);
    }
    if (!a.count("FASTH") && sn_load_read_files(ctx, (head + ".fastb").c_str(), (head + ".qualp").c_str(), (head + ".bci").c_str())) return die("loading the read files");
    sn_params prm;
    prm.min_qual = (uint32_t)atoi(a["MIN_QUAL"].c_str()); prm.min_freq = (uint32_t)atoi(a["MIN_FREQ"].c_str());
    prm.min_bc = (uint32_t)atoi(a["MIN_BC"].c_str()); prm.ign_bc_below = 0;
    const bool paths = truth(a["PATHS"]), index = truth(a["INDEX"]);
    if (sn_build_read_qgraph48(ctx, out.c_str(), &prm, paths ? 1 : 0, /*write_files=*/1)) return die("buildReadQGraph48");
    if (index) {
        if (sn_write_inv(ctx, (out + "/a.inv").c_str()) || sn_write_to_left_right(ctx, (out + "/a.to_left").c_str(), (out + "/a.to_right").c_str())) return die("a.inv / a.to_left / a.to_right");
        if (sn_write_k(ctx, (out + "/a.k").c_str()) || sn_write_hbx(ctx, (out + "/a.hbx").c_str()) || sn_write_edges_fastb(ctx, (out + "/a.fastb").c_str())
            || sn_write_kmers(ctx, (out + "/a.kmers").c_str())) return die("a.k / a.hbx / a.fastb / a.kmers");
        if (paths && (sn_build_paths_index(ctx) || sn_write_paths_index(ctx, (out + "/a.paths.inv").c_str(), (out + "/a.countsb").c_str()))) return die("writePathsIndex");
        if (paths && (sn_build_pathsx(ctx) || sn_write_pathsx(ctx, (out + "/a.pathsX").c_str()))) return die("a.pathsX");
        if (paths) {
            sn_dup_stats ds;
            if (sn_mark_dups(ctx, &ds) || sn_write_dup(ctx, (out + "/a.dup").c_str())) return die("MarkDups");
            const double np = ds.n_pairs ? (double)ds.n_pairs : 1.0;                     // (SecretOps.cc:757-770)
            printf("%.2f%% of pairs appear to be duplicates\n", 100.0 * ds.n_dup_pairs / np);
            printf("%.2f%% of duplicates involve multiple barcodes\n", (100.0 * ds.n_interdups) / (double)ds.n_dups);
            printf("%.2f%% of pairs appear to be artifactual duplicates\n", 100.0 * ds.n_art_pairs / np);
        }
    }
    sn_counts c;
    sn_get_counts(ctx, &c);
    printf("reads %llu bases %llu k-mers %llu (of %llu distinct, %llu occurrences) unipaths %llu hbv %llu vertices / %llu edges path entries %llu\n",
           (unsigned long long)c.n_reads, (unsigned long long)c.n_bases, (unsigned long long)c.n_kmers, (unsigned long long)c.n_kmers_distinct,
           (unsigned long long)c.n_kmer_occurrences, (unsigned long long)c.n_edges, (unsigned long long)c.n_hbv_vertices, (unsigned long long)c.n_hbv_edges,
           (unsigned long long)c.n_path_edges);
    sn_ctx_destroy(ctx);
    return 0;
}
