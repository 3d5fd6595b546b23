// sn_scale -- scale runs of the hot path on device-generated reads (sn_generate_reads), plain C++ over the C ABI:
// no Python, no host copy of the reads.  One host thread and one context per GPU; with NGPU > 1 the sharded path
// (sn_mg_build_graph: NCCL inside the library).
//
//   sn_scale [NGPU=1] [MULT=1] [G=<genome bases>] [PAIRS=<pairs of the whole job>] [NBC=<barcodes>] [SEED=20261017]
//            [PATHS=0] [PASSES=<forced count passes>] [HBV=<file>] [OUT=<json file>]
//
// MULT scales BASELINE config 2 (63 Mbp diploid genome, 4 M pairs = 1.2 Gbp, 1 M barcodes): G = 63 Mbp x MULT etc.; MULT = 150 is
// config 3 (3.2 Gbp-class genome x 56: 180 Gbp), of which every one of 8 ranks generates and counts 22.5 Gbp.  Rank r takes
// the pairs [PAIRS r / N, PAIRS (r + 1) / N).  Prints one JSON line: sizes, wall seconds of the build per rank (max = the job),
// counts.  HBV=<file>: rank 0 writes a.hbv there (compare a 1-GPU and an N-GPU run of the same job with cmp).
#include "../../include/supernova_b200.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

int main(int argc, char** argv)
{
    std::map<std::string, std::string> a = {{"NGPU", "1"}, {"MULT", "1"}, {"SEED", "20261017"}, {"PATHS", "0"}};
    for (int i = 1; i < argc; ++i) {
        const char* eq = strchr(argv[i], '=');
        if (!eq) { fprintf(stderr, "usage: sn_scale [NGPU=1] [MULT=1] [G=] [PAIRS=] [NBC=] [SEED=] [PATHS=0] [PASSES=] [HBV=file] [OUT=file]\n"); return 2; }
        a[std::string(argv[i], eq - argv[i])] = eq + 1;
    }
    const int ngpu = atoi(a["NGPU"].c_str());
    const double mult = atof(a["MULT"].c_str());
    sn_synth spec;
    spec.genome_bases = a.count("G") ? strtoull(a["G"].c_str(), nullptr, 10) : (uint64_t)(63000000.0 * mult);
    spec.total_pairs = a.count("PAIRS") ? strtoull(a["PAIRS"].c_str(), nullptr, 10) : (uint64_t)(4000000.0 * mult);
    spec.n_barcodes = a.count("NBC") ? (uint32_t)strtoul(a["NBC"].c_str(), nullptr, 10) : (uint32_t)std::min(4000000.0, 1000000.0 * mult);
    spec.seed = strtoull(a["SEED"].c_str(), nullptr, 10);
    const bool paths = a["PATHS"] == "1";
    if (ngpu < 1 || sn_device_count() < ngpu) { fprintf(stderr, "sn_scale: NGPU=%d but %d CUDA device(s) visible (there is no CPU fallback)\n", ngpu, sn_device_count()); return 1; }
    if (a.count("PASSES")) { setenv("SN_COUNT_PASSES", a["PASSES"].c_str(), 1); setenv("SN_MG_PASSES", a["PASSES"].c_str(), 1); }
    uint32_t T[150];                                    // substitution probability per position in units of 2^-24 (synth.py: cb_error_thresholds)
    for (uint64_t j = 0; j < 150; ++j) T[j] = (uint32_t)(((1ull << 24) * (3375000ull * 100 + 2000ull * j * j * j)) / (1000ull * 100 * 3375000ull));
    unsigned char uid[128];
    if (ngpu > 1 && sn_nccl_unique_id(uid)) { fprintf(stderr, "sn_scale: %s\n", sn_last_error(nullptr)); return 1; }
    sn_params prm; prm.min_qual = 7; prm.min_freq = 3; prm.min_bc = 2; prm.ign_bc_below = 0;
    std::vector<std::string> errs(ngpu);
    std::vector<sn_counts> cnt(ngpu);
    std::vector<double> gen_s(ngpu), build_s(ngpu), path_s(ngpu);
    std::vector<std::map<std::string, double>> stage(ngpu);
    auto work = [&](int r) {
        sn_ctx* ctx = nullptr;
        auto die = [&](const char* what) { errs[r] = std::string(what) + ": " + sn_last_error(ctx); if (ctx) sn_ctx_destroy(ctx); };
        if (sn_ctx_create(&ctx, r)) { errs[r] = sn_last_error(nullptr); return; }
        if (ngpu > 1 && sn_comm_init_nccl(ctx, r, ngpu, uid)) return die("NCCL");
        const uint64_t p0 = spec.total_pairs * (uint64_t)r / (uint64_t)ngpu, p1 = spec.total_pairs * (uint64_t)(r + 1) / (uint64_t)ngpu;
        auto t0 = std::chrono::steady_clock::now();
        if (sn_generate_reads(ctx, &spec, p0, p1 - p0, T)) return die("sn_generate_reads");
        auto t1 = std::chrono::steady_clock::now();
        gen_s[r] = std::chrono::duration<double>(t1 - t0).count();
        if (ngpu > 1) { if (sn_mg_build_graph(ctx, &prm, 0)) return die("sn_mg_build_graph"); }
        else if (sn_count_kmers(ctx, &prm) || sn_build_edges(ctx) || sn_build_hbv(ctx)) return die("count / edges / hbv");
        auto t2 = std::chrono::steady_clock::now();
        build_s[r] = std::chrono::duration<double>(t2 - t1).count();
        for (const char* s : {"exchange", "ghosts", "goodlen", "msp_hist", "msp_scatter", "bucket_count", "make_dict", "prune", "edges", "hbv_dev", "hbv_host", "hbv_csr"}) stage[r][s] = sn_stage_ms(ctx, s);
        if (paths) {
            if (ngpu > 1) { if (sn_mg_build_graph(ctx, &prm, 1)) return die("sn_mg_build_graph with paths"); }
            else if (sn_path_reads(ctx)) return die("sn_path_reads");
            path_s[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count();
        }
        sn_get_counts(ctx, &cnt[r]);
        if (r == 0 && a.count("HBV") && sn_write_hbv(ctx, a["HBV"].c_str())) return die("a.hbv");
        sn_ctx_destroy(ctx);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < ngpu; ++r) th.emplace_back(work, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < ngpu; ++r) if (!errs[r].empty()) { fprintf(stderr, "sn_scale: rank %d: %s\n", r, errs[r].c_str()); return 1; }
    double gmax = 0, bmax = 0; uint64_t n_kmers = 0, n_reads = 0, n_occ = 0;
    for (int r = 0; r < ngpu; ++r) { gmax = std::max(gmax, gen_s[r]); bmax = std::max(bmax, build_s[r]); n_kmers += cnt[r].n_kmers; n_reads += cnt[r].n_reads; n_occ += cnt[r].n_kmer_occurrences; }
    const double gbp = 300.0 * (double)spec.total_pairs / 1e9;
    std::string js = "{\"n_gpus\": " + std::to_string(ngpu) + ", \"genome_bases\": " + std::to_string(spec.genome_bases) + ", \"pairs\": " + std::to_string(spec.total_pairs) +
        ", \"gbp\": " + std::to_string(gbp) + ", \"n_barcodes\": " + std::to_string(spec.n_barcodes) + ", \"seed\": " + std::to_string(spec.seed) +
        ", \"reads\": " + std::to_string(n_reads) + ", \"kmer_occurrences\": " + std::to_string(n_occ) + ", \"kmers\": " + std::to_string(n_kmers) +
        ", \"unipaths\": " + std::to_string(cnt[0].n_edges) + ", \"edge_bases\": " + std::to_string(cnt[0].n_edge_bases) + ", \"hbv_vertices\": " + std::to_string(cnt[0].n_hbv_vertices) +
        ", \"hbv_edges\": " + std::to_string(cnt[0].n_hbv_edges) + ", \"generate_s\": " + std::to_string(gmax) + ", \"build_s\": " + std::to_string(bmax) +
        ", \"gbp_per_s\": " + std::to_string(gbp / bmax) + ", \"build_s_per_rank\": [";
    for (int r = 0; r < ngpu; ++r) js += (r ? ", " : "") + std::to_string(build_s[r]);
    js += "], \"stage_ms_rank0\": {";
    { bool first = true; for (auto& kv : stage[0]) if (kv.second >= 0) { js += std::string(first ? "" : ", ") + "\"" + kv.first + "\": " + std::to_string(kv.second); first = false; } }
    js += "}";
    if (paths) { double pm = 0; uint64_t pe = 0; for (int r = 0; r < ngpu; ++r) { pm = std::max(pm, path_s[r]); pe += cnt[r].n_path_edges; } js += ", \"paths_s\": " + std::to_string(pm) + ", \"path_entries\": " + std::to_string(pe); }
    js += ", \"kmers_on_edges\": " + std::to_string(cnt[0].n_edge_bases - 47 * cnt[0].n_edges) + ", \"ok\": " + (cnt[0].n_edge_bases - 47 * cnt[0].n_edges == n_kmers ? "true" : "false") + "}";
    printf("%s\n", js.c_str());
    if (a.count("OUT")) { FILE* f = fopen(a["OUT"].c_str(), "w"); if (f) { fprintf(f, "%s\n", js.c_str()); fclose(f); } }
    return 0;
}
