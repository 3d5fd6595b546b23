// sn_build_graph -- the C++ host side of the hot path as one command, over the C ABI only
// (include/supernova_b200.h): what StageBuildGraph + the start of DF do around buildReadQGraph48
// (10X/runstages/RunStages.cc:404-413, 10X/DF.cc:579-590), for a maintainer who wants to try the
// library without touching DF, and for the parity tests (tests/test_gpu_cli.py).
//
//   sn_build_graph HEAD=<dir>/reads OUT=<dir> [FASTH=<file[.gz]>] [MIN_QUAL=7] [MIN_FREQ=3] [MIN_BC=2]
//                  [PATHS=True] [INDEX=False] [DEVICE=0]
//
// HEAD   : reads.fastb / reads.qualp / reads.bci as ParseBarcodedFastqs writes them; with FASTH= they are
//          produced first, on the device, from the barcoded pseudo-FASTQ (and written to HEAD.*).
// OUT    : a.hbv, tmp.paths, stats/histogram_kmer_count.json as the reference leaves them; with INDEX=True
//          also a.inv, a.to_left, a.to_right, a.paths.inv, a.countsb.
// Exit status 0, or 1 with the library's message on stderr (the reference: FatalErr -> exit(1)).
#include "../../include/supernova_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>

static bool truth(const std::string& s) { return s == "True" || s == "true" || s == "1"; }

int main(int argc, char** argv)
{
    std::map<std::string, std::string> a = {{"MIN_QUAL", "7"}, {"MIN_FREQ", "3"}, {"MIN_BC", "2"}, {"PATHS", "True"}, {"INDEX", "False"}, {"DEVICE", "0"}};
    for (int i = 1; i < argc; ++i) {
        const char* eq = strchr(argv[i], '=');
        if (!eq) { fprintf(stderr, "sn_build_graph: argument without '=': %s\n", argv[i]); return 2; }
        a[std::string(argv[i], eq - argv[i])] = eq + 1;
    }
    if (!a.count("HEAD") || !a.count("OUT")) {
        fprintf(stderr, "usage: sn_build_graph HEAD=<dir>/reads OUT=<dir> [FASTH=<file[.gz]>] [MIN_QUAL=7] [MIN_FREQ=3] [MIN_BC=2] [PATHS=True] [INDEX=False] [DEVICE=0]\n");
        return 2;
    }
    const std::string head = a["HEAD"], out = a["OUT"];
    sn_ctx* ctx = nullptr;
    if (sn_ctx_create(&ctx, atoi(a["DEVICE"].c_str()))) { fprintf(stderr, "sn_build_graph: %s\n", sn_last_error(nullptr)); return 1; }
    auto die = [&](const char* what) { fprintf(stderr, "sn_build_graph: %s: %s\n", what, sn_last_error(ctx)); sn_ctx_destroy(ctx); return 1; };
    if (a.count("FASTH")) {
        if (sn_load_fasth_file(ctx, a["FASTH"].c_str())) return die("ingest");
        if (sn_save_read_files(ctx, (head + ".fastb").c_str(), (head + ".qualp").c_str(), (head + ".bci").c_str())) return die("writing the read files");
    } else if (sn_load_read_files(ctx, (head + ".fastb").c_str(), (head + ".qualp").c_str(), (head + ".bci").c_str())) return die("loading the read files");
    sn_params prm;
    prm.min_qual = (uint32_t)atoi(a["MIN_QUAL"].c_str()); prm.min_freq = (uint32_t)atoi(a["MIN_FREQ"].c_str());
    prm.min_bc = (uint32_t)atoi(a["MIN_BC"].c_str()); prm.ign_bc_below = 0;
    const bool paths = truth(a["PATHS"]), index = truth(a["INDEX"]);
    if (sn_build_read_qgraph48(ctx, out.c_str(), &prm, paths ? 1 : 0, /*write_files=*/1)) return die("buildReadQGraph48");
    if (index) {
        if (sn_write_inv(ctx, (out + "/a.inv").c_str()) || sn_write_to_left_right(ctx, (out + "/a.to_left").c_str(), (out + "/a.to_right").c_str())) return die("a.inv / a.to_left / a.to_right");
        if (paths && (sn_build_paths_index(ctx) || sn_write_paths_index(ctx, (out + "/a.paths.inv").c_str(), (out + "/a.countsb").c_str()))) return die("writePathsIndex");
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
