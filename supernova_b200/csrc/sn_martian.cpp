// sn_martian -- the Martian exec-stage adapter of the three graph stages the pipeline runs through `tada`
// (SURVEY §8(b) boundary B1; lib/tada/mro/_asm_stages.mro:20-47: MSP, SHARD_ASM, MAIN_ASM_SN), over the C ABI only.
//
//   sn_martian martian <msp|shard-asm|main-asm-sn> <split|main|join> <metadata_path> <files_path> <run_file>
//
// is the command line Martian gives `tada martian <stage>` (lib/tada/src/main.rs:207-218, argument order
// external/martian/src/lib.rs:142-156); installed under the name `tada` (or with the three `src exec` lines of
// _asm_stages.mro pointing at it) the pipeline above it runs unchanged.  The protocol is the adapter's
// (external/martian/src/lib.rs): _args / _outs / _chunk_defs / _chunk_outs / _jobinfo are read from <metadata_path>;
// _stage_defs (split) or _outs (main, join), _log, _jobinfo (cwd, pid, wall clock, rusage) and _complete are written
// there, _errors on failure (exit status 1); every file written is announced by a journal file
// <run_file>.[<split|join>_]<name> (:181-200) and a heartbeat journal is refreshed every 60 s (:537-543).
//
// Only `asm_graph` (MAIN_ASM_SN's out, a vec<basevector> file) crosses into the C++ side (buildGraphFromMSP); the
// .msp / .sedge_asm / .sedge_bcs files between the stages are private to `tada`.  Here they are small
// text files that hand the inputs on: the device does the whole job -- parse, MSP, count, unipaths -- in one pass in
// MAIN_ASM_SN's main, which is where the GPU is needed; the stages before it cost nothing.
//   MSP        split: the fastq files in chunks of 8 (cmd_msp.rs:251-286); main: a .msp file naming its chunk's files and
//              trim_min_qual; join: the list (cmd_msp.rs:308-320)
//   SHARD_ASM  split: one chunk; main: a .sedge_asm file = the .msp files' content + min_kmer_obs, an empty .sedge_bcs
//              (the reference skips the barcode lists too, cmd_shard_asm.rs:63); join: renamed to chunk<i>.* in the
//              stage directory as there (:138-165)
//   MAIN_ASM_SN split: one chunk (cmd_main_asm.rs:184-194); main: the reads of all files through
//              sn_load_fasth_files -> sn_count_kmers (the tada rule: SN_SEM_TADA, MIN_BC 2 = "num_bcs > 1") ->
//              sn_build_edges -> asm_graph; join: the chunk's file renamed to the stage's (:205-212)
#include "../../include/supernova_b200.h"
#include <sys/resource.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

// ---- a small JSON value (objects keep their keys sorted, like the adapter's BTreeMap) ---------------------------------
struct Json;
using JsonP = std::shared_ptr<Json>;
struct Json {
    enum T { Null, Bool, Int, Float, Str, Arr, Obj } t = Null;
    bool b = false; long long i = 0; double f = 0; std::string s;
    std::vector<JsonP> a; std::map<std::string, JsonP> o;
    static JsonP mk(T t) { auto p = std::make_shared<Json>(); p->t = t; return p; }
    static JsonP str(const std::string& v) { auto p = mk(Str); p->s = v; return p; }
    static JsonP num(long long v) { auto p = mk(Int); p->i = v; return p; }
    static JsonP flt(double v) { auto p = mk(Float); p->f = v; return p; }
    const JsonP& at(const std::string& k) const
    {
        auto it = o.find(k);
        if (t != Obj || it == o.end()) throw std::runtime_error("missing key \"" + k + "\"");
        return it->second;
    }
    bool has(const std::string& k) const { return t == Obj && o.count(k); }
    const std::string& as_str() const { if (t != Str) throw std::runtime_error("expected a string"); return s; }
    long long as_int() const { if (t == Int) return i; if (t == Float) return (long long)f; throw std::runtime_error("expected a number"); }
};
struct JsonReader {
    const std::string& d; size_t p = 0;
    explicit JsonReader(const std::string& s) : d(s) {}
    void ws() { while (p < d.size() && (d[p] == ' ' || d[p] == '\n' || d[p] == '\t' || d[p] == '\r')) ++p; }
    [[noreturn]] void bad(const char* what) { throw std::runtime_error(std::string("bad json: ") + what + " at byte " + std::to_string(p)); }
    std::string string_()
    {
        std::string out; ++p;
        while (p < d.size() && d[p] != '"') {
            if (d[p] == '\\') {
                if (++p >= d.size()) bad("escape");
                switch (d[p]) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': {
                        if (p + 4 >= d.size()) bad("\\u");
                        unsigned c = (unsigned)strtoul(d.substr(p + 1, 4).c_str(), nullptr, 16); p += 4;
                        if (c < 0x80) out += (char)c;
                        else if (c < 0x800) { out += (char)(0xC0 | (c >> 6)); out += (char)(0x80 | (c & 0x3F)); }
                        else { out += (char)(0xE0 | (c >> 12)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); }
                        break; }
                    default: out += d[p];
                }
                ++p;
            } else out += d[p++];
        }
        if (p >= d.size()) bad("unterminated string");
        ++p;
        return out;
    }
    JsonP value()
    {
        ws();
        if (p >= d.size()) bad("end of input");
        const char c = d[p];
        if (c == '{') {
            auto v = Json::mk(Json::Obj); ++p; ws();
            if (p < d.size() && d[p] == '}') { ++p; return v; }
            for (;;) {
                ws(); if (p >= d.size() || d[p] != '"') bad("key");
                std::string k = string_(); ws();
                if (p >= d.size() || d[p] != ':') bad("':'");
                ++p; v->o[k] = value(); ws();
                if (p < d.size() && d[p] == ',') { ++p; continue; }
                if (p < d.size() && d[p] == '}') { ++p; return v; }
                bad("',' or '}'");
            }
        }
        if (c == '[') {
            auto v = Json::mk(Json::Arr); ++p; ws();
            if (p < d.size() && d[p] == ']') { ++p; return v; }
            for (;;) {
                v->a.push_back(value()); ws();
                if (p < d.size() && d[p] == ',') { ++p; continue; }
                if (p < d.size() && d[p] == ']') { ++p; return v; }
                bad("',' or ']'");
            }
        }
        if (c == '"') return Json::str(string_());
        if (!d.compare(p, 4, "null")) { p += 4; return Json::mk(Json::Null); }
        if (!d.compare(p, 4, "true")) { p += 4; auto v = Json::mk(Json::Bool); v->b = true; return v; }
        if (!d.compare(p, 5, "false")) { p += 5; return Json::mk(Json::Bool); }
        size_t q = p; bool isf = false;
        while (q < d.size() && (isdigit((unsigned char)d[q]) || d[q] == '-' || d[q] == '+' || d[q] == '.' || d[q] == 'e' || d[q] == 'E')) { isf = isf || d[q] == '.' || d[q] == 'e' || d[q] == 'E'; ++q; }
        if (q == p) bad("value");
        const std::string n = d.substr(p, q - p); p = q;
        return isf ? Json::flt(strtod(n.c_str(), nullptr)) : Json::num(strtoll(n.c_str(), nullptr, 10));
    }
};
static void json_write(const Json& v, std::string& out, int ind)
{
    auto pad = [&](int n) { out.append((size_t)n * 2, ' '); };
    switch (v.t) {
        case Json::Null: out += "null"; break;
        case Json::Bool: out += v.b ? "true" : "false"; break;
        case Json::Int: out += std::to_string(v.i); break;
        case Json::Float: { char b[64]; snprintf(b, sizeof b, "%.17g", v.f); out += b; if (!strpbrk(b, ".eEn")) out += ".0"; break; }
        case Json::Str:
            out += '"';
            for (unsigned char c : v.s) {
                if (c == '"' || c == '\\') { out += '\\'; out += (char)c; }
                else if (c == '\n') out += "\\n"; else if (c == '\t') out += "\\t"; else if (c == '\r') out += "\\r";
                else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
                else out += (char)c;
            }
            out += '"'; break;
        case Json::Arr:
            if (v.a.empty()) { out += "[]"; break; }
            out += "[\n";
            for (size_t k = 0; k < v.a.size(); ++k) { pad(ind + 1); json_write(*v.a[k], out, ind + 1); out += k + 1 < v.a.size() ? ",\n" : "\n"; }
            pad(ind); out += ']'; break;
        case Json::Obj: {
            if (v.o.empty()) { out += "{}"; break; }
            out += "{\n";
            size_t k = 0;
            for (auto& kv : v.o) { pad(ind + 1); json_write(*Json::str(kv.first), out, 0); out += ": "; json_write(*kv.second, out, ind + 1); out += ++k < v.o.size() ? ",\n" : "\n"; }
            pad(ind); out += '}'; break; }
    }
}

// ---- the chunk's metadata directory (external/martian/src/lib.rs:120-360) ---------------------------------------------
static std::string timestamp()
{
    time_t t = time(nullptr); struct tm tmv; localtime_r(&t, &tmv);
    char b[32]; strftime(b, sizeof b, "%Y-%m-%d %H:%M:%S", &tmv);
    return b;
}
struct Metadata {
    std::string stage_name, stage_type, metadata_path, files_path, run_file;
    std::set<std::string> cache; std::mutex m;
    time_t start = time(nullptr); std::string start_stamp = timestamp();
    std::string path(const std::string& name) const { return metadata_path + "/_" + name; }
    void journal(const std::string& name, bool force = false)                                   // :181-200
    {
        std::lock_guard<std::mutex> lk(m);
        const std::string jn = stage_type != "main" ? stage_type + "_" + name : name;
        if (cache.count(jn) && !force) return;
        const std::string rf = run_file + "." + jn, tmp = rf + ".tmp";
        { std::ofstream f(tmp); f << timestamp(); }
        rename(tmp.c_str(), rf.c_str());
        cache.insert(jn);
    }
    bool write_raw(const std::string& name, const std::string& text)
    {
        std::ofstream f(path(name), std::ios::binary | std::ios::trunc);
        if (!f) return false;
        f << text; f.close();
        journal(name);
        return (bool)f;
    }
    void write_json(const std::string& name, const Json& v) { std::string s; json_write(v, s, 0); if (!write_raw(name, s)) throw std::runtime_error("cannot write " + path(name)); }
    JsonP read_json(const std::string& name) const
    {
        std::ifstream f(path(name), std::ios::binary);
        if (!f) throw std::runtime_error("cannot read " + path(name));
        std::stringstream ss; ss << f.rdbuf();
        const std::string text = ss.str();
        JsonReader r(text);
        return r.value();
    }
    void append(const std::string& name, const std::string& line) { { std::ofstream f(path(name), std::ios::app); f << line << "\n"; } journal(name); }
    void log(const std::string& level, const std::string& msg) { append("log", timestamp() + " [" + level + "] " + msg); }
    JsonP jobinfo() const { try { auto j = read_json("jobinfo"); if (j->t == Json::Obj) return j; } catch (...) {} return Json::mk(Json::Obj); }
    void update_jobinfo(const char* exe)                                                        // :270-312
    {
        auto j = jobinfo();
        j->o["cwd"] = Json::str(files_path); j->o["pid"] = Json::num((long long)getpid()); j->o["exe"] = Json::str(exe);
        write_json("jobinfo", *j);
    }
    static JsonP rusage_json(int who)
    {
        struct rusage r; memset(&r, 0, sizeof r); getrusage(who, &r);
        auto d = Json::mk(Json::Obj);
        d->o["ru_utime"] = Json::flt(r.ru_utime.tv_sec + 1e-6 * r.ru_utime.tv_usec); d->o["ru_stime"] = Json::flt(r.ru_stime.tv_sec + 1e-6 * r.ru_stime.tv_usec);
        d->o["ru_maxrss"] = Json::num(r.ru_maxrss); d->o["ru_minflt"] = Json::num(r.ru_minflt); d->o["ru_majflt"] = Json::num(r.ru_majflt);
        d->o["ru_inblock"] = Json::num(r.ru_inblock); d->o["ru_oublock"] = Json::num(r.ru_oublock); d->o["ru_nvcsw"] = Json::num(r.ru_nvcsw); d->o["ru_nivcsw"] = Json::num(r.ru_nivcsw);
        return d;
    }
    void shutdown()                                                                              // :327-348
    {
        log("time", "__end__");
        auto j = jobinfo();
        auto wc = Json::mk(Json::Obj);
        wc->o["start"] = Json::str(start_stamp); wc->o["end"] = Json::str(timestamp()); wc->o["duration_seconds"] = Json::num((long long)(time(nullptr) - start));
        j->o["wallclock"] = wc;
        auto ru = Json::mk(Json::Obj); ru->o["self"] = rusage_json(RUSAGE_SELF); ru->o["children"] = rusage_json(RUSAGE_CHILDREN);
        j->o["rusage"] = ru;
        write_json("jobinfo", *j);
    }
    void complete() { write_raw("complete", timestamp()); shutdown(); }                          // :321-324
};

// ---- the private files between the stages: "key value" lines --------------------------------------------------------
static const char* HANDOFF_MAGIC = "supernova_b200 stage handoff 1";
struct Handoff { std::vector<std::string> fastqs; long long trim_min_qual = -1, min_kmer_obs = -1; };
static void handoff_read(const std::string& path, Handoff& h)
{
    std::ifstream f(path);
    std::string line;
    if (!f || !std::getline(f, line) || line != HANDOFF_MAGIC) throw std::runtime_error(path + ": not a stage file of this adapter (was an earlier stage run by another `tada`?)");
    while (std::getline(f, line)) {
        const size_t sp = line.find(' ');
        if (sp == std::string::npos) continue;
        const std::string k = line.substr(0, sp), v = line.substr(sp + 1);
        if (k == "fastq") h.fastqs.push_back(v);
        else if (k == "trim_min_qual") { const long long q = atoll(v.c_str()); if (h.trim_min_qual >= 0 && h.trim_min_qual != q) throw std::runtime_error("the MSP chunks disagree on trim_min_qual"); h.trim_min_qual = q; }
        else if (k == "min_kmer_obs") h.min_kmer_obs = atoll(v.c_str());
    }
}
static void handoff_write(const std::string& path, const Handoff& h)
{
    std::ofstream f(path, std::ios::trunc);
    f << HANDOFF_MAGIC << "\n";
    if (h.trim_min_qual >= 0) f << "trim_min_qual " << h.trim_min_qual << "\n";
    if (h.min_kmer_obs >= 0) f << "min_kmer_obs " << h.min_kmer_obs << "\n";
    for (auto& q : h.fastqs) f << "fastq " << q << "\n";
    f.close();
    if (!f) throw std::runtime_error("cannot write " + path);
}
static std::vector<std::string> str_array(const Json& v)
{
    std::vector<std::string> out;
    if (v.t == Json::Str) { out.push_back(v.s); return out; }
    if (v.t != Json::Arr) throw std::runtime_error("expected an array of file names");
    for (auto& e : v.a) out.push_back(e->as_str());
    return out;
}
static JsonP chunk_list(std::vector<JsonP> chunks) { auto d = Json::mk(Json::Obj); auto a = Json::mk(Json::Arr); a->a = std::move(chunks); d->o["chunks"] = a; return d; }
static std::string cwd_file(const std::string& name) { char b[4096]; if (!getcwd(b, sizeof b)) throw std::runtime_error("getcwd failed"); return std::string(b) + "/" + name; }

// ---- the stages ----------------------------------------------------------------------------------------------------
static JsonP msp_split(const Json& args)
{
    const std::vector<std::string> fq = str_array(*args.at("fastqs"));
    const std::string perm = cwd_file("permutation.perm");                  // (the stage's split parameter; minimizer order is fixed on the device)
    { std::ofstream f(perm); f << HANDOFF_MAGIC << "\n"; }
    std::vector<JsonP> chunks;
    for (size_t i = 0; i < fq.size(); i += 8) {                            // cmd_msp.rs:268
        auto c = Json::mk(Json::Obj); auto files = Json::mk(Json::Arr);
        for (size_t j = i; j < std::min(fq.size(), i + 8); ++j) files->a.push_back(Json::str(fq[j]));
        c->o["chunk"] = files; c->o["permutation"] = Json::str(perm); c->o["__mem_gb"] = Json::flt(1.0); c->o["__threads"] = Json::num(1);
        chunks.push_back(c);
    }
    return chunk_list(chunks);
}
static JsonP msp_main(const Json& args, const JsonP& outs)
{
    Handoff h;
    h.fastqs = str_array(*args.at("chunk")); h.trim_min_qual = args.at("trim_min_qual")->as_int();
    handoff_write(outs->at("chunks")->as_str(), h);
    return outs;
}
static JsonP msp_join(const std::vector<JsonP>& chunk_outs)
{
    auto d = Json::mk(Json::Obj); auto a = Json::mk(Json::Arr);
    for (auto& c : chunk_outs) a->a.push_back(c->at("chunks"));
    d->o["chunks"] = a;
    return d;
}
static JsonP shard_split()
{
    auto c = Json::mk(Json::Obj);
    c->o["chunk_id"] = Json::num(0); c->o["total_chunks"] = Json::num(1); c->o["__threads"] = Json::num(1); c->o["__mem_gb"] = Json::flt(1.0);
    return chunk_list({c});
}
static JsonP shard_main(const Json& args, const JsonP& outs)
{
    Handoff h;
    for (auto& f : str_array(*args.at("chunks"))) handoff_read(f, h);
    h.min_kmer_obs = args.at("min_kmer_obs")->as_int();
    handoff_write(outs->at("sedge_asm")->as_str(), h);
    { std::ofstream f(outs->at("sedge_bcs")->as_str(), std::ios::trunc); f << HANDOFF_MAGIC << "\n"; }
    return outs;
}
static JsonP shard_join(const std::vector<JsonP>& chunk_outs)
{
    auto d = Json::mk(Json::Obj); auto a = Json::mk(Json::Arr), b = Json::mk(Json::Arr);
    for (size_t i = 0; i < chunk_outs.size(); ++i)
        for (int which = 0; which < 2; ++which) {
            const char* key = which ? "sedge_bcs" : "sedge_asm";
            const std::string from = chunk_outs[i]->at(key)->as_str(), to = cwd_file("chunk" + std::to_string(i) + "." + key);
            if (rename(from.c_str(), to.c_str())) throw std::runtime_error("cannot move " + from + " to " + to);
            (which ? b : a)->a.push_back(Json::str(to));
        }
    d->o["sedge_asm"] = a; d->o["sedge_bcs"] = b;
    return d;
}
static JsonP main_split()
{
    auto c = Json::mk(Json::Obj);
    c->o["__mem_gb"] = Json::flt(16.0); c->o["__threads"] = Json::num(4);
    return chunk_list({c});
}
static JsonP main_main(Metadata& md, const Json& args, const JsonP& outs)
{
    Handoff h;
    for (auto& f : str_array(*args.at("sedge_asm"))) handoff_read(f, h);
    if (h.fastqs.empty() || h.trim_min_qual < 0 || h.min_kmer_obs < 0) throw std::runtime_error("the stage files name no reads or lack trim_min_qual / min_kmer_obs");
    const std::string out = outs->at("asm_graph")->as_str();
    sn_ctx* ctx = nullptr;
    const char* dev = getenv("SN_DEVICE");
    if (sn_ctx_create(&ctx, dev ? atoi(dev) : 0)) throw std::runtime_error(std::string("no CUDA device (there is no CPU fallback): ") + sn_last_error(nullptr));
    auto die = [&](const char* what) { const std::string m = std::string(what) + ": " + sn_last_error(ctx); sn_ctx_destroy(ctx); throw std::runtime_error(m); };
    std::vector<const char*> paths;
    for (auto& f : h.fastqs) paths.push_back(f.c_str());
    md.log("info", "loading " + std::to_string(paths.size()) + " fastq file(s)");
    if (sn_load_fasth_files(ctx, paths.data(), (uint32_t)paths.size())) die("reading the fastq files");
    if (sn_set_semantics(ctx, SN_SEM_TADA)) die("sn_set_semantics");
    sn_params prm; prm.min_qual = (uint32_t)h.trim_min_qual; prm.min_freq = (uint32_t)h.min_kmer_obs; prm.min_bc = 2; prm.ign_bc_below = 0;   // utils.rs:322-408: num_bcs > 1
    if (sn_count_kmers(ctx, &prm)) die("k-mer count");
    if (sn_build_edges(ctx)) die("unipaths");
    if (sn_write_edges_bv(ctx, out.c_str())) die("asm_graph");
    sn_counts c; sn_get_counts(ctx, &c);
    md.log("info", "reads " + std::to_string(c.n_reads) + " bases " + std::to_string(c.n_bases) + " k-mers " + std::to_string(c.n_kmers) + " edges " + std::to_string(c.n_edges));
    sn_ctx_destroy(ctx);
    return outs;
}
static JsonP main_join(const JsonP& outs, const std::vector<JsonP>& chunk_outs)
{
    if (chunk_outs.empty()) throw std::runtime_error("no chunk outs");
    const std::string from = chunk_outs[0]->at("asm_graph")->as_str(), to = outs->at("asm_graph")->as_str();
    if (rename(from.c_str(), to.c_str())) throw std::runtime_error("cannot move " + from + " to " + to);
    auto d = Json::mk(Json::Obj); d->o["asm_graph"] = Json::str(to);
    return d;
}

int main(int argc, char** argv)
{
    if (argc != 7 || strcmp(argv[1], "martian")) {
        fprintf(stderr, "usage: sn_martian martian <msp|shard-asm|main-asm-sn> <split|main|join> <metadata_path> <files_path> <run_file>\n");
        return 2;
    }
    Metadata md;
    md.stage_name = argv[2]; md.stage_type = argv[3]; md.metadata_path = argv[4]; md.files_path = argv[5]; md.run_file = argv[6];
    try {
        md.update_jobinfo(argv[0]);
        md.log("time", "__start__");
        md.journal("stdout"); md.journal("stderr");
    } catch (const std::exception& e) {                       // (a metadata directory that cannot be written: nothing can be reported there)
        fprintf(stderr, "sn_martian: %s\n", e.what());
        return 1;
    }
    // heartbeat: Martian declares a chunk dead without it (lib.rs:537-543)
    std::mutex hm; std::condition_variable hcv; bool done = false;
    std::thread heart([&] { std::unique_lock<std::mutex> lk(hm); while (!done) { md.journal("heartbeat", true); hcv.wait_for(lk, std::chrono::seconds(60)); } });
    auto stop_heart = [&] { { std::lock_guard<std::mutex> lk(hm); done = true; } hcv.notify_all(); heart.join(); };
    int rc = 0;
    try {
        const std::string& st = md.stage_name, & ty = md.stage_type;
        if (st != "msp" && st != "shard-asm" && st != "main-asm-sn") throw std::runtime_error("unknown stage \"" + st + "\" (msp, shard-asm, main-asm-sn)");
        const JsonP args = md.read_json("args");
        if (ty == "split") {
            const JsonP defs = st == "msp" ? msp_split(*args) : (st == "shard-asm" ? shard_split() : main_split());
            md.write_json("stage_defs", *defs);
        } else if (ty == "main") {
            const JsonP outs = md.read_json("outs");
            const JsonP res = st == "msp" ? msp_main(*args, outs) : (st == "shard-asm" ? shard_main(*args, outs) : main_main(md, *args, outs));
            md.write_json("outs", *res);
        } else if (ty == "join") {
            const JsonP outs = md.read_json("outs");
            const JsonP co = md.read_json("chunk_outs");
            if (co->t != Json::Arr) throw std::runtime_error("_chunk_outs is not an array");
            const JsonP res = st == "msp" ? msp_join(co->a) : (st == "shard-asm" ? shard_join(co->a) : main_join(outs, co->a));
            md.write_json("outs", *res);
        } else throw std::runtime_error("unrecognized stage type \"" + ty + "\"");
        md.complete();
    } catch (const std::exception& e) {
        md.write_raw("errors", std::string("sn_martian ") + md.stage_name + " " + md.stage_type + ": " + e.what() + "\n");   // (the adapter's panic hook, :585-600)
        md.log("error", e.what());
        md.shutdown();
        fprintf(stderr, "sn_martian: %s\n", e.what());
        rc = 1;
    }
    stop_heart();
    return rc;
}
