// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE.  Runs the host+device (`SN_HD`)
// per-item logic of the product kernels (sn_kmer.cuh / sn_graph.cuh / sn_path.cuh) in
// plain loops on the CPU, stage by stage exactly as the kernels in sn_kernels.cuh
// apply it, so that the graph and pathing logic can be checked against the oracle on a
// box without a GPU.  It is NOT a fallback: nothing in the product links this file.
#include "../../supernova_b200/csrc/sn_kmer.cuh"
#include "../../supernova_b200/csrc/sn_graph.cuh"
#include "../../supernova_b200/csrc/sn_path.cuh"
#include "../../supernova_b200/csrc/sn_msp.cuh"
#include "../../supernova_b200/csrc/sn_dfside.cuh"
#include "../../supernova_b200/csrc/sn_synth.cuh"
#include "../../supernova_b200/csrc/sn_hbv.h"
#include "../../supernova_b200/csrc/sn_formats.h"
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

using namespace sn;

struct Sim {
    std::vector<DictEntry> tab;
    std::vector<Link2> cand;
    std::vector<uint32_t> boff; int bits = 4;
    mutable std::vector<uint32_t> hs;   // the entries' hashes on their own (DictView::hs)
    std::vector<uint32_t> perm;      // perm[i] = position in `tab` of the i-th input (k-mer sorted) record
    snh::Edges edges;
    snh::Hbv hbv;
    std::vector<int32_t> poffset, pedges; std::vector<uint64_t> poff;
    DictView view() const { DictView d; d.tab = tab.data(); d.boff = boff.data(); d.n = (uint32_t)tab.size(); d.bits = bits; d.sub_bits = 0; d.b_lo = 0; d.b_n = 1u << bits; d.g_cap = 0;
        if (hs.size() != tab.size()) { hs.resize(tab.size()); for (size_t i = 0; i < tab.size(); ++i) hs[i] = tab[i].h; }
        d.hs = hs.data(); return d; }
};

extern "C" {

// a2 for one read: emits records {w0,w1,w2,ctx<<24|bc24} like k_extract
uint32_t hs_extract_read(const uint8_t* packed, uint32_t goodlen, int32_t bc, uint32_t* out)
{
    if (goodlen < SN_K + 1) return 0;
    uint32_t n = goodlen - SN_K + 1;
    for (uint32_t i = 0; i < n; ++i) {
        Kmer k = kmer_from_packed(packed, i);
        uint32_t ctx = 0;
        if (i > 0) ctx |= 16u << packed_base(packed, i - 1);
        if (i + SN_K < goodlen) ctx |= 1u << packed_base(packed, i + SN_K);
        Kmer rc;
        if (kmer_form(k, &rc) == REV) { k = rc; ctx = ctx_rc(ctx); }
        out[4 * i] = k.w0; out[4 * i + 1] = k.w1; out[4 * i + 2] = k.w2;
        out[4 * i + 3] = (ctx << 24) | (bc < 0 ? 0xFFFFFFu : (uint32_t)bc);
    }
    return n;
}

// a14 for one read, as k_msp_scan<true> + k_bucket_count's expansion apply it: the read is cut
// into super-k-mer records (sk_out, 8 words each) and every record is expanded again into k-mer
// records {w0,w1,w2,ctx<<24|bc24} (out) with the record's bucket hash per occurrence (bh_out).
// Returns the number of k-mer records; *n_sk the number of super-k-mers.
uint32_t hs_msp_read(const uint8_t* packed, uint32_t goodlen, int32_t bc, uint32_t* out, uint32_t* bh_out, uint32_t* sk_out, uint32_t* n_sk)
{
    uint32_t ring[SN_W];
    uint32_t n = 0, ns = 0;
    const uint32_t bc24 = bc < 0 ? 0xFFFFFFu : (uint32_t)bc;
    msp_scan(packed, goodlen, ring, 1, [&](uint32_t start, uint32_t nk, uint32_t minval) {
        uint32_t w[SN_SK_WORDS];
        sk_build(packed, goodlen, start, nk, bc24, bucket_hash(minval), w);
        for (int i = 0; i < SN_SK_WORDS; ++i) sk_out[SN_SK_WORDS * ns + i] = w[i];
        ++ns;
        for (uint32_t i = 0; i < sk_nk(w[0]); ++i) {
            Kmer k; uint32_t ctx;
            sk_occurrence(w, i, &k, &ctx);
            out[4 * n] = k.w0; out[4 * n + 1] = k.w1; out[4 * n + 2] = k.w2; out[4 * n + 3] = (ctx << 24) | (w[0] & 0xFFFFFFu);
            bh_out[n] = w[1];
            ++n;
        }
    });
    *n_sk = ns;
    return n;
}

// what k_bucket_count's cursor expands: n super-k-mer records -> their k-mer records
uint64_t hs_sk_expand(const uint32_t* sk, uint64_t n, uint32_t* out)
{
    uint64_t m = 0;
    for (uint64_t r = 0; r < n; ++r) {
        const uint32_t* w = sk + SN_SK_WORDS * r;
        for (uint32_t i = 0; i < sk_nk(w[0]); ++i) {
            Kmer k; uint32_t ctx;
            sk_occurrence(w, i, &k, &ctx);
            out[4 * m] = k.w0; out[4 * m + 1] = k.w1; out[4 * m + 2] = k.w2; out[4 * m + 3] = (ctx << 24) | (w[0] & 0xFFFFFFu);
            ++m;
        }
    }
    return m;
}

Sim* hs_new(uint64_t n, const uint32_t* recs /* n x {w0,w1,w2,cc}, sorted by k-mer */, int bits)
{
    Sim* s = new Sim();
    s->bits = bits;
    // what k_bucket_count leaves behind: the dictionary in (minimizer bucket, hash, k-mer) order
    std::vector<uint32_t> order(n), hs(n), bk(n);
    for (uint64_t i = 0; i < n; ++i) {
        Kmer k; k.w0 = recs[4 * i]; k.w1 = recs[4 * i + 1]; k.w2 = recs[4 * i + 2];
        hs[i] = kmer_hash(k); bk[i] = bucket_hash(kmer_minimizer(k)) >> (32 - bits); order[i] = (uint32_t)i;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return bk[a] != bk[b] ? bk[a] < bk[b] : hs[a] < hs[b]; });
    s->tab.resize(n); s->perm.resize(n);
    s->boff.assign((1u << bits) + 1, 0);
    for (uint64_t p = 0; p < n; ++p) {
        uint64_t i = order[p];
        s->perm[i] = (uint32_t)p;
        DictEntry& e = s->tab[p];
        e.w0 = recs[4 * i]; e.w1 = recs[4 * i + 1]; e.w2 = recs[4 * i + 2]; e.cc = recs[4 * i + 3];
        e.edge = SN_NULL_EDGE; e.off = 0; e.ctx = e.cc >> 24; e.h = hs[i];
        ++s->boff[bk[i] + 1];
    }
    for (uint32_t b = 0; b < (1u << bits); ++b) s->boff[b + 1] += s->boff[b];
    return s;
}
void hs_free(Sim* s) { delete s; }

void hs_prune(Sim* s)
{
    DictView d = s->view();
    std::vector<uint32_t> ctx(s->tab.size());
    s->cand.resize(s->tab.size());
    for (uint32_t i = 0; i < s->tab.size(); ++i) ctx[i] = prune_ctx(d, i, &s->cand[i]);      // k_prune
    for (uint32_t i = 0; i < s->tab.size(); ++i) s->tab[i].ctx = ctx[i];
}

// k_classify .. k_pack_edges
int hs_edges(Sim* s)
{
    DictView d = s->view();
    const uint32_t n = (uint32_t)s->tab.size();
    std::vector<uint8_t> etype(n), visited(n, 0);
    std::vector<uint32_t> own_n(n, 0);
    std::vector<Link2> links(n);
    for (uint32_t i = 0; i < n; ++i) { int t = classify_links(d, i, s->cand[i], &links[i]); etype[i] = (uint8_t)t; own_n[i] = t == T_SINGLE ? 1u : 0u; }   // k_classify
    for (uint32_t i = 0; i < n; ++i) if (etype[i] == T_END_DOWN || etype[i] == T_END_UP) {          // k_walk_count
        uint32_t last = i; visited[i] = 1;
        uint32_t nk = walk_links(links.data(), i, etype[i] == T_END_UP ? 1u : 0u, [&](uint32_t, uint32_t j, uint32_t) { visited[j] = 1; last = j; });
        if (i <= last) own_n[i] = nk;
    }
    for (uint32_t i = 0; i < n; ++i) if (etype[i] == T_INTERIOR && !visited[i]) {                     // k_circle_count
        uint32_t m = i; Kmer mk = entry_kmer(s->tab[i]);
        uint32_t nk = walk_circle_links(links.data(), i, true, [&](uint32_t, uint32_t j, uint32_t) { Kmer q = entry_kmer(s->tab[j]); if (q < mk) { mk = q; m = j; } });
        if (nk) { own_n[m] = nk; etype[m] = T_CIRCLE; }
    }
    std::vector<uint32_t> owners; std::vector<uint64_t> base_off(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) { base_off[i + 1] = base_off[i] + (own_n[i] ? own_n[i] + SN_K - 1 : 0); if (own_n[i]) owners.push_back(i); }
    std::vector<uint8_t> tmp(base_off[n] + 16);
    const uint32_t nE = (uint32_t)owners.size();
    std::vector<uint32_t> elen(nE); std::vector<uint8_t> eflip(nE); std::vector<uint64_t> etmp(nE);
    for (uint32_t e = 0; e < nE; ++e) {                                                              // k_walk_emit
        uint32_t i = owners[e]; int t = etype[i];
        uint8_t* sq = tmp.data() + base_off[i];
        Kmer k = entry_kmer(s->tab[i]);
        if (t == T_END_UP) k = kmer_rc(k);
        for (int b = 0; b < SN_K; ++b) sq[b] = (uint8_t)kmer_base(k, b);
        s->tab[i].edge = e; s->tab[i].off = 0;
        uint32_t nk = 1;
        auto visit = [&](uint32_t step, uint32_t j, uint32_t o) { sq[SN_K - 1 + step] = (uint8_t)step_base(s->tab[j], o); s->tab[j].edge = e; s->tab[j].off = step; };
        if (t == T_END_DOWN || t == T_END_UP) nk = walk_links(links.data(), i, t == T_END_UP ? 1u : 0u, visit);
        else if (t == T_CIRCLE) nk = walk_circle_links(links.data(), i, false, visit);
        elen[e] = nk + SN_K - 1; etmp[e] = base_off[i];
        eflip[e] = seq_form_u8(sq, elen[e]) == REV ? 1 : 0;
    }
    for (uint32_t i = 0; i < n; ++i) {                                                               // k_fix_offsets
        uint32_t e = s->tab[i].edge;
        if (e == SN_NULL_EDGE) return -1;
        if (eflip[e]) s->tab[i].off = (elen[e] - (SN_K - 1)) - 1 - s->tab[i].off;
    }
    snh::Edges& E = s->edges;
    E.len = elen; E.off.assign(nE + 1, 0);
    for (uint32_t e = 0; e < nE; ++e) E.off[e + 1] = E.off[e] + (elen[e] + 3) / 4;
    E.packed.assign(E.off[nE] + 16, 0);
    for (uint32_t e = 0; e < nE; ++e)                                                                // k_pack_edges
        for (uint32_t b = 0; b < elen[e]; ++b) {
            const uint8_t* sq = tmp.data() + etmp[e];
            uint32_t c = eflip[e] ? (sq[elen[e] - 1 - b] ^ 3u) : sq[b];
            E.packed[E.off[e] + (b >> 2)] |= (uint8_t)(c << (2 * (b & 3)));
        }
    return 0;
}
uint64_t hs_n_edges(Sim* s) { return s->edges.n(); }
uint64_t hs_edges_bytes(Sim* s) { return s->edges.off.back(); }
void hs_get_edges(Sim* s, uint32_t* len, uint64_t* off, uint8_t* packed)
{
    memcpy(len, s->edges.len.data(), 4 * s->edges.n()); memcpy(off, s->edges.off.data(), 8 * s->edges.off.size());
    memcpy(packed, s->edges.packed.data(), s->edges.off.back());
}
void hs_get_graph_info(Sim* s, uint8_t* ctx, uint32_t* edge, uint32_t* off)
{ for (size_t i = 0; i < s->tab.size(); ++i) { const DictEntry& e = s->tab[s->perm[i]]; ctx[i] = (uint8_t)e.ctx; edge[i] = e.edge; off[i] = e.off; } }

int hs_hbv(Sim* s, const char* hbv_path)
{
    snh::build_hbv(s->edges, s->hbv);
    const snh::Hbv& H = s->hbv;
    if (hbv_path) {
        std::string err;
        std::vector<uint8_t> ep; std::vector<uint64_t> eo; std::vector<uint32_t> el;
        snh::hbv_edge_sequences(s->edges, H, ep, eo, el);
        if (!snf::write_hbv(hbv_path, H.K, (uint64_t)H.n_vert, H.from_start.data(), H.from_v.data(), H.from_e.data(), H.to_start.data(),
                            H.to_e.data(), ep.data(), eo.data(), el.data(), el.size(), err)) return -1;
    }
    return 0;
}

// k_path_reads over all reads; reads in fastb packing, quals one byte per base
int hs_paths(Sim* s, uint64_t n_reads, const uint8_t* bases, const uint64_t* boff, const uint32_t* len,
             const uint8_t* quals, const uint64_t* qoff, const char* paths_file)
{
    DictView d = s->view();
    EdgeStore es; es.bases = s->edges.packed.data(); es.off = s->edges.off.data(); es.len = s->edges.len.data();
    const snh::Hbv& H = s->hbv;
    HbvView h; h.fwd_xlat = H.fwd.data(); h.rev_xlat = H.rev.data(); h.to_left = H.to_left.data(); h.to_right = H.to_right.data();
    h.src = H.src.data(); h.from_start = H.from_start.data(); h.from_v = H.from_v.data(); h.from_e = H.from_e.data();
    h.to_start = H.to_start.data(); h.to_v = H.to_v.data(); h.to_e = H.to_e.data();
    s->poffset.assign(n_reads, 0); s->poff.assign(n_reads + 1, 0); s->pedges.clear();
    std::vector<Part> parts(SN_MAX_PARTS);
    RPath* path = new RPath();
    int overflow = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        path_one_read(d, es, h, bases + boff[r], quals + qoff[r], len[r], parts.data(), *path);
        if (path->overflow) ++overflow;
        s->poffset[r] = path->offset;
        s->pedges.insert(s->pedges.end(), path->e, path->e + path->n);
        s->poff[r + 1] = s->pedges.size();
    }
    delete path;
    if (paths_file) {
        std::string err;
        if (!snf::write_paths(paths_file, n_reads, s->poffset.data(), s->poff.data(), s->pedges.data(), err)) return -1;
    }
    return overflow ? -2 : 0;
}

static HbvView hbv_view_of(const snh::Hbv& H)
{
    HbvView h; h.fwd_xlat = H.fwd.data(); h.rev_xlat = H.rev.data(); h.to_left = H.to_left.data(); h.to_right = H.to_right.data();
    h.src = H.src.data(); h.from_start = H.from_start.data(); h.from_v = H.from_v.data(); h.from_e = H.from_e.data();
    h.to_start = H.to_start.data(); h.to_v = H.to_v.data(); h.to_e = H.to_e.data();
    return h;
}

// k_rpx_sizes, the scan, k_rpx_encode over the paths hs_paths left: a.pathsX
int hs_pathsx(Sim* s, const char* file)
{
    const uint64_t n = s->poffset.size();
    const HbvView h = hbv_view_of(s->hbv);
    std::vector<uint64_t> zoff(n + 1, 0);
    for (uint64_t r = 0; r < n; ++r) zoff[r + 1] = zoff[r] + rpx_size((uint32_t)(s->poff[r + 1] - s->poff[r]));
    std::vector<uint8_t> data(zoff[n] + 16, 0xEE);                       // (junk: the encoder must write every byte of a record)
    std::vector<int64_t> index((n + 9) / 10);
    for (uint64_t r = 0; r < n; ++r) {
        rpx_encode(data.data() + zoff[r], s->pedges.data() + s->poff[r], (uint32_t)(s->poff[r + 1] - s->poff[r]), s->poffset[r], h);
        if (r % 10 == 0) index[r / 10] = (int64_t)zoff[r];
    }
    std::string err;
    return snf::write_pathsx(file, n, index.data(), index.size(), data.data(), zoff[n], err) ? 0 : -1;
}

// k_md_records, the sort, k_md_members, k_md_groups, k_md_art_records, the sort, k_md_art; counts = {ndups, interdups}
int hs_mark_dups(Sim* s, uint64_t n_reads, const uint8_t* bases, const uint64_t* boff, const uint32_t* len, const uint8_t* quals, const uint64_t* qoff,
                 const int32_t* bc, uint8_t* dup, uint8_t* art, uint64_t* counts)
{
    const uint32_t n = (uint32_t)n_reads;
    if (n & 1) return -1;
    std::vector<uint4> rec(n);
    for (uint32_t r = 0; r < n; ++r) {
        const uint32_t m = r ^ 1u, np = (uint32_t)(s->poff[r + 1] - s->poff[r]);
        rec[r] = dup_record(np, np ? s->pedges[s->poff[r]] : 0, s->poffset[r], bases + boff[m], len[m], r);
    }
    auto key_less = [](const uint4& a, const uint4& b) { return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); };
    std::sort(rec.begin(), rec.end(), key_less);
    std::vector<uint32_t> qsum(n, 0), tiegrp(n, 0); std::vector<uint8_t> flags(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        if (rec[i].x == SN_DUP_NONE) continue;
        const bool prev = i > 0 && dup_same_key(rec[i - 1], rec[i]), next = i + 1 < n && dup_same_key(rec[i + 1], rec[i]);
        if (!prev) flags[i] |= 1;
        if (prev || next) {
            flags[i] |= 2;
            for (uint32_t t = 0; t < 2; ++t) { const uint32_t r = rec[i].z ^ t; for (uint32_t b = 0; b < len[r]; ++b) qsum[i] += quals[qoff[r] + b]; }
        }
    }
    memset(dup, 0, n / 2); memset(art, 0, n / 2);
    counts[0] = counts[1] = 0;
    for (uint32_t j = 0; j < n; ++j) {
        if ((flags[j] & 3) != 3) continue;
        uint32_t k = j + 1;
        while (k < n && !(flags[k] & 1) && rec[k].x != SN_DUP_NONE) ++k;
        const DupGroupOut o = dup_group(rec.data(), qsum.data(), j, k, [&](uint32_t id) -> int32_t { return bc ? bc[id] : 0; });
        for (uint32_t l = j; l < k; ++l) { if (l != o.best) dup[rec[l].z >> 1] = 1; if (o.tie) tiegrp[l] = j + 1; }
        counts[0] += k - j - 1; if (o.inter) counts[1] += k - j - 1;
    }
    std::vector<uint4> ar(n);
    for (uint32_t i = 0; i < n; ++i) {
        ar[i].x = ar[i].y = SN_DUP_NONE; ar[i].z = rec[i].z; ar[i].w = 0;
        if (tiegrp[i]) { const uint32_t r = rec[i].z; ar[i].x = tiegrp[i]; ar[i].y = read_content_hash(bases + boff[r], quals + qoff[r], len[r]); }
    }
    std::sort(ar.begin(), ar.end(), key_less);
    for (uint32_t i = 1; i < n; ++i) {
        if (ar[i].x == SN_DUP_NONE || !dup_same_key(ar[i], ar[i - 1])) continue;
        const uint32_t a = ar[i].z;
        for (uint32_t p = i; p-- > 0 && dup_same_key(ar[p], ar[i]);) {
            const uint32_t b = ar[p].z;
            if (read_content_equal(bases + boff[a], quals + qoff[a], len[a], bases + boff[b], quals + qoff[b], len[b])) { art[a >> 1] = 1; break; }
        }
    }
    return 0;
}

// sn_msp.cuh: the window / interleaved-pass bucket mapping of the MSP partition kernels
uint32_t hs_window_bucket(uint32_t b, uint32_t b_lo, uint32_t b_n, uint32_t pcfg) { return msp_window_bucket(b, b_lo, b_n, pcfg); }

// sn_synth.cuh: reads [2 * first_pair, 2 * (first_pair + n_pairs)) of the counter-based generator, as k_synth_reads makes them
void hs_synth_reads(uint64_t G, uint64_t total_pairs, uint32_t n_bc, uint64_t seed, const uint32_t* T, uint64_t first_pair, uint64_t n_pairs,
                    uint8_t* bases, uint8_t* quals, int32_t* bc)
{
    SynSpec sp; sp.genome_bases = G; sp.total_pairs = total_pairs; sp.seed = seed; sp.n_barcodes = n_bc;
    for (uint64_t r = 0; r < 2 * n_pairs; ++r) {
        const uint64_t p = first_pair + (r >> 1);
        syn_read(sp, T, p, (uint32_t)(r & 1), bases + r * SN_SYN_L, quals + r * SN_SYN_L);
        bc[r] = syn_barcode(sp, p);
    }
}

}  // extern "C"
