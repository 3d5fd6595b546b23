// sn_graph.cuh -- per-item logic of the graph stages (adjacency pruning and unipath
// edge walking), host+device so tests/hostsim can run the very same code on a CPU.
// References: kmers/ReadPather.h:346-385 (recomputeAdjacencies),
// paths/long/BuildReadQGraph48.cc:327-541 (EdgeBuilder).
#pragma once
#include "sn_kmer.cuh"

namespace sn {

SN_HD Kmer entry_kmer(const DictEntry& e) { Kmer k; k.w0 = e.w0; k.w1 = e.w1; k.w2 = e.w2; return k; }

// a6: AdjProc::operator() -- drop every pred/succ bit whose neighbour k-mer is not
// in the dictionary.  Reads only immutable keys of other entries; writes own ctx.
SN_HD uint32_t prune_ctx(const DictView& d, uint32_t i)
{
    const DictEntry& e = d.tab[i];
    Kmer k = entry_kmer(e);
    uint32_t ctx = e.cc >> 24;
    for (uint32_t c = 0; c < 4; ++c)
        if (ctx & (1u << c)) { if (dict_find(d, kmer_succ(k, c), nullptr) == SN_NULL_EDGE) ctx &= ~(1u << c); }
    for (uint32_t c = 0; c < 4; ++c)
        if (ctx & (16u << c)) { if (dict_find(d, kmer_pred(k, c), nullptr) == SN_NULL_EDGE) ctx &= ~(16u << c); }
    return ctx;
}

// EdgeBuilder::lookup (:466-476): entry index of `k` and its context expressed in
// the orientation of `k`.  The neighbour is guaranteed present after pruning.
SN_HD uint32_t oriented_lookup(const DictView& d, const Kmer& k, uint32_t* ctx)
{
    bool rc;
    uint32_t j = dict_find(d, k, &rc);
    uint32_t c = (j == SN_NULL_EDGE) ? 0u : d.tab[j].ctx;
    *ctx = rc ? ctx_rc(c) : c;
    return j;
}

// "extension possible" toward the successor side of oriented k-mer (k, ctx):
// exactly one successor, successor not a palindrome, successor has exactly one
// predecessor (:419-428; the upstream test :408-417 is this one on the RC strand).
SN_HD bool down_possible(const DictView& d, const Kmer& k, uint32_t ctx)
{
    uint32_t s = ctx_succ(ctx);
    if (!mask_single(s)) return false;
    Kmer nx = kmer_succ(k, mask_code(s));
    if (kmer_is_pal(nx)) return false;
    uint32_t nctx;
    oriented_lookup(d, nx, &nctx);
    return mask_single(ctx_pred(nctx));
}

enum EntryType { T_SINGLE = 0, T_INTERIOR = 1, T_END_DOWN = 2, T_END_UP = 3 };

// EdgeBuilder::buildEdge dispatch (:335-345)
SN_HD int classify_entry(const DictView& d, uint32_t i)
{
    const DictEntry& e = d.tab[i];
    Kmer k = entry_kmer(e);
    if (kmer_is_pal(k)) return T_SINGLE;
    bool up = down_possible(d, kmer_rc(k), ctx_rc(e.ctx));
    bool down = down_possible(d, k, e.ctx);
    if (up) return down ? T_INTERIOR : T_END_UP;
    return down ? T_END_DOWN : T_SINGLE;
}

// EdgeBuilder::extend (:445-456) as a visitor: walks from the end entry `i` in the
// direction given by its type and calls f(step, entry_index, appended_base) for every
// k-mer after the first.  Returns the number of k-mers on the edge.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class F>
SN_HD uint32_t walk_edge(const DictView& d, uint32_t i, int type, F&& f)
{
    const DictEntry& e = d.tab[i];
    Kmer k = entry_kmer(e);
    uint32_t ctx = e.ctx;
    if (type == T_END_UP) { k = kmer_rc(k); ctx = ctx_rc(ctx); }
    uint32_t n = 1;
    while (mask_single(ctx_succ(ctx))) {
        uint32_t c = mask_code(ctx_succ(ctx));
        Kmer nx = kmer_succ(k, c);
        if (kmer_is_pal(nx)) break;
        uint32_t nctx;
        uint32_t j = oriented_lookup(d, nx, &nctx);
        if (!mask_single(ctx_pred(nctx))) break;
        f(n, j, c);
        k = nx; ctx = nctx; ++n;
    }
    return n;
}

// simpleCircle (:348-372).  Leader election: every k-mer of a smooth circle walks its
// successors and gives up (returns 0) as soon as it meets an entry with a smaller table index,
// so exactly one walker per circle completes the loop (`elect` = true).  With `elect` = false
// the walk always completes (used by the owner to emit the edge).  f(step, entry, base) is
// called for every k-mer after the first.  Returns the number of k-mers.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class F>
SN_HD uint32_t walk_circle(const DictView& d, uint32_t i, bool elect, F&& f)
{
    const DictEntry& e = d.tab[i];
    Kmer k = entry_kmer(e);
    uint32_t ctx = e.ctx;
    uint32_t n = 1;
    for (;;) {
        uint32_t c = mask_code(ctx_succ(ctx));
        Kmer nx = kmer_succ(k, c);
        uint32_t nctx;
        uint32_t j = oriented_lookup(d, nx, &nctx);
        if (j == i) break;
        if (elect && j < i) return 0;
        f(n, j, c);
        k = nx; ctx = nctx; ++n;
    }
    return n;
}

// getCanonicalForm(beg,end) for a run-time length over unpacked base codes
// (dna/CanonicalForm.h:34-46): odd lengths use bit 1 of the middle base.
SN_HD int seq_form_u8(const uint8_t* s, uint32_t len)
{
    if (len & 1) return (s[len / 2] & 2) ? REV : FWD;
    uint32_t i = 0, j = len;
    while (i != j) {
        uint32_t f = s[i], r = s[--j] ^ 3u;
        if (f < r) return FWD;
        if (r < f) return REV;
        ++i;
    }
    return PAL;
}

}  // namespace sn
