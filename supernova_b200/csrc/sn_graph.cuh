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

enum EntryType { T_SINGLE = 0, T_INTERIOR = 1, T_END_DOWN = 2, T_END_UP = 3, T_CIRCLE = 4 };

// Unipath links.  For every dictionary entry two links are stored, one per side of its
// canonical k-mer: x = "down" (successor side), y = "up" (predecessor side, i.e. the successor
// side of the reverse complement).  A link exists exactly when EdgeBuilder would extend across
// it (:408-428, :445-456); it holds the neighbour's table index << 1 | the orientation (1 = RC
// of the stored k-mer) in which a walk arriving over this link sees the neighbour.  Walking an
// edge is then one dependent 8-byte load per k-mer instead of a dictionary probe.
#define SN_NO_LINK 0xFFFFFFFFu
struct Link2 { uint32_t x, y; };

SN_HD uint32_t compute_link(const DictView& d, const Kmer& k, uint32_t ctx)
{
    uint32_t s = ctx_succ(ctx);
    if (!mask_single(s)) return SN_NO_LINK;
    Kmer nx = kmer_succ(k, mask_code(s));
    Kmer rc;
    int form = kmer_form(nx, &rc);
    if (form == PAL) return SN_NO_LINK;
    uint32_t j = dict_find_canonical(d, form == REV ? rc : nx);
    if (j == SN_NULL_EDGE) return SN_NO_LINK;               // cannot happen after pruning
    uint32_t c = d.tab[j].ctx;
    uint32_t nctx = form == REV ? ctx_rc(c) : c;
    if (!mask_single(ctx_pred(nctx))) return SN_NO_LINK;
    return (j << 1) | (form == REV ? 1u : 0u);
}

// EdgeBuilder::buildEdge dispatch (:335-345) expressed through the links
SN_HD int classify_links(const DictView& d, uint32_t i, Link2* out)
{
    const DictEntry& e = d.tab[i];
    Kmer k = entry_kmer(e);
    out->x = SN_NO_LINK; out->y = SN_NO_LINK;
    if (kmer_is_pal(k)) return T_SINGLE;
    out->x = compute_link(d, k, e.ctx);
    out->y = compute_link(d, kmer_rc(k), ctx_rc(e.ctx));
    bool down = out->x != SN_NO_LINK, up = out->y != SN_NO_LINK;
    if (up) return down ? T_INTERIOR : T_END_UP;
    return down ? T_END_DOWN : T_SINGLE;
}

// base appended to the edge sequence when a walk steps onto entry j seen in orientation o:
// the last base of the oriented k-mer
SN_HD uint32_t step_base(const DictEntry& e, uint32_t o) { return o ? ((e.w0 >> 30) ^ 3u) : (e.w2 & 3u); }

// EdgeBuilder::extend (:445-456) over the links: starts at entry i in orientation o and calls
// f(step, entry, orientation) for every k-mer after the first.  Returns the number of k-mers.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class F>
SN_HD uint32_t walk_links(const Link2* links, uint32_t i, uint32_t o, F&& f)
{
    uint32_t cur = i, n = 1;
    for (;;) {
        uint32_t l = o ? links[cur].y : links[cur].x;
        if (l == SN_NO_LINK) break;
        cur = l >> 1; o = l & 1u;
        f(n, cur, o);
        ++n;
    }
    return n;
}
// simpleCircle (:348-372).  Leader election: every k-mer of a smooth circle walks its
// successors and gives up (returns 0) as soon as it meets an entry with a smaller table index,
// so exactly one walker per circle completes the loop (`elect` = true).  With `elect` = false
// the walk always completes (used by the owner to emit the edge).
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class F>
SN_HD uint32_t walk_circle_links(const Link2* links, uint32_t i, bool elect, F&& f)
{
    uint32_t cur = i, o = 0, n = 1;
    for (;;) {
        uint32_t l = o ? links[cur].y : links[cur].x;
        cur = l >> 1; o = l & 1u;
        if (cur == i) break;
        if (elect && cur < i) return 0;
        f(n, cur, o);
        ++n;
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
