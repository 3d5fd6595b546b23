// sn_graph.cuh -- per-item logic of the graph stages (adjacency pruning and unipath
// edge walking), host+device so tests/hostsim can run the very same code on a CPU.
// References: kmers/ReadPather.h:346-385 (recomputeAdjacencies),
// paths/long/BuildReadQGraph48.cc:327-541 (EdgeBuilder).
#pragma once
#include "sn_kmer.cuh"

namespace sn {

SN_HD Kmer entry_kmer(const DictEntry& e) { Kmer k; k.w0 = e.w0; k.w1 = e.w1; k.w2 = e.w2; return k; }

#define SN_NO_LINK 0xFFFFFFFFu
struct Link2 { uint32_t x, y; };

// a6: AdjProc::operator() -- drop every pred/succ bit whose neighbour k-mer is not
// in the dictionary.  Reads only immutable keys of other entries; writes own ctx.
// While it is looking the neighbours up anyway it remembers, per side, the table index
// (<< 1 | orientation) of the neighbour when exactly one survives: the candidate unipath
// links that classify_links validates without searching again.
SN_HD uint32_t prune_ctx(const DictView& d, uint32_t i, Link2* cand)
{
    const DictEntry& e = d.tab[i];
    Kmer k = entry_kmer(e);
    uint32_t ctx = e.cc >> 24;
    uint32_t ns = 0, np = 0, ls = SN_NO_LINK, lp = SN_NO_LINK;
    // the neighbours share all but one of k's p-mers: their minimizers (hence their buckets -- nearly
    // always k's own) follow from one pass over k
    const KmerMin km = kmer_minimizer_nb(k);
    // (loops over the SET bits, not over the four bases: the lanes of a warp then look their first,
    // second, ... neighbour up together whatever base it is)
    for (uint32_t m = ctx & 0xFu; m; m &= m - 1u) {
        const uint32_t c = low_bit_index(m);
        Kmer q = kmer_succ(k, c), r;
        const bool rc = kmer_form(q, &r) == REV;
        uint32_t j = dict_find_in_bucket(d, succ_minimizer(km, c), rc ? r : q);
        if (j == SN_NULL_EDGE) ctx &= ~(1u << c); else { ++ns; ls = (j << 1) | (rc ? 1u : 0u); }
    }
    // the predecessor side is the successor side of the reverse complement: a walk that leaves
    // through it sees the neighbour in the orientation of succ(rc(k)) = rc(pred(k))
    for (uint32_t m = (ctx >> 4) & 0xFu; m; m &= m - 1u) {
        const uint32_t c = low_bit_index(m);
        Kmer q = kmer_pred(k, c), r;
        const bool rc = kmer_form(q, &r) == REV;
        uint32_t j = dict_find_in_bucket(d, pred_minimizer(km, c), rc ? r : q);
        if (j == SN_NULL_EDGE) ctx &= ~(16u << c); else { ++np; lp = (j << 1) | (rc ? 0u : 1u); }
    }
    if (cand) { cand->x = ns == 1 ? ls : SN_NO_LINK; cand->y = np == 1 ? lp : SN_NO_LINK; }
    return ctx;
}

enum EntryType { T_SINGLE = 0, T_INTERIOR = 1, T_END_DOWN = 2, T_END_UP = 3, T_CIRCLE = 4 };

// Unipath links.  For every dictionary entry two links are stored, one per side of its
// canonical k-mer: x = "down" (successor side), y = "up" (predecessor side, i.e. the successor
// side of the reverse complement).  A link exists exactly when EdgeBuilder would extend across
// it (:408-428, :445-456); it holds the neighbour's table index << 1 | the orientation (1 = RC
// of the stored k-mer) in which a walk arriving over this link sees the neighbour.  Walking an
// edge is then one dependent 8-byte load per k-mer instead of a dictionary probe.
// A candidate link (the single surviving neighbour on one side, found by prune_ctx) is a
// unipath link exactly when EdgeBuilder would extend across it (:408-428, :445-456): the
// neighbour is not a palindrome and, seen in walk orientation, has exactly one predecessor.
SN_HD uint32_t validate_link(const DictView& d, uint32_t cand)
{
    if (cand == SN_NO_LINK) return SN_NO_LINK;
    const DictEntry& nb = d.tab[cand >> 1];
    if (kmer_is_pal(entry_kmer(nb))) return SN_NO_LINK;
    uint32_t nctx = (cand & 1u) ? ctx_rc(nb.ctx) : nb.ctx;
    return mask_single(ctx_pred(nctx)) ? cand : SN_NO_LINK;
}

// EdgeBuilder::buildEdge dispatch (:335-345) expressed through the links
SN_HD int classify_links(const DictView& d, uint32_t i, const Link2& cand, Link2* out)
{
    out->x = SN_NO_LINK; out->y = SN_NO_LINK;
    if (kmer_is_pal(entry_kmer(d.tab[i]))) return T_SINGLE;
    out->x = validate_link(d, cand.x);
    out->y = validate_link(d, cand.y);
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
