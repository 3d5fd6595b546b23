#include "sn_hbv.h"
#include <algorithm>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <functional>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <string>

namespace snh {
namespace {

constexpr int K = 48;
inline uint32_t base_at(const uint8_t* p, uint64_t i) { return (p[i >> 2] >> (2 * (i & 3))) & 3u; }
inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct Sub { uint64_t hi, lo; };        // (K-1)-mer, 2 bits per base, MSB first, left aligned in 128 bits
inline bool operator==(const Sub& a, const Sub& b) { return a.hi == b.hi && a.lo == b.lo; }

// reverse the order of the 2-bit fields of a 64-bit word
inline uint64_t rev2_64(uint64_t x)
{
    x = __builtin_bswap64(x);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
}
// EdgeEnd(pBV, rc, distal, K-1) (HBVFromEdges.cc:28-56; SwitchHitterIter feudal/BaseVec.h:98-126).
// Vertices only need EQUALITY of (K-1)-mers, so the key is the 94-bit window in fastb bit
// order (base j at bits 2j); the reverse-complement strand is the complemented window with its
// fields reversed.  Reads 16 bytes: the packed store is padded.
Sub edge_end(const uint8_t* s, uint32_t len, bool rc, bool distal)
{
    typedef unsigned __int128 u128;
    const u128 MASK = (((u128)1) << 94) - 1;
    uint32_t pos = distal == rc ? 0 : len - (K - 1);     // rc strand: its first K-1 bases are the last K-1 of the edge
    const uint8_t* p = s + (pos >> 2);
    u128 w = ((u128)load64(p + 8) << 64) | load64(p);
    w = (w >> (2 * (pos & 3))) & MASK;
    if (rc) {
        w = (~w) & MASK;
        u128 r = ((u128)rev2_64((uint64_t)w) << 64) | rev2_64((uint64_t)(w >> 64));
        w = r >> 34;
    }
    return Sub{(uint64_t)(w >> 64), (uint64_t)w};
}
// getCanonicalForm of a whole edge (dna/CanonicalForm.h:34-46): 0 fwd, 1 rev, 2 palindrome
int seq_form(const uint8_t* s, uint32_t len)
{
    if (len & 1) return (base_at(s, len / 2) & 2) ? 1 : 0;
    uint32_t i = 0, j = len;
    while (i != j) {
        uint32_t f = base_at(s, i), r = base_at(s, --j) ^ 3u;
        if (f < r) return 0;
        if (r < f) return 1;
        ++i;
    }
    return 2;
}
// lexicographic comparison of two equally long fastb-packed sequences, 32 bases per step:
// the first differing base is the lowest set 2-bit field of x^y (LSB-first packing)
inline int cmp_packed(const uint8_t* x, const uint8_t* y, uint32_t len)
{
    uint32_t nb = (len + 3) / 4, i = 0;
    for (; i + 8 <= nb; i += 8) {
        uint64_t a = load64(x + i), b = load64(y + i);
        if (a != b) { int sh = __builtin_ctzll(a ^ b) & ~1; uint32_t p = (a >> sh) & 3, q = (b >> sh) & 3; return p < q ? -1 : 1; }
    }
    for (; i < nb; ++i)
        if (x[i] != y[i]) { int sh = __builtin_ctz((unsigned)(x[i] ^ y[i])) & ~1; uint32_t p = (x[i] >> sh) & 3, q = (y[i] >> sh) & 3; return p < q ? -1 : 1; }
    return 0;
}
// static range split over a few host threads (the per-item work here is independent)
void parallel_ranges(uint64_t n, const std::function<void(uint64_t, uint64_t)>& fn)
{
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = n < 65536 ? 1u : std::min(8u, hw ? hw : 1u);
    if (nt <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(fn, n * t / nt, n * (t + 1) / nt);
    for (auto& x : th) x.join();
}
inline uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }

}  // namespace

#ifdef SN_HOSTSIM      // host-only construction: compiled into tests/hostsim only, never into the product library
// HBVBuilder::add / processQueue (HBVFromEdges.cc:189-228) from one start item: FIFO breadth-first
// numbering of everything reachable.  Vertices are numbered from nextV, HBV edges from nH (both
// advanced).  An item is queued at most once (id -2 = queued): the reference queues duplicates
// and skips them when popped, so an item is processed at its FIRST position either way.
// (A palindromic edge is only ever processed with rc=0: its rev id is set with its fwd id.)
static void bfs_from(GroupRec* groups, ERec* er, uint32_t start, int32_t& nextV, uint32_t& nH,
                     uint32_t* src, int32_t* to_left, int32_t* to_right, std::vector<uint32_t>& Q)
{
    size_t qh = 0, qt = 0;
    if (Q.size() < 64) Q.resize(64);
    Q[qt++] = start;
    er[start].id = -2;
    while (qh < qt) {
        const uint32_t it = Q[qh++];
        ERec& r = er[it];
        if (r.id != -2) continue;                           // palindrome twin already numbered
        if (r.g1 < 0 || r.g2 < 0) throw std::runtime_error("HBV: edge end without a vertex");
        GroupRec& G1 = groups[r.g1]; GroupRec& G2 = groups[r.g2];
        if (G1.vid == -1) G1.vid = nextV++;
        if (G2.vid == -1) G2.vid = nextV++;
        const int32_t id = (int32_t)nH++;
        src[id] = it; to_left[id] = G1.vid; to_right[id] = G2.vid;
        r.id = id;
        if (r.pal) er[it ^ 1u].id = id;
        if (qt + 16 > Q.size()) {                           // keep the FIFO compact
            std::copy(Q.begin() + qh, Q.begin() + qt, Q.begin()); qt -= qh; qh = 0;
            if (qt + 16 > Q.size()) Q.resize(2 * Q.size());
        }
        // a vertex pushes its items once (pad[0] = expanded); the loop is bound by cache misses on the
        // vertex and item records, so they are fetched as soon as an item is queued
        if (!G1.pad[0]) { G1.pad[0] = 1;
            for (uint32_t x = 0; x < G1.n; ++x) { const uint32_t t2 = G1.items[x]; ERec& q = er[t2];
                if (q.id == -1) { q.id = -2; Q[qt++] = t2; __builtin_prefetch(&groups[q.g1]); __builtin_prefetch(&groups[q.g2]); } } }
        if (!G2.pad[0]) { G2.pad[0] = 1;
            for (uint32_t x = 0; x < G2.n; ++x) { const uint32_t t2 = G2.items[x]; ERec& q = er[t2];
                if (q.id == -1) { q.id = -2; Q[qt++] = t2; __builtin_prefetch(&groups[q.g1]); __builtin_prefetch(&groups[q.g2]); } } }
        if (qh + 4 < qt) {                                   // and the item records of what those vertices will push
            const ERec& nx = er[Q[qh + 4]];
            const GroupRec& A = groups[nx.g1]; const GroupRec& B = groups[nx.g2];
            if (!A.pad[0]) for (uint32_t x = 0; x < A.n; ++x) __builtin_prefetch(&er[A.items[x]]);
            if (!B.pad[0]) for (uint32_t x = 0; x < B.n; ++x) __builtin_prefetch(&er[B.items[x]]);
        }
    }
}

#endif

// The same traversal over the one-line item records the device prepares (product path).  A popped
// item costs one cache line, and that line was requested when the item was queued.  What the loop
// branches on -- "vertex already numbered" (the first item to reach a vertex numbers it AND pushes
// its items, so one flag serves both) and "item already queued" -- lives in two byte arrays that stay
// in L2 (bytes, not bits: components numbered by different threads share the arrays, never a byte);
// the vertex ids themselves are only loaded, never branched on.  The
// pushes are branch-free (store, then advance the tail by 0 or 1): on a long chain of bubbles
// (queue length 2-4) the loop is otherwise a sequence of mispredicted branches on cache misses.
struct NumberState { int32_t* vid; uint8_t* vbits; uint8_t* seen; int32_t* ids; };
static inline bool bit_test_set(uint8_t* b, uint32_t i) { const bool was = b[i] != 0; b[i] = 1; return was; }
template <bool LAID_OUT>
static void bfs_items(const ItemRec* rec, const GroupRec* groups, NumberState& st, uint32_t start, int32_t& nextV, uint32_t& nH,
                      uint32_t* src, int32_t* to_left, int32_t* to_right, std::vector<uint32_t>& Q)
{
    size_t qh = 0, qt = 0;
    if (Q.size() < 64) Q.resize(64);
    Q[qt++] = start; bit_test_set(st.seen, start);
    while (qh < qt) {
        const uint32_t it = Q[qh++];
        const ItemRec& r = rec[it];
        if (r.g1 < 0 || r.g2 < 0) throw std::runtime_error("HBV: edge end without a vertex");
        const bool new1 = !bit_test_set(st.vbits, (uint32_t)r.g1);
        if (new1) st.vid[r.g1] = nextV++;
        const bool new2 = !bit_test_set(st.vbits, (uint32_t)r.g2);         // (g1 == g2: already set)
        if (new2) st.vid[r.g2] = nextV++;
        const int32_t id = (int32_t)nH++;
        src[id] = LAID_OUT ? r.pad : it; to_left[id] = st.vid[r.g1]; to_right[id] = st.vid[r.g2];
        st.ids[it] = id;                                     // (laid out: indexed by record; number_hbv maps back)
        if (!LAID_OUT && ((r.info >> 8) & 1u)) { st.ids[it ^ 1u] = id; bit_test_set(st.seen, it ^ 1u); }
        if (qt + 24 > Q.size()) {                           // keep the FIFO compact
            std::copy(Q.begin() + qh, Q.begin() + qt, Q.begin()); qt -= qh; qh = 0;
            if (qt + 24 > Q.size()) Q.resize(2 * Q.size());
        }
        uint32_t* q = Q.data();
        for (int side = 0; side < 2; ++side) {
            if (!(side ? new2 : new1)) continue;            // the vertex pushed its items when it was numbered
            const int32_t g = side ? r.g2 : r.g1;
            uint32_t n = (r.info >> (4 * side)) & 15u;
            const uint32_t* items = side ? r.it2 : r.it1;
            if (n == 15u) {                                  // more than 6 edge ends on this vertex
                n = groups[g].n; items = groups[g].items;
                for (uint32_t x = 0; x < n; ++x) {
                    const uint32_t t2 = items[x];
                    if (!bit_test_set(st.seen, t2)) { q[qt++] = t2; __builtin_prefetch(&rec[t2]); }
                }
                continue;
            }
#pragma GCC unroll 6
            for (uint32_t x = 0; x < 6; ++x) {
                const bool live = x < n;
                const uint32_t t2 = live ? items[x] : start;                  // (start is queued: a dead lane pushes nothing)
                __builtin_prefetch(&rec[t2]);
                const bool was = bit_test_set(st.seen, t2);
                q[qt] = t2; qt += was ? 0u : 1u;
            }
        }
    }
}

// The device found the connected components, the item the reference's outer loop
// (HBVFromEdges.cc:277-285) reaches first in each, and -- from their sizes -- the first vertex and
// edge id of every component, so the components are numbered independently, in parallel, with
// their final ids.
void number_hbv(const HbvComponents& C, const ItemRec* items, const GroupRec* groups, uint64_t nV_in, uint64_t nE, Hbv& H, unsigned threads, const uint32_t* layout,
                unsigned part, unsigned n_parts, uint64_t min_items)
{
    const uint64_t nH = C.n_comp ? C.base_e[C.n_comp] : 0, nV = C.n_comp ? C.base_v[C.n_comp] : 0;
    if (nV != nV_in) throw std::runtime_error("HBV: component vertex counts do not add up");
    H.n_vert = (int32_t)nV;
    H.src.resize(nH); H.to_left.resize(nH); H.to_right.resize(nH);      // (no reallocation when the caller sized them already)
    H.fwd.resize(nE); H.rev.resize(nE);
    std::vector<int32_t> vid(nV, -1), ids(2 * nE, -1);
    std::vector<uint8_t> vbits(nV + 1, 0), seen(2 * nE + 1, 0);
    NumberState st{vid.data(), vbits.data(), seen.data(), ids.data()};
    if (!threads) threads = 1;
    // Work units: a big component on its own (both strands of a well-covered genome are one giant
    // component each, and they sit next to each other at the head of the list: they must not share a
    // thread), small components 256 at a time.
    std::vector<uint64_t> unit_start;
    {
        uint64_t c = 0;
        while (c < C.n_comp) {
            unit_start.push_back(c);
            if (C.base_e[c + 1] - C.base_e[c] >= 4096) { ++c; continue; }
            uint64_t e = c, lim = std::min<uint64_t>(C.n_comp, c + 256);
            while (e < lim && C.base_e[e + 1] - C.base_e[e] < 4096) ++e;
            c = e;
        }
        unit_start.push_back(C.n_comp);
    }
    const uint64_t n_units = unit_start.size() - 1;
    std::atomic<uint64_t> next{0};
    std::atomic<int> bad{0};
    std::string what;
    auto worker = [&]() {
        std::vector<uint32_t> Q(1024);
        try {
            for (;;) {
                const uint64_t u = next.fetch_add(1);
                if (u >= n_units || bad.load()) break;
                if (n_parts > 1 && u % n_parts != part) continue;          // another rank numbers this unit (ids are final: the results add up)
                for (uint64_t c = unit_start[u]; c < unit_start[u + 1]; ++c) {
                    if (C.base_e[c + 1] - C.base_e[c] < min_items) continue;           // numbered on the device (k_hbv_number_small)
                    int32_t nextV = (int32_t)C.base_v[c]; uint32_t nh = (uint32_t)C.base_e[c];
                    const uint32_t start = layout ? layout[C.start_item[c]] : C.start_item[c];
                    if (C.base_e[c + 1] - C.base_e[c] >= 32768 && threads > 1) {
                        // A giant component works on private state: the items of the two strand components
                        // interleave in every cache line of the shared arrays, and two threads numbering
                        // them side by side would spend their time passing those lines back and forth.
                        std::vector<int32_t> pvid(nV, -1), pids(2 * nE, -1);
                        std::vector<uint8_t> pvb(nV + 1, 0), pseen(2 * nE + 1, 0);
                        NumberState ps{pvid.data(), pvb.data(), pseen.data(), pids.data()};
                        if (layout) bfs_items<true>(items, groups, ps, start, nextV, nh, H.src.data(), H.to_left.data(), H.to_right.data(), Q);
                        else bfs_items<false>(items, groups, ps, start, nextV, nh, H.src.data(), H.to_left.data(), H.to_right.data(), Q);
                        for (uint64_t h = C.base_e[c]; h < (uint64_t)nh; ++h) {
                            const uint32_t it = H.src[h];                    // the item itself, whatever the layout
                            const uint32_t rec_i = layout ? layout[it] : it;
                            ids[rec_i] = (int32_t)h;
                            if (!layout && ((items[it].info >> 8) & 1u)) ids[it ^ 1u] = (int32_t)h;
                        }
                    } else if (layout) bfs_items<true>(items, groups, st, start, nextV, nh, H.src.data(), H.to_left.data(), H.to_right.data(), Q);
                    else bfs_items<false>(items, groups, st, start, nextV, nh, H.src.data(), H.to_left.data(), H.to_right.data(), Q);
                    if ((uint64_t)nextV != C.base_v[c + 1] || (uint64_t)nh != C.base_e[c + 1])
                        throw std::runtime_error("HBV: a component was not numbered completely");
                }
            }
        } catch (const std::exception& ex) { if (!bad.exchange(1)) what = ex.what(); }
    };
    if (threads <= 1 || n_units < 2) worker();
    else {
        std::vector<std::thread> th;
        const unsigned nt = (unsigned)std::min<uint64_t>(threads, n_units);
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto& x : th) x.join();
    }
    if (bad.load()) throw std::runtime_error(what);
    if (!layout) parallel_ranges(nE, [&](uint64_t e0, uint64_t e1) { for (uint64_t e = e0; e < e1; ++e) { H.fwd[e] = ids[2 * e]; H.rev[e] = ids[2 * e + 1]; } });
    else parallel_ranges(nE, [&](uint64_t e0, uint64_t e1) {
        for (uint64_t e = e0; e < e1; ++e) {
            const uint32_t f = layout[2 * e];
            H.fwd[e] = ids[f];
            H.rev[e] = ((items[f].info >> 8) & 1u) ? ids[f] : ids[layout[2 * e + 1]];      // a palindromic edge has one HBV edge
        }
    });
    // one part of several: what the other parts number stays 0, so that the parts' arrays add up to the whole
    if (n_parts > 1 || min_items > 0) parallel_ranges(nE, [&](uint64_t e0, uint64_t e1) { for (uint64_t e = e0; e < e1; ++e) { if (H.fwd[e] < 0) H.fwd[e] = 0; if (H.rev[e] < 0) H.rev[e] = 0; } });
}

#ifdef SN_HOSTSIM
// Host-only construction (vertex discovery with a hash table, sequential numbering, adjacency):
// what tests/hostsim runs on a CPU box.  The product uses the device stages of sn_hbvdev.cuh with
// number_hbv in between.
void build_hbv(const Edges& E, Hbv& H)
{
    const uint64_t nE = E.n();
    H = Hbv();
    H.fwd.assign(nE, -1); H.rev.assign(nE, -1);
    H.from_start.assign(1, 0); H.to_start.assign(1, 0);
    if (!nE) return;
    const uint8_t* P = E.packed.data();
    // BVComp (HBVFromEdges.cc:106-111): longer first, then lexicographic on bases
    auto bvcomp = [&](uint32_t a, uint32_t b) {
        if (E.len[a] != E.len[b]) return E.len[a] > E.len[b];
        int c = cmp_packed(P + E.off[a], P + E.off[b], E.len[a]);
        return c ? c < 0 : a < b;
    };
    std::vector<uint32_t> order(nE), rank(nE);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), bvcomp);
    for (uint64_t i = 0; i < nE; ++i) rank[order[i]] = (uint32_t)i;
    // VertexDictBuilder (:124-168): 4 ends per edge (2 for a palindromic edge); a vertex is a
    // distinct (K-1)-mer.  item = edge<<2 | rc<<1 | distal.
    std::vector<uint8_t> pal(nE);
    std::vector<int32_t> end_group(4 * nE, -1);
    std::vector<uint32_t> gstart, gitems;
    {
        uint64_t cap = 16; while (cap < 8 * nE) cap <<= 1;
        std::vector<int32_t> slot_group(cap, -1);
        std::vector<Sub> gkey; gkey.reserve(2 * nE);
        std::vector<uint32_t> gcount;
        for (int phase = 0; phase < 2; ++phase) {            // phase 0 counts, phase 1 fills
            for (uint32_t e = 0; e < nE; ++e) {
                const uint8_t* s = P + E.off[e]; uint32_t len = E.len[e];
                if (!phase) pal[e] = seq_form(s, len) == 2;
                for (int rc = 0; rc < (pal[e] ? 1 : 2); ++rc)
                    for (int distal = 0; distal < 2; ++distal) {
                        int32_t g;
                        if (!phase) {
                            Sub k = edge_end(s, len, rc, distal);
                            uint64_t h = mix(k.hi ^ mix(k.lo)) & (cap - 1);
                            for (;;) {
                                g = slot_group[h];
                                if (g < 0) { g = (int32_t)gkey.size(); slot_group[h] = g; gkey.push_back(k); gcount.push_back(0); break; }
                                if (gkey[g] == k) break;
                                h = (h + 1) & (cap - 1);
                            }
                            end_group[4 * (size_t)e + 2 * rc + distal] = g;
                            ++gcount[g];
                        } else {
                            g = end_group[4 * (size_t)e + 2 * rc + distal];
                            gitems[gstart[g] + gcount[g]++] = (e << 2) | (uint32_t)(rc << 1) | (uint32_t)distal;
                        }
                    }
            }
            if (!phase) {
                gstart.assign(gkey.size() + 1, 0);
                for (size_t g = 0; g < gkey.size(); ++g) gstart[g + 1] = gstart[g] + gcount[g];
                gitems.resize(gstart.back());
                std::fill(gcount.begin(), gcount.end(), 0u);
            }
        }
    }
    const int32_t nV = (int32_t)gstart.size() - 1;
    // inside a vertex: EEComp order (:113-121) = edge rank, rc, pos (pos 0 < pos len-(K-1))
    std::vector<GroupRec> groups(nV);
    for (int32_t g = 0; g < nV; ++g) {
        uint32_t* it = &gitems[gstart[g]]; int n = (int)(gstart[g + 1] - gstart[g]);
        if (n > 8) throw std::runtime_error("HBV: a vertex has more than 8 edge ends (HBVFromEdges.cc:83)");
        std::sort(it, it + n, [&](uint32_t x, uint32_t y) { return ((((uint64_t)rank[x >> 2]) << 2) | (x & 3)) < ((((uint64_t)rank[y >> 2]) << 2) | (y & 3)); });
        GroupRec& r = groups[g]; memset(&r, 0, sizeof r); r.vid = -1; r.n = (uint32_t)n;
        for (int i = 0; i < n; ++i) r.items[i] = it[i] >> 1;
    }
    std::vector<ERec> er(2 * nE);
    for (uint64_t e = 0; e < nE; ++e)
        for (uint32_t rc = 0; rc < 2; ++rc) {
            ERec& r = er[2 * e + rc];
            r.g1 = end_group[4 * e + 2 * rc + 0]; r.g2 = end_group[4 * e + 2 * rc + 1]; r.id = -1; r.pal = pal[e];
        }
    H.src.resize(2 * nE); H.to_left.resize(2 * nE); H.to_right.resize(2 * nE);
    std::vector<uint32_t> Q(1024);
    int32_t nextV = 0; uint32_t nH = 0;
    for (uint32_t pass = 0; pass < 2; ++pass)                 // buildHBVFromEdges' outer loop (:277-285)
        for (uint64_t oi = 0; oi < nE; ++oi) {
            const uint32_t it = (order[oi] << 1) | pass;
            if (er[it].id != -1) continue;
            bfs_from(groups.data(), er.data(), it, nextV, nH, H.src.data(), H.to_left.data(), H.to_right.data(), Q);
        }
    if (nextV != nV) throw std::runtime_error("HBV: vertex numbering did not reach every vertex");
    H.n_vert = nV;
    H.src.resize(nH); H.to_left.resize(nH); H.to_right.resize(nH);
    for (uint64_t e = 0; e < nE; ++e) { H.fwd[e] = er[2 * e].id; H.rev[e] = er[2 * e + 1].id; }
    // digraphE::AddEdge (graph/DigraphTemplate.h:2572-2582) inserts edge id n at upper_bound of the
    // neighbour vertex: every list ends up sorted by (neighbour, id).  Counting sort by vertex
    // (ids ascending), then order each short list by neighbour, stably.
    H.from_start.assign(nV + 1, 0); H.to_start.assign(nV + 1, 0);
    for (uint32_t h = 0; h < nH; ++h) { ++H.from_start[H.to_left[h] + 1]; ++H.to_start[H.to_right[h] + 1]; }
    for (int32_t v = 0; v < nV; ++v) { H.from_start[v + 1] += H.from_start[v]; H.to_start[v + 1] += H.to_start[v]; }
    H.from_v.resize(nH); H.from_e.resize(nH); H.to_v.resize(nH); H.to_e.resize(nH);
    {
        std::vector<uint32_t> fc(H.from_start.begin(), H.from_start.end() - 1), tc(H.to_start.begin(), H.to_start.end() - 1);
        for (uint32_t h = 0; h < nH; ++h) {
            uint32_t a = fc[H.to_left[h]]++; H.from_v[a] = H.to_right[h]; H.from_e[a] = (int32_t)h;
            uint32_t b = tc[H.to_right[h]]++; H.to_v[b] = H.to_left[h]; H.to_e[b] = (int32_t)h;
        }
    }
    bool over4 = false;
    auto order_lists = [&](const std::vector<uint32_t>& start, std::vector<int32_t>& nb, std::vector<int32_t>& eo) {
        for (int32_t v = 0; v < nV; ++v) {
            uint32_t s = start[v], n = start[v + 1] - s;
            if (n > 4) over4 = true;
            for (uint32_t i = 1; i < n; ++i) {                       // stable insertion sort by neighbour
                int32_t w = nb[s + i], e = eo[s + i]; uint32_t j = i;
                while (j > 0 && nb[s + j - 1] > w) { nb[s + j] = nb[s + j - 1]; eo[s + j] = eo[s + j - 1]; --j; }
                nb[s + j] = w; eo[s + j] = e;
            }
        }
    };
    order_lists(H.from_start, H.from_v, H.from_e);
    order_lists(H.to_start, H.to_v, H.to_e);
    if (over4) throw std::runtime_error("HBV: more than 4 edges on one side of a vertex");
    // Involution: the reverse complement of HBV edge fwd[e] is rev[e]
    H.inv.assign(nH, -1);
    for (uint64_t e = 0; e < nE; ++e) { H.inv[H.fwd[e]] = H.rev[e]; H.inv[H.rev[e]] = H.fwd[e]; }
}

#endif  // SN_HOSTSIM

void hbv_edge_sequences(const Edges& E, const Hbv& H, std::vector<uint8_t>& epacked, std::vector<uint64_t>& eoff, std::vector<uint32_t>& elen)
{
    const uint64_t nH = H.src.size();
    eoff.assign(nH + 1, 0); elen.resize(nH);
    for (uint64_t h = 0; h < nH; ++h) { elen[h] = E.len[H.src[h] >> 1]; eoff[h + 1] = eoff[h] + (elen[h] + 3) / 4; }
    epacked.assign(eoff[nH] + 16, 0);
    for (uint64_t h = 0; h < nH; ++h) {
        uint32_t u = H.src[h] >> 1, len = E.len[u]; bool rc = H.src[h] & 1;
        const uint8_t* s = E.packed.data() + E.off[u]; uint8_t* d = epacked.data() + eoff[h];
        if (!rc) memcpy(d, s, (len + 3) / 4);
        else for (uint32_t i = 0; i < len; ++i) d[i >> 2] |= (uint8_t)((base_at(s, len - 1 - i) ^ 3u) << (2 * (i & 3)));
    }
}

}  // namespace snh
