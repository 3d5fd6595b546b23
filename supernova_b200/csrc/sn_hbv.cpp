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

void build_hbv(const Edges& E, const HbvPre& pre, Hbv& H)
{
    const uint64_t nE = E.n();
    H = Hbv();
    H.fwd.assign(nE, -1); H.rev.assign(nE, -1);
    H.from_start.assign(1, 0); H.to_start.assign(1, 0);
    if (!nE) return;
    const uint8_t* P = E.packed.data();
    const bool timing = getenv("SN_HBV_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    auto lap = [&](const char* what) { if (timing) { double t = now(); fprintf(stderr, "[hbv] %-10s %.2f ms\n", what, t - t0); t0 = t; } };
    // BVComp (HBVFromEdges.cc:106-111): longer first, then lexicographic on bases
    auto bvcomp = [&](uint32_t a, uint32_t b) {
        if (E.len[a] != E.len[b]) return E.len[a] > E.len[b];
        int c = cmp_packed(P + E.off[a], P + E.off[b], E.len[a]);
        return c ? c < 0 : a < b;
    };
    std::vector<uint32_t> order(nE), rank(nE);
    if (pre.empty()) {
        std::iota(order.begin(), order.end(), 0u);
        std::sort(order.begin(), order.end(), bvcomp);
    } else {
        // device pre-order by (length desc, first 32 bases): only runs of equal prefix need the full comparison
        order = pre.order;
        auto same_prefix = [&](uint32_t a, uint32_t b) {
            return E.len[a] == E.len[b] && load64(P + E.off[a]) == load64(P + E.off[b]);
        };
        std::vector<std::pair<uint64_t, uint64_t>> ties;
        for (uint64_t i = 0; i < nE;) {
            uint64_t j = i + 1;
            while (j < nE && same_prefix(order[i], order[j])) ++j;
            if (j - i > 1) ties.emplace_back(i, j);
            i = j;
        }
        for (auto& t : ties) std::sort(order.begin() + t.first, order.begin() + t.second, bvcomp);
    }
    for (uint64_t i = 0; i < nE; ++i) rank[order[i]] = (uint32_t)i;
    lap("order");
    // VertexDictBuilder (:124-168): 4 ends per edge (2 for a palindromic edge); a vertex is a
    // distinct (K-1)-mer.  item = edge<<2 | rc<<1 | distal.
    std::vector<uint8_t> pal;
    std::vector<int32_t> end_group;
    std::vector<uint32_t> gstart, gitems;
    if (!pre.empty()) { pal = pre.pal; end_group = pre.end_group; gstart = pre.group_start; gitems = pre.group_items; }
    else {
        pal.resize(nE); end_group.assign(4 * nE, -1);
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
    lap("groups");
    // inside a vertex: EEComp order (:113-121) = edge rank, rc, pos (pos 0 < pos len-(K-1)).
    // One 64-byte record per vertex for the numbering loop: id, item count, items (edge<<1|rc).
    struct alignas(64) GroupRec { int32_t vid; uint32_t n; uint32_t items[8]; uint32_t pad[6]; };
    std::vector<GroupRec> groups(nV);
    bool too_many = false;
    parallel_ranges((uint64_t)nV, [&](uint64_t g0, uint64_t g1) {
        for (uint64_t g = g0; g < g1; ++g) {
            uint32_t* it = &gitems[gstart[g]]; int n = (int)(gstart[g + 1] - gstart[g]);
            if (n > 8) { too_many = true; n = 8; }
            for (int i = 1; i < n; ++i) {
                uint32_t x = it[i]; uint64_t kx = ((uint64_t)rank[x >> 2] << 2) | (x & 3);
                int j = i - 1;
                while (j >= 0 && (((uint64_t)rank[it[j] >> 2] << 2) | (it[j] & 3)) > kx) { it[j + 1] = it[j]; --j; }
                it[j + 1] = x;
            }
            GroupRec& r = groups[g]; r.vid = -1; r.n = (uint32_t)n;
            for (int i = 0; i < n; ++i) r.items[i] = it[i] >> 1;
        }
    });
    if (too_many) throw std::runtime_error("HBV: a vertex has more than 8 edge ends (HBVFromEdges.cc:83)");
    lap("groupsort");
    // HBVBuilder::add / processQueue (:189-228).  The loop only assigns ids; the sorted
    // adjacency lists of digraphE::AddEdge are rebuilt afterwards from to_left/to_right.
    // per (edge, rc): its two vertex groups and its HBV id, one 16-byte record
    struct ERec { int32_t g1, g2, id; uint32_t pal; };
    std::vector<ERec> er(2 * nE);
    parallel_ranges(nE, [&](uint64_t e0, uint64_t e1) {
        for (uint64_t e = e0; e < e1; ++e)
            for (uint32_t rc = 0; rc < 2; ++rc) {
                ERec& r = er[2 * e + rc];
                r.g1 = end_group[4 * e + 2 * rc + 0]; r.g2 = end_group[4 * e + 2 * rc + 1]; r.id = -1; r.pal = pal[e];
            }
    });
    H.src.resize(2 * nE); H.to_left.resize(2 * nE); H.to_right.resize(2 * nE);
    std::vector<uint32_t> Q(2 * nE + 16);
    int32_t nextV = 0; uint32_t nH = 0;
    for (uint32_t pass = 0; pass < 2; ++pass)
        for (uint64_t oi = 0; oi < nE; ++oi) {
            const uint32_t e0 = order[oi];
            if (er[2 * (size_t)e0 + pass].id != -1) continue;
            size_t qh = 0, qt = 0;
            Q[qt++] = (e0 << 1) | pass;
            er[2 * (size_t)e0 + pass].id = -2;
            // An item is queued at most once (id -2 = queued): the reference queues duplicates and
            // skips them when popped, so an item is processed at its FIRST position either way.
            while (qh < qt) {
                const uint32_t it = Q[qh++];
                ERec& r = er[it];
                if (r.id != -2) continue;                           // palindrome twin already numbered
                // (a palindromic edge is only ever processed with rc=0: its rev id is set with its fwd id)
                if (r.g1 < 0 || r.g2 < 0) throw std::runtime_error("HBV: edge end without a vertex");
                GroupRec& G1 = groups[r.g1]; GroupRec& G2 = groups[r.g2];
                if (G1.vid == -1) G1.vid = nextV++;
                if (G2.vid == -1) G2.vid = nextV++;
                const int32_t id = (int32_t)nH++;
                H.src[id] = it; H.to_left[id] = G1.vid; H.to_right[id] = G2.vid;
                r.id = id;
                if (r.pal) er[it ^ 1u].id = id;
                if (qt + 16 > Q.size()) {                           // keep the FIFO compact
                    std::copy(Q.begin() + qh, Q.begin() + qt, Q.begin()); qt -= qh; qh = 0;
                    if (qt + 16 > Q.size()) Q.resize(2 * Q.size());
                }
                // the loop is bound by cache misses on the vertex records: fetch them when an item is queued
                for (uint32_t x = 0; x < G1.n; ++x) { const uint32_t t2 = G1.items[x]; ERec& q = er[t2];
                    if (q.id == -1) { q.id = -2; Q[qt++] = t2; __builtin_prefetch(&groups[q.g1]); __builtin_prefetch(&groups[q.g2]); } }
                for (uint32_t x = 0; x < G2.n; ++x) { const uint32_t t2 = G2.items[x]; ERec& q = er[t2];
                    if (q.id == -1) { q.id = -2; Q[qt++] = t2; __builtin_prefetch(&groups[q.g1]); __builtin_prefetch(&groups[q.g2]); } }
                if (qh + 4 < qt) {                                   // and the item records of what those vertices will push
                    const ERec& nx = er[Q[qh + 4]];
                    const GroupRec& A = groups[nx.g1]; const GroupRec& B = groups[nx.g2];
                    for (uint32_t x = 0; x < A.n; ++x) __builtin_prefetch(&er[A.items[x]]);
                    for (uint32_t x = 0; x < B.n; ++x) __builtin_prefetch(&er[B.items[x]]);
                }
            }
        }
    if (nextV != nV) throw std::runtime_error("HBV: vertex numbering did not reach every vertex");
    lap("bfs");
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
      parallel_ranges((uint64_t)nV, [&](uint64_t v0, uint64_t v1) {
        for (uint64_t v = v0; v < v1; ++v) {
            uint32_t s = start[v], n = start[v + 1] - s;
            if (n > 4) over4 = true;
            for (uint32_t i = 1; i < n; ++i) {                       // stable insertion sort by neighbour
                int32_t w = nb[s + i], e = eo[s + i]; uint32_t j = i;
                while (j > 0 && nb[s + j - 1] > w) { nb[s + j] = nb[s + j - 1]; eo[s + j] = eo[s + j - 1]; --j; }
                nb[s + j] = w; eo[s + j] = e;
            }
        }
      });
    };
    order_lists(H.from_start, H.from_v, H.from_e);
    order_lists(H.to_start, H.to_v, H.to_e);
    if (over4) throw std::runtime_error("HBV: more than 4 edges on one side of a vertex");
    // Involution: the reverse complement of HBV edge fwd[e] is rev[e]
    H.inv.assign(nH, -1);
    for (uint64_t e = 0; e < nE; ++e) { H.inv[H.fwd[e]] = H.rev[e]; H.inv[H.rev[e]] = H.fwd[e]; }
    lap("csr+inv");
}

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
