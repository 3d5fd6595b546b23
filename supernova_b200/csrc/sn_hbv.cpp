#include "sn_hbv.h"
#include <algorithm>
#include <cstring>
#include <deque>
#include <numeric>
#include <stdexcept>

namespace snh {
namespace {

constexpr int K = 48;
inline uint32_t base_at(const uint8_t* p, uint64_t i) { return (p[i >> 2] >> (2 * (i & 3))) & 3u; }

struct Sub { uint64_t hi, lo; };        // (K-1)-mer, 2 bits per base, MSB first, left aligned in 128 bits
inline bool operator<(const Sub& a, const Sub& b) { return a.hi != b.hi ? a.hi < b.hi : a.lo < b.lo; }
inline bool operator==(const Sub& a, const Sub& b) { return a.hi == b.hi && a.lo == b.lo; }

// EdgeEnd(pBV, rc, distal, K-1) (HBVFromEdges.cc:28-56; SwitchHitterIter feudal/BaseVec.h:98-126)
Sub edge_end(const uint8_t* s, uint32_t len, bool rc, bool distal)
{
    uint32_t pos = distal ? len - (K - 1) : 0;
    Sub r{0, 0};
    for (int i = 0; i < K - 1; ++i) {
        uint64_t c = rc ? (base_at(s, len - 1 - (pos + i)) ^ 3u) : base_at(s, pos + i);
        if (i < 32) r.hi |= c << (2 * (31 - i)); else r.lo |= c << (2 * (63 - i));
    }
    return r;
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
struct End { Sub key; uint32_t rank, edge, pos; uint8_t rc; };

}  // namespace

void build_hbv(const Edges& E, Hbv& H)
{
    const uint64_t nE = E.n();
    H = Hbv();
    H.fwd.assign(nE, -1); H.rev.assign(nE, -1);
    if (!nE) return;
    // BVComp (HBVFromEdges.cc:106-111): longer first, then lexicographic on bases
    std::vector<uint32_t> order(nE), rank(nE);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (E.len[a] != E.len[b]) return E.len[a] > E.len[b];
        const uint8_t* x = E.packed.data() + E.off[a]; const uint8_t* y = E.packed.data() + E.off[b];
        for (uint32_t i = 0; i < E.len[a]; ++i) { uint32_t p = base_at(x, i), q = base_at(y, i); if (p != q) return p < q; }
        return a < b;
    });
    for (uint64_t i = 0; i < nE; ++i) rank[order[i]] = (uint32_t)i;
    // VertexDictBuilder::map (:141-148): 4 ends per edge, 2 for a palindromic edge
    std::vector<End> ends; ends.reserve(4 * nE);
    std::vector<uint8_t> pal(nE);
    for (uint32_t e = 0; e < nE; ++e) {
        const uint8_t* s = E.packed.data() + E.off[e]; uint32_t len = E.len[e];
        pal[e] = seq_form(s, len) == 2;
        for (int rc = 0; rc < (pal[e] ? 1 : 2); ++rc)
            for (int distal = 0; distal < 2; ++distal)
                ends.push_back(End{edge_end(s, len, rc, distal), rank[e], e, distal ? len - (K - 1) : 0u, (uint8_t)rc});
    }
    // group by (K-1)-mer; inside a vertex EEComp order (:113-121): edge rank, rc, pos
    std::sort(ends.begin(), ends.end(), [](const End& a, const End& b) {
        if (!(a.key == b.key)) return a.key < b.key;
        if (a.rank != b.rank) return a.rank < b.rank;
        if (a.rc != b.rc) return a.rc < b.rc;
        return a.pos < b.pos;
    });
    struct Vert { Sub key; int32_t id; uint32_t beg, end; };
    std::vector<Vert> V;
    for (size_t i = 0; i < ends.size();) {
        size_t j = i; while (j < ends.size() && ends[j].key == ends[i].key) ++j;
        if (j - i > 8) throw std::runtime_error("HBV: a vertex has more than 8 edge ends (HBVFromEdges.cc:83)");
        V.push_back(Vert{ends[i].key, -1, (uint32_t)i, (uint32_t)j});
        i = j;
    }
    const int32_t nV = (int32_t)V.size();
    H.from.resize(nV); H.from_eo.resize(nV); H.to.resize(nV); H.to_eo.resize(nV);
    auto find_vert = [&](const Sub& k) -> Vert& {
        auto it = std::lower_bound(V.begin(), V.end(), k, [](const Vert& v, const Sub& s) { return v.key < s; });
        if (it == V.end() || !(it->key == k)) throw std::runtime_error("HBV: vertex lookup failed");
        return *it;
    };
    auto done = [&](uint32_t e, int rc) { return (rc ? H.rev : H.fwd)[e] != -1; };
    // digraphE::AddEdge (graph/DigraphTemplate.h:2572-2582): insert at upper_bound
    auto add_sorted = [](std::vector<int32_t>& v, std::vector<int32_t>& eo, int32_t w, int32_t e) {
        size_t i = std::upper_bound(v.begin(), v.end(), w) - v.begin();
        v.insert(v.begin() + i, w); eo.insert(eo.begin() + i, e);
    };
    // HBVBuilder::add / processQueue (:189-228)
    std::deque<uint64_t> Q;
    int32_t nextV = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (uint64_t oi = 0; oi < nE; ++oi) {
            uint32_t e0 = order[oi];
            if (done(e0, pass)) continue;
            Q.push_back(((uint64_t)e0 << 1) | (uint64_t)pass);
            while (!Q.empty()) {
                uint64_t it = Q.front(); Q.pop_front();
                uint32_t e = (uint32_t)(it >> 1); int rc = (int)(it & 1);
                if (done(e, rc)) continue;
                const uint8_t* s = E.packed.data() + E.off[e]; uint32_t len = E.len[e];
                Vert& v1 = find_vert(edge_end(s, len, rc, false));
                if (v1.id == -1) v1.id = nextV++;
                Vert& v2 = find_vert(edge_end(s, len, rc, true));
                if (v2.id == -1) v2.id = nextV++;
                int32_t id = (int32_t)H.src.size();
                H.src.push_back((e << 1) | (uint32_t)rc);
                add_sorted(H.from[v1.id], H.from_eo[v1.id], v2.id, id);
                add_sorted(H.to[v2.id], H.to_eo[v2.id], v1.id, id);
                H.to_left.push_back(v1.id); H.to_right.push_back(v2.id);
                if (!rc || pal[e]) H.fwd[e] = id;
                if (rc || pal[e]) H.rev[e] = id;
                for (const Vert* pv : {&v1, &v2})
                    for (uint32_t x = pv->beg; x < pv->end; ++x)
                        if (!done(ends[x].edge, ends[x].rc)) Q.push_back(((uint64_t)ends[x].edge << 1) | ends[x].rc);
            }
        }
    if (nextV != nV) throw std::runtime_error("HBV: vertex numbering did not reach every vertex");
    // edges_ : the oriented sequences
    const uint64_t nH = H.src.size();
    H.eoff.assign(nH + 1, 0); H.elen.resize(nH);
    for (uint64_t h = 0; h < nH; ++h) { H.elen[h] = E.len[H.src[h] >> 1]; H.eoff[h + 1] = H.eoff[h] + (H.elen[h] + 3) / 4; }
    H.epacked.assign(H.eoff[nH] + 16, 0);
    for (uint64_t h = 0; h < nH; ++h) {
        uint32_t u = H.src[h] >> 1, len = E.len[u]; bool rc = H.src[h] & 1;
        const uint8_t* s = E.packed.data() + E.off[u]; uint8_t* d = H.epacked.data() + H.eoff[h];
        if (!rc) memcpy(d, s, (len + 3) / 4);
        else for (uint32_t i = 0; i < len; ++i) d[i >> 2] |= (uint8_t)((base_at(s, len - 1 - i) ^ 3u) << (2 * (i & 3)));
    }
    // Involution: the reverse complement of HBV edge fwd[e] is rev[e]
    H.inv.assign(nH, -1);
    for (uint64_t e = 0; e < nE; ++e) { H.inv[H.fwd[e]] = H.rev[e]; H.inv[H.rev[e]] = H.fwd[e]; }
}

}  // namespace snh
