#include "sn_hbv.h"
#include <algorithm>
#include <cstring>
#include <numeric>
#include <stdexcept>

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
        for (uint64_t i = 0; i < nE;) {
            uint64_t j = i + 1;
            while (j < nE && same_prefix(order[i], order[j])) ++j;
            if (j - i > 1) std::sort(order.begin() + i, order.begin() + j, bvcomp);
            i = j;
        }
    }
    for (uint64_t i = 0; i < nE; ++i) rank[order[i]] = (uint32_t)i;
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
    // inside a vertex: EEComp order (:113-121) = edge rank, rc, pos (pos 0 < pos len-(K-1))
    for (int32_t g = 0; g < nV; ++g) {
        uint32_t* it = &gitems[gstart[g]]; int n = (int)(gstart[g + 1] - gstart[g]);
        if (n > 8) throw std::runtime_error("HBV: a vertex has more than 8 edge ends (HBVFromEdges.cc:83)");
        for (int i = 1; i < n; ++i) {
            uint32_t x = it[i]; uint64_t kx = ((uint64_t)rank[x >> 2] << 2) | (x & 3);
            int j = i - 1;
            while (j >= 0 && (((uint64_t)rank[it[j] >> 2] << 2) | (it[j] & 3)) > kx) { it[j + 1] = it[j]; --j; }
            it[j + 1] = x;
        }
    }
    // HBVBuilder::add / processQueue (:189-228) with digraphE::AddEdge (graph/DigraphTemplate.h:2572-2582)
    std::vector<int32_t> vid(nV, -1);
    std::vector<uint8_t> fcnt(nV, 0), tcnt(nV, 0);
    std::vector<int32_t> fv(4 * (size_t)nV), fe(4 * (size_t)nV), tv(4 * (size_t)nV), te(4 * (size_t)nV);
    auto add_sorted = [](int32_t* v, int32_t* eo, uint8_t& n, int32_t w, int32_t e) {
        if (n >= 4) throw std::runtime_error("HBV: more than 4 edges on one side of a vertex");
        int i = 0; while (i < n && v[i] <= w) ++i;                    // upper_bound
        for (int j = n; j > i; --j) { v[j] = v[j - 1]; eo[j] = eo[j - 1]; }
        v[i] = w; eo[i] = e; ++n;
    };
    auto done = [&](uint32_t e, uint32_t rc) { return (rc ? H.rev : H.fwd)[e] != -1; };
    H.src.reserve(2 * nE); H.to_left.reserve(2 * nE); H.to_right.reserve(2 * nE);
    std::vector<uint32_t> Q; Q.reserve(1024);
    int32_t nextV = 0;
    for (uint32_t pass = 0; pass < 2; ++pass)
        for (uint64_t oi = 0; oi < nE; ++oi) {
            uint32_t e0 = order[oi];
            if (done(e0, pass)) continue;
            Q.clear(); Q.push_back((e0 << 1) | pass);
            for (size_t qh = 0; qh < Q.size(); ++qh) {
                uint32_t e = Q[qh] >> 1, rc = Q[qh] & 1;
                if (done(e, rc)) continue;
                // (a palindromic edge is only ever processed with rc=0: its rev id is set with its fwd id)
                int32_t g1 = end_group[4 * (size_t)e + 2 * rc + 0], g2 = end_group[4 * (size_t)e + 2 * rc + 1];
                if (g1 < 0 || g2 < 0) throw std::runtime_error("HBV: edge end without a vertex");
                if (vid[g1] == -1) vid[g1] = nextV++;
                if (vid[g2] == -1) vid[g2] = nextV++;
                int32_t v1 = vid[g1], v2 = vid[g2];
                int32_t id = (int32_t)H.src.size();
                H.src.push_back((e << 1) | rc);
                add_sorted(&fv[4 * (size_t)v1], &fe[4 * (size_t)v1], fcnt[v1], v2, id);
                add_sorted(&tv[4 * (size_t)v2], &te[4 * (size_t)v2], tcnt[v2], v1, id);
                H.to_left.push_back(v1); H.to_right.push_back(v2);
                if (!rc || pal[e]) H.fwd[e] = id;
                if (rc || pal[e]) H.rev[e] = id;
                for (int32_t g : {g1, g2})
                    for (uint32_t x = gstart[g]; x < gstart[g + 1]; ++x) {
                        uint32_t it = gitems[x];
                        if (!done(it >> 2, (it >> 1) & 1)) Q.push_back(((it >> 2) << 1) | ((it >> 1) & 1));
                    }
            }
        }
    if (nextV != nV) throw std::runtime_error("HBV: vertex numbering did not reach every vertex");
    H.n_vert = nV;
    H.from_start.assign(nV + 1, 0); H.to_start.assign(nV + 1, 0);
    for (int32_t v = 0; v < nV; ++v) { H.from_start[v + 1] = H.from_start[v] + fcnt[v]; H.to_start[v + 1] = H.to_start[v] + tcnt[v]; }
    const size_t nH = H.src.size();
    H.from_v.resize(nH); H.from_e.resize(nH); H.to_v.resize(nH); H.to_e.resize(nH);
    for (int32_t v = 0; v < nV; ++v) {
        for (int i = 0; i < fcnt[v]; ++i) { H.from_v[H.from_start[v] + i] = fv[4 * (size_t)v + i]; H.from_e[H.from_start[v] + i] = fe[4 * (size_t)v + i]; }
        for (int i = 0; i < tcnt[v]; ++i) { H.to_v[H.to_start[v] + i] = tv[4 * (size_t)v + i]; H.to_e[H.to_start[v] + i] = te[4 * (size_t)v + i]; }
    }
    // Involution: the reverse complement of HBV edge fwd[e] is rev[e]
    H.inv.assign(nH, -1);
    for (uint64_t e = 0; e < nE; ++e) { H.inv[H.fwd[e]] = H.rev[e]; H.inv[H.rev[e]] = H.fwd[e]; }
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
