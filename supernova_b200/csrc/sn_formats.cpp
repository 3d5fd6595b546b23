#include "sn_formats.h"
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <algorithm>

namespace snf {
namespace {

struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
    bool open(const std::string& p, const char* mode) { f = fopen(p.c_str(), mode); return f != nullptr; }
};
bool slurp(const std::string& path, std::vector<uint8_t>& d, std::string& err)
{
    File fh;
    if (!fh.open(path, "rb")) { err = "cannot open " + path; return false; }
    fseek(fh.f, 0, SEEK_END); long n = ftell(fh.f); fseek(fh.f, 0, SEEK_SET);
    d.resize((size_t)n);
    if (n && fread(d.data(), 1, (size_t)n, fh.f) != (size_t)n) { err = "short read " + path; return false; }
    return true;
}
#pragma pack(push, 1)
struct FCB { uint32_t n; uint8_t flags, size_fixed, size_x, size_a; uint64_t var_tab, fixed_off; };
#pragma pack(pop)
static_assert(sizeof(FCB) == 24, "FeudalControlBlock is 24 bytes");

bool read_feudal(const std::string& path, std::vector<uint8_t>& d, FCB& h, std::string& err)
{
    if (!slurp(path, d, err)) return false;
    if (d.size() < sizeof(FCB)) { err = path + ": too short for a feudal file"; return false; }
    memcpy(&h, d.data(), sizeof h);
    if ((h.flags & 3) != 1 || h.var_tab + 8ull * ((uint64_t)h.n + 1) != h.fixed_off || h.fixed_off > d.size()) {
        err = path + ": not a single-file feudal file"; return false; }
    return true;
}
bool write_feudal(const std::string& path, uint32_t n, uint8_t size_fixed, uint8_t size_x, uint8_t size_a,
                  const uint8_t* var, uint64_t var_bytes, const uint64_t* rel_off /*n+1, relative to var*/,
                  const uint8_t* fixed, uint64_t fixed_bytes, std::string& err)
{
    File fh;
    if (!fh.open(path, "wb")) { err = "cannot create " + path; return false; }
    FCB h; h.n = n; h.flags = 1; h.size_fixed = size_fixed; h.size_x = size_x; h.size_a = size_a;
    h.var_tab = sizeof(FCB) + var_bytes; h.fixed_off = h.var_tab + 8ull * ((uint64_t)n + 1);
    fwrite(&h, sizeof h, 1, fh.f);
    if (var_bytes) fwrite(var, 1, var_bytes, fh.f);
    std::vector<uint64_t> abs((size_t)n + 1);
    for (uint64_t i = 0; i <= n; ++i) abs[i] = rel_off[i] + sizeof(FCB);
    fwrite(abs.data(), 8, abs.size(), fh.f);
    if (fixed_bytes) fwrite(fixed, 1, fixed_bytes, fh.f);
    return true;
}
const char MAGIC[9] = "BINWRITE";
unsigned ceil_lg2(unsigned x)                                          // math/PowerOf2.h ceilLg2, x in 1..64
{
    static const unsigned char T[65] = {0,0,1,2,2,3,3,3,3,4,4,4,4,4,4,4,4,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
                                        6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6,6};
    if (x <= 64) return T[x];
    unsigned b = 0; while ((1u << b) < x) ++b; return b;
}
unsigned block_size(unsigned nqs, unsigned nbits) { return (nqs * nbits + 17 + 7) >> 3; }   // feudal/PQVec.h:57-58
}  // namespace

bool read_fastb(const std::string& path, Fastb& out, std::string& err)
{
    std::vector<uint8_t> d; FCB h;
    if (!read_feudal(path, d, h, err)) return false;
    uint64_t n = h.n;
    if (h.fixed_off + 4 * n > d.size()) { err = path + ": truncated fixed data"; return false; }
    out.var.assign(d.begin() + sizeof(FCB), d.begin() + h.var_tab);
    out.off.resize(n + 1); out.len.resize(n);
    memcpy(out.off.data(), d.data() + h.var_tab, 8 * (n + 1));
    for (auto& o : out.off) o -= sizeof(FCB);
    memcpy(out.len.data(), d.data() + h.fixed_off, 4 * n);
    return true;
}
bool read_qualp(const std::string& path, Qualp& out, std::string& err)
{
    std::vector<uint8_t> d; FCB h;
    if (!read_feudal(path, d, h, err)) return false;
    uint64_t n = h.n;
    out.var.assign(d.begin() + sizeof(FCB), d.begin() + h.var_tab);
    out.off.resize(n + 1);
    memcpy(out.off.data(), d.data() + h.var_tab, 8 * (n + 1));
    for (auto& o : out.off) o -= sizeof(FCB);
    return true;
}
bool read_bci(const std::string& path, std::vector<int64_t>& bci, std::string& err)
{
    std::vector<uint8_t> d;
    if (!slurp(path, d, err)) return false;
    if (d.size() < 16 || memcmp(d.data(), MAGIC, 8)) { err = path + ": not a BINWRITE file"; return false; }
    uint64_t n; memcpy(&n, d.data() + 8, 8);
    if (16 + 8 * n > d.size()) { err = path + ": truncated"; return false; }
    bci.resize(n); memcpy(bci.data(), d.data() + 16, 8 * n);
    return true;
}
bool write_fastb(const std::string& path, const Fastb& in, std::string& err)
{   // FCB for BaseVec files: sizeofFixed 4, sizeofX 16, sizeofA 1
    return write_feudal(path, (uint32_t)in.len.size(), 4, 16, 1, in.var.data(), in.var.size(), in.off.data(),
                        reinterpret_cast<const uint8_t*>(in.len.data()), 4 * in.len.size(), err);
}
bool write_qualp(const std::string& path, const Qualp& in, std::string& err)
{   // FCB for PQVec files: sizeofFixed 0, sizeofX 8, sizeofA 1
    return write_feudal(path, (uint32_t)(in.off.size() - 1), 0, 8, 1, in.var.data(), in.var.size(), in.off.data(), nullptr, 0, err);
}
bool write_bci(const std::string& path, const std::vector<int64_t>& bci, std::string& err)
{
    File fh;
    if (!fh.open(path, "wb")) { err = "cannot create " + path; return false; }
    uint64_t n = bci.size();
    fwrite(MAGIC, 1, 8, fh.f); fwrite(&n, 8, 1, fh.f); fwrite(bci.data(), 8, n, fh.f);
    return true;
}

void pqvec_encode(const uint8_t* q, uint32_t n, std::vector<uint8_t>& out)
{
    struct Block { unsigned nqs, bits, minq; };
    std::vector<Block> blocks;
    std::vector<unsigned> costs; costs.reserve(n + 1); costs.push_back(1);
    for (uint32_t i = 0; i < n; ++i) {                       // PQVecEncoder::init :17-85
        unsigned minv = std::min(63u, (unsigned)q[i]), maxv = q[i];
        unsigned bits = ceil_lg2(maxv + 1u - minv);
        unsigned nqs = 1;
        unsigned best_cost = costs[i] + block_size(nqs, bits);
        Block best{1, bits, minv};
        uint32_t j = i;
        while (j != 0 && nqs < 255) {
            unsigned v = q[--j];
            if (v > maxv) maxv = v;
            if (v < minv) minv = v;
            bits = ceil_lg2(maxv + 1u - minv);
            unsigned cur = costs[j] + block_size(++nqs, bits);
            if (cur < best_cost) { best_cost = cur; best = Block{nqs, bits, minv}; }
        }
        costs.push_back(best_cost);
        unsigned to_remove = best.nqs - 1;
        if (!to_remove) blocks.push_back(best);
        else {
            while (to_remove > blocks.back().nqs) { to_remove -= blocks.back().nqs; blocks.pop_back(); }
            if (to_remove == blocks.back().nqs) blocks.back() = best;
            else { blocks.back().nqs -= to_remove; blocks.push_back(best); }
        }
    }
    const uint8_t* it = q;                                   // PQVecEncoder::encode :87-127
    for (const Block& b : blocks) {
        uint64_t nqs = b.nqs, nbits = b.bits, minq = b.minq;
        out.push_back((uint8_t)nqs);
        uint64_t bits = nbits | (minq << 3);
        out.push_back((uint8_t)bits);
        bits >>= 8;
        if (!nbits) { out.push_back((uint8_t)bits); it += nqs; }
        else {
            uint64_t off = 1;
            while (nqs--) {
                uint64_t val = *it++ - minq;
                bits |= val << off;
                if ((off += nbits) >= 8) { out.push_back((uint8_t)bits); off -= 8; bits >>= 8; }
            }
            if (off) out.push_back((uint8_t)bits);
        }
    }
    out.push_back(0);
}
uint32_t pqvec_decode(const uint8_t* p, const uint8_t* pend, uint8_t* out, uint32_t cap)
{
    uint32_t i = 0;
    while (p < pend) {
        uint32_t nq = *p++;
        if (!nq) break;
        uint32_t b0 = *p++;
        uint32_t nbits = b0 & 7u, minq = b0 >> 3;
        uint64_t acc = *p++;
        minq |= (uint32_t)(acc & 1u) << 5; acc >>= 1;
        uint32_t have = 7, mask = (1u << nbits) - 1u;
        for (uint32_t k = 0; k < nq; ++k) {
            uint32_t q = minq;
            if (nbits) {
                if (have < nbits) { acc |= (uint64_t)(*p++) << have; have += 8; }
                q += (uint32_t)acc & mask; acc >>= nbits; have -= nbits;
            }
            if (i < cap) out[i] = (uint8_t)q;
            ++i;
        }
    }
    return i;
}
void expand_bci(const std::vector<int64_t>& bci, std::vector<int32_t>& bc)
{
    bc.assign(bci.empty() ? 0 : (size_t)bci.back(), -1);
    for (size_t b = 0; b + 1 < bci.size(); ++b)
        for (int64_t j = bci[b]; j < bci[b + 1]; ++j) bc[(size_t)j] = (int32_t)b;
}

bool write_bv(const std::string& path, const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n, std::string& err)
{
    File fh;
    if (!fh.open(path, "wb")) { err = "cannot create " + path; return false; }
    fwrite(MAGIC, 1, 8, fh.f); fwrite(&n, 8, 1, fh.f);
    for (uint64_t e = 0; e < n; ++e) { fwrite(&len[e], 4, 1, fh.f); fwrite(packed + off[e], 1, (len[e] + 3) / 4, fh.f); }
    return true;
}
bool read_bv(const std::string& path, Fastb& out, std::string& err)
{
    std::vector<uint8_t> d;
    if (!slurp(path, d, err)) return false;
    if (d.size() < 16 || memcmp(d.data(), MAGIC, 8)) { err = path + ": not a BINWRITE file"; return false; }
    uint64_t n; memcpy(&n, d.data() + 8, 8);
    size_t p = 16;
    out.var.clear(); out.off.assign(1, 0); out.len.clear();
    for (uint64_t e = 0; e < n; ++e) {
        if (p + 4 > d.size()) { err = path + ": truncated"; return false; }
        uint32_t l; memcpy(&l, d.data() + p, 4); p += 4;
        size_t nb = (l + 3) / 4;
        if (p + nb > d.size()) { err = path + ": truncated"; return false; }
        out.var.insert(out.var.end(), d.begin() + p, d.begin() + p + nb); p += nb;
        out.len.push_back(l); out.off.push_back(out.var.size());
    }
    return true;
}
static void put_csr(FILE* f, uint64_t n, const uint32_t* start, const int32_t* vals)
{   // vec<vec<int>>: u64 count, then per inner vector u64 count + ints (feudal/BinaryStream.h:486-493)
    std::vector<uint8_t> buf; buf.reserve(8 + n * 12 + 4ull * start[n]);
    auto put = [&](const void* p, size_t k) { const uint8_t* b = (const uint8_t*)p; buf.insert(buf.end(), b, b + k); };
    put(&n, 8);
    for (uint64_t v = 0; v < n; ++v) { uint64_t m = start[v + 1] - start[v]; put(&m, 8); if (m) put(vals + start[v], 4 * m); }
    fwrite(buf.data(), 1, buf.size(), f);
}
bool write_hbv(const std::string& path, int32_t K, uint64_t n_vert, const uint32_t* from_start, const int32_t* from_v,
               const int32_t* from_e, const uint32_t* to_start, const int32_t* to_e,
               const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n_edges, std::string& err)
{
    File fh;
    if (!fh.open(path, "wb")) { err = "cannot create " + path; return false; }
    fwrite(MAGIC, 1, 8, fh.f); fwrite(&K, 4, 1, fh.f);
    put_csr(fh.f, n_vert, from_start, from_v); put_csr(fh.f, n_vert, from_start, from_e); put_csr(fh.f, n_vert, to_start, to_e);
    fwrite(&n_edges, 8, 1, fh.f);
    for (uint64_t e = 0; e < n_edges; ++e) { fwrite(&len[e], 4, 1, fh.f); fwrite(packed + off[e], 1, (len[e] + 3) / 4, fh.f); }
    return true;
}
bool write_paths(const std::string& path, uint64_t n, const int32_t* offset, const uint64_t* poff, const int32_t* edges, std::string& err)
{
    File fh;
    if (!fh.open(path, "wb")) { err = "cannot create " + path; return false; }
    FCB h; h.n = (uint32_t)n; h.flags = 1; h.size_fixed = 0; h.size_x = 24; h.size_a = 4;
    h.var_tab = sizeof(FCB) + 8 * n + 4 * poff[n]; h.fixed_off = h.var_tab + 8 * (n + 1);
    fwrite(&h, sizeof h, 1, fh.f);
    std::vector<uint8_t> buf; buf.reserve(1 << 20);
    for (uint64_t r = 0; r < n; ++r) {
        uint32_t skip = 0; uint64_t m = poff[r + 1] - poff[r];
        size_t at = buf.size(); buf.resize(at + 8 + 4 * m);
        memcpy(&buf[at], &offset[r], 4); memcpy(&buf[at + 4], &skip, 4);
        if (m) memcpy(&buf[at + 8], edges + poff[r], 4 * m);
        if (buf.size() >= (1 << 20)) { fwrite(buf.data(), 1, buf.size(), fh.f); buf.clear(); }
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), fh.f);
    std::vector<uint64_t> tab(n + 1);
    uint64_t pos = sizeof(FCB);
    for (uint64_t r = 0; r <= n; ++r) { tab[r] = pos; if (r < n) pos += 8 + 4 * (poff[r + 1] - poff[r]); }
    fwrite(tab.data(), 8, tab.size(), fh.f);
    return true;
}
bool write_vec_int(const std::string& path, const std::vector<int32_t>& v, std::string& err)
{
    File fh;
    if (!fh.open(path, "wb")) { err = "cannot create " + path; return false; }
    uint64_t n = v.size();
    fwrite(MAGIC, 1, 8, fh.f); fwrite(&n, 8, 1, fh.f); if (n) fwrite(v.data(), 4, n, fh.f);
    return true;
}

bool write_ulongvecs(const std::string& path, uint64_t n, const uint64_t* ids, const uint64_t* off, std::string& err)
{
    std::vector<uint64_t> rel((size_t)n + 1);
    for (uint64_t i = 0; i <= n; ++i) rel[i] = 8 * off[i];
    return write_feudal(path, (uint32_t)n, 0, 16, 8, reinterpret_cast<const uint8_t*>(ids), 8 * off[n], rel.data(), nullptr, 0, err);
}
bool write_vec_vec_int(const std::string& path, const std::vector<int32_t>& v, std::string& err)
{
    File fh;
    if (!fh.open(path, "wb")) { err = "cannot create " + path; return false; }
    uint64_t one = 1, n = v.size();
    fwrite(MAGIC, 1, 8, fh.f); fwrite(&one, 8, 1, fh.f); fwrite(&n, 8, 1, fh.f); if (n) fwrite(v.data(), 4, n, fh.f);
    return true;
}
bool read_text_maybe_gz(const std::string& path, std::vector<char>& out, std::string& err)
{
    gzFile f = gzopen(path.c_str(), "rb");          // reads plain files transparently
    if (!f) { err = "cannot open " + path; return false; }
    gzbuffer(f, 1 << 20);
    out.clear();
    std::vector<char> buf(1 << 24);
    for (;;) {
        int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) { err = path + ": gzip read error"; gzclose(f); return false; }
        if (n == 0) break;
        out.insert(out.end(), buf.begin(), buf.begin() + n);
    }
    gzclose(f);
    return true;
}

}  // namespace snf
