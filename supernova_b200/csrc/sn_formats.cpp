#include "sn_formats.h"
#include <zlib.h>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <unistd.h>

namespace snf {
namespace {

struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
    bool open(const std::string& p, const char* mode) { f = fopen(p.c_str(), mode); return f != nullptr; }
};
// Output file whose every write is checked: the bytes go to "<path>.tmp<pid>" and the file only takes its
// name once everything, including fflush and fclose, succeeded -- a full disk or an I/O error leaves no
// truncated a.hbv / tmp.paths behind for the next stage (the reference's BinaryWriter and
// IncrementalWriter abort on a failed write).
class Out {
public:
    explicit Out(const std::string& path) : path_(path), tmp_(path + ".tmp" + std::to_string((long)getpid()))
    {
        f_ = fopen(tmp_.c_str(), "wb");
        if (!f_) fail("cannot create");
    }
    ~Out() { if (f_) { fclose(f_); remove(tmp_.c_str()); } }
    void put(const void* p, size_t bytes)
    {
        if (!f_ || bad_ || !bytes) return;
        if (fwrite(p, 1, bytes, f_) != bytes) fail("write failed");
    }
    template <class T> void pod(const T& v) { put(&v, sizeof v); }
    bool commit(std::string& err)
    {
        if (f_ && !bad_ && fflush(f_) != 0) fail("flush failed");
        if (f_) { if (fclose(f_) != 0 && !bad_) fail("close failed"); f_ = nullptr; }
        if (!bad_ && rename(tmp_.c_str(), path_.c_str()) != 0) fail("rename failed");
        if (bad_) { remove(tmp_.c_str()); err = msg_; return false; }
        return true;
    }
private:
    void fail(const char* what) { if (!bad_) { bad_ = true; msg_ = path_ + ": " + what + " (" + strerror(errno) + ")"; } }
    std::string path_, tmp_, msg_;
    FILE* f_ = nullptr;
    bool bad_ = false;
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
    if ((h.flags & 3) != 1 || h.var_tab < sizeof(FCB) || h.var_tab > d.size() || h.var_tab + 8ull * ((uint64_t)h.n + 1) != h.fixed_off || h.fixed_off > d.size()) {
        err = path + ": not a single-file feudal file"; return false; }
    // the offset table must start at the variable block, never step back, and end where the table begins
    // (the reference checks the same in FeudalControlBlock::isValid and on access)
    const uint64_t n = h.n;
    uint64_t prev = sizeof(FCB);
    for (uint64_t i = 0; i <= n; ++i) {
        uint64_t o; memcpy(&o, d.data() + h.var_tab + 8 * i, 8);
        if (o < prev || o > h.var_tab || (i == 0 && o != sizeof(FCB))) { err = path + ": corrupt feudal offset table (entry " + std::to_string(i) + ")"; return false; }
        prev = o;
    }
    if (prev != h.var_tab) { err = path + ": feudal offset table does not cover the variable data"; return false; }
    return true;
}
bool write_feudal(const std::string& path, uint32_t n, uint8_t size_fixed, uint8_t size_x, uint8_t size_a,
                  const uint8_t* var, uint64_t var_bytes, const uint64_t* rel_off /*n+1, relative to var*/,
                  const uint8_t* fixed, uint64_t fixed_bytes, std::string& err)
{
    Out o(path);
    FCB h; h.n = n; h.flags = 1; h.size_fixed = size_fixed; h.size_x = size_x; h.size_a = size_a;
    h.var_tab = sizeof(FCB) + var_bytes; h.fixed_off = h.var_tab + 8ull * ((uint64_t)n + 1);
    o.pod(h);
    o.put(var, var_bytes);
    std::vector<uint64_t> abs((size_t)n + 1);
    for (uint64_t i = 0; i <= n; ++i) abs[i] = rel_off[i] + sizeof(FCB);
    o.put(abs.data(), 8 * abs.size());
    o.put(fixed, fixed_bytes);
    return o.commit(err);
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
    if (h.size_fixed != 4) { err = path + ": not a fastb (sizeofFixed != 4)"; return false; }
    for (uint64_t i = 0; i < n; ++i)
        if (((uint64_t)out.len[i] + 3) / 4 > out.off[i + 1] - out.off[i]) { err = path + ": read " + std::to_string(i) + " is longer than its packed bytes"; return false; }
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
    // bci[0] = 0, non-decreasing: reads [bci[b], bci[b+1]) carry barcode ordinal b (ParseBarcodedFastqs.cc:284-293)
    if (n && bci[0] != 0) { err = path + ": bci[0] != 0"; return false; }
    for (uint64_t i = 1; i < n; ++i) if (bci[i] < bci[i - 1]) { err = path + ": barcode index is not sorted"; return false; }
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
    Out o(path);
    const uint64_t n = bci.size();
    o.put(MAGIC, 8); o.pod(n); o.put(bci.data(), 8 * n);
    return o.commit(err);
}

// ---- PQVec encoder ------------------------------------------------------------------------------
// The bytes are fixed by PQVecEncoder (feudal/PQVec.cc:17-127): a left-to-right dynamic programme
// over "cheapest encoding of q[0..i]" whose candidate blocks end at i and reach back at most 255
// quals (strict improvement only: of equally cheap blocks the shortest wins), and a running block
// list that is SPLICED, not re-derived, when the winner swallows earlier quals -- a partly swallowed
// block keeps the bit width and base it was chosen with.  Both quirks decide bytes, so they are kept;
// the device encoder (sn_ingest.cuh, k_fasth_pqvec) follows the same two rules.
namespace {
struct PqBlock { uint32_t n, bits, base; };
// cheapest block ending at q[i], given cost[j] = bytes of the best encoding of q[0..j)
PqBlock pq_best_block(const uint8_t* q, uint32_t i, const std::vector<uint32_t>& cost, uint32_t* total)
{
    uint32_t lo = std::min<uint32_t>(63u, q[i]), hi = q[i];
    PqBlock best{1u, ceil_lg2(hi + 1u - lo), lo};
    uint32_t best_total = cost[i] + block_size(1u, best.bits);
    const uint32_t reach = std::min<uint32_t>(i, 254u);               // blocks of at most 255 quals
    for (uint32_t back = 1; back <= reach; ++back) {
        const uint32_t v = q[i - back];
        hi = std::max(hi, v); lo = std::min(lo, v);
        const uint32_t bits = ceil_lg2(hi + 1u - lo);
        const uint32_t t = cost[i - back] + block_size(back + 1u, bits);
        if (t < best_total) { best_total = t; best = PqBlock{back + 1u, bits, lo}; }
    }
    *total = best_total;
    return best;
}
// the list holds blocks covering q[0..i); make it cover q[0..i] ending with `b`
void pq_splice(std::vector<PqBlock>& list, const PqBlock& b)
{
    uint32_t swallow = b.n - 1u;                                       // quals of the list that b takes over
    while (swallow && swallow > list.back().n) { swallow -= list.back().n; list.pop_back(); }
    if (swallow && swallow == list.back().n) list.pop_back();
    else if (swallow) list.back().n -= swallow;                        // (keeps its bits/base: see above)
    list.push_back(b);
}
// [n][bits:3 | base:6 ...][packed (q - base)] (feudal/PQVec.cc:87-127): a 64-bit shift register, low bits first
const uint8_t* pq_emit(const PqBlock& b, const uint8_t* q, std::vector<uint8_t>& out)
{
    out.push_back((uint8_t)b.n);
    uint64_t reg = (uint64_t)b.bits | ((uint64_t)b.base << 3);        // 9 header bits
    out.push_back((uint8_t)reg);
    reg >>= 8;
    uint32_t held = 1;                                                // bits waiting in reg
    if (b.bits) {
        for (uint32_t k = 0; k < b.n; ++k) {
            reg |= (uint64_t)(q[k] - b.base) << held;
            held += b.bits;
            if (held >= 8) { out.push_back((uint8_t)reg); reg >>= 8; held -= 8; }
        }
    }
    if (held) out.push_back((uint8_t)reg);
    return q + b.n;
}
}  // namespace
void pqvec_encode(const uint8_t* q, uint32_t n, std::vector<uint8_t>& out)
{
    std::vector<uint32_t> cost(1, 1u); cost.reserve((size_t)n + 1);   // the terminating zero byte
    std::vector<PqBlock> list;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t t;
        const PqBlock b = pq_best_block(q, i, cost, &t);
        cost.push_back(t);
        pq_splice(list, b);
    }
    for (const PqBlock& b : list) q = pq_emit(b, q, out);
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
bool expand_bci(const std::vector<int64_t>& bci, std::vector<int32_t>& bc)
{
    if (bci.empty()) { bc.clear(); return true; }
    if (bci[0] != 0) return false;
    for (size_t b = 1; b < bci.size(); ++b) if (bci[b] < bci[b - 1]) return false;      // every range lies inside [0, bci.back())
    bc.assign((size_t)bci.back(), -1);
    for (size_t b = 0; b + 1 < bci.size(); ++b)
        for (int64_t j = bci[b]; j < bci[b + 1]; ++j) bc[(size_t)j] = (int32_t)b;
    return true;
}

static void put_basevectors(Out& o, const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n)
{   // vec<basevector>: u64 count, then per sequence u32 bases + ceil(bases/4) bytes (feudal/FieldVec.h:596-598,762)
    o.pod(n);
    std::vector<uint8_t> buf; buf.reserve(1 << 20);
    for (uint64_t e = 0; e < n; ++e) {
        const uint8_t* l = reinterpret_cast<const uint8_t*>(&len[e]);
        buf.insert(buf.end(), l, l + 4);
        buf.insert(buf.end(), packed + off[e], packed + off[e] + (len[e] + 3) / 4);
        if (buf.size() >= (1 << 20)) { o.put(buf.data(), buf.size()); buf.clear(); }
    }
    o.put(buf.data(), buf.size());
}
bool write_bv(const std::string& path, const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n, std::string& err)
{
    Out o(path);
    o.put(MAGIC, 8);
    put_basevectors(o, packed, off, len, n);
    return o.commit(err);
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
static void put_csr(Out& o, uint64_t n, const uint32_t* start, const int32_t* vals)
{   // vec<vec<int>>: u64 count, then per inner vector u64 count + ints (feudal/BinaryStream.h:486-493)
    std::vector<uint8_t> buf; buf.reserve(8 + n * 12 + 4ull * start[n]);
    auto put = [&](const void* p, size_t k) { const uint8_t* b = (const uint8_t*)p; buf.insert(buf.end(), b, b + k); };
    put(&n, 8);
    for (uint64_t v = 0; v < n; ++v) { uint64_t m = start[v + 1] - start[v]; put(&m, 8); if (m) put(vals + start[v], 4 * m); }
    o.put(buf.data(), buf.size());
}
bool write_hbv(const std::string& path, int32_t K, uint64_t n_vert, const uint32_t* from_start, const int32_t* from_v,
               const int32_t* from_e, const uint32_t* to_start, const int32_t* to_e,
               const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n_edges, std::string& err)
{
    Out o(path);
    o.put(MAGIC, 8); o.pod(K);
    put_csr(o, n_vert, from_start, from_v); put_csr(o, n_vert, from_start, from_e); put_csr(o, n_vert, to_start, to_e);
    put_basevectors(o, packed, off, len, n_edges);
    return o.commit(err);
}
static void put_serfvecs(Out& o, uint64_t n, const uint32_t* start, const int32_t* vals)
{   // MasterVec<SerfVec<int>>: u64 count, then per inner vector u32 count + ints (feudal/OuterVec.h:377-379, feudal/SmallVec.h:355-357)
    std::vector<uint8_t> buf; buf.reserve(8 + n * 4 + 4ull * start[n]);
    auto put = [&](const void* p, size_t k) { const uint8_t* b = (const uint8_t*)p; buf.insert(buf.end(), b, b + k); };
    put(&n, 8);
    for (uint64_t v = 0; v < n; ++v) { uint32_t m = start[v + 1] - start[v]; put(&m, 4); if (m) put(vals + start[v], 4ull * m); }
    o.put(buf.data(), buf.size());
}
bool write_hbx(const std::string& path, int32_t K, uint64_t n_vert, const uint32_t* from_start, const int32_t* from_v, const int32_t* from_e,
               const uint32_t* to_start, const int32_t* to_v, const int32_t* to_e,
               const uint8_t* packed, const uint64_t* off, const uint32_t* len, uint64_t n_edges,
               const int32_t* to_left, const int32_t* to_right, std::string& err)
{
    Out o(path);
    o.put(MAGIC, 8); o.pod(K);
    put_serfvecs(o, n_vert, from_start, from_v); put_serfvecs(o, n_vert, to_start, to_v);
    put_serfvecs(o, n_vert, from_start, from_e); put_serfvecs(o, n_vert, to_start, to_e);
    put_basevectors(o, packed, off, len, n_edges);
    o.pod(n_edges); o.put(to_left, 4 * n_edges);
    o.pod(n_edges); o.put(to_right, 4 * n_edges);
    return o.commit(err);
}
bool write_pathsx(const std::string& path, uint64_t n_reads, const int64_t* index, uint64_t n_index, const uint8_t* data, uint64_t n_bytes, std::string& err)
{
    Out o(path);
    const int64_t hdr[5] = {10, 0, (int64_t)n_reads, (int64_t)n_index, (int64_t)n_bytes};     // skip, start_rid, next_start_rid, sizes
    o.put(hdr, sizeof hdr);
    o.put(index, 8 * n_index);
    o.put(data, n_bytes);
    return o.commit(err);
}
bool write_vec_u8(const std::string& path, const uint8_t* v, uint64_t n, std::string& err)
{
    Out o(path);
    o.put(MAGIC, 8); o.pod(n); o.put(v, n);
    return o.commit(err);
}
bool write_text(const std::string& path, const std::string& text, std::string& err)
{
    Out o(path);
    o.put(text.data(), text.size());
    return o.commit(err);
}
bool write_paths(const std::string& path, uint64_t n, const int32_t* offset, const uint64_t* poff, const int32_t* edges, std::string& err)
{
    Out o(path);
    FCB h; h.n = (uint32_t)n; h.flags = 1; h.size_fixed = 0; h.size_x = 24; h.size_a = 4;
    h.var_tab = sizeof(FCB) + 8 * n + 4 * poff[n]; h.fixed_off = h.var_tab + 8 * (n + 1);
    o.pod(h);
    std::vector<uint8_t> buf; buf.reserve(1 << 20);
    for (uint64_t r = 0; r < n; ++r) {
        uint32_t skip = 0; uint64_t m = poff[r + 1] - poff[r];
        size_t at = buf.size(); buf.resize(at + 8 + 4 * m);
        memcpy(&buf[at], &offset[r], 4); memcpy(&buf[at + 4], &skip, 4);
        if (m) memcpy(&buf[at + 8], edges + poff[r], 4 * m);
        if (buf.size() >= (1 << 20)) { o.put(buf.data(), buf.size()); buf.clear(); }
    }
    o.put(buf.data(), buf.size());
    std::vector<uint64_t> tab(n + 1);
    uint64_t pos = sizeof(FCB);
    for (uint64_t r = 0; r <= n; ++r) { tab[r] = pos; if (r < n) pos += 8 + 4 * (poff[r + 1] - poff[r]); }
    o.put(tab.data(), 8 * tab.size());
    return o.commit(err);
}
bool write_vec_int(const std::string& path, const std::vector<int32_t>& v, std::string& err)
{
    Out o(path);
    const uint64_t n = v.size();
    o.put(MAGIC, 8); o.pod(n); o.put(v.data(), 4 * n);
    return o.commit(err);
}

bool write_ulongvecs(const std::string& path, uint64_t n, const uint64_t* ids, const uint64_t* off, std::string& err)
{
    std::vector<uint64_t> rel((size_t)n + 1);
    for (uint64_t i = 0; i <= n; ++i) rel[i] = 8 * off[i];
    return write_feudal(path, (uint32_t)n, 0, 16, 8, reinterpret_cast<const uint8_t*>(ids), 8 * off[n], rel.data(), nullptr, 0, err);
}
bool write_vec_vec_int(const std::string& path, const std::vector<int32_t>& v, std::string& err)
{
    Out o(path);
    const uint64_t one = 1, n = v.size();
    o.put(MAGIC, 8); o.pod(one); o.pod(n); o.put(v.data(), 4 * n);
    return o.commit(err);
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
