// sn_dfside.cuh -- the two per-read passes DF runs over the ReadPaths right after the hot path:
//   * ReadPathVecX, the compressed paths DF keeps in memory and writes as a.pathsX
//     (InitializePathsXFromPaths, 10X/DfTools.cc:24-78; the record format: 10X/paths/ReadPathParser.cc:17-50,217-229);
//   * MarkDups version 2 (10X/SecretOps.cc:599-774): read PAIRS whose placement (first edge, offset) and whose partner's
//     first five bases agree are duplicates of each other; the pair with the highest quality sum is kept.
// Per-item logic is host+device (tests/hostsim runs it on the CPU against the golden files); the kernels around it are
// one thread per read / per sorted record, HBM-bound scans next to one radix sort (sn_prims.cuh).
#pragma once
#include "sn_path.cuh"

#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; };      // (the CUDA vector type, for the host build of tests/hostsim)
#endif

namespace sn {

// ---- ReadPathX ------------------------------------------------------------------------------------------------------
// bytes of one record: [numEdges u8] and, when there are edges, [offset i16][first edge u32][2 bits per further edge]
SN_HD uint32_t rpx_size(uint32_t n) { return n ? (n - 1 + 3) / 4 + 7 : 1; }                  // ReadPathParser.cc:19
// The record of one path.  The edge count goes through one byte and the offset through 16 bits, as there (:25,31);
// a count whose low byte is 0 leaves the rest of the record zero (:28-30).  Every further edge is written as its
// index in From(ToRight(previous edge)) (:39-48); an edge that is not there is skipped without advancing.
SN_HD void rpx_encode(uint8_t* out, const int32_t* e, uint32_t n, int32_t offset, const HbvView& h)
{
    const uint32_t sz = rpx_size(n);
    for (uint32_t i = 0; i < sz; ++i) out[i] = 0;
    out[0] = (uint8_t)n;
    if (out[0] == 0) return;
    const uint16_t o16 = (uint16_t)(int16_t)offset;
    out[1] = (uint8_t)o16; out[2] = (uint8_t)(o16 >> 8);
    const uint32_t e0 = (uint32_t)e[0];
    out[3] = (uint8_t)e0; out[4] = (uint8_t)(e0 >> 8); out[5] = (uint8_t)(e0 >> 16); out[6] = (uint8_t)(e0 >> 24);
    uint32_t idx = 7, sub = 0;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const int32_t w = h.to_right[e[i]];
        const uint32_t a = h.from_start[w], b = h.from_start[w + 1];
        for (uint32_t j = a; j < b; ++j)
            if (h.from_e[j] == e[i + 1]) {
                out[idx] = (uint8_t)(out[idx] + (uint8_t)((uint32_t)(uint8_t)(j - a) << sub));    // LLencodeBranchId :217-229
                sub += 2;
                if (sub > 7) { ++idx; sub = 0; }
                break;
            }
    }
}
// what a reader of the record sees (LLunzip :108-131): edge count and offset after their trip through 8 and 16 bits
SN_HD uint32_t rpx_seen_edges(uint32_t n) { return n & 0xFFu; }
SN_HD int32_t rpx_seen_offset(int32_t offset) { return (int32_t)(int16_t)offset; }

// ---- MarkDups ---------------------------------------------------------------------------------------------------------
#define SN_DUP_HEAD 5u                      // BHEAD (SecretOps.cc:606)
#define SN_DUP_NONE 0xFFFFFFFFu             // key word of an unplaced read (X = (-1,-1,-1,-1), :621; they sort LAST here and are skipped)
// X[id1] = (first edge, offset, the partner's first five bases, id1) as ONE 96-bit radix-sort key, the read id included
// (the order inside a group is part of the result): x = edge, y = offset (16 bits, biased so that it compares like the
// signed value) << 16 | head (10 bits), z = id1.  MarkDups reads the paths out of the ReadPathVecX (:620), hence the 8-bit
// count and the 16-bit offset.  (A partner shorter than five bases reads as A past its end; the reference indexes past
// the end there.)
SN_HD uint4 dup_record(uint32_t n_edges, int32_t e0, int32_t offset, const uint8_t* mate, uint32_t mate_len, uint32_t id)
{
    uint4 r; r.z = id; r.w = 0;
    if (rpx_seen_edges(n_edges) == 0) { r.x = r.y = SN_DUP_NONE; return r; }
    uint32_t head = 0;
    for (uint32_t j = 0; j < SN_DUP_HEAD; ++j) head = head * 4 + (j < mate_len ? packed_base(mate, j) : 0u);
    r.x = (uint32_t)e0; r.y = (((uint32_t)rpx_seen_offset(offset) ^ 0x8000u) & 0xFFFFu) << 16 | head;
    return r;
}
SN_HD bool dup_same_key(const uint4& a, const uint4& b) { return a.x == b.x && a.y == b.y; }

struct DupGroupOut { uint32_t best; bool tie; bool inter; };
// One group [j,k) of the sorted records (k - j > 1): the winner is the first member with the highest quality sum
// (:725-737; the members come in ascending read id, so "lowest id on a tie" is "first"), `tie` as the scan there sets it,
// `inter` = more than one barcode in the group (:654-658; ordinal 0 = no barcode).
template <class BC>
SN_HD DupGroupOut dup_group(const uint4* rec, const uint32_t* qsum, uint32_t j, uint32_t k, BC bc_of)
{
    DupGroupOut o; o.best = j; o.tie = false; o.inter = false;
    uint32_t q = qsum[j];
    int32_t b = bc_of(rec[j].z);
    for (uint32_t l = j + 1; l < k; ++l) {
        if (qsum[l] == q) o.tie = true;
        else if (qsum[l] > q) { q = qsum[l]; o.best = l; }
        const int32_t bl = bc_of(rec[l].z);
        if (b == 0) b = bl; else if (bl != b) o.inter = true;
    }
    return o;
}
// content hash of a read (length, bases, quals): the sort key that brings equal reads of a tie group together
SN_HD uint32_t read_content_hash(const uint8_t* packed, const uint8_t* quals, uint32_t len)
{
    uint64_t hsh = 0xcbf29ce484222325ull ^ len;
    for (uint32_t i = 0; i < len; ++i) { hsh = (hsh ^ (uint64_t)(packed_base(packed, i) | ((uint32_t)quals[i] << 2))) * 0x100000001b3ull; }
    hsh ^= hsh >> 29; hsh *= 0xbf58476d1ce4e5b9ull; hsh ^= hsh >> 32;
    return (uint32_t)hsh;
}
SN_HD bool read_content_equal(const uint8_t* pa, const uint8_t* qa, uint32_t la, const uint8_t* pb, const uint8_t* qb, uint32_t lb)
{
    if (la != lb) return false;
    for (uint32_t i = 0; i < la; ++i) if (packed_base(pa, i) != packed_base(pb, i) || qa[i] != qb[i]) return false;
    return true;
}

#ifdef __CUDACC__
struct ReadsView {
    const uint8_t* bases; const uint64_t* boff; const uint32_t* len;
    const uint8_t* quals; const uint64_t* qoff;          // one byte per base, or
    const uint8_t* pq; const uint64_t* pq_off;           // PQVec stream
    uint64_t n_reads;
    // the quals of read r: in place, or decoded into buf (SN_MAX_READ_LEN bytes)
    __device__ __forceinline__ const uint8_t* quals_of(uint64_t r, uint8_t* buf) const
    {
        if (!pq) return quals + qoff[r];
        pqvec_decode(pq + pq_off[r], pq + pq_off[r + 1], buf, SN_MAX_READ_LEN);
        return buf;
    }
};

static __global__ void __launch_bounds__(256) k_rpx_sizes(const uint64_t* __restrict__ path_off, uint64_t n_reads, uint32_t* __restrict__ sz)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) sz[r] = rpx_size((uint32_t)(path_off[r + 1] - path_off[r]));
}
static __global__ void __launch_bounds__(256) k_rpx_encode(const int32_t* __restrict__ pedges, const uint64_t* __restrict__ path_off, const int32_t* __restrict__ poffset,
                                                    uint64_t n_reads, HbvView h, const uint64_t* __restrict__ zoff, uint8_t* __restrict__ zdata,
                                                    long long* __restrict__ zindex)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    rpx_encode(zdata + zoff[r], pedges + path_off[r], (uint32_t)(path_off[r + 1] - path_off[r]), poffset[r], h);
    if (r % 10 == 0) zindex[r / 10] = (long long)zoff[r];                        // updateZipIndex, skip = 10 (ReadPathVecX.cc:684-691)
}

static __global__ void __launch_bounds__(256) k_md_records(ReadsView rv, const int32_t* __restrict__ pedges, const uint64_t* __restrict__ path_off,
                                                    const int32_t* __restrict__ poffset, uint4* __restrict__ rec)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rv.n_reads) return;
    const uint64_t m = r ^ 1ull;                                                  // the partner (:619)
    const uint32_t n = (uint32_t)(path_off[r + 1] - path_off[r]);
    rec[r] = dup_record(n, n ? pedges[path_off[r]] : 0, poffset[r], rv.bases + rv.boff[m], rv.len[m], (uint32_t)r);
}
// per sorted record: is it in a group of more than one, and if so the quality sum of its pair (:691-704)
static __global__ void __launch_bounds__(128) k_md_members(ReadsView rv, const uint4* __restrict__ rec, uint32_t n, uint32_t* __restrict__ qsum, uint8_t* __restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 me = rec[i];
    uint8_t f = 0; uint32_t q = 0;
    if (me.x != SN_DUP_NONE) {
        const bool prev = i > 0 && dup_same_key(rec[i - 1], me), next = i + 1 < n && dup_same_key(rec[i + 1], me);
        if (!prev) f |= 1;                        // first of its group
        if (prev || next) {
            f |= 2;                               // member of a group of more than one
            uint8_t buf[SN_MAX_READ_LEN];
            for (int s = 0; s < 2; ++s) {
                const uint64_t r = (uint64_t)me.z ^ (uint64_t)s;
                const uint8_t* qs = rv.quals_of(r, buf);
                const uint32_t len = rv.len[r];
                for (uint32_t t = 0; t < len; ++t) q += qs[t];
            }
        }
    }
    qsum[i] = q; flags[i] = f;
}
// per group of more than one: the winner stays, every other member's pair is a duplicate (:753-754)
static __global__ void __launch_bounds__(128) k_md_groups(const uint4* __restrict__ rec, uint32_t n, const uint32_t* __restrict__ qsum, const uint8_t* __restrict__ flags,
                                                   const int32_t* __restrict__ bc, uint8_t* __restrict__ dup, uint32_t* __restrict__ tiegrp,
                                                   unsigned long long* __restrict__ counters /* ndups, interdups */)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || (flags[j] & 3) != 3) return;
    uint32_t k = j + 1;
    while (k < n && !(flags[k] & 1) && rec[k].x != SN_DUP_NONE) ++k;
    const DupGroupOut o = dup_group(rec, qsum, j, k, [&](uint32_t id) -> int32_t { return bc ? bc[id] : 0; });
    for (uint32_t l = j; l < k; ++l) {
        if (l != o.best) dup[rec[l].z >> 1] = 1;
        if (o.tie) tiegrp[l] = j + 1;
    }
    atomicAdd(counters, (unsigned long long)(k - j - 1));
    if (o.inter) atomicAdd(counters + 1, (unsigned long long)(k - j - 1));
}
// members of the groups that saw a tie: key {group, content hash, read id} -- sorted next, equal reads of one group end up
// adjacent, in ascending read id
static __global__ void __launch_bounds__(128) k_md_art_records(ReadsView rv, const uint4* __restrict__ rec, uint32_t n, const uint32_t* __restrict__ tiegrp, uint4* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 o = make_uint4(SN_DUP_NONE, SN_DUP_NONE, rec[i].z, 0u);
    if (tiegrp[i]) {
        uint8_t buf[SN_MAX_READ_LEN];
        const uint64_t r = rec[i].z;
        o.x = tiegrp[i]; o.y = read_content_hash(rv.bases + rv.boff[r], rv.quals_of(r, buf), rv.len[r]);
    }
    out[i] = o;
}
// "artifactual duplicates" (:738-752): the members of a tie group are sorted by (bases, quals, pair) and in every run of
// identical bases AND quals all but the first are flagged -- i.e. a member is flagged iff an identical member with a
// smaller read id exists.  Here: walk back over the members with the same content hash until one is identical.
static __global__ void __launch_bounds__(128) k_md_art(ReadsView rv, const uint4* __restrict__ srt, uint32_t n, uint8_t* __restrict__ art)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 || i >= n) return;
    const uint4 me = srt[i];
    if (me.x == SN_DUP_NONE || !dup_same_key(me, srt[i - 1])) return;
    uint8_t ba[SN_MAX_READ_LEN], bb[SN_MAX_READ_LEN];
    const uint64_t a = me.z;
    const uint8_t* qa = rv.quals_of(a, ba);
    for (uint32_t p = i; p-- > 0 && dup_same_key(srt[p], me);) {
        const uint64_t b = srt[p].z;
        if (read_content_equal(rv.bases + rv.boff[a], qa, rv.len[a], rv.bases + rv.boff[b], rv.quals_of(b, bb), rv.len[b])) { art[a >> 1] = 1; break; }
    }
}
#endif  // __CUDACC__

}  // namespace sn
