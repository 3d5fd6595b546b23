// sn_ingest.cuh -- the barcoded pseudo-FASTQ ("fasth") text of the pipeline -> the reads of a context,
// on the device.  SURVEY.md §8(f) row 2: what 10X/ParseBarcodedFastqs.cc does on one host thread
// (newUnpackBarcodeSortedFastq :56-146 + main :284-303): 9 lines per record
//     @name / R1 / Q1 / R2 / Q2 / BARCODE-gemgroup[,raw] / bcQ / SI / SIQ
// n,N -> A on every line; a record is barcoded iff its barcode line holds a '-' and does not start with
// one; barcode ordinals count the CHANGES of the barcode string (text before the first ',') along the
// barcoded records, starting at 1; the unbarcoded reads come first in the output, then the barcoded ones,
// both in file order; quals are Phred+33, PQVec-encoded (feudal/PQVec.cc:17-127).
// Output = the context's own read layout (.fastb / .qualp variable blocks + offsets, lengths, barcode
// ordinals), byte-identical to the files the reference binary writes (tests/test_gpu_ingest.py).
#pragma once
#include "sn_prims.cuh"

namespace sn {

#define SN_ING_SEG 256u              // bytes of text per thread in the newline passes
#define SN_ING_MAXLEN 256u           // SN_MAX_READ_LEN

// errors (bits of *err)
#define SN_ING_E_NAME 1u             // a record does not start with '@'          ("out of sync reading line")
#define SN_ING_E_QLEN 2u             // qual line and base line differ in length
#define SN_ING_E_BASE 4u             // a base character outside ACGTNacgtn (the reference draws ambiguity codes at random)
#define SN_ING_E_LONG 8u             // read longer than SN_ING_MAXLEN
#define SN_ING_E_QUAL 16u            // a quality character below '!' or above Q63 (PQVecEncoder::init refuses it, feudal/PQVec.cc:30-35)

// ---- lines -----------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_nl_count(const uint8_t* __restrict__ text, uint64_t n, uint32_t* __restrict__ cnt, uint64_t n_seg)
{
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const uint64_t a = s * SN_ING_SEG, b = min(n, a + SN_ING_SEG);
    uint32_t c = 0;
    if (b - a == SN_ING_SEG) {
        const uint4* p = reinterpret_cast<const uint4*>(text + a);
#pragma unroll 4
        for (uint32_t i = 0; i < SN_ING_SEG / 16; ++i) {
            const uint4 v = p[i];
            c += (__popc(__vcmpeq4(v.x, 0x0A0A0A0Au)) + __popc(__vcmpeq4(v.y, 0x0A0A0A0Au)) + __popc(__vcmpeq4(v.z, 0x0A0A0A0Au)) + __popc(__vcmpeq4(v.w, 0x0A0A0A0Au))) >> 3;
        }
    } else for (uint64_t i = a; i < b; ++i) c += text[i] == '\n';
    cnt[s] = c;
}
static __global__ void __launch_bounds__(256) k_nl_fill(const uint8_t* __restrict__ text, uint64_t n, const uint64_t* __restrict__ seg_off, uint64_t n_seg, uint64_t* __restrict__ line_start)
{
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    if (s == 0) line_start[0] = 0;
    const uint64_t a = s * SN_ING_SEG, b = min(n, a + SN_ING_SEG);
    uint64_t k = seg_off[s];
    for (uint64_t i = a; i < b; ++i) if (text[i] == '\n') line_start[++k] = i + 1;
}

// line index of a byte offset that is a line start (file boundaries): one thread per file
static __global__ void k_fasth_file_lines(const uint64_t* __restrict__ line_start, uint64_t n_lines, const uint64_t* __restrict__ file_first_byte, uint32_t n_files, uint64_t* __restrict__ file_first_line)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_files) return;
    const uint64_t b = file_first_byte[f];
    uint64_t lo = 0, hi = n_lines;                       // first line with line_start >= b
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (line_start[mid] < b) lo = mid + 1; else hi = mid; }
    file_first_line[f] = line_start[lo] == b ? lo : ~0ull;
}

__device__ __forceinline__ uint8_t ing_fold(uint8_t c) { return (c == 'n' || c == 'N') ? (uint8_t)'A' : c; }

// ---- records -----------------------------------------------------------------------------------
// thread per record: barcoded?  (hasGemGroup(buf) && buf[0] != '-', ParseBarcodedFastqs.cc:107)
static __global__ void __launch_bounds__(256) k_fasth_flags(const uint8_t* __restrict__ text, const uint64_t* __restrict__ ls, uint64_t n_rec, uint32_t* __restrict__ barcoded, uint32_t* err)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t* L = ls + 9 * r;
    if (text[L[0]] != '@' || L[1] - L[0] < 2) atomicOr(err, SN_ING_E_NAME);
    const uint64_t s = L[5], e = L[6] - 1;            // barcode line without its '\n'
    bool dash = false;
    for (uint64_t i = s; i < e; ++i) dash = dash || text[i] == '-';
    barcoded[r] = (dash && e > s && text[s] != '-') ? 1u : 0u;
}
static __global__ void __launch_bounds__(256) k_fasth_blist(const uint32_t* __restrict__ barcoded, const uint64_t* __restrict__ nbc_before, uint64_t n_rec, uint32_t* __restrict__ blist)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rec && barcoded[r]) blist[nbc_before[r]] = (uint32_t)r;
}
// barcoded record j starts a new barcode iff its key differs from the previous barcoded record's
// (key = the line after n/N -> A, up to the first ','  -- :108-111)
// Several input files (FASTQS={a,b,...}, :262-264): the comparison string starts empty in every file (:67), so the
// first barcoded record of a file always opens a new barcode.  file_first_rec[f] = first record of file f.
static __global__ void __launch_bounds__(256) k_fasth_newbc(const uint8_t* __restrict__ text, const uint64_t* __restrict__ ls, const uint32_t* __restrict__ blist, uint64_t n_bc, uint32_t* __restrict__ isnew,
                                                     const uint32_t* __restrict__ file_first_rec, uint32_t n_files)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_bc) return;
    if (j == 0) { isnew[0] = 1u; return; }
    for (uint32_t f = 1; f < n_files; ++f)
        if (blist[j] >= file_first_rec[f] && blist[j - 1] < file_first_rec[f]) { isnew[j] = 1u; return; }
    const uint64_t a = ls[9ull * blist[j] + 5], ae = ls[9ull * blist[j] + 6] - 1;
    const uint64_t b = ls[9ull * blist[j - 1] + 5], be = ls[9ull * blist[j - 1] + 6] - 1;
    uint64_t i = 0;
    uint32_t differ = 0;
    for (;; ++i) {
        const bool ea = a + i >= ae || text[a + i] == ',', eb = b + i >= be || text[b + i] == ',';
        if (ea || eb) { differ = ea != eb; break; }
        if (ing_fold(text[a + i]) != ing_fold(text[b + i])) { differ = 1; break; }
    }
    isnew[j] = differ;
}
// thread per record: where its two reads go and what they are
static __global__ void __launch_bounds__(256) k_fasth_layout(const uint8_t* __restrict__ text, const uint64_t* __restrict__ ls, uint64_t n_rec, const uint32_t* __restrict__ barcoded,
                                                      const uint64_t* __restrict__ nbc_before, uint64_t n_un_rec, const uint64_t* __restrict__ new_before, const uint32_t* __restrict__ isnew,
                                                      uint32_t* __restrict__ len, int32_t* __restrict__ bc, uint64_t* __restrict__ bpos, uint64_t* __restrict__ qpos,
                                                      uint32_t* __restrict__ nbytes, uint32_t* __restrict__ pqcap, uint32_t* err)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t* L = ls + 9 * r;
    const uint64_t j = nbc_before[r];
    const bool isbc = barcoded[r] != 0;
    const uint64_t out = 2 * (isbc ? n_un_rec + j : r - j);
    const int32_t ord = isbc ? (int32_t)(new_before[j] + isnew[j]) : 0;
    for (int m = 0; m < 2; ++m) {
        const uint64_t bs = L[1 + 2 * m], be = L[2 + 2 * m] - 1, qs = L[2 + 2 * m], qe = L[3 + 2 * m] - 1;
        uint32_t n = (uint32_t)(be - bs);
        if (qe - qs != be - bs) atomicOr(err, SN_ING_E_QLEN);
        if (n > SN_ING_MAXLEN) { atomicOr(err, SN_ING_E_LONG); n = SN_ING_MAXLEN; }
        len[out + m] = n; bc[out + m] = ord; bpos[out + m] = bs; qpos[out + m] = qs;
        nbytes[out + m] = (n + 3) >> 2;
        pqcap[out + m] = n + 8;                          // >= any PQVec of n quals (worst case: 6 bits each + headers)
    }
}
// ---- reads -----------------------------------------------------------------------------------
// thread per read: bases -> fastb packing (4 per byte, LSB first)
static __global__ void __launch_bounds__(128) k_fasth_pack(const uint8_t* __restrict__ text, const uint64_t* __restrict__ bpos, const uint32_t* __restrict__ len, const uint64_t* __restrict__ boff,
                                                    uint64_t n_reads, uint8_t* __restrict__ bases, uint32_t* err)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint8_t* s = text + bpos[r];
    uint8_t* d = bases + boff[r];
    const uint32_t n = len[r];
    bool bad = false;
    for (uint32_t i = 0; i < n; i += 4) {
        uint32_t byte = 0;
        for (uint32_t k = 0; k < 4 && i + k < n; ++k) {
            const uint8_t c = s[i + k] & 0xDFu;                              // upper case
            uint32_t v = 0;
            if (c == 'A' || c == 'N') v = 0; else if (c == 'C') v = 1; else if (c == 'G') v = 2; else if (c == 'T') v = 3; else bad = true;
            byte |= v << (2 * k);
        }
        d[i >> 2] = (uint8_t)byte;
    }
    if (bad) atomicOr(err, SN_ING_E_BASE);
}

__device__ __forceinline__ uint32_t pq_ceil_lg2(uint32_t x) { return x <= 1u ? 0u : 32u - (uint32_t)__clz((int)(x - 1u)); }     // math/PowerOf2.h ceilLg2
__device__ __forceinline__ uint32_t pq_block_size(uint32_t nqs, uint32_t nbits) { return (nqs * nbits + 17u + 7u) >> 3; }          // feudal/PQVec.h:57-58

// thread per read: PQVecEncoder (feudal/PQVec.cc:17-127) -- the dynamic programme over block ends
// (`costs`), the block stack it maintains, then the bit packing.  The encoding goes to the read's slot
// of a scratch array; its size to pqsize.
static __global__ void __launch_bounds__(128) k_fasth_pqvec(const uint8_t* __restrict__ text, const uint64_t* __restrict__ qpos, const uint32_t* __restrict__ len, const uint64_t* __restrict__ slot_off,
                                                     uint64_t n_reads, uint8_t* __restrict__ scratch, uint32_t* __restrict__ pqsize, uint32_t* err)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t n = len[r];
    uint8_t q[SN_ING_MAXLEN];
    uint16_t costs[SN_ING_MAXLEN + 1];
    uint8_t b_nqs[SN_ING_MAXLEN], b_bits[SN_ING_MAXLEN], b_minq[SN_ING_MAXLEN];
    const uint8_t* s = text + qpos[r];
    // n/N -> A on EVERY line of a record, the quality lines included (ParseBarcodedFastqs.cc:84-85), then convertPhred (:37-44):
    // a quality character 'N' (Q45) is read as 'A' (Q32) by the reference
    bool badq = false;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ch = ing_fold(s[i]);
        badq = badq || ch < 33u || ch > 33u + 63u;
        q[i] = (uint8_t)min(63u, ch - 33u);
    }
    if (badq) atomicOr(err, SN_ING_E_QUAL);
    uint32_t nblk = 0;
    costs[0] = 1;
    for (uint32_t i = 0; i < n; ++i) {                                       // PQVecEncoder::init :17-85
        uint32_t minv = min(63u, (uint32_t)q[i]), maxv = q[i];
        uint32_t bits = pq_ceil_lg2(maxv + 1u - minv);
        uint32_t nqs = 1;
        uint32_t best_cost = costs[i] + pq_block_size(nqs, bits);
        uint32_t best_n = 1, best_bits = bits, best_min = minv;
        uint32_t j = i;
        while (j != 0 && nqs < 255u) {
            const uint32_t v = q[--j];
            maxv = max(maxv, v); minv = min(minv, v);
            bits = pq_ceil_lg2(maxv + 1u - minv);
            const uint32_t cur = costs[j] + pq_block_size(++nqs, bits);
            if (cur < best_cost) { best_cost = cur; best_n = nqs; best_bits = bits; best_min = minv; }
        }
        costs[i + 1] = (uint16_t)best_cost;
        uint32_t to_remove = best_n - 1;
        if (!to_remove) { b_nqs[nblk] = (uint8_t)best_n; b_bits[nblk] = (uint8_t)best_bits; b_minq[nblk] = (uint8_t)best_min; ++nblk; }
        else {
            while (to_remove > b_nqs[nblk - 1]) { to_remove -= b_nqs[nblk - 1]; --nblk; }
            if (to_remove == b_nqs[nblk - 1]) { b_nqs[nblk - 1] = (uint8_t)best_n; b_bits[nblk - 1] = (uint8_t)best_bits; b_minq[nblk - 1] = (uint8_t)best_min; }
            else { b_nqs[nblk - 1] -= (uint8_t)to_remove; b_nqs[nblk] = (uint8_t)best_n; b_bits[nblk] = (uint8_t)best_bits; b_minq[nblk] = (uint8_t)best_min; ++nblk; }
        }
    }
    uint8_t* out = scratch + slot_off[r];
    uint32_t o = 0, it = 0;
    for (uint32_t b = 0; b < nblk; ++b) {                                    // PQVecEncoder::encode :87-127
        uint32_t nqs = b_nqs[b];
        const uint32_t nbits = b_bits[b], minq = b_minq[b];
        out[o++] = (uint8_t)nqs;
        uint32_t bits = nbits | (minq << 3);
        out[o++] = (uint8_t)bits;
        bits >>= 8;
        if (!nbits) { out[o++] = (uint8_t)bits; it += nqs; }
        else {
            uint32_t off = 1;
            while (nqs--) {
                const uint32_t val = q[it++] - minq;
                bits |= val << off;
                if ((off += nbits) >= 8) { out[o++] = (uint8_t)bits; off -= 8; bits >>= 8; }
            }
            if (off) out[o++] = (uint8_t)bits;
        }
    }
    out[o++] = 0;
    pqsize[r] = o;
}
static __global__ void __launch_bounds__(256) k_fasth_pq_compact(const uint8_t* __restrict__ scratch, const uint64_t* __restrict__ slot_off, const uint32_t* __restrict__ pqsize,
                                                          const uint64_t* __restrict__ pqoff, uint64_t n_reads, uint8_t* __restrict__ pq)
{
    // a warp per read
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_reads) return;
    const uint8_t* s = scratch + slot_off[r];
    uint8_t* d = pq + pqoff[r];
    for (uint32_t i = threadIdx.x & 31u; i < pqsize[r]; i += 32) d[i] = s[i];
}

}  // namespace sn
