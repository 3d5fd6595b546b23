// sn_synth.cuh -- the synthetic linked reads of SURVEY.md §8(d), generated ON the device.
//
// Counter-based: every value is a pure function of (seed, what, index), so a rank generates its slice of the pairs
// with no state and no genome in memory, and the numpy twin (supernova_b200/synth.py: make_reads_cb) produces the same
// reads bit for bit (tests/test_synth_cb.py on the CPU through tests/hostsim, tests/test_gpu_synth.py on the device).
// This is the generator for workloads the host one cannot produce in reasonable time (BASELINE configs 3-5: 22.5 Gbp
// per GPU); bench.py's C2 keeps the numpy/Philox generator its golden digests were made with.
//
//   mix64         SplitMix64 step;  h(stream, a, b) = mix64(mix64(mix64(seed ^ stream * C) ^ a) ^ b)
//   genome        haplotype A: base i = 2 bits of h(1, i / 32, 0); haplotype B = A with one SNP per 1000-base window w at
//                 w * 1000 + (s mod 2^32) mod 1000, base + 1 + (s >> 32) mod 3, s = h(2, w, 0)
//   pair p        u = h(3, p, 0): haplotype = bit 0, strand flip = bit 1, insert = 300 + ((u >> 8) mod 2^24) mod 200,
//                 start = h(3, p, 1) mod (G - insert); R1 = first 150 bases of the fragment, R2 = first 150 of its reverse
//                 complement (flip swaps the two)
//   errors        e = h(4, read, j): substitution iff (e mod 2^24) < T[j], T[j] = floor(2^24 * (0.001 + 0.02 (j/150)^3))
//                 (integer table from the host), new base + 1 + ((e >> 24) mod 256) mod 3, qual {2,12,20}[((e >> 32) mod 256)
//                 mod 3]; otherwise Q30 with 5 % ((e >> 40) mod 2^16 mod 100 < 5), else Q37
//   barcode       ordinal 1 + floor(p * n_barcodes / total_pairs): the reads come sorted by barcode, as the pipeline's do
#pragma once
#include "sn_kmer.cuh"

#define SN_SYN_L 150u

namespace sn {

SN_HD uint64_t syn_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
SN_HD uint64_t syn_key(uint64_t seed, uint64_t stream) { return syn_mix64(seed ^ (stream * 0xD6E8FEB86659FD93ull)); }
SN_HD uint64_t syn_h(uint64_t key, uint64_t a, uint64_t b) { return syn_mix64(syn_mix64(key ^ a) ^ b); }

struct SynSpec { uint64_t genome_bases, total_pairs, seed; uint32_t n_barcodes; };

// base i of haplotype `hap` (0 = A, 1 = B)
SN_HD uint32_t syn_hap_base(uint64_t kg, uint64_t ks, uint32_t hap, uint64_t i)
{
    uint32_t b = (uint32_t)(syn_h(kg, i >> 5, 0) >> (2 * (i & 31))) & 3u;
    if (hap) {
        const uint64_t w = i / 1000, s = syn_h(ks, w, 0);
        if (i == w * 1000 + (s & 0xFFFFFFFFull) % 1000) b = (b + 1u + (uint32_t)((s >> 32) % 3)) & 3u;
    }
    return b;
}
// read `m` (0/1) of pair p: bases (codes 0..3) and quals, SN_SYN_L each
SN_HD void syn_read(const SynSpec& sp, const uint32_t* T, uint64_t p, uint32_t m, uint8_t* bases, uint8_t* quals)
{
    const uint64_t kg = syn_key(sp.seed, 1), ks = syn_key(sp.seed, 2), kp = syn_key(sp.seed, 3), ke = syn_key(sp.seed, 4);
    const uint64_t u = syn_h(kp, p, 0);
    const uint32_t hap = (uint32_t)(u & 1), flip = (uint32_t)((u >> 1) & 1);
    const uint64_t insert = 300 + ((u >> 8) & 0xFFFFFFull) % 200;
    const uint64_t start = syn_h(kp, p, 1) % (sp.genome_bases - insert);
    const bool forward = (m == 0) != (flip != 0);
    const uint64_t kr = syn_mix64(ke ^ (2 * p + m));
    const uint8_t eq[3] = {2, 12, 20};
    for (uint32_t j = 0; j < SN_SYN_L; ++j) {
        uint32_t b = forward ? syn_hap_base(kg, ks, hap, start + j) : 3u - syn_hap_base(kg, ks, hap, start + insert - 1 - j);
        const uint64_t e = syn_mix64(kr ^ j);
        uint32_t q;
        if ((uint32_t)(e & 0xFFFFFFull) < T[j]) { b = (b + 1u + (uint32_t)((e >> 24) & 0xFF) % 3u) & 3u; q = eq[((e >> 32) & 0xFF) % 3]; }
        else q = (((e >> 40) & 0xFFFF) % 100 < 5) ? 30u : 37u;
        bases[j] = (uint8_t)b; quals[j] = (uint8_t)q;
    }
}
SN_HD int32_t syn_barcode(const SynSpec& sp, uint64_t p) { return (int32_t)(1 + (p * sp.n_barcodes) / sp.total_pairs); }

#ifdef __CUDACC__
// thread per read of the chunk [r0, r0 + n): packed bases (fastb layout, 38 bytes per read), length, barcode ordinal, and the
// quals as Phred+33 text for the PQVec encoder of the ingest path (k_fasth_pqvec)
static __global__ void __launch_bounds__(128) k_synth_reads(SynSpec sp, const uint32_t* __restrict__ T, uint64_t first_pair, uint64_t r0, uint64_t n,
                                                     uint8_t* __restrict__ bases, uint64_t* __restrict__ boff, uint32_t* __restrict__ len, int32_t* __restrict__ bc,
                                                     uint8_t* __restrict__ qtext, uint64_t* __restrict__ qpos, uint64_t* __restrict__ slot_off)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t r = r0 + t, p = first_pair + (r >> 1);
    uint8_t b[SN_SYN_L], q[SN_SYN_L];
    syn_read(sp, T, p, (uint32_t)(r & 1), b, q);
    constexpr uint32_t NB = (SN_SYN_L + 3) / 4;
    uint8_t* out = bases + r * NB;
    for (uint32_t k = 0; k < NB; ++k) {
        uint32_t v = 0;
        for (uint32_t x = 0; x < 4 && 4 * k + x < SN_SYN_L; ++x) v |= (uint32_t)b[4 * k + x] << (2 * x);
        out[k] = (uint8_t)v;
    }
    boff[r] = r * NB; len[r] = SN_SYN_L; bc[r] = syn_barcode(sp, p);
    uint8_t* qt = qtext + t * SN_SYN_L;
    for (uint32_t j = 0; j < SN_SYN_L; ++j) qt[j] = (uint8_t)(q[j] + 33u);
    qpos[t] = t * SN_SYN_L; slot_off[t] = t * (uint64_t)(SN_SYN_L + 8);
}
#endif

}  // namespace sn
