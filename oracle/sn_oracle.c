/*
 * sn_oracle.c -- TEST INFRASTRUCTURE.  NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's k-mer count ->
 * unipath edges -> HyperBasevector -> ReadPath hot path
 * (10XGenomics/supernova lib/assembly, buildReadQGraph48).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the
 * product library (supernova_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every stage of
 * this file byte-for-byte against outputs of the reference's own C++ compiled from
 * /root/reference (oracle/_ref/OracleProbe: kmers.kvec, a.hbv, tmp.paths) and the
 * outputs for the committed fixtures live in tests/golden/.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/lib/assembly/src).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define K 48
#define KW 3 /* u32 words per k-mer: kmers/KMer.h:344-350 (16 bases/word, MSB first) */

typedef struct { uint32_t w[KW]; } kmer_t;

/* ------------------------------------------------------------------------- */
/* k-mer primitives: kmers/KMer.h                                             */
/* ------------------------------------------------------------------------- */

/* KMer::assign kmers/KMer.h:154-160 : base i at bits 2*(15-i%16) of word i/16 */
static kmer_t kmer_from_codes(const uint8_t* b)
{
    kmer_t k;
    for (int w = 0; w < KW; ++w) {
        uint32_t v = 0;
        for (int i = 0; i < 16; ++i) v = (v << 2) | (b[16 * w + i] & 3u);
        k.w[w] = v;
    }
    return k;
}
static inline unsigned kmer_base(const kmer_t* k, int i)
{ return (k->w[i >> 4] >> (2 * (15 - (i & 15)))) & 3u; }

/* KMer::toSuccessor kmers/KMer.h:189-201 (K=48: no unused trailing bits) */
static inline void kmer_to_succ(kmer_t* k, unsigned code)
{
    k->w[0] = (k->w[0] << 2) | (k->w[1] >> 30);
    k->w[1] = (k->w[1] << 2) | (k->w[2] >> 30);
    k->w[2] = (k->w[2] << 2) | (code & 3u);
}
/* KMer::toPredecessor kmers/KMer.h:174-187 */
static inline void kmer_to_pred(kmer_t* k, unsigned code)
{
    k->w[2] = (k->w[2] >> 2) | (k->w[1] << 30);
    k->w[1] = (k->w[1] >> 2) | (k->w[0] << 30);
    k->w[0] = (k->w[0] >> 2) | ((code & 3u) << 30);
}
static inline uint32_t rc_word(uint32_t x)
{   /* reverse the order of the 16 2-bit fields and complement */
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
    return ~x;
}
/* KMer::rc kmers/KMer.h:203-225 */
static inline kmer_t kmer_rc(const kmer_t* k)
{
    kmer_t r;
    r.w[0] = rc_word(k->w[2]); r.w[1] = rc_word(k->w[1]); r.w[2] = rc_word(k->w[0]);
    return r;
}
/* compare(KMer,KMer) kmers/KMer.h:299-306 */
static inline int kmer_cmp(const kmer_t* a, const kmer_t* b)
{
    for (int i = 0; i < KW; ++i)
        if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}
/* CF<K>::getForm dna/CanonicalForm.h:57-67 (K even): 0=FWD 1=REV 2=PALINDROME.
 * Comparing base i with the complement of base K-1-i outside-in is the
 * lexicographic comparison of the k-mer with its reverse complement. */
enum { FWD = 0, REV = 1, PAL = 2 };
static inline int kmer_form(const kmer_t* k)
{
    kmer_t r = kmer_rc(k);
    int c = kmer_cmp(k, &r);
    return c < 0 ? FWD : (c > 0 ? REV : PAL);
}
/* getCanonicalForm(beg,end) dna/CanonicalForm.h:34-46 for run-time length */
static int seq_form(const uint8_t* s, size_t len)
{
    if (len & 1) return (s[len / 2] & 2) ? REV : FWD;
    size_t i = 0, j = len;
    while (i != j) {
        unsigned f = s[i], r = s[--j] ^ 3u;
        if (f < r) return FWD;
        if (r < f) return REV;
        ++i;
    }
    return PAL;
}
static void seq_rc_inplace(uint8_t* s, size_t len)
{
    for (size_t i = 0, j = len; i < j; ) {
        --j;
        uint8_t a = s[i] ^ 3u, b = s[j] ^ 3u;
        s[i] = b; s[j] = a;
        ++i;
    }
}

/* KMerContext: kmers/KMerContext.h:23-121 ; rc = bit reversal of the byte
 * (kmers/KMerContext.cc:18-36). pred mask in the high nibble, succ in the low. */
static inline uint8_t ctx_rc(uint8_t c)
{
    c = (uint8_t)(((c >> 1) & 0x55) | ((c & 0x55) << 1));
    c = (uint8_t)(((c >> 2) & 0x33) | ((c & 0x33) << 2));
    return (uint8_t)((c >> 4) | (c << 4));
}
static const uint8_t SIDE_COUNT[16] = {0,1,1,2,1,2,2,3,1,2,2,3,2,3,3,4};
static const uint8_t BITS2VAL[16]   = {4,0,1,4,2,4,4,4,3,4,4,4,4,4,4,4};
#define CTX_PRED(c) ((c) >> 4)
#define CTX_SUCC(c) ((c) & 0xF)

/* ------------------------------------------------------------------------- */
/* public structures                                                          */
/* ------------------------------------------------------------------------- */
typedef struct {
    kmer_t   kmer;     /* canonical */
    uint32_t count;    /* saturating 2^24-1: kmers/ReadPather.h:128-129,145 */
    uint8_t  ctx0;     /* context before recomputeAdjacencies (== kmers.kvec) */
    uint8_t  ctx;      /* context after recomputeAdjacencies */
    uint32_t edge;     /* ~0u = null */
    uint32_t off;
} kent_t;

typedef struct {
    /* inputs */
    uint64_t n_reads;
    const uint8_t*  bases;   /* base codes 0..3, ragged */
    const uint8_t*  quals;
    const uint64_t* off;     /* n_reads+1 */
    const int32_t*  bc;      /* per-read barcode ordinal (10X/DF.cc:464-469) or NULL */
    unsigned min_qual, min_freq, min_bc;
    int64_t ign_bc_below;
    int count_len_k;          /* tada variant (lib/tada/src/cmd_msp.rs:109-110): a read trimmed to exactly K bases gives its one k-mer */
    /* stage outputs */
    uint32_t* good_len;
    uint64_t  n_occ;         /* k-mer occurrences emitted by Kmerizer::map */
    uint64_t  n_kmers; kent_t* kmers;         /* sorted by k-mer */
    uint64_t  n_edges; uint8_t** edge_seq; uint32_t* edge_len;   /* canonical unipaths */
    /* HBV */
    int32_t   n_vert; int32_t n_hbv_edges;
    int32_t** from;  int32_t** from_eo; int32_t* from_n;
    int32_t** to;    int32_t** to_eo;   int32_t* to_n;
    uint8_t** hbv_seq; uint32_t* hbv_len;
    int32_t*  fwd_xlat; int32_t* rev_xlat;
    int32_t*  to_left;  int32_t* to_right;
    /* paths */
    int32_t*  path_offset; uint64_t* path_off; int32_t* path_edges; /* CSR over reads */
} sn_oracle_t;

/* ------------------------------------------------------------------------- */
/* a1: GoodLenTailFinder  paths/long/BuildReadQGraph48.cc:65-89               */
/* ------------------------------------------------------------------------- */
static uint32_t good_len(const uint8_t* q, uint32_t n, unsigned min_qual)
{
    unsigned good = 0;
    uint32_t i = n;
    while (i != 0) {
        --i;
        if (q[i] < min_qual) good = 0;
        else if (++good == K) return i + K;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a2: Kmerizer::map  BuildReadQGraph48.cc:155-172                            */
/* ------------------------------------------------------------------------- */
typedef struct { kmer_t kmer; int32_t bc; uint8_t ctx; } occ_t;

static inline void emit_occ(occ_t* o, const kmer_t* k, uint8_t ctx, int32_t bc)
{
    if (kmer_form(k) == REV) { o->kmer = kmer_rc(k); o->ctx = ctx_rc(ctx); }
    else { o->kmer = *k; o->ctx = ctx; }
    o->bc = bc;
}
static uint64_t kmerize_read(const uint8_t* b, uint32_t len, int32_t bc, occ_t* out, int count_len_k)
{
    if (count_len_k && len == K) {                                /* msp_read, cmd_msp.rs:109-110: one k-mer, no neighbour on either side */
        kmer_t one = kmer_from_codes(b);
        emit_occ(&out[0], &one, 0, bc);
        return 1;
    }
    if (len < K + 1) return 0;
    uint64_t n = 0;
    kmer_t kkk = kmer_from_codes(b);
    uint32_t itr = K, last = len - 1;
    emit_occ(&out[n++], &kkk, (uint8_t)(1u << b[itr]), bc);       /* initialContext */
    while (itr != last) {
        unsigned pred = kmer_base(&kkk, 0);
        kmer_to_succ(&kkk, b[itr]); ++itr;
        emit_occ(&out[n++], &kkk, (uint8_t)((1u << pred) << 4 | (1u << b[itr])), bc);
    }
    {
        unsigned pred = kmer_base(&kkk, 0);
        kmer_to_succ(&kkk, b[last]);
        emit_occ(&out[n++], &kkk, (uint8_t)((1u << pred) << 4), bc);  /* finalContext */
    }
    return n;
}
static int occ_cmp(const void* a, const void* b)
{
    const occ_t* x = (const occ_t*)a; const occ_t* y = (const occ_t*)b;
    int c = kmer_cmp(&x->kmer, &y->kmer);
    if (c) return c;
    return (x->bc > y->bc) - (x->bc < y->bc);
}

/* lookup in the sorted table == KmerDict::findEntryCanonical */
static kent_t* find_canonical(sn_oracle_t* o, const kmer_t* k)
{
    uint64_t lo = 0, hi = o->n_kmers;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        int c = kmer_cmp(&o->kmers[mid].kmer, k);
        if (c == 0) return &o->kmers[mid];
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return NULL;
}
/* KmerDict::findEntry kmers/ReadPather.h:238-241 */
static kent_t* find_entry(sn_oracle_t* o, const kmer_t* k)
{
    if (kmer_form(k) == REV) { kmer_t r = kmer_rc(k); return find_canonical(o, &r); }
    return find_canonical(o, k);
}

/* ------------------------------------------------------------------------- */
/* a4/a5: MapReduceEngine sort+group, Kmerizer::reduce, summarizeEntries,      */
/* areIgnoredBarcodes, areEnoughBarcodes  BuildReadQGraph48.cc:91-137,174-181  */
/* ------------------------------------------------------------------------- */
int sn_oracle_count(sn_oracle_t* o)
{
    uint64_t n = o->n_reads;
    o->good_len = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint64_t tot = 0;
    for (uint64_t r = 0; r < n; ++r) {
        uint32_t len = (uint32_t)(o->off[r + 1] - o->off[r]);
        uint32_t g = good_len(o->quals + o->off[r], len, o->min_qual);
        o->good_len[r] = g;
        if (g >= K + 1 || (o->count_len_k && g == K)) tot += g - K + 1;
    }
    o->n_occ = tot;
    occ_t* occ = (occ_t*)malloc(sizeof(occ_t) * (tot ? tot : 1));
    uint64_t p = 0;
    for (uint64_t r = 0; r < n; ++r) {
        int32_t bc = -1;                                        /* :158-159 */
        if ((int64_t)r >= o->ign_bc_below && o->bc) bc = o->bc[r];
        p += kmerize_read(o->bases + o->off[r], o->good_len[r], bc, occ + p, o->count_len_k);
    }
    qsort(occ, tot, sizeof(occ_t), occ_cmp);
    /* group */
    uint64_t cap = 1024, nk = 0;
    kent_t* tab = (kent_t*)malloc(sizeof(kent_t) * cap);
    for (uint64_t i = 0; i < tot; ) {
        uint64_t j = i;
        uint64_t count = 0; uint8_t ctx = 0; int ignored = 0;
        unsigned distinct = 0; int32_t lastbc = 0;   /* sorted by bc within group */
        while (j < tot && kmer_cmp(&occ[j].kmer, &occ[i].kmer) == 0) {
            count += 1;                                          /* max(getCount(),1) : :100 */
            ctx |= occ[j].ctx;
            if (occ[j].bc == -1) ignored = 1;                    /* :110-112 */
            if (occ[j].bc > 0 && (distinct == 0 || occ[j].bc != lastbc)) { ++distinct; lastbc = occ[j].bc; }
            ++j;
        }
        if (count > 0xFFFFFFu) count = 0xFFFFFFu;                 /* KDef::setCount */
        int bc_test = 1;
        if (o->bc) bc_test = ignored || distinct >= o->min_bc;   /* :176-178 */
        if (count >= o->min_freq && bc_test) {
            if (nk == cap) { cap *= 2; tab = (kent_t*)realloc(tab, sizeof(kent_t) * cap); }
            kent_t* e = &tab[nk++];
            e->kmer = occ[i].kmer; e->count = (uint32_t)count; e->ctx0 = ctx; e->ctx = ctx;
            e->edge = ~0u; e->off = 0;
        }
        i = j;
    }
    free(occ);
    o->kmers = tab; o->n_kmers = nk;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a6: KmerDict::recomputeAdjacencies  kmers/ReadPather.h:346-385             */
/* ------------------------------------------------------------------------- */
int sn_oracle_prune(sn_oracle_t* o)
{
    for (uint64_t i = 0; i < o->n_kmers; ++i) {
        kent_t* e = &o->kmers[i];
        uint8_t ctx = e->ctx0;
        for (unsigned c = 0; c < 4; ++c) if (CTX_SUCC(ctx) & (1u << c)) {
            kmer_t k = e->kmer; kmer_to_succ(&k, c);
            if (!find_entry(o, &k)) ctx &= (uint8_t)~(1u << c);
        }
        for (unsigned c = 0; c < 4; ++c) if (CTX_PRED(ctx) & (1u << c)) {
            kmer_t k = e->kmer; kmer_to_pred(&k, c);
            if (!find_entry(o, &k)) ctx &= (uint8_t)~((1u << c) << 4);
        }
        e->ctx = ctx;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a7: EdgeBuilder  BuildReadQGraph48.cc:327-541                              */
/* ------------------------------------------------------------------------- */
typedef struct {
    sn_oracle_t* o;
    uint8_t* seq; size_t seq_n, seq_cap;
    kent_t** ents; size_t ent_n, ent_cap;
    size_t edge_cap;
} ebuild_t;

static void eb_push_base(ebuild_t* b, uint8_t c)
{
    if (b->seq_n == b->seq_cap) { b->seq_cap = b->seq_cap ? 2 * b->seq_cap : 256; b->seq = (uint8_t*)realloc(b->seq, b->seq_cap); }
    b->seq[b->seq_n++] = c;
}
static void eb_push_ent(ebuild_t* b, kent_t* e)
{
    if (b->ent_n == b->ent_cap) { b->ent_cap = b->ent_cap ? 2 * b->ent_cap : 256; b->ents = (kent_t**)realloc(b->ents, sizeof(kent_t*) * b->ent_cap); }
    b->ents[b->ent_n++] = e;
}
static void eb_assign_kmer(ebuild_t* b, const kmer_t* k)
{
    b->seq_n = 0;
    for (int i = 0; i < K; ++i) eb_push_base(b, (uint8_t)kmer_base(k, i));
}
/* EdgeBuilder::lookup :466-476 : context in the orientation of `k` */
static kent_t* eb_lookup(ebuild_t* b, const kmer_t* k, uint8_t* ctx)
{
    kent_t* r;
    if (kmer_form(k) == REV) {
        kmer_t rc = kmer_rc(k);
        r = find_canonical(b->o, &rc);
        if (!r) { fprintf(stderr, "oracle: lookup failed (rev)\n"); abort(); }
        *ctx = ctx_rc(r->ctx);
    } else {
        r = find_canonical(b->o, k);
        if (!r) { fprintf(stderr, "oracle: lookup failed\n"); abort(); }
        *ctx = r->ctx;
    }
    return r;
}
/* EdgeBuilder::addEdge :478-506 */
static void eb_add_edge(ebuild_t* b)
{
    sn_oracle_t* o = b->o;
    if (seq_form(b->seq, b->seq_n) == REV) {
        seq_rc_inplace(b->seq, b->seq_n);
        for (size_t i = 0, j = b->ent_n; i + 1 < j; ++i) { --j; kent_t* t = b->ents[i]; b->ents[i] = b->ents[j]; b->ents[j] = t; }
    }
    if (o->n_edges == b->edge_cap) {
        b->edge_cap = b->edge_cap ? 2 * b->edge_cap : 1024;
        o->edge_seq = (uint8_t**)realloc(o->edge_seq, sizeof(uint8_t*) * b->edge_cap);
        o->edge_len = (uint32_t*)realloc(o->edge_len, sizeof(uint32_t) * b->edge_cap);
    }
    uint32_t id = (uint32_t)o->n_edges++;
    o->edge_seq[id] = (uint8_t*)malloc(b->seq_n);
    memcpy(o->edge_seq[id], b->seq, b->seq_n);
    o->edge_len[id] = (uint32_t)b->seq_n;
    for (size_t i = 0; i < b->ent_n; ++i) {
        kent_t* e = b->ents[i];
        if (e->edge != ~0u) { fprintf(stderr, "oracle: preoccupied kmer\n"); abort(); }
        e->edge = id; e->off = (uint32_t)i;
    }
    b->seq_n = 0; b->ent_n = 0;
}
static int eb_is_pal(const kmer_t* k) { return kmer_form(k) == PAL; }  /* :399-406, K even */

/* :408-417 */
static int eb_up_possible(ebuild_t* b, kent_t* e)
{
    uint8_t ctx = e->ctx;
    if (SIDE_COUNT[CTX_PRED(ctx)] != 1) return 0;
    kmer_t p = e->kmer; kmer_to_pred(&p, BITS2VAL[CTX_PRED(ctx)]);
    if (eb_is_pal(&p)) return 0;
    eb_lookup(b, &p, &ctx);
    return SIDE_COUNT[CTX_SUCC(ctx)] == 1;
}
/* :419-428 */
static int eb_down_possible(ebuild_t* b, kent_t* e)
{
    uint8_t ctx = e->ctx;
    if (SIDE_COUNT[CTX_SUCC(ctx)] != 1) return 0;
    kmer_t s = e->kmer; kmer_to_succ(&s, BITS2VAL[CTX_SUCC(ctx)]);
    if (eb_is_pal(&s)) return 0;
    eb_lookup(b, &s, &ctx);
    return SIDE_COUNT[CTX_PRED(ctx)] == 1;
}
/* EdgeBuilder::extend :445-464 */
static void eb_extend(ebuild_t* b, const kmer_t* start, uint8_t ctx)
{
    kmer_t next = *start;
    while (SIDE_COUNT[CTX_SUCC(ctx)] == 1) {
        unsigned succ = BITS2VAL[CTX_SUCC(ctx)];
        kmer_to_succ(&next, succ);
        if (eb_is_pal(&next)) break;
        kent_t* e = eb_lookup(b, &next, &ctx);
        if (SIDE_COUNT[CTX_PRED(ctx)] != 1) break;
        eb_push_base(b, (uint8_t)succ);
        eb_push_ent(b, e);
    }
    switch (seq_form(b->seq, b->seq_n)) {
    case PAL: case FWD: eb_add_edge(b); break;
    default: b->seq_n = 0; b->ent_n = 0; break;
    }
}
/* EdgeBuilder::buildEdge :335-345 */
static void eb_build_edge(ebuild_t* b, kent_t* e)
{
    if (eb_is_pal(&e->kmer)) { eb_assign_kmer(b, &e->kmer); eb_push_ent(b, e); eb_add_edge(b); }
    else if (eb_up_possible(b, e)) {
        if (eb_down_possible(b, e)) return;
        kmer_t rc = kmer_rc(&e->kmer);                      /* extendUpstream :435-438 */
        eb_assign_kmer(b, &rc); eb_push_ent(b, e);
        eb_extend(b, &rc, ctx_rc(e->ctx));
    } else if (eb_down_possible(b, e)) {                    /* extendDownstream :440-443 */
        eb_assign_kmer(b, &e->kmer); eb_push_ent(b, e);
        eb_extend(b, &e->kmer, e->ctx);
    } else { eb_assign_kmer(b, &e->kmer); eb_push_ent(b, e); eb_add_edge(b); }
}
/* canonicalizeCircle :375-397 */
static void eb_canon_circle(ebuild_t* b)
{
    size_t idx = 0;
    for (size_t i = 1; i < b->ent_n; ++i)
        if (kmer_cmp(&b->ents[i]->kmer, &b->ents[idx]->kmer) < 0) idx = i;
    kmer_t at = kmer_from_codes(b->seq + idx);
    if (kmer_form(&at) == REV) {
        seq_rc_inplace(b->seq, b->seq_n);
        for (size_t i = 0, j = b->ent_n; i + 1 < j; ++i) { --j; kent_t* t = b->ents[i]; b->ents[i] = b->ents[j]; b->ents[j] = t; }
        idx = b->seq_n - idx - K;
    }
    if (!idx) return;
    size_t n = b->seq_n;
    uint8_t* bv = (uint8_t*)malloc(n);
    size_t m = 0;
    for (size_t i = idx; i < n; ++i) bv[m++] = b->seq[i];
    for (size_t i = K - 1; i < K + idx - 1; ++i) bv[m++] = b->seq[i];
    memcpy(b->seq, bv, m); b->seq_n = m; free(bv);
    kent_t** ne = (kent_t**)malloc(sizeof(kent_t*) * b->ent_n);
    size_t q = 0;
    for (size_t i = idx; i < b->ent_n; ++i) ne[q++] = b->ents[i];
    for (size_t i = 0; i < idx; ++i) ne[q++] = b->ents[i];
    memcpy(b->ents, ne, sizeof(kent_t*) * b->ent_n); free(ne);
}
/* simpleCircle :348-372 */
static void eb_simple_circle(ebuild_t* b, kent_t* first)
{
    eb_assign_kmer(b, &first->kmer); eb_push_ent(b, first);
    uint8_t ctx = first->ctx;
    kmer_t k = first->kmer;
    for (;;) {
        if (SIDE_COUNT[CTX_PRED(ctx)] != 1 || SIDE_COUNT[CTX_SUCC(ctx)] != 1) { fprintf(stderr, "oracle: circle assert\n"); abort(); }
        unsigned succ = BITS2VAL[CTX_SUCC(ctx)];
        kmer_to_succ(&k, succ);
        kent_t* e = eb_lookup(b, &k, &ctx);
        if (e == first) break;
        if (e->edge != ~0u) { fprintf(stderr, "oracle: failed to close circle\n"); abort(); }
        eb_push_base(b, (uint8_t)succ);
        eb_push_ent(b, e);
    }
    eb_canon_circle(b);
    eb_add_edge(b);
}
/* buildEdges :514-541 */
int sn_oracle_edges(sn_oracle_t* o)
{
    ebuild_t b; memset(&b, 0, sizeof b); b.o = o;
    o->n_edges = 0; o->edge_seq = NULL; o->edge_len = NULL;
    for (uint64_t i = 0; i < o->n_kmers; ++i)
        if (o->kmers[i].edge == ~0u) eb_build_edge(&b, &o->kmers[i]);
    for (uint64_t i = 0; i < o->n_kmers; ++i)
        if (o->kmers[i].edge == ~0u) eb_simple_circle(&b, &o->kmers[i]);
    free(b.seq); free(b.ents);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a8: buildHBVFromEdges  paths/long/HBVFromEdges.cc:244-296                  */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t w[KW]; } sub_t;   /* (K-1)-mer, last 2 bits zero */
static sub_t sub_from(const uint8_t* s, size_t len, int rc, int distal)
{   /* EdgeEnd(pBV,rc,distal,K-1) :32-33 with SwitchHitterIter feudal/BaseVec.h:98-126 */
    sub_t r; memset(&r, 0, sizeof r);
    size_t pos = distal ? len - (K - 1) : 0;
    for (int i = 0; i < K - 1; ++i) {
        unsigned c = rc ? (s[len - 1 - (pos + i)] ^ 3u) : s[pos + i];
        r.w[i >> 4] |= (uint32_t)c << (2 * (15 - (i & 15)));
    }
    return r;
}
static inline int sub_cmp(const sub_t* a, const sub_t* b)
{
    for (int i = 0; i < KW; ++i) if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}
typedef struct { sub_t key; uint32_t rank; uint32_t edge; uint8_t rc; uint32_t pos; } eend_t;
/* EEComp :113-121 (edge order = BVComp :106-111, then rc, then pos) */
static int eend_cmp(const void* a, const void* b)
{
    const eend_t* x = (const eend_t*)a; const eend_t* y = (const eend_t*)b;
    int c = sub_cmp(&x->key, &y->key); if (c) return c;
    if (x->rank != y->rank) return x->rank < y->rank ? -1 : 1;
    if (x->rc != y->rc) return x->rc < y->rc ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return 0;
}
static sn_oracle_t* g_sort_o;
/* BVComp :106-111 : longer first, then lexicographic (feudal/FieldVec.h:830-833) */
static int edge_order_cmp(const void* a, const void* b)
{
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    uint32_t lx = g_sort_o->edge_len[x], ly = g_sort_o->edge_len[y];
    if (lx != ly) return lx > ly ? -1 : 1;
    int c = memcmp(g_sort_o->edge_seq[x], g_sort_o->edge_seq[y], lx);
    if (c) return c;
    return (x > y) - (x < y);
}
typedef struct { sub_t key; int32_t id; uint32_t beg, end; } vert_t;

static void ins_sorted(int32_t** arr, int32_t** eo, int32_t* n, int32_t w, int32_t e)
{   /* digraphE::AddEdge graph/DigraphTemplate.h:2572-2582 : upper_bound insertion */
    int32_t m = *n;
    *arr = (int32_t*)realloc(*arr, sizeof(int32_t) * (m + 1));
    *eo = (int32_t*)realloc(*eo, sizeof(int32_t) * (m + 1));
    int32_t i = 0; while (i < m && (*arr)[i] <= w) ++i;
    memmove(*arr + i + 1, *arr + i, sizeof(int32_t) * (m - i));
    memmove(*eo + i + 1, *eo + i, sizeof(int32_t) * (m - i));
    (*arr)[i] = w; (*eo)[i] = e; *n = m + 1;
}

int sn_oracle_hbv(sn_oracle_t* o)
{
    uint64_t nE = o->n_edges;
    o->n_vert = 0; o->n_hbv_edges = 0;
    o->fwd_xlat = (int32_t*)malloc(sizeof(int32_t) * (nE ? nE : 1));
    o->rev_xlat = (int32_t*)malloc(sizeof(int32_t) * (nE ? nE : 1));
    if (!nE) return 0;
    /* edge order */
    uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * nE);
    uint32_t* rank = (uint32_t*)malloc(sizeof(uint32_t) * nE);
    for (uint64_t i = 0; i < nE; ++i) order[i] = (uint32_t)i;
    g_sort_o = o;
    qsort(order, nE, sizeof(uint32_t), edge_order_cmp);
    for (uint64_t i = 0; i < nE; ++i) rank[order[i]] = (uint32_t)i;
    /* VertexDictBuilder::map :141-148 */
    eend_t* ee = (eend_t*)malloc(sizeof(eend_t) * 4 * nE);
    uint64_t ne = 0;
    uint8_t* is_pal = (uint8_t*)malloc(nE);
    for (uint64_t e = 0; e < nE; ++e) {
        const uint8_t* s = o->edge_seq[e]; size_t len = o->edge_len[e];
        is_pal[e] = seq_form(s, len) == PAL;
        for (int rc = 0; rc < (is_pal[e] ? 1 : 2); ++rc)
            for (int distal = 0; distal < 2; ++distal) {
                eend_t* x = &ee[ne++];
                x->key = sub_from(s, len, rc, distal); x->rank = rank[e]; x->edge = (uint32_t)e;
                x->rc = (uint8_t)rc; x->pos = distal ? (uint32_t)(len - (K - 1)) : 0;
            }
    }
    qsort(ee, ne, sizeof(eend_t), eend_cmp);
    /* vertices = groups of equal keys (VertexDictBuilder::reduce :150-159) */
    vert_t* V = (vert_t*)malloc(sizeof(vert_t) * ne);
    int32_t nV = 0;
    for (uint64_t i = 0; i < ne; ) {
        uint64_t j = i; while (j < ne && sub_cmp(&ee[j].key, &ee[i].key) == 0) ++j;
        if (j - i > 8) { fprintf(stderr, "oracle: vertex with >8 edge ends\n"); abort(); } /* MAX_EDGES :83 */
        V[nV].key = ee[i].key; V[nV].id = -1; V[nV].beg = (uint32_t)i; V[nV].end = (uint32_t)j; ++nV;
        i = j;
    }
    o->n_vert = nV;
    o->from = (int32_t**)calloc(nV, sizeof(int32_t*)); o->from_eo = (int32_t**)calloc(nV, sizeof(int32_t*)); o->from_n = (int32_t*)calloc(nV, sizeof(int32_t));
    o->to = (int32_t**)calloc(nV, sizeof(int32_t*)); o->to_eo = (int32_t**)calloc(nV, sizeof(int32_t*)); o->to_n = (int32_t*)calloc(nV, sizeof(int32_t));
    o->hbv_seq = (uint8_t**)malloc(sizeof(uint8_t*) * 2 * nE); o->hbv_len = (uint32_t*)malloc(sizeof(uint32_t) * 2 * nE);
    o->to_left = (int32_t*)malloc(sizeof(int32_t) * 2 * nE); o->to_right = (int32_t*)malloc(sizeof(int32_t) * 2 * nE);
    for (uint64_t i = 0; i < nE; ++i) o->fwd_xlat[i] = o->rev_xlat[i] = -1;
    /* HBVBuilder :172-241 */
    uint64_t qcap = 1024, qh = 0, qt = 0;
    uint64_t* Q = (uint64_t*)malloc(sizeof(uint64_t) * qcap);   /* (edge<<1)|rc */
    int32_t nextV = 0;
#define DONE(e, rc) (((rc) ? o->rev_xlat : o->fwd_xlat)[e] != -1)
#define QPUSH(v) do { if (qt == qcap) { if (qh > 0) { memmove(Q, Q + qh, sizeof(uint64_t) * (qt - qh)); qt -= qh; qh = 0; } \
        if (qt == qcap) { qcap *= 2; Q = (uint64_t*)realloc(Q, sizeof(uint64_t) * qcap); } } Q[qt++] = (v); } while (0)
    for (int pass = 0; pass < 2; ++pass)
        for (uint64_t oi = 0; oi < nE; ++oi) {
            uint32_t e0 = order[oi];
            if (DONE(e0, pass)) continue;
            QPUSH(((uint64_t)e0 << 1) | (uint64_t)pass);
            while (qh < qt) {
                uint64_t it = Q[qh++];
                uint32_t e = (uint32_t)(it >> 1); int rc = (int)(it & 1);
                if (DONE(e, rc)) continue;
                const uint8_t* s = o->edge_seq[e]; size_t len = o->edge_len[e];
                vert_t* pv[2];
                for (int distal = 0; distal < 2; ++distal) {
                    sub_t key = sub_from(s, len, rc, distal);
                    int32_t lo = 0, hi = nV; vert_t* f = NULL;
                    while (lo < hi) { int32_t mid = (lo + hi) >> 1; int c = sub_cmp(&V[mid].key, &key);
                        if (c == 0) { f = &V[mid]; break; } if (c < 0) lo = mid + 1; else hi = mid; }
                    if (!f) { fprintf(stderr, "oracle: vertex lookup failed\n"); abort(); }
                    if (f->id == -1) f->id = nextV++;
                    pv[distal] = f;
                }
                int32_t v1 = pv[0]->id, v2 = pv[1]->id;
                int32_t ne_id = o->n_hbv_edges++;
                uint8_t* cp = (uint8_t*)malloc(len); memcpy(cp, s, len);
                if (rc) seq_rc_inplace(cp, len);
                o->hbv_seq[ne_id] = cp; o->hbv_len[ne_id] = (uint32_t)len;
                ins_sorted(&o->from[v1], &o->from_eo[v1], &o->from_n[v1], v2, ne_id);
                ins_sorted(&o->to[v2], &o->to_eo[v2], &o->to_n[v2], v1, ne_id);
                o->to_left[ne_id] = v1; o->to_right[ne_id] = v2;
                if (!rc || is_pal[e]) o->fwd_xlat[e] = ne_id;
                if (rc || is_pal[e]) o->rev_xlat[e] = ne_id;
                for (int d = 0; d < 2; ++d)
                    for (uint32_t x = pv[d]->beg; x < pv[d]->end; ++x)
                        if (!DONE(ee[x].edge, ee[x].rc)) QPUSH(((uint64_t)ee[x].edge << 1) | ee[x].rc);
            }
            qh = qt = 0;
        }
    if (nextV != nV) { fprintf(stderr, "oracle: vertex count mismatch %d %d\n", nextV, nV); abort(); }
    free(Q); free(V); free(ee); free(order); free(rank); free(is_pal);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a10: Pather::path  BuildReadQGraph48.cc:705-747 ; PathPart :605-690        */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t edge; int rc; uint32_t off; uint32_t len; uint32_t elen; } part_t;  /* elen==0 => gap */
#define IS_GAP(p) ((p)->elen == 0)
static part_t mk_gap(uint32_t len) { part_t p; p.edge = ~0u; p.rc = 0; p.off = 0; p.len = len; p.elen = 0; return p; }

/* CF<K>::isRC dna/CanonicalForm.h:84-91 */
static int cf_is_rc(const uint8_t* a, const uint8_t* b)
{
    int i = 0, j = K;
    while (i != K) {
        --j;
        if (a[i] != (a[j] ^ 3u)) return a[i] != b[i];
        ++i;
    }
    return 0;
}
static size_t path_read(sn_oracle_t* o, const uint8_t* rd, uint32_t n, part_t* parts)
{
    size_t np = 0;
    if (n < K) { parts[np++] = mk_gap(n); return np; }
    uint32_t itr = 0, end = n - K + 1;
    while (itr != end) {
        kmer_t kmer = kmer_from_codes(rd + itr);
        kent_t* ent = find_entry(o, &kmer);
        if (!ent) {
            uint32_t gap = 1; uint32_t itr2 = itr + K; ++itr;
            while (itr2 != n) {
                kmer_to_succ(&kmer, rd[itr2]); ++itr2;
                if ((ent = find_entry(o, &kmer))) break;
                ++gap; ++itr;
            }
            parts[np++] = mk_gap(gap);
        }
        if (ent) {
            const uint8_t* edge = o->edge_seq[ent->edge]; uint32_t esz = o->edge_len[ent->edge];
            int offset = (int)ent->off;
            uint32_t len = 1;
            int rc = cf_is_rc(rd + itr, edge + offset);
            if (!rc) {
                uint32_t a = itr + K, b = (uint32_t)offset + K;
                while (a < n && b < esz && rd[a] == edge[b]) { ++len; ++a; ++b; }
            } else {
                offset = (int)esz - offset;
                uint32_t a = itr + K, b = (uint32_t)offset;     /* rc coordinate */
                while (a < n && b < esz && rd[a] == (edge[esz - 1 - b] ^ 3u)) { ++len; ++a; ++b; }
                offset -= K;
            }
            part_t p; p.edge = ent->edge; p.rc = rc; p.off = (uint32_t)offset; p.len = len; p.elen = esz - K + 1;
            parts[np++] = p;
            itr += len;
        }
    }
    return np;
}
static inline uint32_t part_end_off(const part_t* p) { return p->off + p->len; }
static inline int part_same_edge(const part_t* a, const part_t* b) { return a->edge == b->edge && a->rc == b->rc; }
/* isConformingCapturedGap :669-676 */
static int conforming_gap(const part_t* p, unsigned max_jitter)
{
    const part_t* prev = p - 1; const part_t* next = p + 1;
    unsigned dist = next->off - part_end_off(prev);
    if (!part_same_edge(prev, next)) dist += prev->elen;
    int d = (int)(p->len - dist);
    return (unsigned)(d < 0 ? -d : d) <= max_jitter;
}
/* Pather::isJoinable :810-816 (compares the LAST K-1 bases of both oriented edges) */
static int joinable(sn_oracle_t* o, const part_t* a, const part_t* b)
{
    if (a->edge == b->edge) return 1;
    sub_t k1 = sub_from(o->edge_seq[a->edge], o->edge_len[a->edge], a->rc, 1);
    sub_t k2 = sub_from(o->edge_seq[b->edge], o->edge_len[b->edge], b->rc, 1);
    return sub_cmp(&k1, &k2) == 0;
}
static inline int32_t part_hbv_edge(sn_oracle_t* o, const part_t* p)
{ return p->rc ? o->rev_xlat[p->edge] : o->fwd_xlat[p->edge]; }

/* ------------------------------------------------------------------------- */
/* a12: ExtendReadPath  paths/long/ExtendReadPath.cc                           */
/* ------------------------------------------------------------------------- */
typedef struct { int32_t offset; int32_t n; int32_t cap; int32_t* e; } rpath_t;
static void rp_push(rpath_t* p, int32_t e)
{ if (p->n == p->cap) { p->cap = p->cap ? 2 * p->cap : 8; p->e = (int32_t*)realloc(p->e, sizeof(int32_t) * p->cap); } p->e[p->n++] = e; }

/* scoreRightOverlap :15-62 */
static unsigned score_right(const uint8_t* b, const uint8_t* q, uint32_t n, size_t start, const uint8_t* edge, uint32_t esz)
{
    size_t bi = n - start, ei = K - 1;
    unsigned qsum = 0, penalty = 0;
    while (bi != n && ei != esz) {
        if (b[bi] != edge[ei]) { unsigned qs = (q[bi] == 2) ? 20u : q[bi]; penalty += qs; qsum += penalty; }
        else if (penalty > 0) { penalty -= (0.2 * penalty); }
        ++bi; ++ei;
    }
    while (bi++ != n) qsum += 10;
    return qsum;
}
/* scoreLeftOverlap :65-115 */
static unsigned score_left(const uint8_t* b, const uint8_t* q, uint32_t n, size_t start, const uint8_t* edge, uint32_t esz)
{
    (void)n;
    long bi = (long)start - 1, ei = (long)esz - K;
    unsigned qsum = 0, penalty = 0;
    while (bi >= 0 && ei >= 0) {
        if (b[bi] != edge[ei]) { unsigned qs = (q[bi] == 2) ? 20u : q[bi]; penalty += qs; qsum += penalty; }
        else if (penalty > 0) { penalty -= (0.2 * penalty); }
        --bi; --ei;
    }
    while (bi-- >= 0) qsum += 10;
    return qsum;
}
static int cmp_i32(const void* a, const void* b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return (x > y) - (x < y); }

/* attemptLeftwardExtension :133-236 */
static int extend_left(sn_oracle_t* o, rpath_t* p, const uint8_t* b, const uint8_t* q, uint32_t n)
{
    if (!p->n) return 0;
    if (p->offset >= 0) return 0;
    size_t last_gap = (size_t)(-(long)p->offset);
    if (last_gap < 10) return 0;
    int32_t vleft = o->to_left[p->e[0]];
    int32_t ne = o->to_n[vleft]; const int32_t* edges = o->to_eo[vleft]; const int32_t* vdest = o->to[vleft];
    int32_t short_dest[8]; int ns = 0; int hanging[8], elong[8]; int nlong = 0;
    for (int i = 0; i < ne; ++i) {
        hanging[i] = (o->to_n[vdest[i]] == 0 && o->from_n[vdest[i]] == 1);
        elong[i] = ((size_t)o->hbv_len[edges[i]] - (K - 1) >= last_gap);
        nlong += elong[i];
        if (!elong[i] && !hanging[i]) short_dest[ns++] = vdest[i];
    }
    if (ne != 1) {
        if (ns > 0) {
            if (nlong > 0) return 0;
            qsort(short_dest, ns, sizeof(int32_t), cmp_i32);
            int u = 1; for (int i = 1; i < ns; ++i) if (short_dest[i] != short_dest[i - 1]) ++u;
            if (u != 1) return 0;
            if (o->to_n[short_dest[ns - 1]] != 1) return 0;
        }
    }
    int least_edge = -1; unsigned least = 0xFFFFFFFFu;
    for (int i = 0; i < ne; ++i) if (!hanging[i] || ne == 1) {
        unsigned sc = score_left(b, q, n, last_gap, o->hbv_seq[edges[i]], o->hbv_len[edges[i]]);
        if (sc < least) { least_edge = edges[i]; least = sc; }
    }
    if (least_edge == -1 || least > last_gap * 10) return 0;
    int edge_size = (int)o->hbv_len[least_edge] - K + 1;
    p->offset += edge_size;
    rp_push(p, 0);
    memmove(p->e + 1, p->e, sizeof(int32_t) * (p->n - 1));
    p->e[0] = least_edge;
    return 1;
}
/* attemptRightwardExtension :239-358 */
static int extend_right(sn_oracle_t* o, rpath_t* p, const uint8_t* b, const uint8_t* q, uint32_t n)
{
    if (!p->n) return 0;
    int ilast = (int)n; ilast += p->offset;
    for (int i = 0; i < p->n; ++i) ilast -= (int)o->hbv_len[p->e[i]] - K + 1;
    ilast -= (K - 1);
    if (ilast < 10) return 0;
    size_t last_gap = (size_t)ilast;
    int32_t vright = o->to_right[p->e[p->n - 1]];
    int32_t ne = o->from_n[vright]; const int32_t* edges = o->from_eo[vright]; const int32_t* vdest = o->from[vright];
    int32_t short_dest[8]; int ns = 0; int hanging[8], elong[8]; int nlong = 0;
    for (int i = 0; i < ne; ++i) {
        hanging[i] = (o->from_n[vdest[i]] == 0 && o->to_n[vdest[i]] == 1);
        elong[i] = ((size_t)o->hbv_len[edges[i]] - (K - 1) >= last_gap);
        nlong += elong[i];
        if (!elong[i] && !hanging[i]) short_dest[ns++] = vdest[i];
    }
    if (ne != 1) {
        if (ns > 0) {
            if (nlong > 0) return 0;
            qsort(short_dest, ns, sizeof(int32_t), cmp_i32);
            int u = 1; for (int i = 1; i < ns; ++i) if (short_dest[i] != short_dest[i - 1]) ++u;
            if (u != 1) return 0;
            if (o->from_n[short_dest[ns - 1]] != 1) return 0;
        }
    }
    int least_edge = -1; unsigned least = 0xFFFFFFFFu;
    for (int i = 0; i < ne; ++i) if (!hanging[i] || ne == 1) {
        unsigned sc = score_right(b, q, n, last_gap, o->hbv_seq[edges[i]], o->hbv_len[edges[i]]);
        if (sc < least) { least_edge = edges[i]; least = sc; }
    }
    if (least_edge == -1 || least > last_gap * 10) return 0;
    rp_push(p, least_edge);
    return 1;
}

/* ------------------------------------------------------------------------- */
/* a11: HBVPather::algorithmTwo + pathPartsToReadPath :1217-1336,1385-1420     */
/* ------------------------------------------------------------------------- */
static void parts_to_path(sn_oracle_t* o, const part_t* parts, size_t np, rpath_t* path)
{
    path->n = 0;
    const part_t* last = NULL;
    for (size_t i = 0; i < np; ++i) {
        const part_t* p = &parts[i];
        if (IS_GAP(p)) continue;
        if (last && part_same_edge(last, p)) continue;
        rp_push(path, part_hbv_edge(o, p));
        last = p;
    }
    if (path->n == 0) path->offset = 0;
    else if (!IS_GAP(&parts[0])) path->offset = (int)parts[0].off;
    else path->offset = (int)parts[1].off - (int)parts[0].len;
}
static void path_one_read(sn_oracle_t* o, const uint8_t* rd, const uint8_t* q, uint32_t n, rpath_t* path,
                          part_t* parts, part_t* np_buf)
{
    size_t np = path_read(o, rd, n, parts);
    /* seeds on short hanging edges -> gaps :1236-1258 */
    size_t nn = 0;
    for (size_t i = 0; i < np; ++i) {
        part_t part = parts[i];
        if (!IS_GAP(&part)) {
            int32_t e = part_hbv_edge(o, &part);
            int32_t vl = o->to_left[e], vr = o->to_right[e];
            if (o->to_n[vl] == 0 && o->to_n[vr] > 1 && o->from_n[vr] > 0 && part.elen <= 100)
                part = mk_gap(part.len);
        }
        if (IS_GAP(&part) && nn && IS_GAP(&np_buf[nn - 1])) np_buf[nn - 1].len += part.len;
        else np_buf[nn++] = part;
    }
    memcpy(parts, np_buf, sizeof(part_t) * nn); np = nn;
    /* non-conforming captured gap :1264-1288 */
    if (np >= 3) {
        size_t seeds = IS_GAP(&parts[0]) ? 0u : 1u;
        for (size_t i = 1; i + 1 < np; ++i) {
            if (!IS_GAP(&parts[i])) { seeds++; continue; }
            if (!conforming_gap(&parts[i], 3) || !joinable(o, &parts[i - 1], &parts[i + 1])) {
                if (seeds > 1) {
                    part_t tmp = mk_gap(parts[i - 1].len);
                    for (size_t j = i; j < np; ++j) tmp.len += parts[j].len;
                    np = i - 1; parts[np++] = tmp;
                } else {
                    for (size_t j = i + 1; j < np; ++j) parts[i].len += parts[j].len;
                    np = i + 1;
                }
                break;
            }
        }
    }
    /* back off short terminal seeds :1293-1307 */
    if (IS_GAP(&parts[np - 1]) && np > 1) {
        part_t* last2 = &parts[np - 2];
        if (last2->off == 0 && last2->len <= 5) {
            part_t last = parts[np - 1];
            last.len += last2->len;
            np -= 2; parts[np++] = last;
        }
    } else if (!IS_GAP(&parts[np - 1])) {
        part_t* last = &parts[np - 1];
        if (last->off == 0 && last->len <= 5) *last = mk_gap(last->len);
    }
    parts_to_path(o, parts, np, path);
    /* graph connectivity :1313-1320 */
    if (path->n >= 2)
        for (int i = 0; i + 1 < path->n; ++i)
            if (o->to_right[path->e[i]] != o->to_left[path->e[i + 1]]) { path->n = i + 1; break; }
    /* ExtendReadPath::attemptLeftRightExtension ExtendReadPath.cc:121-129 */
    while (extend_left(o, path, rd, q, n)) {}
    while (extend_right(o, path, rd, q, n)) {}
}

int sn_oracle_paths(sn_oracle_t* o)
{
    uint64_t n = o->n_reads;
    o->path_offset = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    o->path_off = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
    uint64_t cap = 4 * n + 16, tot = 0;
    o->path_edges = (int32_t*)malloc(sizeof(int32_t) * cap);
    uint32_t maxlen = 0;
    for (uint64_t r = 0; r < n; ++r) { uint32_t l = (uint32_t)(o->off[r + 1] - o->off[r]); if (l > maxlen) maxlen = l; }
    part_t* parts = (part_t*)malloc(sizeof(part_t) * (maxlen + 2));
    part_t* buf = (part_t*)malloc(sizeof(part_t) * (maxlen + 2));
    rpath_t path; memset(&path, 0, sizeof path);
    for (uint64_t r = 0; r < n; ++r) {
        uint32_t len = (uint32_t)(o->off[r + 1] - o->off[r]);
        path.n = 0; path.offset = 0;
        path_one_read(o, o->bases + o->off[r], o->quals + o->off[r], len, &path, parts, buf);
        o->path_offset[r] = path.offset;
        o->path_off[r] = tot;
        if (tot + path.n > cap) { cap = 2 * (tot + path.n); o->path_edges = (int32_t*)realloc(o->path_edges, sizeof(int32_t) * cap); }
        memcpy(o->path_edges + tot, path.e, sizeof(int32_t) * path.n); tot += path.n;
    }
    o->path_off[n] = tot;
    free(parts); free(buf); free(path.e);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a9: HyperBasevector::Involution paths/HyperBasevector.cc:685-697            */
/* inv[e] = id of the edge whose sequence is the reverse complement of e       */
/* ------------------------------------------------------------------------- */
int sn_oracle_involution(sn_oracle_t* o, int32_t* inv)
{
    for (uint64_t e = 0; e < o->n_edges; ++e) {
        inv[o->fwd_xlat[e]] = o->rev_xlat[e];
        inv[o->rev_xlat[e]] = o->fwd_xlat[e];
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* driver + lifetime                                                           */
/* ------------------------------------------------------------------------- */
sn_oracle_t* sn_oracle_new(uint64_t n_reads, const uint8_t* bases, const uint8_t* quals, const uint64_t* off,
                           const int32_t* bc, unsigned min_qual, unsigned min_freq, unsigned min_bc)
{
    sn_oracle_t* o = (sn_oracle_t*)calloc(1, sizeof(sn_oracle_t));
    o->n_reads = n_reads; o->bases = bases; o->quals = quals; o->off = off; o->bc = bc;
    o->min_qual = min_qual; o->min_freq = min_freq; o->min_bc = min_bc; o->ign_bc_below = 0; o->count_len_k = 0;
    return o;
}
/* ignBcBelow of buildReadQGraph48 (BuildReadQGraph48.cc:158-159): reads with id below it count as barcode -1 */
void sn_oracle_set_ign_bc_below(sn_oracle_t* o, int64_t v) { o->ign_bc_below = v; }
void sn_oracle_set_count_len_k(sn_oracle_t* o, int v) { o->count_len_k = v; }
int sn_oracle_run(sn_oracle_t* o, int with_paths)
{
    sn_oracle_count(o); sn_oracle_prune(o); sn_oracle_edges(o); sn_oracle_hbv(o);
    if (with_paths) sn_oracle_paths(o);
    return 0;
}
void sn_oracle_free(sn_oracle_t* o)
{
    if (!o) return;
    free(o->good_len); free(o->kmers);
    for (uint64_t i = 0; i < o->n_edges; ++i) free(o->edge_seq[i]);
    free(o->edge_seq); free(o->edge_len);
    for (int32_t v = 0; v < o->n_vert; ++v) { free(o->from[v]); free(o->from_eo[v]); free(o->to[v]); free(o->to_eo[v]); }
    free(o->from); free(o->from_eo); free(o->from_n); free(o->to); free(o->to_eo); free(o->to_n);
    for (int32_t e = 0; e < o->n_hbv_edges; ++e) free(o->hbv_seq[e]);
    free(o->hbv_seq); free(o->hbv_len); free(o->fwd_xlat); free(o->rev_xlat); free(o->to_left); free(o->to_right);
    free(o->path_offset); free(o->path_off); free(o->path_edges);
    free(o);
}

/* accessors for ctypes (tests) */
uint64_t sn_oracle_n_occ(sn_oracle_t* o) { return o->n_occ; }
uint64_t sn_oracle_n_kmers(sn_oracle_t* o) { return o->n_kmers; }
uint64_t sn_oracle_n_edges(sn_oracle_t* o) { return o->n_edges; }
int32_t  sn_oracle_n_vert(sn_oracle_t* o) { return o->n_vert; }
int32_t  sn_oracle_n_hbv_edges(sn_oracle_t* o) { return o->n_hbv_edges; }
const uint32_t* sn_oracle_good_len(sn_oracle_t* o) { return o->good_len; }
/* out: n_kmers x {w0,w1,w2,count,ctx0,ctx} as u32[6] */
void sn_oracle_get_kmers(sn_oracle_t* o, uint32_t* out)
{
    for (uint64_t i = 0; i < o->n_kmers; ++i) {
        kent_t* e = &o->kmers[i];
        out[6 * i + 0] = e->kmer.w[0]; out[6 * i + 1] = e->kmer.w[1]; out[6 * i + 2] = e->kmer.w[2];
        out[6 * i + 3] = e->count; out[6 * i + 4] = e->ctx0; out[6 * i + 5] = e->ctx;
    }
}
uint32_t sn_oracle_edge_len(sn_oracle_t* o, uint64_t e) { return o->edge_len[e]; }
const uint8_t* sn_oracle_edge_seq(sn_oracle_t* o, uint64_t e) { return o->edge_seq[e]; }
const int32_t* sn_oracle_path_offset(sn_oracle_t* o) { return o->path_offset; }
const uint64_t* sn_oracle_path_off(sn_oracle_t* o) { return o->path_off; }
const int32_t* sn_oracle_path_edges(sn_oracle_t* o) { return o->path_edges; }
const int32_t* sn_oracle_fwd_xlat(sn_oracle_t* o) { return o->fwd_xlat; }
const int32_t* sn_oracle_rev_xlat(sn_oracle_t* o) { return o->rev_xlat; }

/* ---- serializers (reference byte layouts) ---------------- */
typedef struct { uint8_t* p; size_t n, cap; } buf_t;
static void bput(buf_t* b, const void* s, size_t n)
{ if (b->n + n > b->cap) { b->cap = (b->n + n) * 2 + 64; b->p = (uint8_t*)realloc(b->p, b->cap); } memcpy(b->p + b->n, s, n); b->n += n; }
static void bput_u64(buf_t* b, uint64_t v) { bput(b, &v, 8); }
static void bput_bases(buf_t* b, const uint8_t* s, uint32_t len)
{   /* feudal/FieldVec.h:596-598,762 : u32 nbases + ceil(n/4) bytes, base j at bits 2*(j%4) */
    bput(b, &len, 4);
    for (uint32_t i = 0; i < len; i += 4) {
        uint8_t x = 0;
        for (uint32_t j = 0; j < 4 && i + j < len; ++j) x |= (uint8_t)(s[i + j] << (2 * j));
        bput(b, &x, 1);
    }
}
/* a.hbv : paths/HyperBasevector.cc:121-125, graph/DigraphTemplate.h:3091-3097 */
int sn_oracle_write_hbv(sn_oracle_t* o, const char* path)
{
    buf_t b = {0, 0, 0};
    bput(&b, "BINWRITE", 8);
    int32_t k = K; bput(&b, &k, 4);
    int32_t** arrs[3] = { o->from, o->from_eo, o->to_eo };
    int32_t* ns[3] = { o->from_n, o->from_n, o->to_n };
    for (int a = 0; a < 3; ++a) {
        bput_u64(&b, (uint64_t)o->n_vert);
        for (int32_t v = 0; v < o->n_vert; ++v) { bput_u64(&b, (uint64_t)ns[a][v]); bput(&b, arrs[a][v], 4 * (size_t)ns[a][v]); }
    }
    bput_u64(&b, (uint64_t)o->n_hbv_edges);
    for (int32_t e = 0; e < o->n_hbv_edges; ++e) bput_bases(&b, o->hbv_seq[e], o->hbv_len[e]);
    FILE* f = fopen(path, "wb"); if (!f) { free(b.p); return -1; }
    fwrite(b.p, 1, b.n, f); fclose(f); free(b.p);
    return 0;
}
/* tmp.paths : feudal file of ReadPath (paths/long/ReadPath.h:61-63,
 * feudal/FeudalFileWriter.cc:100-121, feudal/FeudalControlBlock.h:159-165) */
int sn_oracle_write_paths(sn_oracle_t* o, const char* path)
{
    FILE* f = fopen(path, "wb"); if (!f) return -1;
    uint64_t n = o->n_reads;
    uint64_t var = 24 + 8 * n + 4 * o->path_off[n];
    uint8_t hdr[24]; memset(hdr, 0, 24);
    uint32_t n32 = (uint32_t)n; memcpy(hdr, &n32, 4);
    hdr[4] = 1; hdr[5] = 0; hdr[6] = 24; hdr[7] = 4;
    uint64_t fixed = var + 8 * (n + 1);
    memcpy(hdr + 8, &var, 8); memcpy(hdr + 16, &fixed, 8);
    fwrite(hdr, 1, 24, f);
    for (uint64_t r = 0; r < n; ++r) {
        int32_t off = o->path_offset[r]; uint32_t skip = 0;
        fwrite(&off, 4, 1, f); fwrite(&skip, 4, 1, f);
        fwrite(o->path_edges + o->path_off[r], 4, o->path_off[r + 1] - o->path_off[r], f);
    }
    uint64_t pos = 24;
    for (uint64_t r = 0; r <= n; ++r) {
        fwrite(&pos, 8, 1, f);
        if (r < n) pos += 8 + 4 * (o->path_off[r + 1] - o->path_off[r]);
    }
    fclose(f);
    return 0;
}
