/*
 * supernova_b200.h -- C ABI of the B200-native k-mer count -> unipath graph
 * (HyperBasevector) -> ReadPath hot path.
 *
 * Drop-in boundary (boundary B2): these entry points sit directly
 * under the reference's C++ call
 *     buildReadQGraph48(work_dir, read_head, "", reads, quals, False, False, minQual,
 *                       minFreq, ignBcBelow, minBC, &bc, .75, 0, "", True, False,
 *                       &hbv, &paths)
 *   lib/assembly/src/paths/long/BuildReadQGraph48.h:28-40, called from
 *   StageBuildGraph, lib/assembly/src/10X/runstages/RunStages.cc:404-406
 * and replace what it does between "reads/quals/barcodes in memory" and
 * "a.hbv + tmp.paths on disk".  Plain pointers and sizes only; every buffer handed
 * in is HOST memory in the reference's own in-memory/on-disk layout; every function
 * returns 0 on success or a negative sn_status and leaves a message in
 * sn_last_error().  No exceptions cross the ABI.  There is no CPU fallback: without a
 * CUDA device sn_ctx_create fails.
 */
#ifndef SUPERNOVA_B200_H_
#define SUPERNOVA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sn_ctx sn_ctx;

enum sn_status {
    SN_OK = 0,
    SN_ERR_CUDA = -1,        /* a CUDA call failed (message has the call and the error) */
    SN_ERR_ARG = -2,         /* invalid argument / unsupported configuration            */
    SN_ERR_STATE = -3,       /* stage called out of order                               */
    SN_ERR_IO = -4,          /* file could not be read / written                        */
    SN_ERR_DATA = -5         /* inconsistent input (e.g. PQVec length != read length)   */
};

/* Thresholds of Kmerizer / GoodLenTailFinder; the values StageBuildGraph passes are
 * K=48 (fixed), minQual=7, minFreq=3, minBC=2, ignBcBelow=0
 * (10X/DF.cc:138-141, 10X/runstages/RunStages.cc:405). */
typedef struct sn_params {
    uint32_t min_qual;
    uint32_t min_freq;
    uint32_t min_bc;         /* 0, 1 or 2 */
    int64_t  ign_bc_below;   /* reads with id below this count as barcode -1 */
} sn_params;

/* k-mer record of sn_get_kmers == the {KMer<48>, KDef} part of a kmers.kvec entry
 * (kmers/ReadPather.h:100-150): w[] MSB-first 2-bit k-mer, count_ctx = count:24 | ctx<<24
 * with ctx the context BEFORE recomputeAdjacencies.  Sorted by k-mer. */
typedef struct sn_kmer_rec { uint32_t w[3]; uint32_t count_ctx; } sn_kmer_rec;

typedef struct sn_counts {
    uint64_t n_reads, n_bases;
    uint64_t n_kmer_occurrences;   /* records emitted by Kmerizer::map                */
    uint64_t n_kmers_distinct;     /* before the frequency / barcode filter           */
    uint64_t n_kmers;              /* dictionary size (pVec->size())                  */
    uint64_t n_edges;              /* unipath edges (before HBV doubling)             */
    uint64_t n_edge_bases;
    uint64_t n_hbv_vertices, n_hbv_edges;
    uint64_t n_path_edges;         /* total ReadPath entries                          */
    uint64_t n_superkmers;         /* super-k-mer records this context counted (MSP)  */
} sn_counts;

int  sn_ctx_create(sn_ctx** out, int device);
void sn_ctx_destroy(sn_ctx* ctx);
const char* sn_last_error(const sn_ctx* ctx);       /* ctx may be NULL: last create error */
int  sn_device_count(void);

/* ---- ingest: replaces vecbvec reads / VecPQVec quals / vec<int32_t> bc ------------ */
/* bases   : the variable-data block of a .fastb (feudal/FieldVec.h:596-598): 2 bits per
 *           base, 4 bases per byte LSB first, each read starting on a byte boundary.
 * base_off: n_reads+1 byte offsets into bases.   len: bases per read.
 * pq      : the variable-data block of a .qualp (PQVec blocks, feudal/PQVec.cc:87-127).
 * pq_off  : n_reads+1 byte offsets into pq.
 * bc      : per-read barcode ordinal as expanded in 10X/DF.cc:464-469, or NULL.        */
int sn_load_reads(sn_ctx* ctx, uint64_t n_reads, const uint8_t* bases, const uint64_t* base_off,
                  const uint32_t* len, const uint8_t* pq, const uint64_t* pq_off, const int32_t* bc);
/* sn_load_reads with the good-length scan and the first MSP pass running under the copies
 * (chunked copies on a second stream; host buffers should be page-locked).  params fixes
 * min_qual for that work; the next sn_count_kmers / sn_build_read_qgraph48 with the same
 * min_qual continues from there.  with_hist = 0: overlap the good lengths only (multi-GPU).  */
int sn_load_reads_streamed(sn_ctx* ctx, uint64_t n_reads, const uint8_t* bases, const uint64_t* base_off,
                           const uint32_t* len, const uint8_t* pq, const uint64_t* pq_off, const int32_t* bc,
                           const sn_params* params, int with_hist);
/* same with one unpacked Phred byte per base (qual_off = n_reads+1 element offsets)   */
int sn_load_reads_q8(sn_ctx* ctx, uint64_t n_reads, const uint8_t* bases, const uint64_t* base_off,
                     const uint32_t* len, const uint8_t* quals, const uint64_t* qual_off, const int32_t* bc);
/* ---- ingest on the device: replaces ParseBarcodedFastqs
 * (10X/ParseBarcodedFastqs.cc:56-146,284-303) -- the barcoded pseudo-FASTQ of the pipeline, 9 lines per
 * record (@name, R1, Q1, R2, Q2, BARCODE-gemgroup[,raw], bcQ, SI, SIQ), parsed, 2-bit packed and
 * PQVec-encoded by CUDA kernels straight into the context.  `text` is the decompressed content.       */
int sn_load_fasth_text(sn_ctx* ctx, const char* text, uint64_t n_bytes);
int sn_load_fasth_file(sn_ctx* ctx, const char* path /* plain or .gz */);
/* FASTQS={a,b,...} (:258-264): several barcode-sorted files; a file never continues the barcode of the one before it */
int sn_load_fasth_files(sn_ctx* ctx, const char* const* paths, uint32_t n_files);
/* the loaded reads as reads.fastb / reads.qualp / reads.bci (any may be NULL)                        */
int sn_save_read_files(sn_ctx* ctx, const char* fastb, const char* qualp, const char* bci);
/* synthetic linked reads generated ON the device (counter-based: every base is a pure function of seed and
 * index; no genome in memory), for workloads a host generator cannot produce in reasonable time.  The context receives the
 * reads of pairs [first_pair, first_pair + n_pairs) of a job of spec->total_pairs pairs (a rank's slice), 2 x 150 bases each,
 * barcode ordinals ascending, quals PQVec-encoded on the device.  err_thresholds: 150 integers, substitution probability at
 * position j in units of 2^-24 (supernova_b200/synth.py: cb_error_thresholds).  The numpy twin (synth.make_reads_cb) gives
 * the same reads bit for bit.                                                                                          */
typedef struct sn_synth { uint64_t genome_bases, total_pairs, seed; uint32_t n_barcodes; } sn_synth;
int sn_generate_reads(sn_ctx* ctx, const sn_synth* spec, uint64_t first_pair, uint64_t n_pairs, const uint32_t* err_thresholds);

/* the three files ParseBarcodedFastqs writes (10X/ParseBarcodedFastqs.cc:284-303)      */
int sn_load_read_files(sn_ctx* ctx, const char* fastb, const char* qualp, const char* bci);
/* reads [first_read, first_read + n_reads) of the files (n_reads = 0: to the end): one rank's shard; bci may be NULL */
int sn_load_read_files_range(sn_ctx* ctx, const char* fastb, const char* qualp, const char* bci, uint64_t first_read, uint64_t n_reads);
/* the same with the per-read barcode ordinals already in memory (the vec<int32_t> buildReadQGraph48's caller
 * passes as bcp, 10X/runstages/RunStages.cc:405); bc may be NULL (n_bc ignored)            */
int sn_load_read_files_bc(sn_ctx* ctx, const char* fastb, const char* qualp, const int32_t* bc, uint64_t n_bc);

/* Which of the reference's two implementations of the count the context follows where they differ (SURVEY §8 a14-a16,
 * equivalence note): SN_SEM_CXX (default) = the C++ path, a read needs goodLen >= K + 1 (BuildReadQGraph48.cc:160);
 * SN_SEM_TADA = the Rust stages (lib/tada), a read trimmed to exactly K bases still gives its one k-mer, without
 * neighbours (cmd_msp.rs:109-110, find_trim_len :129-146).  Thresholds and barcode rule are the same in both
 * (min_kmer_obs = MIN_FREQ, num_bcs > 1 = MIN_BC 2; utils.rs:322-408).  Set before sn_count_kmers / a streamed load.  */
enum { SN_SEM_CXX = 0, SN_SEM_TADA = 1 };
int sn_set_semantics(sn_ctx* ctx, int semantics);

/* ---- stages (must run in this order) -------------------------------------------------- */
/* createDict up to the KmerVec (BuildReadQGraph48.cc:218-292): good lengths, k-mer records,
 * sort, count, filter.                                                                  */
int sn_count_kmers(sn_ctx* ctx, const sn_params* params);
/* recomputeAdjacencies + buildEdges (BuildReadQGraph48.cc:320-321, 514-541)             */
int sn_build_edges(sn_ctx* ctx);
/* buildHBVFromEdges + Involution (paths/long/HBVFromEdges.cc:244-296)                   */
int sn_build_hbv(sn_ctx* ctx);
/* pathReads with useNewAligner=True (BuildReadQGraph48.cc:1440-1469)                    */
int sn_path_reads(sn_ctx* ctx);

/* DF side, SURVEY §8(f) row 1: writePathsIndex (10X/PathsIndex.cc:23-143, called at 10X/DF.cc:588) -- for
 * every HBV edge the reads whose path crosses it (sorted; twice if it crosses twice) and countsb[e] =
 * reads on e + reads on inv[e].  After sn_path_reads.                                               */
int sn_build_paths_index(sn_ctx* ctx);
int sn_get_paths_index(sn_ctx* ctx, uint64_t* off /* n_hbv_edges+1 */, uint64_t* read_ids /* n_path_edges */, int32_t* countsb /* n_hbv_edges */);
int sn_write_paths_index(sn_ctx* ctx, const char* paths_inv /* a.paths.inv */, const char* countsb /* a.countsb */);

/* DF side, the rest of what DF does with the graph and the paths before it goes on (10X/DF.cc:573-600):
 * -- the ReadPathVecX DF keeps the paths in (InitializePathsXFromPaths, 10X/DfTools.cc:24-78; record format
 *    10X/paths/ReadPathParser.cc:17-50): per read [edge count u8][offset i16][first edge u32][2 bits per further edge],
 *    every 10th record indexed.  After sn_path_reads; the buffers stay owned by the context.
 * -- MarkDups, version 2 (10X/SecretOps.cc:599-774): read pairs placed alike (first edge, offset, first five bases of
 *    the partner read) are duplicates of each other, the one with the highest quality sum stays.  dup[p] / art[p] per PAIR
 *    p = read / 2 (`art`: the "artifactual duplicates" it only counts).  After sn_path_reads; needs an even number of
 *    reads.  The percentages it prints are n_dup_pairs / n_pairs, n_interdups / n_dups, n_art_pairs / n_pairs.       */
typedef struct sn_dup_stats { uint64_t n_pairs, n_dup_pairs, n_dups, n_interdups, n_art_pairs; } sn_dup_stats;
int sn_build_pathsx(sn_ctx* ctx);
int sn_get_pathsx(sn_ctx* ctx, uint64_t* n_index, const int64_t** zip_index, uint64_t* n_bytes, const uint8_t** zipped_data);
int sn_write_pathsx(sn_ctx* ctx, const char* path);          /* a.pathsX */
int sn_mark_dups(sn_ctx* ctx, sn_dup_stats* stats /* may be NULL */);
int sn_get_dups(sn_ctx* ctx, uint8_t* dup /* n_reads / 2 */, uint8_t* art /* n_reads / 2, may be NULL */);
int sn_write_dup(sn_ctx* ctx, const char* path);             /* a.dup (vec<Bool>) */

/* buildGraphFromMSP (paths/long/BuildReadQGraph48.h:24-26, .cc:1631-1684), the production boundary when the
 * edges come from the tada stages (MSPEDGES = _ASM_SN.asm_graph, mro/_assembler.mro:57): the vec<basevector>
 * file becomes the edge set (any orientation, any order), the HyperBasevector is built from it
 * (mspEdgesToHBV) and every edge k-mer enters the dictionary with its (edge, offset) (:1656-1664; a k-mer that
 * occurs twice keeps the last).  With reads loaded, sn_path_reads + sn_write_paths then do :1680.          */
int sn_build_graph_from_edges(sn_ctx* ctx, const char* edges_bv);

/* ---- results ------------------------------------------------------------------------- */
int sn_get_counts(const sn_ctx* ctx, sn_counts* out);
int sn_get_good_lengths(sn_ctx* ctx, uint32_t* out /* n_reads */);
int sn_get_kmers(sn_ctx* ctx, sn_kmer_rec* out /* n_kmers */);
/* dictionary after the graph stages: pruned context, unipath id and offset per k-mer      */
int sn_get_kmer_graph_info(sn_ctx* ctx, uint8_t* ctx_pruned, uint32_t* edge, uint32_t* offset);
/* unipath edges: len[n_edges], byte offsets off[n_edges+1], packed bases (fastb layout)   */
int sn_get_edges(sn_ctx* ctx, uint32_t* len, uint64_t* off, uint8_t* packed /* off[n_edges] bytes */);
int sn_get_edges_bytes(const sn_ctx* ctx, uint64_t* packed_bytes);
/* HBV: CSR adjacency.  from_start/to_start: n_vertices+1; from_v/from_e/to_v/to_e: n_hbv_edges */
int sn_get_hbv(sn_ctx* ctx, uint32_t* from_start, int32_t* from_v, int32_t* from_e,
               uint32_t* to_start, int32_t* to_v, int32_t* to_e,
               int32_t* fwd_xlat /* n_edges */, int32_t* rev_xlat /* n_edges */, int32_t* inv /* n_hbv_edges */);
/* ReadPaths: offset[n_reads], path_off[n_reads+1], edges[n_path_edges]                    */
int sn_get_paths(sn_ctx* ctx, int32_t* offset, uint64_t* path_off, int32_t* edges);

/* ---- files in the reference's formats -------------------------------------------------- */
int sn_write_hbv(sn_ctx* ctx, const char* path);             /* a.hbv                              */
int sn_write_paths(sn_ctx* ctx, const char* path);           /* tmp.paths (feudal ReadPathVec)     */
int sn_write_edges_bv(sn_ctx* ctx, const char* path);        /* vec<basevector> (== MSPEDGES file) */
int sn_write_inv(sn_ctx* ctx, const char* path);             /* a.inv  (vec<int>)                  */
int sn_write_to_left_right(sn_ctx* ctx, const char* to_left, const char* to_right);   /* a.to_left, a.to_right (vec<int>) */
int sn_write_kmer_spectrum(sn_ctx* ctx, const char* json);   /* stats/histogram_kmer_count.json    */
/* the other files WriteAssemblyFiles leaves next to a.hbv (10X/WriteFiles.cc:33-51)                    */
int sn_write_hbx(sn_ctx* ctx, const char* path);             /* a.hbx   (HyperBasevectorX)         */
int sn_write_edges_fastb(sn_ctx* ctx, const char* path);     /* a.fastb (the HBV edges, feudal)    */
int sn_write_kmers(sn_ctx* ctx, const char* path);           /* a.kmers (vec<int>: k-mers per edge) */
int sn_write_k(sn_ctx* ctx, const char* path);               /* a.k     ("48")                     */

/* One call == buildReadQGraph48: the four stages, then work_dir/a.hbv (when write_hbv),
 * work_dir/tmp.paths (when with_paths) and work_dir/stats/histogram_kmer_count.json.      */
int sn_build_read_qgraph48(sn_ctx* ctx, const char* work_dir, const sn_params* params, int with_paths, int write_files);

/* number of minimizer-bucket bits for a job with that many k-mer occurrences in total (768..1536 occurrences per bucket) */
int   sn_msp_bucket_bits(uint64_t n_occ_total);

/* ---- multi-GPU with the collectives issued by the library (C++ host, NCCL over NVLink / NVSwitch) ---------------
 * One context per rank (= per GPU).  sn_mg_build_graph is the whole hot path over the ranks of the communicator:
 *   reads sharded over the ranks -> super-k-mers routed to the owner of their minimizer bucket (ONE alltoallv;
 *   owner(b) = b * N >> bits, the role of `shard % total_chunks`, lib/tada/src/cmd_shard_asm.rs:40) -> count + filter
 *   per owner -> the dictionary STAYS sharded: the neighbours of a rank's k-mers on other ranks are resolved by a
 *   query/answer exchange (cf. fix_sedge_exts, lib/tada/src/debruijn.rs:785-826), the unipath chains are cut where they
 *   cross ranks and stitched over a gathered table of those cuts, the edge bases are completed by one all-reduce ->
 *   every rank holds all edges and the whole HyperBasevector; ReadPaths of the rank's own reads (with_paths: the
 *   finished k-mer table is gathered first).
 * sn_comm_init_nccl: `unique_id128` = the 128 bytes rank 0 got from sn_nccl_unique_id, handed to every rank by the
 * launcher (MPI / torch.distributed / a file).  sn_comm_init_local: n ranks as n contexts of ONE process on one
 * device, one host thread per rank -- test infrastructure for the rank logic on a single-GPU box.                  */
int   sn_nccl_unique_id(void* out128);
int   sn_comm_init_nccl(sn_ctx* ctx, int rank, int n_ranks, const void* unique_id128);
void* sn_local_group_create(int n_ranks);
void  sn_local_group_destroy(void* group);
void  sn_local_group_abort(void* group);          /* a rank failed: the other ranks' collectives return an error instead of waiting */
int   sn_comm_init_local(sn_ctx* ctx, void* group, int rank);
void  sn_comm_free(sn_ctx* ctx);
int   sn_mg_build_graph(sn_ctx* ctx, const sn_params* params, int with_paths);
/* 1 while the context holds only its rank's shard of the k-mer table (sn_get_kmers etc. then return that shard) */
int   sn_mg_dict_is_sharded(const sn_ctx* ctx);

/* ---- measurement ------------------------------------------------------------------------ */
/* Device time (CUDA events on the context's stream) of the most recent run of a stage or
 * kernel group, in milliseconds; names: "h2d","goodlen","msp_hist","msp_scatter","bucket_count","make_dict",
 * "prune","edges","hbv_dev","hbv_host","hbv_csr","path".  Returns a negative value for an unknown name.          */
double sn_stage_ms(const sn_ctx* ctx, const char* name);
/* number of kernel launches issued by this context so far */
uint64_t sn_kernel_launches(const sn_ctx* ctx);
/* the cudaStream_t every kernel and copy of this context is issued on (for event timing) */
void* sn_stream(const sn_ctx* ctx);

/* ---- host-side format helpers (no device needed) ------------------------------------------ */
/* PQVec codec; out must hold at least 2*n+8 bytes; returns bytes written */
uint64_t sn_pqvec_encode(const uint8_t* quals, uint32_t n, uint8_t* out);
uint32_t sn_pqvec_decode(const uint8_t* pq, uint64_t pq_bytes, uint8_t* out, uint32_t cap);
/* Pack n_reads reads given as one base code (0..3) per byte into the .fastb variable-data
 * layout and PQVec-encode their quals with `threads` host threads.  Buffers are malloc'ed
 * and owned by the caller (release with sn_free). */
int sn_pack_reads(uint64_t n_reads, const uint8_t* codes, const uint8_t* quals, const uint64_t* off, int threads,
                  uint8_t** bases, uint64_t** base_off, uint32_t** len, uint8_t** pq, uint64_t** pq_off);
void sn_free(void* p);
/* tmp.paths (feudal ReadPathVec) from arrays: the ranks' ReadPaths of a multi-GPU job, concatenated in read order by the caller */
int sn_write_paths_arrays(const char* path, uint64_t n_reads, const int32_t* offset, const uint64_t* path_off, const int32_t* edges);
/* write the reference's input files from the in-memory layout above */
int sn_write_read_files(const char* fastb, const char* qualp, const char* bci, uint64_t n_reads,
                        const uint8_t* bases, const uint64_t* base_off, const uint32_t* len,
                        const uint8_t* pq, const uint64_t* pq_off, const int32_t* bc);

#ifdef __cplusplus
}
#endif
#endif /* SUPERNOVA_B200_H_ */
