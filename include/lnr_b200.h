/* lnr_b200.h -- C ABI of the B200-native approximate-map path of `linear filter` (xp3i4/linear).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Every entry point names the reference interface it replaces (paths relative to the reference tree).
 * INTEGRATION.md shows the shim a maintainer adds on the reference side (SeqAn String <-> these buffers).
 *
 * Conventions
 *   - every function returns 0 on success, a negative LNR_E_* code otherwise; lnr_last_error() gives text.
 *     The reference's functions on this path always return 0 and never throw (pmpfinder.cpp:2709).
 *   - sequences are the reference's Dna5 ordinals, one byte per base: A,C,G,T,N = 0..4 (base.h:106-116).
 *   - `threads_sem` is the reference's -t. It is a SEMANTIC parameter of the index and of the genome
 *     feature builder (chunk seams: index_util.cpp:1654-1670, pmpfinder.cpp:603-650), not a degree of
 *     parallelism here.
 *   - cords / anchors / hits are the reference's 64-bit encodings (include/cords.h:23-39).
 *   - there is no CPU fallback: without a CUDA device every call fails with LNR_E_CUDA.
 */
#ifndef LNR_B200_H
#define LNR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LNR_OK 0
#define LNR_E_CUDA (-1)      /* CUDA runtime error or no device */
#define LNR_E_ARG (-2)       /* invalid argument */
#define LNR_E_CAPACITY (-3)  /* caller buffer too small; *needed outputs say how much */
#define LNR_E_UNSUPPORTED (-4)
#define LNR_E_LIMIT (-5)     /* input outside the reference's own limits (contig >= 2^30, > 1024 contigs, read >= 2^20) */

typedef struct lnr_ctx lnr_ctx;
typedef struct lnr_genome lnr_genome;
typedef struct lnr_feats lnr_feats;
typedef struct lnr_index lnr_index;

/* ---- context -------------------------------------------------------------------------------------*/
int lnr_ctx_create(int device, lnr_ctx ** out);
void lnr_ctx_destroy(lnr_ctx *);
const char * lnr_last_error(const lnr_ctx *);
/* record a CUDA-event pair around every kernel launched through this context (bench/roofline) */
int lnr_ctx_set_profiling(lnr_ctx *, int on);
/* kernels launched since the last reset: n names/ms/counts copied (up to cap); returns total distinct */
int lnr_ctx_kernel_times(lnr_ctx *, int cap, const char ** names, float * total_ms, uint64_t * launches);
int lnr_ctx_reset_kernel_times(lnr_ctx *);

/* ---- genome: replaces Mapper::loadGenomes' in-memory StringSet<String<Dna5>> (mapper.cpp:389) --------*/
int lnr_genome_upload(lnr_ctx *, uint32_t n_contigs, const uint8_t * const * dna5, const uint64_t * len,
                      lnr_genome ** out);
/* same, but the bases are already in device memory, contigs back to back in one buffer (bench "value" leg) */
int lnr_genome_from_device(lnr_ctx *, uint32_t n_contigs, const uint8_t * dev_concat, const uint64_t * len,
                           lnr_genome ** out);
void lnr_genome_destroy(lnr_genome *);

/* ---- genome features: createFeatures(StringSet<String<Dna5>>&, StringSet<FeaturesDynamic>&, int, unsigned)
 *      pmpfinder.cpp:775 -> createFeatures2_48 (parallel builder) :589 for feature_type 2 = 2-mer/48 (-f 2), createFeatures1_32
 *      :393 for feature_type 1 = 1-mer/32 (-f 1, one short per 16 bases). With -f 1 the reference leaves the last entries of
 *      a string unwritten and reads up to 10 entries past its end; here both are defined as 0. */
int lnr_features_build(lnr_ctx *, const lnr_genome *, int feature_type, unsigned threads_sem, lnr_feats ** out);
int lnr_features_count(const lnr_feats *, uint32_t contig, uint64_t * n_entries);
/* dst: int32[3*n] (int96 entries) for feature_type 2, int16[n] for feature_type 1 */
int lnr_features_download(const lnr_feats *, uint32_t contig, void * dst, uint64_t cap_entries, uint64_t * n_entries);
void lnr_features_destroy(lnr_feats *);

/* ---- index: createIndexDynamic(..., threads, efficient) index_util.cpp:2478 -> createDIndex :1628 -------
 *      index_type 1 = DIndex (-i 1, the default), 2 = HIndex (-i 2) */
int lnr_index_build(lnr_ctx *, const lnr_genome *, int index_type, unsigned threads_sem, lnr_index ** out);
/* DIndex contents: dir = int32[2^26+1], hs = uint64[n_hs] (include/index_util.h:99-120).
 * Pass NULL for dir/hs to query n_hs only. */
int lnr_index_export_dindex(const lnr_index *, int32_t * dir, uint64_t * hs, uint64_t hs_cap, uint64_t * n_hs);
/* HIndex contents (index_type 2; include/index_util.h:139-248): ysa words (heads + bodies + two zero terminators),
 * emptyDir, and the open-addressing directory as (val1, val2) pairs sorted by val1 -- the reference's own physical
 * table layout differs from run to run, so only the key -> value map is part of the contract.
 * Any pointer may be NULL (size query). */
int lnr_index_export_hindex(const lnr_index *, uint64_t * ysa, uint64_t ysa_cap, uint64_t * n_ysa, uint64_t * keyvals /* 2*n_kv */,
                            uint64_t kv_cap, uint64_t * n_kv, uint64_t * empty_dir, uint64_t * table_len);
/* Multi-GPU build: shard s of n keeps only the minimizers X in [s*2^26/n, (s+1)*2^26/n). The result is a
 * DIndex whose dir counts only those buckets and whose hs holds only their records; the > 400 omission rule and the
 * order inside a bucket are bucket-local, so concatenating the shards' hs in shard order and rebasing each shard's dir
 * slice by the records of the shards before it gives exactly the index lnr_index_build produces. The exchange
 * (one all-gather of counts, hs slices and dir slices) is done by the caller over NCCL on the device buffers below. */
/* (index_type 1 only; the HIndex shards through lnr_index_build_sharded below.) */
int lnr_index_build_shard(lnr_ctx *, const lnr_genome *, int index_type, unsigned threads_sem, unsigned shard, unsigned n_shards,
                          lnr_index ** out);
/* The same build with the exchange inside the library (lnr_index_build_sharded). One lnr_comm per rank wraps
 * an NCCL communicator: either created here from a unique id the caller distributes (lnr_nccl_unique_id on one rank,
 * lnr_comm_create on all), or an existing ncclComm_t of the host program (lnr_comm_from_nccl; not destroyed by
 * lnr_comm_destroy). NCCL is bound at run time (dlopen libnccl.so.2); LNR_E_UNSUPPORTED when it is not installed.
 * lnr_index_build_sharded is collective: every rank calls it with the same genome and threads_sem. Every rank hashes the
 * genome and derives, from the same sampled minimizer histogram, the same n bucket ranges of ~equal record count (equal-
 * width ranges would leave 99.6 % of a random genome's records on the first of two ranks); rank r keeps the minimizers of
 * range r, learns all ranks' record counts (one 8-byte all-gather),
 * builds its buckets directly inside its slice of the final hs / dir arrays, and one grouped exchange (in-place
 * ncclBroadcast of every rank's hs slice and dir slice at their displacements -- no padding, no staging copy) completes
 * the identical DIndex on every rank.
 * index_type 2 (HIndex, createHIndex index_util.cpp:1471): the same scheme on the 18-bit X axis of the 17-base shape --
 * every rank counts the (head, body) pairs per X over the whole genome, derives the same n X ranges of equal pair count,
 * sorts and assembles the blocks of its own range straight into its slice of the final ysa (a block is one X, so no
 * block straddles a cut), one 16-byte all-gather of (pairs, blocks) gives the displacements, one grouped exchange of the
 * ysa slices completes the array on every rank, and the directory is derived locally from the head words of the assembled
 * ysa. The result equals lnr_index_build's for any number of ranks. */
typedef struct lnr_comm lnr_comm;
int lnr_nccl_unique_id(uint8_t id[128]);
int lnr_comm_create(lnr_ctx *, const uint8_t id[128], int rank, int n_ranks, lnr_comm ** out);
int lnr_comm_from_nccl(lnr_ctx *, void * nccl_comm /* ncclComm_t */, int rank, int n_ranks, lnr_comm ** out);
void lnr_comm_destroy(lnr_comm *);
int lnr_index_build_sharded(lnr_ctx *, const lnr_genome *, int index_type, unsigned threads_sem, lnr_comm *, lnr_index ** out);
/* The X ranges the sharded HIndex build uses (host arithmetic, no device needed): from the number of (head, body) pairs per
 * X, cuts[0..n_ranks] with cuts[0] = 0, cuts[n_ranks] = n_x and rank r owning X in [cuts[r], cuts[r+1]); cut r is the first
 * X at which the running pair count reaches r/n_ranks of all pairs. */
int lnr_hindex_shard_cuts(const uint32_t * pairs_per_x, uint32_t n_x, int n_ranks, uint32_t * cuts);
/* device-to-device copies of the index arrays (dev_dir: int32[2^26+1], dev_hs: uint64[>= n_hs]; either may be NULL) */
int lnr_index_export_dindex_device(const lnr_index *, int32_t * dev_dir, uint64_t * dev_hs, uint64_t hs_cap);
/* wraps assembled device arrays into an index (copies them) */
int lnr_index_from_device(lnr_ctx *, const int32_t * dev_dir, const uint64_t * dev_hs, uint64_t n_hs, lnr_index ** out);
/* ---- index serialisation. The reference has no on-disk index: `linear` hashes the genome again on every
 * run (createIndexDynamic, 21 s at 3.1 Gbase). lnr_index_save writes what createIndexDynamic produced -- DIndex: dir + hs;
 * HIndex: ysa + the directory table + emptyDir -- behind a 64-byte header (magic "LNRIDX1", index type, element counts, one
 * 64-bit checksum per array); lnr_index_load reads it back onto ctx's device and re-derives the lookup arrays of the
 * seeding kernels. The file does not name the genome or -t it was built from: that pairing is the caller's (store the
 * file next to the FASTA). Errors: LNR_E_ARG = cannot open / not an index file / checksum or size mismatch. */
int lnr_index_save(const lnr_index *, const char * path);
int lnr_index_load(lnr_ctx *, const char * path, lnr_index ** out);
void lnr_index_destroy(lnr_index *);

/* ---- per-read approximate mapping -------------------------------------------------------------------
 * Replaces, for every read j of a block (Mapper::p_calRecords mapper.cpp:438-447, map_ :825-849):
 *     _compltRvseStr(read, comStr); createFeatures(read) ; createFeatures(comStr);      base.cpp:335, pmpfinder.cpp:724
 *     apxMap(index, read, anchors, crhit, f1, f2, apx_gaps, cords_str, cords_end, cords_info,
 *            f_chain=1, pm_g, pm_pmp);                                                   pmpfinder.cpp:2709
 * Reads with length <= 200 get an empty cord list (mapper.cpp:440). cords_end[i] = cords_str[i] +
 * ((W<<20)|W), W = 96 for -f 2 and 192 for -f 1, is reconstructed by the caller (pmpfinder.cpp:2801). */
typedef struct lnr_params
{
    int preset;       /* -p: 0 => thd_stop_chain_len_ratio 0.7, 1/2 => 0 (mapper.cpp:174-197); code default 1 */
    int feature_type; /* -f: 2 or 1; 0 = whatever the genome features were built with */
    int no_chain;     /* 1 = -c 0: apxMap with f_chain = 0 (alg_type 1, pmpfinder.cpp:2773-2787): anchors grouped by sorting
                       * (getDAnchorList :2185, getDHitList :2246), path_dst_1 :1269, a second attempt at k-mer step 7 when the
                       * longest block covers < 0.7 of the read; 0 = the default chaining path (f_chain = 1) */
    int gdl_state;    /* -c 0 only. The reference keeps one PMPParms per thread and this mode changes it for good:
                       * GetDHitListParms starts as (thd_list_n 20, thd_best_n 1) and is (10, 999) from the first read that
                       * needed the second attempt on (toggle(0), :2784). 0 = the state as constructed, 1 = the later one. */
    int reserved[4];
} lnr_params;

/* optional stage checkpoints; any pointer may be NULL. Offsets arrays hold n_reads+1 entries. */
typedef struct lnr_debug_out
{
    uint64_t * raw_anchors; uint64_t raw_anchors_cap; uint64_t * raw_anchors_off;   /* getDIndexMatchAll, primary call, without the anchors[0] sentinel */
    uint64_t * hits;        uint64_t hits_cap;        uint64_t * hits_off;          /* hits after getAnchorHitsChains (primary call) */
    uint64_t * cords1;      uint64_t cords1_cap;      uint64_t * cords1_off;        /* cords after the first apxMap_ */
} lnr_debug_out;

/* host buffers in, host buffers out (the drop-in call; copies are part of the call) */
int lnr_apxmap_batch(lnr_ctx *, const lnr_index *, const lnr_feats * f2, const lnr_params *,
                     uint32_t n_reads, const uint8_t * dna5_concat, const uint64_t * read_off /* n+1 */,
                     uint64_t * cords_str_concat, uint64_t * cords_off /* n+1 */, uint64_t cords_capacity,
                     lnr_debug_out * dbg /* may be NULL */);
/* The same call with the reads 2-bit packed (BASELINE north_star: "2-bit-packed"): base i of the batch (reads back to
 * back, read_off in bases) is bits 2(i&3)..2(i&3)+1 of packed2[i >> 2], A,C,G,T = 0..3. n_mask is NULL for a batch
 * without N, else a bitmap with bit (i & 7) of n_mask[i >> 3] set where base i is N (its two packed bits are 0). A
 * quarter of the bytes of the Dna5 form cross PCIe; the device expands them once. Results are identical to
 * lnr_apxmap_batch on the corresponding Dna5 string. lnr_pack_dna5 is the host-side converter a shim calls once per
 * read block (String<Dna5> -> packed2 / n_mask; buffers of (n+3)/4 and (n+7)/8 bytes; *has_n tells whether the bitmap is
 * needed). */
int lnr_pack_dna5(const uint8_t * dna5, uint64_t n_bases, uint8_t * packed2, uint8_t * n_mask, int * has_n);
int lnr_apxmap_batch_packed(lnr_ctx *, const lnr_index *, const lnr_feats * f2, const lnr_params *,
                            uint32_t n_reads, const uint8_t * packed2, const uint8_t * n_mask /* may be NULL */,
                            const uint64_t * read_off /* n+1, in bases */,
                            uint64_t * cords_str_concat, uint64_t * cords_off /* n+1 */, uint64_t cords_capacity,
                            lnr_debug_out * dbg /* may be NULL */);
/* device buffers in, device buffers out (inputs already resident in HBM; bench "value" leg).
 * n_cords_total (host) receives the number of cords written. */
int lnr_apxmap_batch_device(lnr_ctx *, const lnr_index *, const lnr_feats * f2, const lnr_params *,
                            uint32_t n_reads, const uint8_t * dev_dna5_concat, const uint64_t * host_read_off,
                            uint64_t * dev_cords, uint64_t * dev_cords_off, uint64_t cords_capacity,
                            uint64_t * n_cords_total);
/* algorithmic-byte counters of the last batch: S seeds, H bucket records scanned, A raw anchors,
 * Hits, W window candidates evaluated, C cords */
int lnr_last_batch_counters(lnr_ctx *, uint64_t counters[8]);
/* fallback-path diagnostics of the last batch (tests assert that the rare paths really ran): 0 tasks taken by the big-arena
 * hit pass, 1 reads finished by the big-arena finish pass, 2 seeding samples re-scanned because their match list did not
 * fit the pool, 3 tasks with more anchors than a warp's share of the scratch holds that the big-arena warps of the hit-section
 * kernels served (the "heavy lane"), 4..7 reserved */
int lnr_last_batch_diag(lnr_ctx *, uint64_t diag[8]);
/* profiling aid: SM cycles spent per stage of the warp-per-read pipeline in the last batch (lane 0, summed over warps):
 * 0 binning, 1 ascending sort, 2 run filter, 3 x sort, 4 chaining DP, 5 traceback, 6 hit blocks, 7 hit window filter,
 * 8 window extension, 9 clean/gaps, 10 cord block chaining, 11 reads needing the sequential tie-order sort, 12 reads */
int lnr_last_batch_stage_cycles(lnr_ctx *, uint64_t cycles[16]);

/* read features of one read (both strands), for the host gap stage and for parity tests:
 * createFeatures(begin(read), end(read), f1[s]) pmpfinder.cpp:724 -> createFeatures2_48 (serial) :556 */
int lnr_read_features(lnr_ctx *, const uint8_t * dna5, uint64_t len, int feature_type,
                      void * dst_fwd, void * dst_rev, uint64_t cap_entries, uint64_t * n_entries);

/* ---- read ingest -------------------------------------------------------------------------------
 * FASTA / FASTQ text -> Dna5 ordinals back to back + read offsets + the span of every record id inside the text, on the
 * device. Replaces seqan's readRecords + Dna5 conversion in front of p_calRecords (loadRecords base.cpp:154,
 * readRecords4FinPool2_ parallel_io.cpp:466; Dna5 table: A/a C/c G/g T/t/U/u -> 0..3, any other byte -> 4).
 * FASTA: '>' at a line start opens a record, sequence lines may be wrapped, '\r' and ' ' are dropped. FASTQ: four lines
 * per record. The text must start with '>' or '@'. cut_id_at_space: ids end at their first blank (loadRecords). */
typedef struct lnr_reads lnr_reads;
int lnr_reads_parse(lnr_ctx *, const char * text, uint64_t n_bytes, int cut_id_at_space, lnr_reads ** out);
/* text already in device memory; first_byte = text[0] (decides the format) */
int lnr_reads_parse_device(lnr_ctx *, const char * dev_text, uint64_t n_bytes, int first_byte, int cut_id_at_space, lnr_reads ** out);
int lnr_reads_info(const lnr_reads *, uint64_t * n_reads, uint64_t * total_bases);
/* any destination may be NULL; read_off has n_reads + 1 entries, id_off / id_len index the ORIGINAL text */
int lnr_reads_download(const lnr_reads *, uint8_t * bases, uint64_t * read_off, uint64_t * id_off, uint32_t * id_len);
/* device views for lnr_apxmap_batch_device (bases) -- valid until lnr_reads_destroy */
int lnr_reads_device(const lnr_reads *, const uint8_t ** dev_bases, const uint64_t ** dev_read_off);
/* lnr_apxmap_batch over reads [first, first + n_reads) of a parsed set: no host copy of the bases at all */
int lnr_apxmap_reads(lnr_ctx *, const lnr_index *, const lnr_feats *, const lnr_params *, const lnr_reads *, uint32_t first,
                     uint32_t n_reads, uint64_t * cords, uint64_t * cords_off, uint64_t cords_capacity, lnr_debug_out * dbg);
void lnr_reads_destroy(lnr_reads *);

/* ---- SAM* / BAM* record construction ---------------------------------------------------------------
 * Replaces cords2BamLink (f_io.cpp:883; cord2cigar_ :758, ifCreateNew_ :673, insertNewBamRecord align_util.cpp:301) for a block
 * of reads, the step Mapper::p_calRecords runs on the cords when no alignment is requested (mapper.cpp:463-468): the cords
 * of a read become its records -- contig, leftmost position, flag (16 reverse strand, 2048 supplementary), the reference's
 * cigar* (operations = I D X S, adjacent equal operations merged) and its three score counters. cords may be what
 * lnr_apxmap_batch returned or what the host's mapGaps made of them; cords_end = cords + ((window << 20) | window).
 * Outputs: rec_off / cigar_off (n_reads + 1 entries each), recs[rec_off[r] .. rec_off[r+1]) the records of read r, their
 * cigar_begin / cigar_end relative to cigar_off[r]; a cigar element is (operation << 32) | count. LNR_E_CAPACITY leaves the
 * needed sizes in rec_off[n_reads] and cigar_off[n_reads]. fillBamRecords (names, SEQ, SA:Z text) stays host code. */
typedef struct lnr_bam_parms
{
    uint32_t window;      /* 96 (-f 2) or 192 (-f 1); 0 = 96 */
    uint32_t reserved;
    uint64_t thd_large_x; /* 8000, mapper.cpp:466 */
    int64_t thd_di, thd_x;/* FIOParms: (1 << 60) - 1 each by default, 80 / 200 with -p 0 (f_io.cpp:16, mapper.cpp:185) */
} lnr_bam_parms;
typedef struct lnr_bam_rec
{
    int32_t rid, begin_pos;
    uint32_t flag;
    int32_t s1, s2, s3;             /* BamAlignmentRecordLinkScore */
    uint32_t cigar_begin, cigar_end;
} lnr_bam_rec;
int lnr_cords_to_records(lnr_ctx *, uint32_t n_reads, const uint64_t * cords, const uint64_t * cords_off /* n+1 */,
                         const uint64_t * read_len /* n */, const lnr_bam_parms *, lnr_bam_rec * recs, uint64_t rec_cap,
                         uint64_t * rec_off /* n+1 */, uint64_t * cigars, uint64_t cigar_cap, uint64_t * cigar_off /* n+1 */);

/* self-test of the warp-cooperative std::sort emulation used by chainAnchorsHits (pmpfinder.cpp:2465 sorts by
 * AnchorX only, so tied anchors end up in libstdc++'s introsort order): sorts n 64-bit records in place, ascending by
 * their high 32 bits (30 significant), with one warp. Parity tests compare the result with std::sort on the host. */
int lnr_selftest_sort(lnr_ctx *, uint64_t * records, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
