/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the approximate-map hot path of xp3i4/linear (`linear filter`).
 * Plain sequential C++ (std::vector + libstdc++ std::sort), written from the behaviour of the
 * reference; every function cites the reference file:line it follows (paths relative to
 * /root/reference). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (linear_b200/) never does.
 *
 * Pinning: the reference ships no golden vectors, so this oracle is pinned
 * against the reference itself, compiled unmodified into oracle/_ref/libref_harness.so by
 * oracle/build_ref.sh; tests/test_oracle_vs_ref.py compares every stage on seeded inputs and
 * tests/golden/ holds digests produced by that real reference.
 */
#ifndef LNR_ORACLE_H
#define LNR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* index_type 1 = DIndex (-i 1), 2 = HIndex (-i 2); feature_type 2 = 2-mer/48 (-f 2), 1 = 1-mer/32 (-f 1);
 * threads = the reference's -t (semantic, index_util.cpp:1654); preset = -p (0: stop ratio 0.7, else 0). */
orc_ctx * orc_create(int n_contigs, const uint8_t * const * dna5, const uint64_t * lens,
                     int index_type, int feature_type, int threads, int preset, int build_index);
void orc_destroy(orc_ctx *);

int64_t orc_dindex_dir(orc_ctx *, const int32_t ** dir);
int64_t orc_dindex_hs(orc_ctx *, const uint64_t ** hs);
int64_t orc_hindex_ysa(orc_ctx *, const uint64_t ** ysa, uint64_t * empty_dir);
int64_t orc_hindex_dir_kv(orc_ctx *, const uint64_t ** kv, uint64_t * table_len);
int64_t orc_genome_features(orc_ctx *, int contig, const int32_t ** f);
int64_t orc_read_features(orc_ctx *, const uint8_t * read, uint64_t len, int strand, const int32_t ** f);
/* stage numbering identical to ref_read_stage in oracle/ref_harness.cpp */
int64_t orc_read_stage(orc_ctx *, const uint8_t * read, uint64_t len, int stage,
                       uint64_t str, uint64_t end, int toggle, const uint64_t ** out);
int orc_map_batch(orc_ctx *, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int map_threads,
                  uint64_t * cords, uint64_t * cords_off, uint64_t cords_cap);

/* -c 0 (f_chain = 0, alg_type 1: getDAnchorList / getDHitList / path_dst_1). gdl_state 0 = a PMPParms as constructed
 * (thd_list_n 20, thd_best_n 1), 1 = as any earlier toggle(0) leaves it (10, 999). */
int orc_map_batch_c0(orc_ctx *, uint32_t n_reads, const uint8_t * bases, const uint64_t * read_off, int map_threads, int gdl_state,
                     uint64_t * cords, uint64_t * cords_off, uint64_t cords_cap);

#ifdef __cplusplus
}
#endif
#endif
