/*
 * floria_oracle.h — C API of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a C++ restatement of floria's read-to-haplotype scoring / local clustering path
 * (see floria_oracle.cpp for the per-function reference citations).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or expected outputs for this path
 *  and cannot be compiled here (no Rust toolchain), so the oracle is pinned only by
 * line-by-line correspondence with the cited Rust and by the hand-derived known answers of SURVEY.md
 * Appendix D (tests/test_oracle_known_answers.py).
 *
 * The POD structs (fb_frags, fb_params, fb_block_results, fb_parts) are shared with the product
 * boundary so the parity tests can compare field by field.
 */
#ifndef FLORIA_ORACLE_H
#define FLORIA_ORACLE_H
#include "../include/floria_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void orc_phred_lut(float *out256);
double orc_stable_binom_cdf_p_rev(uint64_t n, uint64_t k, double p, double div_factor);
double orc_log_sum_exp(const double *probs, uint64_t n);
double orc_mec_threshold(uint32_t ploidy, double epsilon, uint32_t sensitivity);

int64_t orc_get_range_with_lengths(const uint64_t *snp_to_genome_pos, uint64_t n_snps, uint64_t block_length,
                                   uint64_t overlap_len, double minimal_density, uint32_t *lo, uint32_t *hi,
                                   uint64_t cap);
int64_t orc_find_reads_in_interval(uint32_t start, uint32_t end, uint64_t n_reads, const uint32_t *first,
                                   const uint32_t *last, uint32_t *out_ids, uint64_t cap);

int orc_score_reads(const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap, uint32_t ploidy,
                    const fb_params *, double *same, double *diff);
/* variant (a4) utils_frags.rs:77-108 distance_read_haplo; outputs rounded usize pairs */
int orc_score_reads_noeps(const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                          uint32_t ploidy, const fb_params *, uint64_t *same, uint64_t *diff);
int orc_hap_block_from_partition(const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                                 uint32_t ploidy, int use_qual, const fb_params *, uint32_t pos_lo, uint32_t n_pos,
                                 double *counts, uint8_t *key_mask);
int orc_get_mec_stats_epsilon(const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                              uint32_t ploidy, int use_phred, const fb_params *, double *bases, double *errors);
int orc_beam_search_phasing(const fb_frags *, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy,
                            const fb_params *, uint8_t *hap_out, double *best_score, double *tap_same,
                            double *tap_diff, double *tap_logp, uint64_t tap_cap, uint64_t *tap_n);
int orc_optimize_clustering(const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap_in,
                            uint32_t ploidy, const fb_params *, uint8_t *hap_out, double *score,
                            uint32_t *n_rounds);
/* mirror of fb_phase_block (sel == NULL: every read) */
int orc_phase_block(const fb_frags *, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy, const fb_params *,
                    uint8_t *hap_out, double *mec_bases, double *mec_errors, fb_block_phase *out);
/* n_threads workers over blocks, mirroring rayon's par_iter (graph_processing.rs:345-362). */
int orc_phase_blocks(const fb_frags *, uint64_t n_blocks, const uint32_t *blk_lo, const uint32_t *blk_hi,
                     const fb_params *, uint32_t n_threads, fb_block_results **out);
void orc_free_block_results(fb_block_results *);
int orc_process_reads_for_final_parts(const fb_frags *, uint64_t n_parts, const uint64_t *part_ptr,
                                      const uint32_t *part_reads, const uint32_t *range_lo,
                                      const uint32_t *range_hi, const fb_params *, fb_parts **out);
void orc_free_parts(fb_parts *);
int orc_get_hapq(const fb_frags *, uint64_t n_parts, const uint64_t *part_ptr, const uint32_t *part_reads,
                 const uint32_t *range_lo, const uint32_t *range_hi, const uint64_t *snp_to_genome_pos,
                 uint64_t n_snps, const fb_params *, uint8_t *hapq, double *rel_err, double *avg_err);
int orc_update_hap_graph(const fb_frags *, uint64_t n_cols, const uint64_t *col_ptr, const uint64_t *node_ptr,
                         const uint32_t *node_reads, const uint32_t *node_lo, const uint32_t *node_hi,
                         const fb_params *, double *out_weights);
const char *orc_last_error(void);
int orc_heap_trace(const double *scores, int n, int width, int *data_out, int *sorted_out);

#ifdef __cplusplus
}
#endif
#endif
