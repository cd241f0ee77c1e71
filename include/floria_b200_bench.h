/*
 * floria_b200_bench.h — measurement helpers of libfloria_b200.so.  NOT part of the reference-facing boundary
 * (include/floria_b200.h): these exist so bench.py / tests can build the 100k x 50k "roofline" block of
 * BASELINE.json configs[2] directly in HBM (5e9 cells do not fit a host CSR round trip) and time the
 * bandwidth-bound kernels on it.
 */
#ifndef FLORIA_B200_BENCH_H
#define FLORIA_B200_BENCH_H
#include "floria_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Full-span synthetic reads (first = 1, last = n_snps) generated on the device with the counter-based PRNG of
 * floria_b200/synth.py (make_contig(..., full_span=True, qual_mode="long")): identical cells for identical
 * (seed, truth, nall, src).  truth: [ploidy*n_snps] alleles, nall: [n_snps] alleles per column (2 or 3),
 * src: [n_reads] source haplotype of each read. */
int fb_bench_synth_dense(fb_ctx *, uint64_t n_reads, uint32_t n_snps, uint32_t ploidy, uint64_t seed, double present,
                         double flip, const uint8_t *truth, const uint8_t *nall, const uint8_t *src, fb_dfrags **out);

/* One block made of ALL reads of `df`, partition `hap`: builds the haplotype table (k_hist), then `iters` times
 * runs the scoring sweep of opt_iterate (k_sweep, every read x every haplotype) and a histogram rebuild (k_hist),
 * timing each launch with CUDA events.  cells = stored cells streamed by one launch. */
int fb_bench_sweep_hist(fb_ctx *, const fb_dfrags *df, uint32_t ploidy, const uint8_t *hap, const fb_params *,
                        uint32_t iters, float *sweep_ms, float *hist_ms, uint64_t *cells);

/* Test: one block made of ALL reads of `df` with partition `hap`: the raw count table of k_hist
 * ([ploidy][n_pos][4] words, value in the low 62 bits in units of 2^-26, bit 62 = allele key present) and the
 * SCORE-mode sweep of every read against every haplotype (same / diff weight sums in units of 2^-26, number of cells on
 * empty positions; each [n_reads][ploidy]).  With all output pointers NULL only *n_pos is returned (size query).
 * Used by the full-size property tests (checksums that tie k_hist to k_sweep, linearity in the partition). */
int fb_bench_block_tables(fb_ctx *, const fb_dfrags *df, uint32_t ploidy, const uint8_t *hap, const fb_params *,
                          uint64_t *n_pos, uint64_t *counts, int64_t *same_q26, int64_t *diff_q26, uint32_t *n_empty);

/* The host CSR (fb_frags layout) of a resident contig: row_ptr [n_reads+1], first / last [n_reads], pos / allele / qual
 * [nnz] (query nnz with fb_dfrags_nnz).  bench.py builds the HOST buffers of the end-to-end leg from the device-generated
 * block this way. */
int fb_bench_export_csr(fb_ctx *, const fb_dfrags *df, uint64_t *row_ptr, uint32_t *first, uint32_t *last, uint32_t *pos,
                        uint8_t *allele, uint8_t *qual);
uint64_t fb_dfrags_nnz(const fb_dfrags *);
uint64_t fb_dfrags_n_reads(const fb_dfrags *);

/* Debug/test: copy the packed planes of a resident contig back to the host (any pointer may be NULL). */
int fb_bench_download_planes(fb_ctx *, const fb_dfrags *df, uint64_t *n_groups, uint8_t *qual /*[16*ng]*/,
                             uint32_t *allele /*[ng]*/, uint16_t *present /*[ng]*/);
#ifdef __cplusplus
}
#endif
#endif
