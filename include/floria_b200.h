/*
 * floria_b200.h — C-ABI boundary of the B200-native floria hot path.
 *
 * The reference (bluenote-1577/floria, Rust) has no FFI of its own.  The entry
 * points below are what a `floria` maintainer would bind (see INTEGRATION.md
 * for the `extern "C"` block and the patched call sites); each one names the
 * Rust `pub fn` (file:line under /root/reference) it replaces.  Only plain
 * pointers and sizes cross the boundary.  All pointers are HOST memory owned
 * by the caller unless a type says otherwise; the library never frees caller
 * memory; results are library-allocated POD arrays released by
 * fb_free_*().  Every function returns 0 on success and a non-zero status
 * otherwise (message via fb_last_error); nothing unwinds across the boundary.
 * There is NO CPU fallback: without a CUDA device fb_init fails.
 *
 * Conventions shared with the reference:
 *   - a read ("Frag", src/types_structs.rs:68-85) is identified by its
 *     counter_id == index in the contig's sorted Vec<Frag> (src/bin/floria.rs:289-293);
 *   - SNP positions are 1-based u32 (src/types_structs.rs:12, utils_frags.rs:461);
 *   - reads must arrive sorted by Frag::cmp (types_structs.rs:87-93):
 *     first_position asc, last_position desc, counter_id asc;
 *   - alleles are VCF allele indices 0..3 (file_reader.rs:702-710); quals are raw
 *     phred bytes 0..255 (file_reader.rs:711).
 */
#ifndef FLORIA_B200_H
#define FLORIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_OK 0
#define FB_ERR_ARG 1      /* invalid argument / unsorted input / allele > 3 */
#define FB_ERR_CUDA 2     /* CUDA runtime error */
#define FB_ERR_NODEV 3    /* no CUDA device: the product path has no CPU fallback */
#define FB_ERR_LIMIT 4    /* an internal capacity was exceeded */

typedef struct fb_ctx fb_ctx;
typedef struct fb_dfrags fb_dfrags; /* a contig's reads resident in HBM (packed planes) */

/* Knobs of src/types_structs.rs:20-51 (Options) and src/constants.rs:3-22 that
 * reach the hot path.  fb_params_default() fills the reference defaults
 * (parse_cmd_line.rs:34-43,160; constants.rs:3-6). */
typedef struct {
    double epsilon;              /* -e */
    double div_factor;           /* constants.rs:5   DIV_FACTOR = 0.25 */
    double prob_cutoff_ln;       /* constants.rs:6   ln(PROB_CUTOFF = 0.01) */
    uint32_t max_number_solns;   /* -n, default 10 (beam width) */
    uint32_t max_ploidy;         /* -p, default 5 */
    uint32_t num_iter_optimize;  /* constants.rs:3   NUM_ITER_OPTIMIZE = 20 */
    uint32_t ploidy_sensitivity; /* -s, 1..3, default 2 */
    uint32_t stopping_heuristic; /* !--no-stop-heuristic */
    uint32_t order_model;        /* 0 = canonical (ascending key) iteration order; only 0 is implemented */
    uint32_t block_length;       /* -l, bases; used by fb_get_hapq */
    uint32_t reassign_short;     /* hidden --reassign-short; must be 0 (not implemented on device) */
    const float *phred_lut;      /* 256 floats: (1f32 - 10f32.powf(q/-10)) (utils_frags.rs:702-711);
                                    NULL = computed by the library with the host libm */
} fb_params;

/* A contig's fragments in CSR form.  pos is ascending within a read. */
typedef struct {
    uint64_t n_reads;
    uint64_t nnz;
    const uint64_t *row_ptr; /* [n_reads+1] */
    const uint32_t *first;   /* [n_reads] Frag::first_position */
    const uint32_t *last;    /* [n_reads] Frag::last_position  */
    const uint32_t *pos;     /* [nnz] keys of seq_dict / qual_dict / positions */
    const uint8_t *allele;   /* [nnz] seq_dict values, 0..3 */
    const uint8_t *qual;     /* [nnz] qual_dict values */
} fb_frags;

/* Result of fb_phase_blocks: what get_local_hap_blocks (graph_processing.rs:103-304)
 * computes per block before HapNode construction. */
typedef struct {
    uint64_t n_blocks;
    uint32_t max_ploidy;
    uint32_t _pad;
    uint32_t *best_ploidy;     /* [n_blocks]; 0 = no reads in the interval (reference returns None, :129-131) */
    uint32_t *ploidies_run;    /* [n_blocks]; how many ploidies the reference loop evaluates before its break */
    double *mec_vector;        /* [n_blocks*max_ploidy] graph_processing.rs:117,159; 0 where not evaluated */
    double *expected_errors;   /* [n_blocks*max_ploidy] graph_processing.rs:196 */
    uint64_t *read_ptr;        /* [n_blocks+1] CSR over the arrays below */
    uint32_t *read_ids;        /* counter_id of each read of the block (ascending) */
    uint8_t *hap;              /* haplotype index of that read in the best-ploidy partition */
    /* work counters: stored cells visited by scoring / histogram passes and beam steps,
       counted only for the ploidies the reference loop would evaluate */
    uint64_t cells_sweep, cells_hist, cells_beam;
    uint64_t *block_cells;     /* [n_blocks] the three counters summed, per block (lets a batched call be split again) */
} fb_block_results;

/* Result of fb_process_reads_for_final_parts. */
typedef struct {
    uint64_t n_parts;
    uint64_t *part_ptr;   /* [n_parts+1] */
    uint32_t *read_ids;   /* ascending counter_id within each part */
    uint32_t *range_lo;   /* [n_parts] snp range (inclusive, 1-based) */
    uint32_t *range_hi;
} fb_parts;

/* Device-time breakdown of the last batched call on this context (CUDA events, ms). */
typedef struct {
    float upload_ms, pack_ms, beam_ms, sweep_ms, hist_ms, mec_ms, select_ms, total_ms, download_ms;
    uint64_t n_launches;      /* kernels of this library launched by the call */
    uint64_t n_sweep_launches, n_hist_launches, n_beam_launches;
    uint64_t sweep_cells, hist_cells; /* stored cells streamed by those launches (all ploidies) */
} fb_timings;

/* ---- context ---------------------------------------------------------------------------- */
int fb_init(int device, fb_ctx **out);
void fb_destroy(fb_ctx *);
const char *fb_last_error(const fb_ctx *); /* ctx may be NULL: last error of a failed fb_init */
void fb_params_default(fb_params *);
int fb_last_timings(const fb_ctx *, fb_timings *out);
/* stream the library launches on (a cudaStream_t), so callers can bracket it with their own events */
void *fb_stream(const fb_ctx *);

/* ---- several devices in one host process (fb_init(n_devices, device_ids)) ------------- */
/* Contigs are independent units (src/bin/floria.rs:229 loops over them; graph_processing.rs:345-362 over their
 * blocks), so a contig list is dealt to the devices by a static longest-processing-time-first queue and every device
 * phases its share in ONE batched call (its contigs concatenated along the SNP axis), driven by one host thread per
 * device.  The results come back to the caller's process in contig order: no collective is needed when one process
 * owns all devices.  (One process per GPU, e.g. under torchrun: every rank opens fb_init_multi(1, {local_rank}),
 * calls fb_lpt_assign to learn its share and gathers the partition records with NCCL, see floria_b200/shard.py.) */
typedef struct fb_multi fb_multi;
typedef struct fb_dcontigs fb_dcontigs; /* a contig list resident in the HBM of the devices that own its contigs */
int fb_init_multi(int n_devices, const int *device_ids, fb_multi **out);
void fb_destroy_multi(fb_multi *);
int fb_multi_size(const fb_multi *);
fb_ctx *fb_multi_ctx(fb_multi *, int i);
const char *fb_multi_last_error(const fb_multi *);
/* deterministic LPT: units by descending cost (ties by index) onto the least loaded bin (first minimum) */
void fb_lpt_assign(const double *costs, uint64_t n_units, uint32_t n_bins, uint32_t *owner_out);
/* blk_ptr [n_contigs+1] indexes blk_lo / blk_hi (1-based SNP ranges in each contig's own coordinates) */
int fb_contigs_upload(fb_multi *, uint64_t n_contigs, const fb_frags *contigs, const uint64_t *blk_ptr,
                      const uint32_t *blk_lo, const uint32_t *blk_hi, fb_dcontigs **out);
void fb_contigs_free(fb_multi *, fb_dcontigs *);
/* out [n_contigs]: one fb_block_results per contig (free each with fb_free_block_results), identical to what
 * fb_phase_blocks returns for that contig alone (except that cells_beam carries the sum of the three work counters: a
 * batched call only knows their per-block sum, block_cells); device_of [n_contigs] and device_ms [n_devices] (CUDA-event time of
 * each device's batch) may be NULL */
int fb_phase_contigs_resident(fb_multi *, const fb_dcontigs *, const fb_params *, fb_block_results **out,
                              uint32_t *device_of, float *device_ms);
int fb_phase_contigs(fb_multi *, uint64_t n_contigs, const fb_frags *contigs, const uint64_t *blk_ptr,
                     const uint32_t *blk_lo, const uint32_t *blk_hi, const fb_params *, fb_block_results **out,
                     uint32_t *device_of, float *device_ms);

/* ---- data movement ------------------------------------------------------------------------- */
/* Validates (sorted by Frag::cmp, allele <= 3, pos within [first,last]) and packs the CSR reads into
 * the HBM layout of DESIGN.md (2-bit allele planes, 8-bit quals, 1-bit presence, 16-position groups). */
int fb_frags_upload(fb_ctx *, const fb_frags *, fb_dfrags **out);
/* Several fragment sets as ONE resident contig: the reads of part k follow those of part k - 1 and their SNP positions
 * are shifted by pos_shift[k] (the caller chooses shifts that keep the merged reads sorted, e.g. past the last position of
 * the previous part).  The cells go from the caller's buffers straight to the device; this is how fb_contigs_upload batches
 * the contigs of a device without a host-side merge. */
int fb_frags_upload_parts(fb_ctx *, uint64_t n_parts, const fb_frags *parts, const uint32_t *pos_shift, fb_dfrags **out);
void fb_frags_free(fb_ctx *, fb_dfrags *);
uint64_t fb_dfrags_bytes(const fb_dfrags *); /* bytes of packed planes resident in HBM */

/* ---- host-side helpers that define the work units ------------------------------------------------ */
/* utils_frags.rs:405-463 get_range_with_lengths; returns the number of ranges (written up to cap). */
int64_t fb_get_range_with_lengths(const uint64_t *snp_to_genome_pos, uint64_t n_snps, uint64_t block_length,
                                  uint64_t overlap_len, double minimal_density, uint32_t *lo, uint32_t *hi,
                                  uint64_t cap);
/* local_clustering.rs:12-59 find_reads_in_interval (max_num_reads = usize::MAX); returns the count. */
int64_t fb_find_reads_in_interval(uint32_t start, uint32_t end, uint64_t n_reads, const uint32_t *first,
                                  const uint32_t *last, uint32_t *out_ids, uint64_t cap);

/* ---- batched hot path ----------------------------------------------------------------------------------- */
/* Replaces the body of the par_iter in generate_hap_graph (graph_processing.rs:345-362) minus HapNode
 * construction: for every block j, get_local_hap_blocks' read selection, ploidy loop
 * (beam_search_phasing -> optimize_clustering -> get_mec_stats_epsilon_no_phred) and stopping rule. */
int fb_phase_blocks(fb_ctx *, const fb_frags *, uint64_t n_blocks, const uint32_t *blk_lo, const uint32_t *blk_hi,
                    const fb_params *, fb_block_results **out);
/* Same, with the contig already resident in HBM (no host<->device copy of read data). */
int fb_phase_blocks_resident(fb_ctx *, const fb_dfrags *, uint64_t n_blocks, const uint32_t *blk_lo,
                             const uint32_t *blk_hi, const fb_params *, fb_block_results **out);
void fb_free_block_results(fb_block_results *);

/* One block at a FIXED ploidy on an explicit read list: the body of the ploidy loop of get_local_hap_blocks
 * (graph_processing.rs:140-162): beam_search_phasing -> optimize_clustering -> get_mec_stats_epsilon_no_phred.
 * This is how a block is phased whose reads find_reads_in_interval would drop (spans above 10000 SNPs,
 * local_clustering.rs:44), e.g. the 100k-read x 50k-SNP roofline block of BASELINE.json.  `sel` = ascending counter_ids
 * (NULL = every read of the contig).  hap_out [n_sel], mec_bases / mec_errors [ploidy] (the (good, bad) pairs of
 * graph_processing.rs:156-162) may be NULL. */
typedef struct {
    double beam_score;     /* score of the winning beam node */
    double opt_score;      /* optimize_clustering's returned score */
    uint32_t n_rounds;     /* accepted opt_iterate rounds */
    uint32_t ploidy;
    uint64_t cells_sweep, cells_hist, cells_beam; /* work counters, SURVEY.md section 8d */
} fb_block_phase;
int fb_phase_block(fb_ctx *, const fb_frags *, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy, const fb_params *,
                   uint8_t *hap_out, double *mec_bases, double *mec_errors, fb_block_phase *out);
int fb_phase_block_resident(fb_ctx *, const fb_dfrags *, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy,
                            const fb_params *, uint8_t *hap_out, double *mec_bases, double *mec_errors,
                            fb_block_phase *out);

/* ---- fine-grained entry points (same semantics as the preserved Rust pub fns) ---------------------------------- */
/* `sel` = ascending counter_ids of the reads of one block; `hap[i]` = haplotype of sel[i] (0..ploidy-1). */

/* utils_frags.rs:177 hap_block_from_partition(partition, true) followed by
 * utils_frags.rs:32 distance_read_haplo_epsilon_empty(read, block[h], eps) for every read x haplotype.
 * Outputs are [n_sel*ploidy], row-major by read. same_q26/diff_q26 are the exact weight sums in units
 * of 2^-26 (diff_q26 excludes the epsilon terms, n_empty counts them). Any output may be NULL. */
int fb_score_reads(fb_ctx *, const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                   uint32_t ploidy, const fb_params *, double *same, double *diff, int64_t *same_q26,
                   int64_t *diff_q26, uint32_t *n_empty);

/* utils_frags.rs:160-184 set_to_seq_dict / hap_block_from_partition as a dense table over SNP positions
 * [pos_lo, pos_lo+n_pos): counts[h][p][a] (f64) and key_mask[h][p] (bit a set = allele key present). */
int fb_hap_block_from_partition(fb_ctx *, const fb_frags *, uint64_t n_sel, const uint32_t *sel,
                                const uint8_t *hap, uint32_t ploidy, int use_qual, const fb_params *,
                                uint32_t pos_lo, uint32_t n_pos, double *counts, uint8_t *key_mask);

/* local_clustering.rs:218-260 get_mec_stats_epsilon (use_phred=1, on hap_block_from_partition(.., true)) and
 * local_clustering.rs:187-215 get_mec_stats_epsilon_no_phred (use_phred=0). bases/errors are [ploidy]. */
int fb_get_mec_stats_epsilon(fb_ctx *, const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                             uint32_t ploidy, int use_phred, const fb_params *, double *bases, double *errors);

/* global_clustering.rs:10-179 beam_search_phasing(vec![empty; ploidy], reads, eps, div_factor, cutoff,
 * max_number_solns, true, false). hap_out[i] = set index sel[i] lands in; best_score = the winning node's score.
 * tap_* (optional, may be NULL) receive the first `tap_cap` (same, diff, log-p) triples in evaluation order
 * (step, node in heap order, haplotype) for tolerance checks of the log-likelihoods. */
int fb_beam_search_phasing(fb_ctx *, const fb_frags *, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy,
                           const fb_params *, uint8_t *hap_out, double *best_score, double *tap_same,
                           double *tap_diff, double *tap_logp, uint64_t tap_cap, uint64_t *tap_n);

/* local_clustering.rs:71-130 optimize_clustering(partition, eps, num_iter_optimize).
 * n_rounds = accepted opt_iterate rounds. */
int fb_optimize_clustering(fb_ctx *, const fb_frags *, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap_in,
                           uint32_t ploidy, const fb_params *, uint8_t *hap_out, double *score,
                           uint32_t *n_rounds);

/* part_block_manip.rs:174-274 process_reads_for_final_parts (haplosets given as CSR of counter_ids). */
int fb_process_reads_for_final_parts(fb_ctx *, const fb_frags *, uint64_t n_parts, const uint64_t *part_ptr,
                                     const uint32_t *part_reads, const uint32_t *range_lo, const uint32_t *range_hi,
                                     const fb_params *, fb_parts **out);
void fb_free_parts(fb_parts *);
/* The three contig-level calls with the contig already resident in HBM (fb_frags_upload once per contig; a pipeline
 * fb_phase_blocks_resident -> fb_update_hap_graph_resident -> fb_process_reads_for_final_parts_resident ->
 * fb_get_hapq_resident packs the reads once instead of four times). */
int fb_process_reads_for_final_parts_resident(fb_ctx *, const fb_dfrags *, uint64_t n_parts, const uint64_t *part_ptr,
                                              const uint32_t *part_reads, const uint32_t *range_lo,
                                              const uint32_t *range_hi, const fb_params *, fb_parts **out);
int fb_get_hapq_resident(fb_ctx *, const fb_dfrags *, uint64_t n_parts, const uint64_t *part_ptr,
                         const uint32_t *part_reads, const uint32_t *range_lo, const uint32_t *range_hi,
                         const uint64_t *snp_to_genome_pos, uint64_t n_snps, const fb_params *, uint8_t *hapq,
                         double *rel_err, double *avg_err);
int fb_update_hap_graph_resident(fb_ctx *, const fb_dfrags *, uint64_t n_cols, const uint64_t *col_ptr,
                                 const uint64_t *node_ptr, const uint32_t *node_reads, const uint32_t *node_lo,
                                 const uint32_t *node_hi, const fb_params *, double *out_weights);

/* part_block_manip.rs:517-620 get_hapq. hapq/rel_err are [n_parts].
 * Order dependence (declared): get_errors_cov_from_frags compares every allele count with the RUNNING SUM of the
 * position (utils_frags.rs:616-624), so at sites with three or more alleles `rel_err` / `avg_err` depend on the hash
 * iteration order of the real binary; this library uses ascending allele order (HAPQ itself is not affected). */
int fb_get_hapq(fb_ctx *, const fb_frags *, uint64_t n_parts, const uint64_t *part_ptr, const uint32_t *part_reads,
                const uint32_t *range_lo, const uint32_t *range_hi, const uint64_t *snp_to_genome_pos,
                uint64_t n_snps, const fb_params *, uint8_t *hapq, double *rel_err, double *avg_err);

/* graph_processing.rs:22-100 update_hap_graph edge weights between consecutive columns of the block graph:
 * reads of node (col i, row r) scored with utils_frags.rs:77 distance_read_haplo against every node of col i+1.
 * Nodes are given as CSR of counter_ids plus their snp_endpoints (HapNode::new restricts hap_map to them,
 * types_structs.rs:169-180). out_weights is [sum_i rows(i)*rows(i+1)] in (i, r, l) order, before the
 * MIN_SHARED_READS_UNAMBIG filter. */
int fb_update_hap_graph(fb_ctx *, const fb_frags *, uint64_t n_cols, const uint64_t *col_ptr /*[n_cols+1] node ids*/,
                        const uint64_t *node_ptr, const uint32_t *node_reads, const uint32_t *node_lo,
                        const uint32_t *node_hi, const fb_params *, double *out_weights);

#ifdef __cplusplus
}
#endif
#endif /* FLORIA_B200_H */
