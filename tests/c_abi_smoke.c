/* c_abi_smoke.c — a plain C client of include/floria_b200.h (what a cgo / Rust FFI binding sees).
 *   c_abi_smoke layout            prints sizeof / offsetof of every struct of the boundary (compared with the ctypes
 *                                 mirror floria_b200/_cdefs.py by tests/test_c_abi.py; runs without a GPU)
 *   c_abi_smoke run               init -> upload -> phase_blocks_resident -> process_reads_for_final_parts -> get_hapq on a
 *                                 small deterministic contig; prints the results (compared with the Python binding)
 * Build: gcc -std=c99 -Iinclude tests/c_abi_smoke.c -Lfloria_b200 -lfloria_b200 -Wl,-rpath,$PWD/floria_b200 -lm */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "floria_b200.h"

#define OFF(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))

static int layout(void) {
    printf("fb_params %zu\n", sizeof(fb_params));
    OFF(fb_params, epsilon); OFF(fb_params, div_factor); OFF(fb_params, prob_cutoff_ln); OFF(fb_params, max_number_solns);
    OFF(fb_params, max_ploidy); OFF(fb_params, num_iter_optimize); OFF(fb_params, ploidy_sensitivity);
    OFF(fb_params, stopping_heuristic); OFF(fb_params, order_model); OFF(fb_params, block_length);
    OFF(fb_params, reassign_short); OFF(fb_params, phred_lut);
    printf("fb_frags %zu\n", sizeof(fb_frags));
    OFF(fb_frags, n_reads); OFF(fb_frags, nnz); OFF(fb_frags, row_ptr); OFF(fb_frags, first); OFF(fb_frags, last);
    OFF(fb_frags, pos); OFF(fb_frags, allele); OFF(fb_frags, qual);
    printf("fb_block_results %zu\n", sizeof(fb_block_results));
    OFF(fb_block_results, n_blocks); OFF(fb_block_results, max_ploidy); OFF(fb_block_results, best_ploidy);
    OFF(fb_block_results, ploidies_run); OFF(fb_block_results, mec_vector); OFF(fb_block_results, expected_errors);
    OFF(fb_block_results, read_ptr); OFF(fb_block_results, read_ids); OFF(fb_block_results, hap);
    OFF(fb_block_results, cells_sweep); OFF(fb_block_results, cells_hist); OFF(fb_block_results, cells_beam);
    OFF(fb_block_results, block_cells);
    printf("fb_block_phase %zu\n", sizeof(fb_block_phase));
    OFF(fb_block_phase, beam_score); OFF(fb_block_phase, opt_score); OFF(fb_block_phase, n_rounds);
    OFF(fb_block_phase, ploidy); OFF(fb_block_phase, cells_sweep); OFF(fb_block_phase, cells_hist);
    OFF(fb_block_phase, cells_beam);
    printf("fb_parts %zu\n", sizeof(fb_parts));
    OFF(fb_parts, n_parts); OFF(fb_parts, part_ptr); OFF(fb_parts, read_ids); OFF(fb_parts, range_lo);
    OFF(fb_parts, range_hi);
    printf("fb_timings %zu\n", sizeof(fb_timings));
    OFF(fb_timings, upload_ms); OFF(fb_timings, download_ms); OFF(fb_timings, n_launches); OFF(fb_timings, hist_cells);
    return 0;
}

/* splitmix64: the contig is a pure function of the seed (the Python side regenerates it with the same code) */
static uint64_t sm64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

#define CK(call)                                                                   \
    do {                                                                           \
        int rc_ = (call);                                                          \
        if (rc_) {                                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fb_last_error(ctx));     \
            return 1;                                                              \
        }                                                                          \
    } while (0)

static int run(void) {
    enum { R = 240, S = 200, SPAN = 40 };
    uint64_t seed = 12345;
    static uint64_t row_ptr[R + 1], g2[S];
    static uint32_t first[R], last[R], pos[R * SPAN];
    static uint8_t allele[R * SPAN], qual[R * SPAN], truth[2][S];
    for (int s = 0; s < S; ++s) {
        truth[0][s] = (uint8_t)(sm64(&seed) & 1);
        truth[1][s] = (uint8_t)(1 - truth[0][s]);
        g2[s] = 100ull * s + 7;
    }
    /* reads sorted by (first asc, last desc): starts are non-decreasing, all spans equal */
    uint64_t nnz = 0;
    row_ptr[0] = 0;
    for (int r = 0; r < R; ++r) {
        int f = 1 + (r * (S - SPAN)) / R;
        int h = (int)(sm64(&seed) & 1);
        first[r] = (uint32_t)f;
        last[r] = (uint32_t)(f + SPAN - 1);
        for (int k = 0; k < SPAN; ++k) {
            pos[nnz] = (uint32_t)(f + k);
            int a = truth[h][f + k - 1];
            if (sm64(&seed) % 25 == 0) a = 1 - a;
            allele[nnz] = (uint8_t)a;
            qual[nnz] = (uint8_t)(10 + sm64(&seed) % 30);
            nnz++;
        }
        row_ptr[r + 1] = nnz;
    }
    fb_frags fr = {R, nnz, row_ptr, first, last, pos, allele, qual};
    fb_ctx *ctx = NULL;
    CK(fb_init(0, &ctx));
    fb_params prm;
    fb_params_default(&prm);
    prm.max_ploidy = 3;
    uint32_t lo[64], hi[64];
    int64_t nb = fb_get_range_with_lengths(g2, S, 5000, 5000 / 3, 0.0005, lo, hi, 64);
    if (nb <= 0 || nb > 64) return 2;
    fb_dfrags *df = NULL;
    CK(fb_frags_upload(ctx, &fr, &df));
    fb_block_results *res = NULL;
    CK(fb_phase_blocks_resident(ctx, df, (uint64_t)nb, lo, hi, &prm, &res));
    printf("blocks %lld\n", (long long)nb);
    for (int64_t j = 0; j < nb; ++j)
        printf("block %lld ploidy %u mec %.17g %.17g %.17g reads %llu\n", (long long)j, res->best_ploidy[j],
               res->mec_vector[j * 3], res->mec_vector[j * 3 + 1], res->mec_vector[j * 3 + 2],
               (unsigned long long)(res->read_ptr[j + 1] - res->read_ptr[j]));
    /* haplosets = the haplotypes of every block's partition, ranges = the block ranges */
    uint64_t np = 0, tot = 0;
    static uint64_t part_ptr[256];
    static uint32_t part_reads[R * 8], rlo[256], rhi[256];
    part_ptr[0] = 0;
    for (int64_t j = 0; j < nb; ++j)
        for (uint32_t h = 0; h < res->best_ploidy[j]; ++h) {
            for (uint64_t k = res->read_ptr[j]; k < res->read_ptr[j + 1]; ++k)
                if (res->hap[k] == h) part_reads[tot++] = res->read_ids[k];
            rlo[np] = lo[j];
            rhi[np] = hi[j];
            part_ptr[++np] = tot;
        }
    fb_parts *parts = NULL;
    CK(fb_process_reads_for_final_parts(ctx, &fr, np, part_ptr, part_reads, rlo, rhi, &prm, &parts));
    printf("parts %llu\n", (unsigned long long)parts->n_parts);
    static uint8_t hapq[512];
    static double rel[512];
    double avg = 0;
    CK(fb_get_hapq(ctx, &fr, parts->n_parts, parts->part_ptr, parts->read_ids, parts->range_lo, parts->range_hi, g2, S, &prm,
                   hapq, rel, &avg));
    for (uint64_t i = 0; i < parts->n_parts; ++i)
        printf("part %llu range %u-%u reads %llu hapq %u rel %.17g\n", (unsigned long long)i, parts->range_lo[i],
               parts->range_hi[i], (unsigned long long)(parts->part_ptr[i + 1] - parts->part_ptr[i]), hapq[i], rel[i]);
    printf("avg_err %.17g\n", avg);
    fb_free_parts(parts);
    fb_free_block_results(res);
    fb_frags_free(ctx, df);
    fb_destroy(ctx);
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && strcmp(argv[1], "layout") == 0) return layout();
    if (argc > 1 && strcmp(argv[1], "run") == 0) return run();
    fprintf(stderr, "usage: %s layout|run\n", argv[0]);
    return 2;
}
