// fb_beam_host.cuh — host driver of the beam-search kernel: work queue, per-CTA scratch slots, launch.
#pragma once
#include "fb_beam.cuh"
#include "fb_engine.cuh"
#include <stdlib.h>

struct BeamRun {
    std::vector<unsigned long long> cells_beam, tap_n;
    std::vector<double> best_score;
    float beam_ms = 0;
};

static int fb_run_beam(fb_ctx *ctx, Engine &e, const fb_params *prm, const BeamTapDev *tap, BeamRun &br) {
    const int n_inst = e.n_inst();
    br.cells_beam.assign(n_inst, 0);
    br.tap_n.assign(n_inst, 0);
    br.best_score.assign(n_inst, 0.0);
    std::vector<int> order;
    uint32_t maxP = 1, maxR = 1;
    uint64_t max_pool = 0;
    const uint32_t B = prm->max_number_solns;
    if (B < 1) FB_FAIL(FB_ERR_ARG, "max_number_solns must be >= 1");
    for (int i = 0; i < n_inst; ++i) {
        const InstDev &in = e.inst[i];
        if (in.ploidy < 2) continue;  // ploidy 1: every read lands in haplotype 0 (assign buffers start zeroed)
        order.push_back(i);
        maxP = std::max(maxP, in.ploidy);
        maxR = std::max(maxR, in.n_reads);
        const uint64_t NS = (uint64_t)in.ploidy * B * (in.ploidy + 1) + 1;
        const uint64_t state_words = ((uint64_t)in.ng * 64 + in.ng + 1) & ~1ULL;
        max_pool = std::max(max_pool, NS * state_words * 8);
    }
    if (order.empty()) return FB_OK;
    const uint32_t maxW = maxP * B;
    const uint32_t maxNS = maxP * B * (maxP + 1) + 1;
    if (maxW > FB_BEAM_THREADS) FB_FAIL(FB_ERR_LIMIT, "ploidy*max_number_solns = %u exceeds %d", maxW, FB_BEAM_THREADS);
    if (maxNS > 65535) FB_FAIL(FB_ERR_LIMIT, "too many haplotype states (%u)", maxNS);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const uint64_t ca = e.blocks[e.inst[a].block].nnz * e.inst[a].ploidy;
        const uint64_t cb = e.blocks[e.inst[b].block].nnz * e.inst[b].ploidy;
        return ca > cb;
    });
    BeamSmem L;
    L.layout(maxP, maxW, maxNS);
    if (L.total > 200 * 1024) FB_FAIL(FB_ERR_LIMIT, "beam search needs %u bytes of shared memory", L.total);
    // many instances: 128-thread CTAs, two per SM (the per-read dependency chains of two instances interleave);
    // few instances: 256-thread CTAs, one per SM (shortest chain per step)
    const int forced = getenv("FB_BEAM_CTA") ? atoi(getenv("FB_BEAM_CTA")) : 0;  // tests / A-B runs: 128 or 256
    const bool small_cta = maxW <= FB_BEAM_THREADS_SMALL && forced != FB_BEAM_THREADS &&
                           (forced == FB_BEAM_THREADS_SMALL || order.size() >= (size_t)ctx->sm_count * 2);
    const int nt = small_cta ? FB_BEAM_THREADS_SMALL : FB_BEAM_THREADS;
    int occ = 1;
    FB_CK((cudaError_t)fb_beam_occupancy(nt, L.total, &occ));
    if (occ < 1) occ = 1;
    const uint64_t pool_bytes = (max_pool + 255) & ~255ULL;
    const uint64_t hist_bytes = (((uint64_t)maxR * maxW * 4) + 255) & ~255ULL;
    const uint64_t slot_bytes = pool_bytes + hist_bytes;
    size_t free_b = 0, total_b = 0;
    FB_CK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t n_slots = std::min<uint64_t>(order.size(), (uint64_t)ctx->sm_count * occ);
    const uint64_t budget = (uint64_t)(free_b * 0.85);
    if (slot_bytes > budget) FB_FAIL(FB_ERR_LIMIT, "beam search scratch (%llu bytes) does not fit", (unsigned long long)slot_bytes);
    n_slots = std::max<uint64_t>(1, std::min<uint64_t>(n_slots, budget / slot_bytes));

    uint8_t *d_scratch = nullptr;
    int *d_order = nullptr, *d_counter = nullptr;
    unsigned long long *d_cells = nullptr, *d_tapn = nullptr;
    double *d_best = nullptr;
    int rc;
    auto cleanup = [&]() {
        fb_cache_free(d_scratch);
        fb_cache_free(d_order);
        fb_cache_free(d_counter);
        fb_cache_free(d_cells);
        fb_cache_free(d_tapn);
        fb_cache_free(d_best);
    };
    if ((rc = fb_dalloc(ctx, &d_scratch, n_slots * slot_bytes)) || (rc = fb_upload(ctx, &d_order, order)) ||
        (rc = fb_dalloc(ctx, &d_counter, 1)) || (rc = fb_dalloc(ctx, &d_cells, (size_t)n_inst)) ||
        (rc = fb_dalloc(ctx, &d_tapn, (size_t)n_inst)) || (rc = fb_dalloc(ctx, &d_best, (size_t)n_inst))) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream);
    cudaMemsetAsync(d_cells, 0, sizeof(unsigned long long) * n_inst, ctx->stream);
    cudaMemsetAsync(d_tapn, 0, sizeof(unsigned long long) * n_inst, ctx->stream);
    cudaMemsetAsync(d_best, 0, sizeof(double) * n_inst, ctx->stream);
    BeamParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.fr = e.df->dev();
    bp.inst = e.d_inst;
    bp.rinfo = e.d_rinfo;
    bp.rextra = e.d_rextra;
    bp.lut = ctx->d_lut;
    bp.order = d_order;
    bp.n_work = (int)order.size();
    bp.work_counter = d_counter;
    bp.assign_out = e.d_assign[0];
    bp.eps = prm->epsilon;
    bp.div_factor = prm->div_factor;
    bp.cutoff = prm->prob_cutoff_ln;
    bp.eps_safe = fb_eps_is_safe(prm->epsilon);
    bp.B = B;
    bp.maxP = maxP;
    bp.maxW = maxW;
    bp.maxNS = maxNS;
    bp.scratch = d_scratch;
    bp.slot_bytes = slot_bytes;
    bp.hist_off = pool_bytes;
    bp.cells_out = d_cells;
    bp.best_out = d_best;
    bp.tapn_out = d_tapn;
    if (tap) bp.tap = *tap;
    unsigned long long *d_prof = nullptr;
    const bool prof = getenv("FB_BEAM_PROF") != nullptr;
    if (prof) {
        if ((rc = fb_dalloc(ctx, &d_prof, 24))) {
            cleanup();
            return rc;
        }
        cudaMemsetAsync(d_prof, 0, 192, ctx->stream);
        bp.prof = d_prof;
    }
    cudaEvent_t e0 = fb_event(ctx);
    cudaError_t launch_err = (cudaError_t)fb_beam_launch(nt, (unsigned)n_slots, L.total, ctx->stream, bp);
    cudaEvent_t e1 = fb_event(ctx);
    ctx->tim.n_launches++;
    ctx->tim.n_beam_launches++;
    cudaError_t ce = launch_err;
    if (ce == cudaSuccess) {
        cudaMemcpyAsync(br.cells_beam.data(), d_cells, sizeof(unsigned long long) * n_inst, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(br.tap_n.data(), d_tapn, sizeof(unsigned long long) * n_inst, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(br.best_score.data(), d_best, sizeof(double) * n_inst, cudaMemcpyDeviceToHost, ctx->stream);
        ce = cudaStreamSynchronize(ctx->stream);
    }
    if (ce != cudaSuccess) {
        ctx->err = std::string("beam search: ") + cudaGetErrorString(ce);
        cleanup();
        return FB_ERR_CUDA;
    }
    cudaEventElapsedTime(&br.beam_ms, e0, e1);
    if (prof) {
        unsigned long long h[24];
        cudaMemcpy(h, d_prof, 192, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[k_beam prof] scoring warp 0, cycles/step: loop %.0f | reductions %.0f | diff_f %.0f | p-value %.0f\n", h[16] / (double)std::max<unsigned long long>(h[12], 1), h[17] / (double)std::max<unsigned long long>(h[12], 1), h[18] / (double)std::max<unsigned long long>(h[12], 1), h[19] / (double)std::max<unsigned long long>(h[12], 1));
        fb_cache_free(d_prof);
        double steps = (double)std::max<unsigned long long>(h[12], 1);
        fprintf(stderr,
                "[k_beam prof] %.3f ms, %d instances on %llu CTAs, %.0f steps; cycles/step: phase1(score) %.0f | warp0: lse %.0f "
                "compact+fold+dups+classes %.0f heap %.0f nextgen %.0f | phase3(copy) %.0f | children/step %.1f survivors/step %.1f "
                "copyjobs/step %.2f inplace/step %.2f live states/step %.2f nodes/step %.2f | backtrack total %.0f\n",
                br.beam_ms, (int)order.size(), (unsigned long long)n_slots, steps, h[0] / steps, h[6] / steps, h[7] / steps,
                h[8] / steps, h[1] / steps, h[2] / steps, h[10] / steps, h[11] / steps, h[13] / steps, h[14] / steps,
                h[15] / steps, h[9] / steps, (double)h[5]);
    }
    cleanup();
    return FB_OK;
}
