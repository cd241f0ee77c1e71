// fb_beam_host.cuh — host driver of the beam-search kernels: work queue, scratch slots, launch.
//   k_beam      (fb_beam.cuh)       one CTA per (block, ploidy) instance, persistent CTAs pull from a queue
//   k_beam_wide (fb_beam_wide.cuh)  the whole grid on one instance at a time (blocks whose reads span thousands of SNPs)
#pragma once
#include "fb_beam.cuh"
#include "fb_beam_wide.cuh"
#include "fb_engine.cuh"
#include <stdlib.h>
#include <chrono>

struct BeamRun {
    std::vector<unsigned long long> cells_beam, tap_n;
    std::vector<double> best_score;
    float beam_ms = 0;
    int n_wide = 0, n_narrow = 0;
};

// Which instances go to the grid-wide kernel.  A step of k_beam costs about (7 + 0.009 g) us for reads of g groups
// (scoring and materialisation of a state are one warp / one CTA wide), a step of k_beam_wide about 5.5 us whatever g
// (profiles/README.md); k_beam runs min(n, SMs) instances at once, k_beam_wide one.  FB_BEAM_WIDE=0/1 forces the choice
// (tests, A/B runs).
static void fb_beam_split(const fb_ctx *ctx, const Engine &e, std::vector<int> &narrow, std::vector<int> &wide) {
    const char *env = getenv("FB_BEAM_WIDE");
    const int forced = env ? atoi(env) : -1;
    std::vector<int> cand;
    double g_sum = 0;
    for (int i = 0; i < e.n_inst(); ++i) {
        const InstDev &in = e.inst[i];
        if (in.ploidy < 2 || in.n_reads == 0) continue;  // ploidy 1: every read lands in haplotype 0
        const double g_mean = (double)e.blocks[in.block].groups / in.n_reads;
        if (forced == 1 || (forced != 0 && g_mean >= 192.0)) {
            cand.push_back(i);
            g_sum += g_mean;
        } else {
            narrow.push_back(i);
        }
    }
    if (cand.empty()) return;
    if (forced != 1) {
        const double g_mean = g_sum / cand.size();
        const double par = (double)std::min<size_t>(cand.size(), (size_t)ctx->sm_count);
        if ((7.0 + 0.009 * g_mean) / par <= 5.5 * 1.2) {  // enough instances to fill the GPU one CTA each
            narrow.insert(narrow.end(), cand.begin(), cand.end());
            std::sort(narrow.begin(), narrow.end());
            return;
        }
    }
    wide = cand;
}

static int fb_run_beam(fb_ctx *ctx, Engine &e, const fb_params *prm, const BeamTapDev *tap, BeamRun &br) {
    const bool hprof = getenv("FB_HOST_PROF") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_enter = now();
    double t_split = 0, t_alloc = 0;
    const int n_inst = e.n_inst();
    br.cells_beam.assign(n_inst, 0);
    br.tap_n.assign(n_inst, 0);
    br.best_score.assign(n_inst, 0.0);
    const uint32_t B = prm->max_number_solns;
    if (B < 1) FB_FAIL(FB_ERR_ARG, "max_number_solns must be >= 1");
    std::vector<int> order_n, order_w;
    fb_beam_split(ctx, e, order_n, order_w);
    br.n_narrow = (int)order_n.size();
    br.n_wide = (int)order_w.size();
    t_split = now();
    if (order_n.empty() && order_w.empty()) return FB_OK;

    unsigned long long *d_cells = nullptr, *d_tapn = nullptr, *d_prof = nullptr;
    double *d_best = nullptr;
    int *d_counter = nullptr;
    std::vector<void *> temps;  // per-launch scratch, released at the end
    int rc = FB_OK;
    auto cleanup = [&]() {
        fb_cache_free(d_cells);
        fb_cache_free(d_tapn);
        fb_cache_free(d_best);
        fb_cache_free(d_counter);
        fb_cache_free(d_prof);
        for (void *p : temps) fb_cache_free(p);
    };
    if ((rc = fb_dalloc(ctx, &d_counter, 1)) || (rc = fb_dalloc(ctx, &d_cells, (size_t)n_inst)) ||
        (rc = fb_dalloc(ctx, &d_tapn, (size_t)n_inst)) || (rc = fb_dalloc(ctx, &d_best, (size_t)n_inst))) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream);
    cudaMemsetAsync(d_cells, 0, sizeof(unsigned long long) * n_inst, ctx->stream);
    cudaMemsetAsync(d_tapn, 0, sizeof(unsigned long long) * n_inst, ctx->stream);
    cudaMemsetAsync(d_best, 0, sizeof(double) * n_inst, ctx->stream);
    const bool prof = getenv("FB_BEAM_PROF") != nullptr;
    if (prof) {
        if ((rc = fb_dalloc(ctx, &d_prof, 48))) {
            cleanup();
            return rc;
        }
        cudaMemsetAsync(d_prof, 0, 48 * 8, ctx->stream);
    }
    // memory budget of the scratch slots: what was free when the context was opened minus what this context holds now
    // (cudaMemGetInfo itself costs 1-50 ms per call on a context with many allocations: measured, profiles/README.md)
    size_t free_b = ctx->mem_free_at_init > ctx->cache.live_bytes ? ctx->mem_free_at_init - ctx->cache.live_bytes : 0;
    if (free_b == 0) {
        size_t total_b = 0;
        FB_CK(cudaMemGetInfo(&free_b, &total_b));
    }
    const uint64_t budget = (uint64_t)(free_b * 0.85);

    t_alloc = now();
    cudaEvent_t e0 = fb_event(ctx);
    cudaError_t launch_err = cudaSuccess;
    uint64_t slots_n = 0;
    int grid_w = 0;
    for (int kind = 0; kind < 2 && launch_err == cudaSuccess; ++kind) {
        std::vector<int> &order = kind == 0 ? order_n : order_w;
        if (order.empty()) continue;
        uint32_t maxP = 1, maxR = 1;
        uint64_t max_pool = 0;
        for (int i : order) {
            const InstDev &in = e.inst[i];
            maxP = std::max(maxP, in.ploidy);
            maxR = std::max(maxR, in.n_reads);
            const uint64_t NS = (uint64_t)in.ploidy * B * (in.ploidy + 1) + 1;
            const uint64_t state_words = ((uint64_t)in.ng * 64 + in.ng + 1) & ~1ULL;
            max_pool = std::max(max_pool, NS * state_words * 8);
        }
        const uint32_t maxW = maxP * B;
        const uint32_t maxNS = maxP * B * (maxP + 1) + 1;
        if (maxW > FB_BEAM_THREADS) {
            cleanup();
            FB_FAIL(FB_ERR_LIMIT, "ploidy*max_number_solns = %u exceeds %d", maxW, FB_BEAM_THREADS);
        }
        if (maxNS > 65535) {
            cleanup();
            FB_FAIL(FB_ERR_LIMIT, "too many haplotype states (%u)", maxNS);
        }
        if (kind == 0) {
            // Largest first, so that the queue drains evenly; and one ploidy after the other (highest first): every ploidy is
            // its own instantiation of the kernel body, a step's instructions (~30 KB) just fit the SM's 32 KB instruction
            // cache, and CTAs of different ploidies sharing an SM evict each other's code (ncu: 1.4 no_instruction stalls per
            // issue in a mixed launch against 0.45 in a single-ploidy one).  FB_BEAM_MIXED=1 restores the pure cost order.
            const bool mixed = getenv("FB_BEAM_MIXED") && atoi(getenv("FB_BEAM_MIXED")) != 0;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                if (!mixed && e.inst[a].ploidy != e.inst[b].ploidy) return e.inst[a].ploidy > e.inst[b].ploidy;
                const uint64_t ca = e.blocks[e.inst[a].block].nnz * e.inst[a].ploidy;
                const uint64_t cb = e.blocks[e.inst[b].block].nnz * e.inst[b].ploidy;
                return ca > cb;
            });
        }
        BeamSmem L;
        L.layout(maxP, maxW, maxNS);
        if (L.total > 200 * 1024) {
            cleanup();
            FB_FAIL(FB_ERR_LIMIT, "beam search needs %u bytes of shared memory", L.total);
        }
        const uint64_t pool_bytes = (max_pool + 255) & ~255ULL;
        const uint64_t hist_bytes = (((uint64_t)maxR * maxW * 4) + 255) & ~255ULL;
        const uint64_t slot_bytes = pool_bytes + hist_bytes;
        if (slot_bytes > budget) {
            cleanup();
            FB_FAIL(FB_ERR_LIMIT, "beam search scratch (%llu bytes) does not fit", (unsigned long long)slot_bytes);
        }
        uint64_t n_slots = 1;
        int nt = FB_BW_THREADS;
        if (kind == 0) {
            // many instances: 128-thread CTAs, three per SM (the per-read dependency chains of several instances
            // interleave); few instances: 256-thread CTAs, one per SM (shortest chain per step)
            const int forced = getenv("FB_BEAM_CTA") ? atoi(getenv("FB_BEAM_CTA")) : 0;  // tests / A-B runs: 128 or 256
            const bool small_cta = maxW <= FB_BEAM_THREADS_SMALL && forced != FB_BEAM_THREADS &&
                                   (forced == FB_BEAM_THREADS_SMALL || order.size() >= (size_t)ctx->sm_count * 2);
            nt = small_cta ? FB_BEAM_THREADS_SMALL : FB_BEAM_THREADS;
            int occ = 1;
            cudaError_t oe = (cudaError_t)fb_beam_occupancy(nt, L.total, &occ);
            if (oe != cudaSuccess) {
                launch_err = oe;
                break;
            }
            if (occ < 1) occ = 1;
            n_slots = std::min<uint64_t>(order.size(), (uint64_t)ctx->sm_count * occ);
            n_slots = std::max<uint64_t>(1, std::min<uint64_t>(n_slots, budget / slot_bytes));
            slots_n = n_slots;
        } else {
            cudaError_t oe = (cudaError_t)fb_beam_wide_max_grid(L.total, ctx->sm_count, &grid_w);
            if (oe != cudaSuccess) {
                launch_err = oe;
                break;
            }
            if (grid_w < 1) {
                cleanup();
                FB_FAIL(FB_ERR_LIMIT, "k_beam_wide does not fit an SM (%u bytes of shared memory)", L.total);
            }
            // pipelined upload: leave SMs to the k_pack launches that run beside this kernel
            if (e.df->pipelined) {
                const char *ps = getenv("FB_PIPE_SMS");
                grid_w = std::max(1, grid_w - (ps ? std::max(1, atoi(ps)) : 32));
            }
            if (getenv("FB_BEAM_WIDE_GRID")) grid_w = std::max(1, std::min(grid_w, atoi(getenv("FB_BEAM_WIDE_GRID"))));
        }
        uint8_t *d_scratch = nullptr;
        int *d_order = nullptr;
        BeamWideAcc *d_wacc = nullptr;
        BeamWideStep *d_wstep = nullptr;
        unsigned long long *d_wbar = nullptr;
        if ((rc = fb_dalloc(ctx, &d_scratch, n_slots * slot_bytes)) || (temps.push_back(d_scratch), 0) ||
            (rc = fb_upload(ctx, &d_order, order)) || (temps.push_back(d_order), 0)) {
            cleanup();
            return rc;
        }
        if (kind == 1) {
            if ((rc = fb_dalloc(ctx, &d_wacc, (size_t)FB_BW_SLOTS * maxNS)) || (temps.push_back(d_wacc), 0) ||
                (rc = fb_dalloc(ctx, &d_wstep, FB_BW_SLOTS)) || (temps.push_back(d_wstep), 0) ||
                (rc = fb_dalloc(ctx, &d_wbar, 1)) || (temps.push_back(d_wbar), 0)) {
                cleanup();
                return rc;
            }
            cudaMemsetAsync(d_wbar, 0, sizeof(unsigned long long), ctx->stream);
        }
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
        bp.L = L;
        bp.scratch = d_scratch;
        bp.slot_bytes = slot_bytes;
        bp.hist_off = pool_bytes;
        bp.cells_out = d_cells;
        bp.best_out = d_best;
        bp.tapn_out = d_tapn;
        if (tap) bp.tap = *tap;
        bp.prof = prof ? d_prof + 24 * kind : nullptr;
        bp.wacc = d_wacc;
        bp.wstep = d_wstep;
        bp.wbar = d_wbar;
        if (kind == 0) {
            // k_beam takes its reads in any order: it needs the whole contig
            if (e.df->pipelined) cudaStreamWaitEvent(ctx->stream, e.df->ev_done, 0);
            launch_err = (cudaError_t)fb_beam_launch(nt, (unsigned)n_slots, L.total, ctx->stream, bp);
        } else {
            if (e.df->pipelined) {
                bp.ready = e.df->d_ready;  // reads are consumed in order: start as soon as the first chunk is packed
                cudaStreamWaitEvent(ctx->stream, e.df->ev_first, 0);
            }
            launch_err = (cudaError_t)fb_beam_wide_launch((unsigned)grid_w, L.total, ctx->stream, bp);
        }
        ctx->tim.n_launches++;
        ctx->tim.n_beam_launches++;
    }
    if (e.df->pipelined) cudaStreamWaitEvent(ctx->stream, e.df->ev_done, 0);  // the rest of the call reads the whole contig
    cudaEvent_t e1 = fb_event(ctx);
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
    if (hprof)
        fprintf(stderr, "[fb_run_beam host ms] split %.2f | result buffers %.2f | scratch + launch + sync %.2f (kernels %.2f)\n",
                t_split - t_enter, t_alloc - t_split, now() - t_alloc, br.beam_ms);
    if (prof) {
        unsigned long long hh[48];
        cudaMemcpy(hh, d_prof, 48 * 8, cudaMemcpyDeviceToHost);
        if (!order_n.empty()) {
            const unsigned long long *h = hh;
            const double steps = (double)std::max<unsigned long long>(h[12], 1);
            fprintf(stderr, "[k_beam prof] scoring warp 0, cycles/step: loop %.0f | reductions %.0f | diff_f %.0f | p-value %.0f\n",
                    h[16] / steps, h[17] / steps, h[18] / steps, h[19] / steps);
            fprintf(stderr,
                    "[k_beam prof] %.3f ms (both kernels), %d instances on %llu CTAs, %.0f steps; cycles/step: phase1(score) %.0f | warp0: lse %.0f "
                    "compact+fold+dups+classes %.0f heap %.0f nextgen %.0f | phase3(copy) %.0f | children/step %.1f survivors/step %.1f "
                    "copyjobs/step %.2f inplace/step %.2f live states/step %.2f nodes/step %.2f | uniform steps %.1f %% | backtrack total %.0f\n",
                    br.beam_ms, (int)order_n.size(), (unsigned long long)slots_n, steps, h[0] / steps, h[6] / steps, h[7] / steps,
                    h[8] / steps, h[1] / steps, h[2] / steps, h[10] / steps, h[11] / steps, h[13] / steps, h[14] / steps,
                    h[15] / steps, h[9] / steps, 100.0 * h[21] / steps, (double)h[5]);
        }
        if (!order_w.empty()) {
            const unsigned long long *h = hh + 24;
            const double steps = (double)std::max<unsigned long long>(h[12], 1);
            fprintf(stderr, "[k_beam_wide prof] p-value step of CTA 0 (cycles after the grid barrier): sums loaded %.0f, p-values formed %.0f\n",
                    h[16] / steps, h[17] / steps);
            fprintf(stderr,
                    "[k_beam_wide prof] %.3f ms (both kernels), %d instances on a grid of %d CTAs, %.0f steps; CTA 0 cycles/step: "
                    "warp 0: bookkeeping after the live list %.0f | grid barrier %.0f | p-values %.0f | lse %.0f "
                    "compact+fold+dups+classes %.0f heap %.0f (jobs + live list: rest) || warp 1: phase A %.0f | stage next read + "
                    "phase C %.0f | waiting for warp 0 %.0f || children/step %.1f survivors/step %.1f copyjobs/step %.2f "
                    "inplace/step %.2f live states/step %.2f nodes/step %.2f | uniform steps %.1f %% | backtrack total %.0f\n",
                    br.beam_ms, (int)order_w.size(), grid_w, steps, h[1] / steps, h[3] / steps, h[4] / steps, h[6] / steps,
                    h[7] / steps, h[8] / steps, h[0] / steps, h[2] / steps, h[20] / steps, h[10] / steps, h[11] / steps,
                    h[13] / steps, h[14] / steps, h[15] / steps, h[9] / steps, 100.0 * h[21] / steps, (double)h[5]);
        }
    }
    cleanup();
    return FB_OK;
}
