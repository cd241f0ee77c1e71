// fb_multi.cuh — several devices in one host process, and batching of a contig list.
//
// floria phases contigs one after the other (src/bin/floria.rs:229) and the blocks of a contig in a rayon par_iter
// (graph_processing.rs:345-362); both are independent units until process_chunks.  Here a contig list is dealt to the
// devices by a static longest-processing-time-first queue; every device concatenates its contigs along the SNP axis
// (positions shifted so that no read of one contig can touch a block of another) and phases them in ONE
// fb_phase_blocks_resident call, driven by its own host thread and stream.  The per-contig results are cut out of the
// batched result afterwards: they are identical, block for block, to per-contig calls (tests/test_gpu_multi.py).
#pragma once
#include <thread>

#include "fb_engine.cuh"

struct fb_multi {
    std::vector<fb_ctx *> ctx;
    std::string err;
};

struct FbContigSlot {  // where a contig sits inside its device's batch
    uint32_t device = 0;
    uint64_t blk_off = 0, n_blocks = 0;  // into the device's merged block list
    uint64_t read_off = 0;               // counter_id shift
    uint32_t pos_off = 0;                // SNP position shift
};
struct fb_dcontigs {
    uint64_t n_contigs = 0;
    std::vector<FbContigSlot> slot;                 // [n_contigs]
    std::vector<fb_dfrags *> df;                    // [n_devices] merged contigs resident on that device (may be null)
    std::vector<std::vector<uint32_t>> lo, hi;      // [n_devices] merged block ranges
    std::vector<std::vector<uint64_t>> members;     // [n_devices] contig indices in batch order
};

static thread_local std::string g_multi_err;

static void fb_lpt(const double *costs, uint64_t n, uint32_t bins, uint32_t *owner) {
    std::vector<uint64_t> order(n);
    for (uint64_t i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return costs[a] > costs[b]; });
    std::vector<double> load(bins, 0.0);
    for (uint64_t u : order) {
        uint32_t best = 0;
        for (uint32_t b = 1; b < bins; ++b)
            if (load[b] < load[best]) best = b;  // first minimum
        owner[u] = best;
        load[best] += costs[u];
    }
}

extern "C" {

int fb_init_multi(int n_devices, const int *device_ids, fb_multi **out) {
    if (!out) return FB_ERR_ARG;
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) {
        g_init_err = "fb_init_multi: n_devices must be in 1..64";
        return FB_ERR_ARG;
    }
    fb_multi *m = new fb_multi();
    for (int i = 0; i < n_devices; ++i) {
        fb_ctx *c = nullptr;
        const int rc = fb_init(device_ids ? device_ids[i] : i, &c);
        if (rc) {
            for (fb_ctx *x : m->ctx) fb_destroy(x);
            delete m;
            return rc;  // message in fb_last_error(NULL)
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return FB_OK;
}

void fb_destroy_multi(fb_multi *m) {
    if (!m) return;
    for (fb_ctx *c : m->ctx) fb_destroy(c);
    delete m;
}

int fb_multi_size(const fb_multi *m) { return m ? (int)m->ctx.size() : 0; }
fb_ctx *fb_multi_ctx(fb_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
const char *fb_multi_last_error(const fb_multi *m) { return m ? m->err.c_str() : g_init_err.c_str(); }

void fb_lpt_assign(const double *costs, uint64_t n_units, uint32_t n_bins, uint32_t *owner_out) {
    if (!costs || !owner_out || n_bins == 0) return;
    fb_lpt(costs, n_units, n_bins, owner_out);
}

void fb_contigs_free(fb_multi *m, fb_dcontigs *dc) {
    if (!dc) return;
    for (size_t d = 0; d < dc->df.size(); ++d)
        if (dc->df[d]) fb_frags_free(m && d < m->ctx.size() ? m->ctx[d] : nullptr, dc->df[d]);
    delete dc;
}

int fb_contigs_upload(fb_multi *m, uint64_t n_contigs, const fb_frags *contigs, const uint64_t *blk_ptr,
                      const uint32_t *blk_lo, const uint32_t *blk_hi, fb_dcontigs **out) {
    if (!m) return FB_ERR_ARG;
    if (!out || (n_contigs && (!contigs || !blk_ptr))) {
        m->err = "null argument";
        return FB_ERR_ARG;
    }
    *out = nullptr;
    const uint32_t D = (uint32_t)m->ctx.size();
    std::unique_ptr<fb_dcontigs> dc(new fb_dcontigs());
    dc->n_contigs = n_contigs;
    dc->slot.resize(n_contigs);
    dc->df.assign(D, nullptr);
    dc->lo.resize(D);
    dc->hi.resize(D);
    dc->members.resize(D);
    // cost of a contig: stored cells of its reads, counted once per block they fall in (blocks overlap by a third,
    // utils_frags.rs:405-463) ~ nnz * blocks-per-read; the ploidy factor is the same for every contig of a call
    std::vector<double> cost(n_contigs);
    for (uint64_t k = 0; k < n_contigs; ++k) cost[k] = (double)contigs[k].nnz + 1e-3 * (double)(blk_ptr[k + 1] - blk_ptr[k]);
    std::vector<uint32_t> owner(n_contigs);
    fb_lpt(cost.data(), n_contigs, D, owner.data());
    for (uint64_t k = 0; k < n_contigs; ++k) {
        dc->slot[k].device = owner[k];
        dc->members[owner[k]].push_back(k);
    }
    std::vector<int> rcs(D, FB_OK);
    auto work = [&](uint32_t d) {
        fb_ctx *ctx = m->ctx[d];
        const std::vector<uint64_t> &mem = dc->members[d];
        if (mem.empty()) return;
        // positions of contig k are shifted past everything of the contigs before it; the cells themselves go from the
        // caller's buffers straight to the device and are shifted there (fb_frags_upload_parts)
        std::vector<fb_frags> parts;
        std::vector<uint32_t> shift;
        uint64_t r0 = 0, p0 = 0;
        for (uint64_t k : mem) {
            const fb_frags &f = contigs[k];
            FbContigSlot &sl = dc->slot[k];
            sl.read_off = r0;
            sl.pos_off = (uint32_t)p0;
            sl.blk_off = dc->lo[d].size();
            sl.n_blocks = blk_ptr[k + 1] - blk_ptr[k];
            uint32_t span = 0;
            for (uint64_t i = 0; i < f.n_reads; ++i) span = std::max(span, f.last[i]);
            for (uint64_t j = blk_ptr[k]; j < blk_ptr[k + 1]; ++j) {
                dc->lo[d].push_back(blk_lo[j] + (uint32_t)p0);
                dc->hi[d].push_back(blk_hi[j] + (uint32_t)p0);
                span = std::max(span, blk_hi[j]);
            }
            parts.push_back(f);
            shift.push_back((uint32_t)p0);
            // the next contig starts on a fresh 16-position group beyond everything of this one
            p0 = (p0 + span + 64 + 15) & ~15ULL;
            r0 += f.n_reads;
            if (p0 >= (1ULL << 32) - (1ULL << 20)) {
                ctx->err = "batched contigs exceed the 32-bit SNP index space";
                rcs[d] = FB_ERR_LIMIT;
                return;
            }
        }
        rcs[d] = fb_frags_upload_parts(ctx, parts.size(), parts.data(), shift.data(), &dc->df[d]);
    };
    std::vector<std::thread> th;
    for (uint32_t d = 1; d < D; ++d) th.emplace_back(work, d);
    work(0);
    for (auto &t : th) t.join();
    for (uint32_t d = 0; d < D; ++d)
        if (rcs[d]) {
            m->err = "device " + std::to_string(d) + ": " + m->ctx[d]->err;
            fb_contigs_free(m, dc.release());
            return rcs[d];
        }
    *out = dc.release();
    return FB_OK;
}

int fb_phase_contigs_resident(fb_multi *m, const fb_dcontigs *dc, const fb_params *prm, fb_block_results **out,
                              uint32_t *device_of, float *device_ms) {
    if (!m) return FB_ERR_ARG;
    if (!dc || !out || !prm) {
        m->err = "null argument";
        return FB_ERR_ARG;
    }
    const uint32_t D = (uint32_t)m->ctx.size();
    for (uint64_t k = 0; k < dc->n_contigs; ++k) out[k] = nullptr;
    std::vector<int> rcs(D, FB_OK);
    std::vector<fb_block_results *> batch(D, nullptr);
    auto work = [&](uint32_t d) {
        if (device_ms) device_ms[d] = 0.f;
        if (!dc->df[d]) return;
        fb_ctx *ctx = m->ctx[d];
        const float t0 = ctx->tim.total_ms + ctx->tim.download_ms;
        rcs[d] = fb_phase_blocks_resident(ctx, dc->df[d], dc->lo[d].size(), dc->lo[d].data(), dc->hi[d].data(), prm, &batch[d]);
        if (device_ms) device_ms[d] = ctx->tim.total_ms + ctx->tim.download_ms - t0;
        if (rcs[d]) return;
        // cut the batched result into per-contig results (the batch lists contigs, blocks and reads in order)
        const fb_block_results *b = batch[d];
        const uint32_t mp = b->max_ploidy;
        for (uint64_t k : dc->members[d]) {
            const FbContigSlot &sl = dc->slot[k];
            const uint64_t nb = sl.n_blocks, j0 = sl.blk_off;
            fb_block_results *r = (fb_block_results *)calloc(1, sizeof(fb_block_results));
            r->n_blocks = nb;
            r->max_ploidy = mp;
            r->best_ploidy = (uint32_t *)calloc(nb + 1, sizeof(uint32_t));
            r->ploidies_run = (uint32_t *)calloc(nb + 1, sizeof(uint32_t));
            r->mec_vector = (double *)calloc(nb * mp + 1, sizeof(double));
            r->expected_errors = (double *)calloc(nb * mp + 1, sizeof(double));
            r->read_ptr = (uint64_t *)calloc(nb + 1, sizeof(uint64_t));
            r->block_cells = (uint64_t *)calloc(nb + 1, sizeof(uint64_t));
            const uint64_t a0 = b->read_ptr[j0], a1 = b->read_ptr[j0 + nb];
            r->read_ids = (uint32_t *)calloc(a1 - a0 + 1, sizeof(uint32_t));
            r->hap = (uint8_t *)calloc(a1 - a0 + 1, 1);
            for (uint64_t j = 0; j < nb; ++j) {
                r->best_ploidy[j] = b->best_ploidy[j0 + j];
                r->ploidies_run[j] = b->ploidies_run[j0 + j];
                r->block_cells[j] = b->block_cells[j0 + j];
                r->cells_beam += b->block_cells[j0 + j];  // a batched call only knows the sum of the three counters
                r->read_ptr[j] = b->read_ptr[j0 + j] - a0;
            }
            r->read_ptr[nb] = a1 - a0;
            memcpy(r->mec_vector, b->mec_vector + j0 * mp, nb * mp * sizeof(double));
            memcpy(r->expected_errors, b->expected_errors + j0 * mp, nb * mp * sizeof(double));
            for (uint64_t x = a0; x < a1; ++x) r->read_ids[x - a0] = b->read_ids[x] - (uint32_t)sl.read_off;
            if (a1 > a0) memcpy(r->hap, b->hap + a0, a1 - a0);
            out[k] = r;
            if (device_of) device_of[k] = d;
        }
    };
    std::vector<std::thread> th;
    for (uint32_t d = 1; d < D; ++d) th.emplace_back(work, d);
    work(0);
    for (auto &t : th) t.join();
    int rc = FB_OK;
    for (uint32_t d = 0; d < D; ++d) {
        if (batch[d]) fb_free_block_results(batch[d]);
        if (rcs[d] && !rc) {
            rc = rcs[d];
            m->err = "device " + std::to_string(d) + ": " + m->ctx[d]->err;
        }
    }
    if (rc)
        for (uint64_t k = 0; k < dc->n_contigs; ++k) {
            fb_free_block_results(out[k]);
            out[k] = nullptr;
        }
    return rc;
}

int fb_phase_contigs(fb_multi *m, uint64_t n_contigs, const fb_frags *contigs, const uint64_t *blk_ptr,
                     const uint32_t *blk_lo, const uint32_t *blk_hi, const fb_params *prm, fb_block_results **out,
                     uint32_t *device_of, float *device_ms) {
    fb_dcontigs *dc = nullptr;
    int rc = fb_contigs_upload(m, n_contigs, contigs, blk_ptr, blk_lo, blk_hi, &dc);
    if (rc) return rc;
    rc = fb_phase_contigs_resident(m, dc, prm, out, device_of, device_ms);
    fb_contigs_free(m, dc);
    return rc;
}

}  // extern "C"
