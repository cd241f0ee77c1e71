// fb_lib.cu — the C-ABI of include/floria_b200.h.  Host logic only; kernels live in fb_kernels.cuh / fb_beam.cuh.
#include <chrono>
#include <math.h>
#include <stdlib.h>

#include <memory>

#include "fb_engine.cuh"
#include "fb_beam_host.cuh"
#include "fb_final.cuh"
#include "fb_graph.cuh"

// ======================================================================================================================
// context
// ======================================================================================================================
static int fb_set_lut(fb_ctx *ctx, const fb_params *prm) {
    float lut[256];
    if (prm && prm->phred_lut) {
        memcpy(lut, prm->phred_lut, sizeof(lut));
    } else {
        // utils_frags.rs:702-711 phred_scale: 1f32 - 10f32.powf(q as f32 / -10.)
        for (int q = 0; q < 256; ++q) lut[q] = 1.0f - powf(10.0f, (float)q / -10.0f);
    }
    if (ctx->lut_valid && memcmp(lut, ctx->h_lut_f, sizeof(lut)) == 0) return FB_OK;
    for (int q = 0; q < 256; ++q) {
        double x = (double)lut[q] * FB_Q26;
        if (!(x >= 0.0) || x > FB_Q26 || x != floor(x))
            FB_FAIL(FB_ERR_ARG, "phred_lut[%d] = %g is not a multiple of 2^-26 in [0,1]", q, (double)lut[q]);
        ctx->h_lut[q] = (uint32_t)x;
    }
    memcpy(ctx->h_lut_f, lut, sizeof(lut));
    FB_CK(cudaMemcpyAsync(ctx->d_lut, ctx->h_lut, sizeof(ctx->h_lut), cudaMemcpyHostToDevice, ctx->stream));
    FB_CK(cudaStreamSynchronize(ctx->stream));
    ctx->lut_valid = true;
    return FB_OK;
}

static int fb_check_params(fb_ctx *ctx, const fb_params *p, uint32_t ploidy_needed) {
    if (!p) FB_FAIL(FB_ERR_ARG, "params is NULL");
    if (p->order_model != 0) FB_FAIL(FB_ERR_ARG, "order_model %u is not implemented (only 0 = canonical)", p->order_model);
    if (!(p->epsilon > 0.0) || !(p->epsilon < 1.0)) FB_FAIL(FB_ERR_ARG, "epsilon must be in (0,1)");
    if (ploidy_needed > FB_MAXP) FB_FAIL(FB_ERR_LIMIT, "ploidy %u exceeds the compiled maximum %d", ploidy_needed, FB_MAXP);
    if (p->reassign_short) FB_FAIL(FB_ERR_ARG, "reassign_short is not implemented on the device path");
    return fb_set_lut(ctx, p);
}

extern "C" {

int fb_init(int device, fb_ctx **out) {
    if (!out) return FB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_init_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                     " (floria_b200 has no CPU fallback)";
        return FB_ERR_NODEV;
    }
    if (device < 0 || device >= n) {
        g_init_err = "device index out of range";
        return FB_ERR_ARG;
    }
    fb_ctx *ctx = new fb_ctx();
    ctx->device = device;
    memset(&ctx->tim, 0, sizeof(ctx->tim));
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreate(&ctx->stream)) != cudaSuccess ||
        (e = cudaMalloc((void **)&ctx->d_lut, 256 * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc((void **)&ctx->d_n_active, sizeof(int))) != cudaSuccess ||
        (e = cudaMallocHost((void **)&ctx->h_n_active, 2 * sizeof(int))) != cudaSuccess) {
        g_init_err = std::string("fb_init: ") + cudaGetErrorString(e);
        delete ctx;
        return FB_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) ctx->mem_free_at_init = free_b;
    }
    *out = ctx;
    return FB_OK;
}

void fb_destroy(fb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_n_active);
    ctx->cache.trim();
    FbCacheRegistry::get().orphan(&ctx->cache);
    if (ctx->h_n_active) cudaFreeHost(ctx->h_n_active);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *fb_last_error(const fb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_err.c_str(); }

void fb_params_default(fb_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->epsilon = 0.04;
    p->div_factor = 0.25;        // constants.rs:5
    p->prob_cutoff_ln = log(0.01);  // constants.rs:6
    p->max_number_solns = 10;    // parse_cmd_line.rs:34
    p->max_ploidy = 5;           // parse_cmd_line.rs:43
    p->num_iter_optimize = 20;   // constants.rs:3
    p->ploidy_sensitivity = 2;   // parse_cmd_line.rs:160
    p->stopping_heuristic = 1;
    p->order_model = 0;
    p->block_length = 10000;
    p->reassign_short = 0;
    p->phred_lut = nullptr;
}

int fb_last_timings(const fb_ctx *ctx, fb_timings *out) {
    if (!ctx || !out) return FB_ERR_ARG;
    *out = ctx->tim;
    return FB_OK;
}

void *fb_stream(const fb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// ======================================================================================================================
// data movement
// ======================================================================================================================
// The packed planes of `n_parts` CSR fragment sets concatenated into one contig: the reads of part k follow those of part
// k - 1 and their SNP positions are shifted by pos_shift[k] (batched contigs, fb_multi.cuh; one part with shift 0 is a plain
// upload).  The cells go straight from the caller's buffers to the device (one copy per array and part) and are shifted by
// k_pack: the host only touches per-read metadata.
}  // extern "C"

// pipelined = 1: see fb_frags_upload_pipelined below (one part only)
static int fb_frags_upload_impl(fb_ctx *ctx, uint64_t n_parts, const fb_frags *parts, const uint32_t *pos_shift, int pipelined,
                                fb_dfrags **out) {
    if (!ctx) return FB_ERR_ARG;
    if (!parts || !out || n_parts == 0) FB_FAIL(FB_ERR_ARG, "null argument");
    *out = nullptr;
    FB_CK(cudaSetDevice(ctx->device));
    uint64_t R = 0, NNZ = 0;
    for (uint64_t k = 0; k < n_parts; ++k) {
        R += parts[k].n_reads;
        NNZ += parts[k].nnz;
    }
    if (R >= (1ull << 31)) FB_FAIL(FB_ERR_LIMIT, "too many reads");
    std::unique_ptr<fb_dfrags> df(new fb_dfrags());
    df->ctx = ctx;
    df->n_reads = R;
    df->nnz = NNZ;
    df->h_first.resize(R);
    df->h_last.resize(R);
    df->h_nnz.resize(R);
    df->h_gstart.resize(R);
    df->h_gptr.resize(R + 1);
    df->h_gnum.resize(R);
    df->h_prefmax_last.resize(R);
    std::vector<uint64_t> row(R + 1);
    std::vector<uint32_t> rshift(n_parts > 1 ? R : 0);
    uint64_t ng = 0, r0 = 0, c0 = 0;
    uint32_t pm = 0, pf = 0, pl = 0;
    row[0] = 0;
    for (uint64_t k = 0; k < n_parts; ++k) {
        const fb_frags *fr = &parts[k];
        const uint32_t sh = pos_shift ? pos_shift[k] : 0u;
        const uint64_t Rk = fr->n_reads;
        if (Rk && fr->row_ptr[0] != 0) FB_FAIL(FB_ERR_ARG, "row_ptr[0] must be 0");
        if (Rk && fr->row_ptr[Rk] != fr->nnz) FB_FAIL(FB_ERR_ARG, "row_ptr[n_reads] must equal nnz");
        for (uint64_t i = 0; i < Rk; ++i) {
            const uint64_t a = fr->row_ptr[i], b = fr->row_ptr[i + 1];
            if (b <= a) FB_FAIL(FB_ERR_ARG, "read %llu has no cells", (unsigned long long)(r0 + i));
            if (fr->first[i] < 1 || fr->last[i] < fr->first[i] || fr->last[i] > 0xFFFFFFFFu - sh)
                FB_FAIL(FB_ERR_ARG, "read %llu: bad first/last", (unsigned long long)(r0 + i));
            const uint32_t f = fr->first[i] + sh, l = fr->last[i] + sh;
            // per-cell checks (strictly ascending positions inside [first, last], allele <= 3) run in k_pack on the device
            if (r0 + i > 0) {
                // Frag::cmp (types_structs.rs:87-93): first asc, last desc, counter_id asc
                if (f < pf || (f == pf && l > pl))
                    FB_FAIL(FB_ERR_ARG, "reads must be sorted by Frag::cmp (violated at read %llu)", (unsigned long long)(r0 + i));
            }
            pf = f;
            pl = l;
            const uint64_t x = r0 + i;
            df->h_first[x] = f;
            df->h_last[x] = l;
            df->h_nnz[x] = (uint32_t)(b - a);
            row[x + 1] = c0 + b;
            if (n_parts > 1) rshift[x] = sh;
            const uint32_t g0 = (f - 1) >> 4, g1 = (l - 1) >> 4;
            df->h_gstart[x] = g0;
            ng = (ng + 7) & ~7ULL;  // every read starts on an 8-group boundary (16-byte aligned planes)
            df->h_gptr[x] = (uint32_t)ng;
            df->h_gnum[x] = g1 - g0 + 1;
            ng += (uint64_t)(g1 - g0 + 1);
            if (ng >= (1ull << 32) - 64) FB_FAIL(FB_ERR_LIMIT, "more than 2^32 groups");
            pm = std::max(pm, l);
            df->h_prefmax_last[x] = pm;
        }
        r0 += Rk;
        c0 += fr->nnz;
    }
    ng = (ng + 7) & ~7ULL;
    df->h_gptr[R] = (uint32_t)ng;
    df->n_groups = ng;
    int rc;
    if (pipelined) {
        // ---- chunked, on ctx->stream2: H2D of a chunk of whole reads -> k_pack of that chunk -> *d_ready = reads done.  The caller
        //      computes on ctx->stream meanwhile (k_beam_wide polls d_ready); fb_frags_finish waits and reports cell errors.
        if (!ctx->stream2) FB_CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        cudaStream_t s2 = ctx->stream2;
        const fb_frags *fr = &parts[0];
        const uint64_t chunk_cells = 192ull << 20;  // ~1.2 GB of CSR per chunk (about 20 ms of PCIe)
        uint64_t max_chunk = 0;
        std::vector<uint64_t> cuts(1, 0);  // read indices where chunks start
        for (uint64_t r = 0; r < R;) {
            uint64_t r1 = r + 1;
            while (r1 < R && row[r1 + 1] - row[r] <= chunk_cells) ++r1;
            max_chunk = std::max(max_chunk, row[r1] - row[r]);
            cuts.push_back(r1);
            r = r1;
        }
        uint64_t *d_row = nullptr;
        uint32_t *d_pos[2] = {nullptr, nullptr};
        uint8_t *d_al[2] = {nullptr, nullptr}, *d_q[2] = {nullptr, nullptr};
        auto track = [&](void *p) { df->pipe_temps.push_back(p); };
        df->pipelined = true;
        if ((rc = fb_dalloc(ctx, &d_row, R + 1)) || (track(d_row), 0) || (rc = fb_dalloc(ctx, &df->d_ready, 1)) ||
            (rc = fb_dalloc(ctx, &df->d_pack_err, 1)) || (rc = fb_dalloc(ctx, &df->d_first, R)) || (rc = fb_dalloc(ctx, &df->d_last, R)) ||
            (rc = fb_dalloc(ctx, &df->d_nnz, R)) || (rc = fb_dalloc(ctx, &df->d_gstart, R)) || (rc = fb_dalloc(ctx, &df->d_gptr, R + 1)) ||
            (rc = fb_dalloc(ctx, &df->d_gnum, R)) || (rc = fb_dalloc(ctx, &df->d_qual, ng + 1)) ||
            (rc = fb_dalloc(ctx, &df->d_allele, ng + 1)) || (rc = fb_dalloc(ctx, &df->d_present, ng + 2))) {
            fb_frags_free(ctx, df.release());
            return rc;
        }
        for (int b = 0; b < 2; ++b)
            if ((rc = fb_dalloc(ctx, &d_pos[b], max_chunk)) || (track(d_pos[b]), 0) || (rc = fb_dalloc(ctx, &d_al[b], max_chunk)) ||
                (track(d_al[b]), 0) || (rc = fb_dalloc(ctx, &d_q[b], max_chunk)) || (track(d_q[b]), 0)) {
                fb_frags_free(ctx, df.release());
                return rc;
            }
        cudaEventCreateWithFlags(&df->ev_begin, cudaEventDefault);
        cudaEventCreateWithFlags(&df->ev_first, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&df->ev_done, cudaEventDefault);
        // everything of this upload is ordered on stream2; the memory comes from this context's cache, whose blocks were last
        // used on ctx->stream: order stream2 behind what ctx->stream has queued so far
        cudaEvent_t ev_prev;
        cudaEventCreateWithFlags(&ev_prev, cudaEventDisableTiming);
        cudaEventRecord(ev_prev, ctx->stream);
        cudaStreamWaitEvent(s2, ev_prev, 0);
        cudaEventDestroy(ev_prev);
        cudaEventRecord(df->ev_begin, s2);
        cudaMemcpyAsync(d_row, row.data(), (R + 1) * 8, cudaMemcpyHostToDevice, s2);
        cudaMemcpyAsync(df->d_first, df->h_first.data(), R * 4, cudaMemcpyHostToDevice, s2);
        cudaMemcpyAsync(df->d_last, df->h_last.data(), R * 4, cudaMemcpyHostToDevice, s2);
        cudaMemcpyAsync(df->d_nnz, df->h_nnz.data(), R * 4, cudaMemcpyHostToDevice, s2);
        cudaMemcpyAsync(df->d_gstart, df->h_gstart.data(), R * 4, cudaMemcpyHostToDevice, s2);
        cudaMemcpyAsync(df->d_gptr, df->h_gptr.data(), (R + 1) * 4, cudaMemcpyHostToDevice, s2);
        cudaMemcpyAsync(df->d_gnum, df->h_gnum.data(), R * 4, cudaMemcpyHostToDevice, s2);
        cudaMemsetAsync(df->d_qual, 0xFF, (ng + 1) * sizeof(uint4), s2);
        cudaMemsetAsync(df->d_allele, 0, (ng + 1) * sizeof(uint32_t), s2);
        cudaMemsetAsync(df->d_present, 0, (ng + 2) * sizeof(uint16_t), s2);
        cudaMemsetAsync(df->d_ready, 0, sizeof(unsigned int), s2);
        cudaMemsetAsync(df->d_pack_err, 0xFF, sizeof(unsigned long long), s2);
        // `row` (pageable) must stay alive until its copy has run: the small metadata copies above are synchronous with respect
        // to the host for pageable sources (staged), so they have been consumed when cudaMemcpyAsync returns
        cudaEvent_t ev_buf[2] = {nullptr, nullptr};
        for (size_t k = 0; k + 1 < cuts.size(); ++k) {
            const int b = (int)(k & 1);
            const uint64_t ra = cuts[k], rb = cuts[k + 1], ca = row[ra], cb = row[rb], n = cb - ca;
            if (ev_buf[b]) cudaStreamWaitEvent(s2, ev_buf[b], 0);  // (same stream: already ordered; kept for clarity)
            cudaMemcpyAsync(d_pos[b], fr->pos + ca, n * 4, cudaMemcpyHostToDevice, s2);
            cudaMemcpyAsync(d_al[b], fr->allele + ca, n, cudaMemcpyHostToDevice, s2);
            cudaMemcpyAsync(d_q[b], fr->qual + ca, n, cudaMemcpyHostToDevice, s2);
            k_pack<<<(unsigned)((n + 255) / 256), 256, 0, s2>>>(n, ca, R, d_row, d_pos[b], d_al[b], d_q[b], df->d_gstart, df->d_gptr,
                                                              df->d_first, df->d_last, reinterpret_cast<uint8_t *>(df->d_qual),
                                                              df->d_allele, reinterpret_cast<uint32_t *>(df->d_present),
                                                              df->d_pack_err, nullptr);
            k_set_ready<<<1, 1, 0, s2>>>(df->d_ready, (unsigned int)rb);
            ctx->tim.n_launches += 2;
            if (k == 0) cudaEventRecord(df->ev_first, s2);
        }
        cudaEventRecord(df->ev_done, s2);
        FB_CK(cudaGetLastError());
        df->bytes = ng * 22;
        *out = df.release();
        return FB_OK;
    }
    cudaEvent_t e0 = fb_event(ctx);
    // temporary CSR on the device
    uint64_t *d_row = nullptr;
    uint32_t *d_pos = nullptr, *d_rshift = nullptr;
    uint8_t *d_al = nullptr, *d_q = nullptr;
    unsigned long long *d_err = nullptr;
    unsigned long long h_err = 0;
    auto cleanup = [&]() {
        fb_cache_free(d_err);
        fb_cache_free(d_row);
        fb_cache_free(d_pos);
        fb_cache_free(d_al);
        fb_cache_free(d_q);
        fb_cache_free(d_rshift);
    };
    if ((rc = fb_upload(ctx, &d_row, row)) || (rc = fb_dalloc(ctx, &d_pos, NNZ)) || (rc = fb_dalloc(ctx, &d_al, NNZ)) ||
        (rc = fb_dalloc(ctx, &d_q, NNZ)) || (n_parts > 1 && (rc = fb_upload(ctx, &d_rshift, rshift))) ||
        (rc = fb_upload(ctx, &df->d_first, df->h_first)) || (rc = fb_upload(ctx, &df->d_last, df->h_last)) ||
        (rc = fb_upload(ctx, &df->d_nnz, df->h_nnz)) || (rc = fb_upload(ctx, &df->d_gstart, df->h_gstart)) ||
        (rc = fb_upload(ctx, &df->d_gptr, df->h_gptr)) || (rc = fb_upload(ctx, &df->d_gnum, df->h_gnum)) ||
        (rc = fb_dalloc(ctx, &df->d_qual, ng + 1)) ||
        (rc = fb_dalloc(ctx, &df->d_allele, ng + 1)) || (rc = fb_dalloc(ctx, &df->d_present, ng + 2)) ||
        (rc = fb_dalloc(ctx, &d_err, 1))) {
        cleanup();
        fb_frags_free(ctx, df.release());
        return rc;
    }
    c0 = 0;
    for (uint64_t k = 0; k < n_parts; ++k) {
        const fb_frags *fr = &parts[k];
        if (fr->nnz) {
            cudaMemcpyAsync(d_pos + c0, fr->pos, fr->nnz * 4, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(d_al + c0, fr->allele, fr->nnz, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(d_q + c0, fr->qual, fr->nnz, cudaMemcpyHostToDevice, ctx->stream);
        }
        c0 += fr->nnz;
    }
    cudaEvent_t e1 = fb_event(ctx);
    // absent cells carry quality byte 0xFF (never read: masked by `present`), so a zero byte always is a real q = 0 cell
    cudaMemsetAsync(df->d_qual, 0xFF, (ng + 1) * sizeof(uint4), ctx->stream);
    cudaMemsetAsync(df->d_allele, 0, (ng + 1) * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(df->d_present, 0, (ng + 2) * sizeof(uint16_t), ctx->stream);
    cudaMemsetAsync(d_err, 0xFF, sizeof(unsigned long long), ctx->stream);
    if (NNZ) {
        k_pack<<<(unsigned)((NNZ + 255) / 256), 256, 0, ctx->stream>>>(
            NNZ, 0, R, d_row, d_pos, d_al, d_q, df->d_gstart, df->d_gptr, df->d_first, df->d_last,
            reinterpret_cast<uint8_t *>(df->d_qual), df->d_allele, reinterpret_cast<uint32_t *>(df->d_present), d_err, d_rshift);
        ctx->tim.n_launches++;
    }
    cudaMemcpyAsync(&h_err, d_err, sizeof(h_err), cudaMemcpyDeviceToHost, ctx->stream);
    cudaEvent_t e2 = fb_event(ctx);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    cleanup();
    if (ce != cudaSuccess) {
        ctx->err = std::string("fb_frags_upload: ") + cudaGetErrorString(ce);
        fb_frags_free(ctx, df.release());
        return FB_ERR_CUDA;
    }
    if (h_err != ~0ULL) {
        const uint64_t c = h_err - 1;
        const uint64_t r = std::upper_bound(row.begin(), row.end(), c) - row.begin() - 1;
        uint64_t k = 0, cc = c;
        while (k + 1 < n_parts && cc >= parts[k].nnz) cc -= parts[k++].nnz;
        char b_[256];
        snprintf(b_, sizeof(b_),
                 "read %llu: invalid cell (allele %u at position %u): alleles must be 0..3 and positions strictly "
                 "ascending with first/last_position their min/max",
                 (unsigned long long)r, (unsigned)parts[k].allele[cc], parts[k].pos[cc]);
        ctx->err = b_;
        fb_frags_free(ctx, df.release());
        return FB_ERR_ARG;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ctx->tim.upload_ms += ms;
    cudaEventElapsedTime(&ms, e1, e2);
    ctx->tim.pack_ms += ms;
    df->bytes = ng * 22;
    *out = df.release();
    return FB_OK;
}

// waits for a pipelined upload, reports invalid cells, releases the upload's temporaries (no-op for a plain upload)
static int fb_frags_finish(fb_ctx *ctx, fb_dfrags *df, const fb_frags *fr) {
    if (!df || !df->pipelined) return FB_OK;
    df->pipelined = false;
    cudaError_t ce = cudaStreamSynchronize(ctx->stream2);
    unsigned long long h_err = ~0ULL;
    if (ce == cudaSuccess) ce = cudaMemcpy(&h_err, df->d_pack_err, sizeof(h_err), cudaMemcpyDeviceToHost);
    float ms = 0;
    if (ce == cudaSuccess && cudaEventElapsedTime(&ms, df->ev_begin, df->ev_done) == cudaSuccess) ctx->tim.upload_ms += ms;
    for (void *p : df->pipe_temps) fb_cache_free(p);
    df->pipe_temps.clear();
    if (ce != cudaSuccess) {
        ctx->err = std::string("pipelined upload: ") + cudaGetErrorString(ce);
        return FB_ERR_CUDA;
    }
    if (h_err != ~0ULL) {
        const uint64_t c = h_err - 1;
        const uint64_t r = std::upper_bound(fr->row_ptr, fr->row_ptr + fr->n_reads + 1, c) - fr->row_ptr - 1;
        char b_[256];
        snprintf(b_, sizeof(b_),
                 "read %llu: invalid cell (allele %u at position %u): alleles must be 0..3 and positions strictly "
                 "ascending with first/last_position their min/max",
                 (unsigned long long)r, (unsigned)fr->allele[c], fr->pos[c]);
        ctx->err = b_;
        return FB_ERR_ARG;
    }
    return FB_OK;
}

extern "C" {

int fb_frags_upload_parts(fb_ctx *ctx, uint64_t n_parts, const fb_frags *parts, const uint32_t *pos_shift, fb_dfrags **out) {
    return fb_frags_upload_impl(ctx, n_parts, parts, pos_shift, 0, out);
}

int fb_frags_upload(fb_ctx *ctx, const fb_frags *fr, fb_dfrags **out) {
    if (!ctx) return FB_ERR_ARG;
    if (!fr || !out) FB_FAIL(FB_ERR_ARG, "null argument");
    return fb_frags_upload_impl(ctx, 1, fr, nullptr, 0, out);
}

void fb_frags_free(fb_ctx *ctx, fb_dfrags *df) {
    if (!df) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (df->pipelined && ctx && ctx->stream2) cudaStreamSynchronize(ctx->stream2);  // never free under a running upload
    for (void *p : df->pipe_temps) fb_cache_free(p);
    fb_cache_free(df->d_ready);
    fb_cache_free(df->d_pack_err);
    if (df->ev_first) cudaEventDestroy(df->ev_first);
    if (df->ev_done) cudaEventDestroy(df->ev_done);
    if (df->ev_begin) cudaEventDestroy(df->ev_begin);
    fb_cache_free(df->d_first);
    fb_cache_free(df->d_last);
    fb_cache_free(df->d_nnz);
    fb_cache_free(df->d_gstart);
    fb_cache_free(df->d_gptr);
    fb_cache_free(df->d_gnum);
    fb_cache_free(df->d_qual);
    fb_cache_free(df->d_allele);
    fb_cache_free(df->d_present);
    delete df;
}

uint64_t fb_dfrags_bytes(const fb_dfrags *df) { return df ? df->bytes : 0; }
uint64_t fb_dfrags_nnz(const fb_dfrags *df) { return df ? df->nnz : 0; }
uint64_t fb_dfrags_n_reads(const fb_dfrags *df) { return df ? df->n_reads : 0; }

// ======================================================================================================================
// host-side helpers
// ======================================================================================================================
// utils_frags.rs:405-463 get_range_with_lengths
int64_t fb_get_range_with_lengths(const uint64_t *snp_to_genome_pos, uint64_t n, uint64_t block_length,
                                  uint64_t overlap_len, double minimal_density, uint32_t *lo, uint32_t *hi,
                                  uint64_t cap) {
    if (n == 0) return 0;
    int64_t cnt = 0;
    auto emit = [&](uint32_t a, uint32_t b) {
        if ((uint64_t)cnt < cap) {
            lo[cnt] = a + 1;  // line 461: 1-indexed
            hi[cnt] = b + 1;
        }
        cnt++;
    };
    uint64_t cum_pos = 0, last_pos = snp_to_genome_pos[0];
    uint32_t left_endpoint = 0, new_left_end = 0;
    bool hit_new_left = false;
    for (uint64_t ii = 0; ii < n; ++ii) {
        const uint64_t pos = snp_to_genome_pos[ii];
        const uint32_t i = (uint32_t)ii;
        if (ii == n - 1) {
            emit(left_endpoint, i);
            break;
        }
        if (pos < last_pos) return -1;  // "VCF malformed. Positions are not increasing" (process::exit in the reference)
        cum_pos += pos - last_pos;
        last_pos = pos;
        if (cum_pos > block_length - overlap_len && !hit_new_left) {
            new_left_end = i;
            hit_new_left = true;
        }
        if (cum_pos > block_length) {
            cum_pos = 0;
            const double snp_density = (double)(i - left_endpoint) / (double)block_length;
            if (snp_density > minimal_density) emit(left_endpoint, i - 1);
            if (snp_to_genome_pos[new_left_end] + block_length < snp_to_genome_pos[new_left_end + 1])
                left_endpoint = new_left_end;
            else
                left_endpoint = new_left_end + 1;
            last_pos = snp_to_genome_pos[left_endpoint];
            hit_new_left = false;
        }
    }
    return cnt;
}

// local_clustering.rs:12-59 find_reads_in_interval
int64_t fb_find_reads_in_interval(uint32_t start, uint32_t end, uint64_t n_reads, const uint32_t *first,
                                  const uint32_t *last, uint32_t *out_ids, uint64_t cap) {
    int64_t cnt = 0;
    for (uint64_t i = 0; i < n_reads; ++i) {
        if (last[i] < start) continue;
        if (first[i] > end) break;
        if (last[i] - first[i] > 10000) continue;
        if ((uint64_t)cnt < cap) out_ids[cnt] = (uint32_t)i;
        cnt++;
    }
    return cnt;
}

}  // extern "C"

// ======================================================================================================================
// single-instance helpers for the fine-grained entry points
// ======================================================================================================================
struct Single {
    fb_ctx *ctx;
    fb_dfrags *df = nullptr;
    Engine eng;
    bool own_df = false;
    ~Single() {
        eng.release();
        if (own_df && df) fb_frags_free(ctx, df);
    }
};

static int fb_check_sel(fb_ctx *ctx, const fb_dfrags *df, uint64_t n_sel, const uint32_t *sel) {
    if (n_sel && !sel) FB_FAIL(FB_ERR_ARG, "sel is NULL");
    for (uint64_t i = 0; i < n_sel; ++i) {
        if (sel[i] >= df->n_reads) FB_FAIL(FB_ERR_ARG, "sel[%llu] out of range", (unsigned long long)i);
        if (i && sel[i] <= sel[i - 1]) FB_FAIL(FB_ERR_ARG, "sel must be strictly ascending");
    }
    return FB_OK;
}

// upload frags, build a one-block / one-instance engine, upload the given assignment into buffer 0
static int fb_single_setup(Single &s, fb_ctx *ctx, const fb_frags *fr, uint64_t n_sel, const uint32_t *sel,
                           const uint8_t *hap, uint32_t ploidy, const fb_params *prm) {
    s.ctx = ctx;
    if (!ctx) return FB_ERR_ARG;
    FB_CK(cudaSetDevice(ctx->device));
    ctx->ev_used = 0;
    int rc = fb_check_params(ctx, prm, ploidy);
    if (rc) return rc;
    if (ploidy < 1) FB_FAIL(FB_ERR_ARG, "ploidy must be >= 1");
    if ((rc = fb_frags_upload(ctx, fr, &s.df))) return rc;
    s.own_df = true;
    if ((rc = fb_check_sel(ctx, s.df, n_sel, sel))) return rc;
    s.eng.ctx = ctx;
    s.eng.df = s.df;
    std::vector<uint32_t> reads(sel, sel + n_sel);
    int b = s.eng.add_block(reads);
    s.eng.add_instance(b, ploidy);
    if ((rc = s.eng.finalize_and_upload(prm->epsilon))) return rc;
    if (hap && n_sel) {
        FB_CK(cudaMemcpyAsync(s.eng.d_assign[0], hap, n_sel, cudaMemcpyHostToDevice, ctx->stream));
    }
    return FB_OK;
}

extern "C" {

int fb_score_reads(fb_ctx *ctx, const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                   uint32_t ploidy, const fb_params *prm, double *same, double *diff, int64_t *same_q26,
                   int64_t *diff_q26, uint32_t *n_empty) {
    Single s;
    int rc = fb_single_setup(s, ctx, fr, n_sel, sel, hap, ploidy, prm);
    if (rc) return rc;
    Engine &e = s.eng;
    if ((rc = e.launch_hist(0, 1, 0))) return rc;
    const uint64_t n = n_sel * ploidy;
    double *d_same = nullptr, *d_diff = nullptr;
    long long *d_sq = nullptr, *d_dq = nullptr;
    uint32_t *d_ne = nullptr;
    if ((rc = fb_dalloc(ctx, &d_same, n)) || (rc = fb_dalloc(ctx, &d_diff, n)) || (rc = fb_dalloc(ctx, &d_sq, n)) ||
        (rc = fb_dalloc(ctx, &d_dq, n)) || (rc = fb_dalloc(ctx, &d_ne, n)))
        return rc;
    SweepArgs a = e.sweep_args(FB_SWEEP_SCORE);
    a.o_same = d_same;
    a.o_diff = d_diff;
    a.o_same_q26 = d_sq;
    a.o_diff_q26 = d_dq;
    a.o_nempty = d_ne;
    rc = e.launch_sweep(a);
    if (!rc) {
        if (same) cudaMemcpyAsync(same, d_same, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (diff) cudaMemcpyAsync(diff, d_diff, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (same_q26) cudaMemcpyAsync(same_q26, d_sq, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (diff_q26) cudaMemcpyAsync(diff_q26, d_dq, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (n_empty) cudaMemcpyAsync(n_empty, d_ne, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        if (ce != cudaSuccess) {
            ctx->err = std::string("fb_score_reads: ") + cudaGetErrorString(ce);
            rc = FB_ERR_CUDA;
        }
    }
    fb_cache_free(d_same);
    fb_cache_free(d_diff);
    fb_cache_free(d_sq);
    fb_cache_free(d_dq);
    fb_cache_free(d_ne);
    e.collect_timings();
    return rc;
}

int fb_hap_block_from_partition(fb_ctx *ctx, const fb_frags *fr, uint64_t n_sel, const uint32_t *sel,
                                const uint8_t *hap, uint32_t ploidy, int use_qual, const fb_params *prm,
                                uint32_t pos_lo, uint32_t n_pos, double *counts, uint8_t *key_mask) {
    Single s;
    int rc = fb_single_setup(s, ctx, fr, n_sel, sel, hap, ploidy, prm);
    if (rc) return rc;
    Engine &e = s.eng;
    if ((rc = e.launch_hist(0, use_qual ? 1 : 0, 0))) return rc;
    std::vector<uint64_t> h(e.tot_cnt);
    FB_CK(cudaMemcpyAsync(h.data(), e.d_cnt[0], e.tot_cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CK(cudaStreamSynchronize(ctx->stream));
    const InstDev &in = e.inst[0];
    const uint64_t npos_blk = (uint64_t)in.ng * 16;
    const uint64_t base0 = (uint64_t)in.ag0 * 16;  // position0 of table row 0
    if (counts) memset(counts, 0, sizeof(double) * (size_t)ploidy * n_pos * 4);
    if (key_mask) memset(key_mask, 0, (size_t)ploidy * n_pos);
    for (uint32_t hh = 0; hh < ploidy; ++hh)
        for (uint64_t p = 0; p < npos_blk; ++p) {
            uint64_t pos1 = base0 + p + 1;  // 1-based SNP position
            if (pos1 < pos_lo || pos1 >= (uint64_t)pos_lo + n_pos) continue;
            uint64_t o = pos1 - pos_lo;
            for (int a = 0; a < 4; ++a) {
                uint64_t w = h[((uint64_t)hh * npos_blk + p) * 4 + a];
                if (counts) counts[((uint64_t)hh * n_pos + o) * 4 + a] = fb_q26_to_f64((int64_t)(w & FB_CNT_MASK));
                if (key_mask && (w & FB_PRESENT)) key_mask[(uint64_t)hh * n_pos + o] |= (uint8_t)(1u << a);
            }
        }
    e.collect_timings();
    return FB_OK;
}

int fb_get_mec_stats_epsilon(fb_ctx *ctx, const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                             uint32_t ploidy, int use_phred, const fb_params *prm, double *bases, double *errors) {
    Single s;
    int rc = fb_single_setup(s, ctx, fr, n_sel, sel, hap, ploidy, prm);
    if (rc) return rc;
    Engine &e = s.eng;
    if ((rc = e.launch_hist(0, use_phred ? 1 : 0, 0))) return rc;
    if ((rc = e.launch_mec(0, 0))) return rc;
    std::vector<double> h(ploidy * 2);
    FB_CK(cudaMemcpyAsync(h.data(), e.d_mec[0], ploidy * 2 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < ploidy; ++i) {
        if (bases) bases[i] = h[i * 2];
        if (errors) errors[i] = h[i * 2 + 1];
    }
    e.collect_timings();
    return FB_OK;
}

int fb_optimize_clustering(fb_ctx *ctx, const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap_in,
                           uint32_t ploidy, const fb_params *prm, uint8_t *hap_out, double *score,
                           uint32_t *n_rounds) {
    Single s;
    int rc = fb_single_setup(s, ctx, fr, n_sel, sel, hap_in, ploidy, prm);
    if (rc) return rc;
    Engine &e = s.eng;
    for (uint64_t i = 0; i < n_sel; ++i)
        if (hap_in[i] >= ploidy) FB_FAIL(FB_ERR_ARG, "hap_in[%llu] >= ploidy", (unsigned long long)i);
    if ((rc = e.run_optimize(prm->num_iter_optimize))) return rc;
    InstState st;
    FB_CK(cudaMemcpyAsync(&st, e.d_st, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CK(cudaStreamSynchronize(ctx->stream));
    if (hap_out && n_sel) FB_CK(cudaMemcpy(hap_out, e.d_assign[st.cur], n_sel, cudaMemcpyDeviceToHost));
    if (score) *score = st.prev_score;
    if (n_rounds) *n_rounds = st.accepted;
    e.collect_timings();
    return FB_OK;
}

int fb_beam_search_phasing(fb_ctx *ctx, const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy,
                           const fb_params *prm, uint8_t *hap_out, double *best_score, double *tap_same,
                           double *tap_diff, double *tap_logp, uint64_t tap_cap, uint64_t *tap_n) {
    Single s;
    int rc = fb_single_setup(s, ctx, fr, n_sel, sel, nullptr, ploidy, prm);
    if (rc) return rc;
    Engine &e = s.eng;
    BeamTapDev tap;
    memset(&tap, 0, sizeof(tap));
    double *d_ts = nullptr, *d_td = nullptr, *d_tp = nullptr;
    if (tap_cap) {
        if ((rc = fb_dalloc(ctx, &d_ts, tap_cap)) || (rc = fb_dalloc(ctx, &d_td, tap_cap)) ||
            (rc = fb_dalloc(ctx, &d_tp, tap_cap))) {
            fb_cache_free(d_ts);  // every early return releases what was allocated
            fb_cache_free(d_td);
            fb_cache_free(d_tp);
            return rc;
        }
        tap.same = d_ts;
        tap.diff = d_td;
        tap.logp = d_tp;
        tap.cap = tap_cap;
    }
    BeamRun br;
    rc = fb_run_beam(ctx, e, prm, &tap, br);
    if (!rc) {
        if (hap_out && n_sel) cudaMemcpyAsync(hap_out, e.d_assign[0], n_sel, cudaMemcpyDeviceToHost, ctx->stream);
        if (tap_cap) {
            uint64_t n = std::min<uint64_t>(tap_cap, br.tap_n.empty() ? 0 : br.tap_n[0]);
            if (tap_same) cudaMemcpyAsync(tap_same, d_ts, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (tap_diff) cudaMemcpyAsync(tap_diff, d_td, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (tap_logp) cudaMemcpyAsync(tap_logp, d_tp, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        }
        cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        if (ce != cudaSuccess) {
            ctx->err = std::string("fb_beam_search_phasing: ") + cudaGetErrorString(ce);
            rc = FB_ERR_CUDA;
        }
        if (best_score) *best_score = br.best_score.empty() ? 0.0 : br.best_score[0];
        if (tap_n) *tap_n = br.tap_n.empty() ? 0 : br.tap_n[0];
    }
    fb_cache_free(d_ts);
    fb_cache_free(d_td);
    fb_cache_free(d_tp);
    e.collect_timings();
    return rc;
}

// ======================================================================================================================
// batched hot path: get_local_hap_blocks for every block (graph_processing.rs:103-304, 345-362)
// ======================================================================================================================
int fb_phase_blocks_resident(fb_ctx *ctx, const fb_dfrags *df, uint64_t n_blocks, const uint32_t *blk_lo,
                             const uint32_t *blk_hi, const fb_params *prm, fb_block_results **out) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || !out || (n_blocks && (!blk_lo || !blk_hi))) FB_FAIL(FB_ERR_ARG, "null argument");
    *out = nullptr;
    FB_CK(cudaSetDevice(ctx->device));
    int rc = fb_check_params(ctx, prm, prm ? prm->max_ploidy : 0);
    if (rc) return rc;
    const uint32_t mp = prm->max_ploidy;
    if (mp < 1) FB_FAIL(FB_ERR_ARG, "max_ploidy must be >= 1");
    ctx->ev_used = 0;
    cudaEvent_t ev_start = fb_event(ctx);
    const bool hprof = getenv("FB_HOST_PROF") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tp[6] = {0, 0, 0, 0, 0, 0};  // plan | finalize+upload | beam | optimize | final hist/mec + download | stopping rule
    double t_last = now();
    auto lap = [&](int k) {
        const double t = now();
        tp[k] += t - t_last;
        t_last = t;
    };

    // graph_processing.rs:121-131: the reads of every block; blocks without reads return None.  Planned once (read
    // selection, extents, per-read descriptors), by a few host threads, and appended to the engine of every ploidy wave.
    std::vector<PlannedBlock> planned(n_blocks);
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const uint64_t nt = std::max<uint64_t>(1, std::min<uint64_t>(std::min<unsigned>(8u, hw), n_blocks / 8));
        auto work = [&](uint64_t t) {
            std::vector<uint32_t> reads;
            for (uint64_t j = n_blocks * t / nt; j < n_blocks * (t + 1) / nt; ++j) {
                fb_find_reads(df, blk_lo[j], blk_hi[j], reads);
                fb_plan_block(df, std::move(reads), planned[j]);
                reads = std::vector<uint32_t>();
            }
        };
        std::vector<std::thread> th;
        for (uint64_t t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
    }
    auto block_reads = [&](uint64_t j) -> const std::vector<uint32_t> & { return planned[j].plan.reads; };
    // the per-read descriptors of all blocks, concatenated and uploaded once
    SharedReads shared;
    std::vector<uint32_t> shared_off(n_blocks + 1, 0);
    {
        uint64_t tot_r = 0;
        for (uint64_t j = 0; j < n_blocks; ++j) {
            shared_off[j] = (uint32_t)tot_r;
            tot_r += planned[j].ri.size();
        }
        if (tot_r >= (1ull << 32)) {
            ctx->err = "more than 2^32 (block, read) pairs in one call";
            return FB_ERR_LIMIT;
        }
        shared_off[n_blocks] = (uint32_t)tot_r;
        shared.rinfo.resize(tot_r);
        shared.rextra.resize(tot_r);
        for (uint64_t j = 0; j < n_blocks; ++j) {
            if (planned[j].ri.empty()) continue;
            memcpy(shared.rinfo.data() + shared_off[j], planned[j].ri.data(), planned[j].ri.size() * sizeof(RInfo));
            memcpy(shared.rextra.data() + shared_off[j], planned[j].rx.data(), planned[j].rx.size() * sizeof(RExtra));
            std::vector<RInfo>().swap(planned[j].ri);
            std::vector<RExtra>().swap(planned[j].rx);
        }
        if ((rc = fb_upload(ctx, &shared.d_rinfo, shared.rinfo)) || (rc = fb_upload(ctx, &shared.d_rextra, shared.rextra))) return rc;
    }

    fb_block_results *r = (fb_block_results *)calloc(1, sizeof(fb_block_results));
    r->n_blocks = n_blocks;
    r->max_ploidy = mp;
    r->best_ploidy = (uint32_t *)calloc(n_blocks + 1, sizeof(uint32_t));
    r->ploidies_run = (uint32_t *)calloc(n_blocks + 1, sizeof(uint32_t));
    r->mec_vector = (double *)calloc(n_blocks * mp + 1, sizeof(double));
    r->expected_errors = (double *)calloc(n_blocks * mp + 1, sizeof(double));
    r->read_ptr = (uint64_t *)calloc(n_blocks + 1, sizeof(uint64_t));
    r->block_cells = (uint64_t *)calloc(n_blocks + 1, sizeof(uint64_t));
    uint64_t tot = 0;
    for (uint64_t j = 0; j < n_blocks; ++j) {
        r->read_ptr[j] = tot;
        tot += block_reads(j).size();
    }
    r->read_ptr[n_blocks] = tot;
    r->read_ids = (uint32_t *)calloc(tot + 1, sizeof(uint32_t));
    r->hap = (uint8_t *)calloc(tot + 1, 1);
    for (uint64_t j = 0; j < n_blocks; ++j)
        if (!block_reads(j).empty())
            memcpy(r->read_ids + r->read_ptr[j], block_reads(j).data(), block_reads(j).size() * sizeof(uint32_t));
    auto fail = [&](int code) {
        fb_free_block_results(r);
        return code;
    };

    // ---- the ploidy loop of get_local_hap_blocks (graph_processing.rs:132-252), in WAVES ----------------------------------
    // The reference evaluates ploidy 1, 2, ... for a block until its stopping rule breaks the loop.  Running every ploidy of
    // every block speculatively keeps the GPU full but wastes the ploidies beyond the break (on the mixed-ploidy metagenome of
    // BASELINE.json configs[4]: 35 % of the beam-search and optimize work, and the highest ploidies are the expensive ones).
    // The loop therefore advances in waves: the first wave carries ploidies 1..FB_PLOIDY_WAVE (default 3) of every block,
    // every later wave ONE more ploidy of the blocks whose loop is still running; the stopping rule is applied between waves
    // on the host.  Only the ploidies the reference would evaluate are counted in the work counters, as before.
    const char *wenv = getenv("FB_PLOIDY_WAVE");
    uint32_t wave0 = wenv ? (uint32_t)atoi(wenv) : 3u;
    if (wave0 == 0 || wave0 > mp) wave0 = mp;  // 0: everything in one wave (round-1 behaviour)
    const double epsilon = prm->epsilon;
    std::vector<uint8_t> running(n_blocks, 0);
    std::vector<std::vector<uint8_t>> hap_prev(n_blocks);  // partition of the last evaluated ploidy (the break may fall back to it)
    uint64_t n_running = 0;
    for (uint64_t j = 0; j < n_blocks; ++j)
        if (!block_reads(j).empty()) {
            running[j] = 1;
            ++n_running;
        }
    lap(0);
    float download_ms = 0.f;
    uint64_t sweep_cells_all = 0, hist_cells_all = 0;
    for (uint32_t p_lo = 1; p_lo <= mp && n_running; ) {
        // later waves: ONE more ploidy (FB_PLOIDY_STEP = 2 / 3 measured slower at every share size of configs[4]:
        // profiles/README.md)
        uint32_t wstep = 1u;
        if (const char *se = getenv("FB_PLOIDY_STEP")) wstep = (uint32_t)std::max(1, atoi(se));
        const uint32_t p_hi = p_lo == 1 ? wave0 : std::min(mp, p_lo + wstep - 1);  // inclusive
        Engine e;
        e.ctx = ctx;
        e.df = df;
        e.shared = &shared;
        std::vector<int> first_inst(n_blocks, -1), blk_index(n_blocks, -1);
        for (uint64_t j = 0; j < n_blocks; ++j) {
            if (!running[j]) continue;
            const int b = e.add_shared(planned[j].plan, shared_off[j]);
            blk_index[j] = b;
            for (uint32_t p = p_lo; p <= p_hi; ++p) {
                const int ii = e.add_instance(b, p);
                if (p == p_lo) first_inst[j] = ii;
            }
        }
        lap(0);
        if ((rc = e.finalize_and_upload(prm->epsilon))) return fail(rc);
        lap(1);
        // beam_search_phasing for every instance with ploidy > 1 (ploidy 1: every read lands in haplotype 0)
        BeamRun br;
        if ((rc = fb_run_beam(ctx, e, prm, nullptr, br))) return fail(rc);
        lap(2);
        if ((rc = e.run_optimize(prm->num_iter_optimize))) return fail(rc);
        lap(3);
        // get_mec_stats_epsilon_no_phred on the optimized partition: unweighted histogram into the spare buffer
        if ((rc = e.launch_hist(1, 0, 0, 0, 1))) return fail(rc);
        if ((rc = e.launch_mec(1, 0))) return fail(rc);
        cudaEvent_t ev_compute = fb_event(ctx);
        const int n_inst = e.n_inst();
        std::vector<InstState> st(n_inst);
        std::vector<double> mec0(e.tot_mec * 2), mec1(e.tot_mec * 2);
        std::vector<uint8_t> as0(e.tot_assign), as1(e.tot_assign);
        if (n_inst) {
            cudaError_t ce = cudaMemcpyAsync(st.data(), e.d_st, sizeof(InstState) * n_inst, cudaMemcpyDeviceToHost, ctx->stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(mec0.data(), e.d_mec[0], mec0.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(mec1.data(), e.d_mec[1], mec1.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(as0.data(), e.d_assign[0], as0.size(), cudaMemcpyDeviceToHost, ctx->stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(as1.data(), e.d_assign[1], as1.size(), cudaMemcpyDeviceToHost, ctx->stream);
            if (ce != cudaSuccess) {
                ctx->err = std::string("fb_phase_blocks: ") + cudaGetErrorString(ce);
                return fail(FB_ERR_CUDA);
            }
        }
        cudaEvent_t ev_end = fb_event(ctx);
        cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        if (ce == cudaSuccess) ce = cudaGetLastError();
        if (ce != cudaSuccess) {
            ctx->err = std::string("fb_phase_blocks: ") + cudaGetErrorString(ce);
            return fail(FB_ERR_CUDA);
        }
        lap(4);
        {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev_compute, ev_end);
            download_ms += ms;
        }
        // ---- the loop body and stopping rule for the ploidies of this wave (graph_processing.rs:132-252), on the host ----
        for (uint64_t j = 0; j < n_blocks; ++j) {
            if (!running[j]) continue;
            const BlockPlan &b = e.blocks[blk_index[j]];
            double *mec_vector = r->mec_vector + j * mp;
            double *expected = r->expected_errors + j * mp;
            const uint64_t o = r->read_ptr[j];
            bool broke = false;
            uint32_t best_ploidy = 0;
            for (uint32_t ploidy = p_lo; ploidy <= p_hi; ++ploidy) {
                const int ii = first_inst[j] + (int)(ploidy - p_lo);
                const InstDev &in = e.inst[ii];
                const InstState &s = st[ii];
                best_ploidy = ploidy;
                r->ploidies_run[j] += 1;
                // the no-phred stats were written to the buffer opposite to the accepted one
                const std::vector<double> &mec = (s.cur ^ 1) == 0 ? mec0 : mec1;
                double num_alleles = 0.0;
                for (uint32_t h = 0; h < ploidy; ++h) {
                    const double good = mec[((uint64_t)in.mec_off + h) * 2 + 0];
                    const double bad = mec[((uint64_t)in.mec_off + h) * 2 + 1];
                    mec_vector[ploidy - 1] += bad;  // :159
                    num_alleles += good;
                    num_alleles += bad;
                }
                expected[ploidy - 1] = num_alleles * epsilon;  // :196
                const uint64_t cs = (uint64_t)s.n_opt_iterate * b.nnz, ch = (uint64_t)(s.n_hist + 1) * b.nnz;
                const uint64_t cb = ploidy == 1 ? b.nnz : br.cells_beam[ii];
                r->cells_sweep += cs;
                r->cells_hist += ch;
                r->cells_beam += cb;
                r->block_cells[j] += cs + ch + cb;
                bool fall_back = false;
                if (ploidy > 1) {
                    const double thr = fb_mec_threshold(ploidy, epsilon, prm->ploidy_sensitivity);
                    if ((mec_vector[ploidy - 1] / mec_vector[ploidy - 2]) < thr) {
                    } else if (prm->stopping_heuristic) {
                        best_ploidy -= 1;
                        fall_back = true;
                        broke = true;
                    }
                    if (!broke && mec_vector[ploidy - 1] < expected[ploidy - 1]) broke = true;
                } else {
                    if (mec_vector[ploidy - 1] < expected[ploidy - 1]) broke = true;
                }
                if (!fall_back) {  // this ploidy's partition is the current candidate
                    const std::vector<uint8_t> &as = s.cur == 0 ? as0 : as1;
                    hap_prev[j].assign(as.begin() + in.assign_off, as.begin() + in.assign_off + b.n_reads);
                }
                if (broke) break;
            }
            if (broke || p_hi == mp) {  // loop finished: `best_ploidy` and its partition (hap_prev holds it in both cases)
                r->best_ploidy[j] = best_ploidy;
                if (!hap_prev[j].empty()) memcpy(r->hap + o, hap_prev[j].data(), hap_prev[j].size());
                running[j] = 0;
                --n_running;
                std::vector<uint8_t>().swap(hap_prev[j]);
            }
        }
        for (int ii = 0; ii < n_inst; ++ii) {
            const uint64_t nnz = e.blocks[e.inst[ii].block].nnz;
            sweep_cells_all += (uint64_t)st[ii].n_opt_iterate * nnz;
            hist_cells_all += (uint64_t)(st[ii].n_hist + 1) * nnz;
        }
        e.collect_timings();
        ctx->tim.beam_ms += br.beam_ms;
        lap(5);
        p_lo = p_hi + 1;
    }
    ctx->tim.sweep_cells += sweep_cells_all;
    ctx->tim.hist_cells += hist_cells_all;
    cudaEvent_t ev_done = fb_event(ctx);
    FB_CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_start, ev_done);
    ctx->tim.total_ms += ms - download_ms;
    ctx->tim.download_ms += download_ms;
    *out = r;
    if (hprof)
        fprintf(stderr, "[fb_phase_blocks_resident host ms] plan %.2f | finalize+upload %.2f | beam (launch..sync) %.2f | optimize %.2f | "
                        "final hist/mec + download %.2f | stopping rule + results %.2f\n",
                tp[0], tp[1], tp[2], tp[3], tp[4], tp[5]);
    return FB_OK;
}

int fb_phase_blocks(fb_ctx *ctx, const fb_frags *fr, uint64_t n_blocks, const uint32_t *blk_lo, const uint32_t *blk_hi,
                    const fb_params *prm, fb_block_results **out) {
    if (!ctx) return FB_ERR_ARG;
    fb_dfrags *df = nullptr;
    int rc = fb_frags_upload(ctx, fr, &df);
    if (rc) return rc;
    rc = fb_phase_blocks_resident(ctx, df, n_blocks, blk_lo, blk_hi, prm, out);
    fb_frags_free(ctx, df);
    return rc;
}

// One block at a fixed ploidy on an explicit read list (graph_processing.rs:140-162).
int fb_phase_block_resident(fb_ctx *ctx, const fb_dfrags *df, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy,
                            const fb_params *prm, uint8_t *hap_out, double *mec_bases, double *mec_errors,
                            fb_block_phase *out) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || !out) FB_FAIL(FB_ERR_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    FB_CK(cudaSetDevice(ctx->device));
    int rc = fb_check_params(ctx, prm, ploidy);
    if (rc) return rc;
    if (ploidy < 1) FB_FAIL(FB_ERR_ARG, "ploidy must be >= 1");
    ctx->ev_used = 0;
    cudaEvent_t ev_start = fb_event(ctx);
    std::vector<uint32_t> reads;
    if (sel) {
        if ((rc = fb_check_sel(ctx, df, n_sel, sel))) return rc;
        reads.assign(sel, sel + n_sel);
    } else {
        n_sel = df->n_reads;
        reads.resize(n_sel);
        for (uint64_t i = 0; i < n_sel; ++i) reads[i] = (uint32_t)i;
    }
    if (n_sel == 0) FB_FAIL(FB_ERR_ARG, "a block needs at least one read");
    Engine e;
    e.ctx = ctx;
    e.df = df;
    const int b = e.add_block(reads);
    e.add_instance(b, ploidy);
    if ((rc = e.finalize_and_upload(prm->epsilon))) return rc;
    BeamRun br;
    if ((rc = fb_run_beam(ctx, e, prm, nullptr, br))) return rc;
    if (df->pipelined) cudaStreamWaitEvent(ctx->stream, df->ev_done, 0);  // everything below reads the whole contig
    if ((rc = e.run_optimize(prm->num_iter_optimize))) return rc;
    // get_mec_stats_epsilon_no_phred on the optimized partition: unweighted histogram into the spare buffer
    if ((rc = e.launch_hist(1, 0, 0, 0, 1))) return rc;
    if ((rc = e.launch_mec(1, 0))) return rc;
    cudaEvent_t ev_compute = fb_event(ctx);
    InstState st;
    std::vector<double> mec(ploidy * 2);
    FB_CK(cudaMemcpyAsync(&st, e.d_st, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CK(cudaStreamSynchronize(ctx->stream));
    FB_CK(cudaMemcpyAsync(mec.data(), e.d_mec[st.cur ^ 1], mec.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (hap_out) FB_CK(cudaMemcpyAsync(hap_out, e.d_assign[st.cur], n_sel, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t ev_end = fb_event(ctx);
    FB_CK(cudaStreamSynchronize(ctx->stream));
    FB_CK(cudaGetLastError());
    for (uint32_t h = 0; h < ploidy; ++h) {
        if (mec_bases) mec_bases[h] = mec[h * 2];
        if (mec_errors) mec_errors[h] = mec[h * 2 + 1];
    }
    const uint64_t nnz = e.blocks[b].nnz;
    out->beam_score = ploidy > 1 ? br.best_score[0] : 0.0;
    out->opt_score = st.prev_score;
    out->n_rounds = st.accepted;
    out->ploidy = ploidy;
    out->cells_sweep = (uint64_t)st.n_opt_iterate * nnz;
    out->cells_hist = (uint64_t)(st.n_hist + 1) * nnz;
    out->cells_beam = ploidy == 1 ? nnz : br.cells_beam[0];
    ctx->tim.sweep_cells += out->cells_sweep;
    ctx->tim.hist_cells += out->cells_hist;
    e.collect_timings();
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_start, ev_compute);
    ctx->tim.total_ms += ms;
    cudaEventElapsedTime(&ms, ev_compute, ev_end);
    ctx->tim.download_ms += ms;
    ctx->tim.beam_ms += br.beam_ms;
    return FB_OK;
}

int fb_phase_block(fb_ctx *ctx, const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy,
                   const fb_params *prm, uint8_t *hap_out, double *mec_bases, double *mec_errors, fb_block_phase *out) {
    if (!ctx) return FB_ERR_ARG;
    if (!fr) FB_FAIL(FB_ERR_ARG, "null argument");
    fb_dfrags *df = nullptr;
    // Large inputs: the upload is pipelined with the beam search (chunks of reads are copied and packed on a second stream
    // while k_beam_wide, which consumes the reads in order, already runs on the leading ones).  FB_PIPELINE_UPLOAD=0/1 forces it.
    const char *env = getenv("FB_PIPELINE_UPLOAD");
    const bool pipelined = ploidy >= 2 && (env ? atoi(env) != 0 : fr->nnz >= (256ull << 20));
    int rc = fb_frags_upload_impl(ctx, 1, fr, nullptr, pipelined ? 1 : 0, &df);
    if (rc) return rc;
    rc = fb_phase_block_resident(ctx, df, n_sel, sel, ploidy, prm, hap_out, mec_bases, mec_errors, out);
    const int rc2 = fb_frags_finish(ctx, df, fr);
    fb_frags_free(ctx, df);
    return rc ? rc : rc2;
}

void fb_free_block_results(fb_block_results *r) {
    if (!r) return;
    free(r->best_ploidy);
    free(r->ploidies_run);
    free(r->mec_vector);
    free(r->expected_errors);
    free(r->read_ptr);
    free(r->read_ids);
    free(r->hap);
    free(r->block_cells);
    free(r);
}

// ======================================================================================================================
// rows a14 / a15: final read refinement and HAPQ (part_block_manip.rs)
// ======================================================================================================================
}  // extern "C" (helpers below are C++)

// one engine instance (ploidy 1) per part; returns the per-part sorted read lists
static int fb_parts_engine(fb_ctx *ctx, Engine &e, const fb_dfrags *df, uint64_t n_parts, const uint64_t *part_ptr,
                           const uint32_t *part_reads, std::vector<std::vector<uint32_t>> &parts) {
    parts.assign(n_parts, std::vector<uint32_t>());
    for (uint64_t i = 0; i < n_parts; ++i) {
        std::vector<uint32_t> &v = parts[i];
        v.assign(part_reads + part_ptr[i], part_reads + part_ptr[i + 1]);
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());  // a FxHashSet holds a read once
        for (uint32_t r : v)
            if (r >= df->n_reads) FB_FAIL(FB_ERR_ARG, "part %llu: read id %u out of range", (unsigned long long)i, r);
        int b = e.add_block(v);
        e.add_instance(b, 1);
    }
    return FB_OK;
}

// part_block_manip.rs:27-98 separate_broken_haplogroups (host logic, restated on counter_ids)
static void fb_separate_broken_haplogroups(const fb_dfrags *df, std::vector<std::vector<uint32_t>> &parts,
                                           std::vector<std::pair<uint32_t, uint32_t>> &ranges) {
    const std::vector<uint32_t> &first = df->h_first, &last = df->h_last;
    auto sorted_by_first = [&](const std::vector<uint32_t> &part) {
        std::vector<uint32_t> v = part;  // ascending counter_id == canonical set order
        std::stable_sort(v.begin(), v.end(), [&](uint32_t x, uint32_t y) { return first[x] < first[y]; });
        return v;
    };
    std::vector<std::pair<size_t, std::vector<uint32_t>>> all_breaks;
    for (size_t i = 0; i < ranges.size(); ++i) {
        std::vector<uint32_t> v = sorted_by_first(parts[i]);
        uint32_t current_lastest_pos = 0;
        std::vector<uint32_t> breaks;
        for (uint32_t f : v) {
            if (current_lastest_pos != 0 && first[f] > current_lastest_pos) {
                if (current_lastest_pos >= ranges[i].first && current_lastest_pos < ranges[i].second)
                    breaks.push_back(current_lastest_pos);
            }
            if (last[f] > current_lastest_pos) current_lastest_pos = last[f];
        }
        if (!breaks.empty()) all_breaks.push_back(std::make_pair(i, breaks));
    }
    std::vector<std::vector<uint32_t>> new_parts;
    std::vector<std::pair<uint32_t, uint32_t>> new_ranges;
    for (auto &bi : all_breaks) {
        size_t spot_index = 0;
        const std::vector<uint32_t> &break_spots = bi.second;
        uint32_t break_start = ranges[bi.first].first;
        std::vector<uint32_t> v = sorted_by_first(parts[bi.first]);
        uint32_t end_spot = break_spots[spot_index];
        std::vector<uint32_t> new_part;
        for (uint32_t f : v) {
            if (last[f] <= end_spot) {
                new_part.push_back(f);
            } else {
                // faithful to :71-84: the fragment that triggers the switch is not inserted anywhere
                new_ranges.push_back(std::make_pair(break_start, end_spot));
                new_parts.push_back(std::move(new_part));
                break_start = end_spot + 1;
                spot_index += 1;
                end_spot = spot_index != break_spots.size() ? break_spots[spot_index] : 0xFFFFFFFFu;
                new_part = std::vector<uint32_t>();
            }
        }
        new_ranges.push_back(std::make_pair(break_start, ranges[bi.first].second));
        new_parts.push_back(std::move(new_part));
    }
    for (auto &bi : all_breaks) parts[bi.first].clear();
    for (size_t i = 0; i < new_parts.size(); ++i) {
        std::sort(new_parts[i].begin(), new_parts[i].end());
        parts.push_back(std::move(new_parts[i]));
        ranges.push_back(new_ranges[i]);
    }
}

extern "C" {

int fb_process_reads_for_final_parts_resident(fb_ctx *ctx, const fb_dfrags *df, uint64_t n_parts,
                                              const uint64_t *part_ptr, const uint32_t *part_reads,
                                              const uint32_t *range_lo, const uint32_t *range_hi, const fb_params *prm,
                                              fb_parts **out) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || !out || (n_parts && (!part_ptr || !range_lo || !range_hi))) FB_FAIL(FB_ERR_ARG, "null argument");
    *out = nullptr;
    FB_CK(cudaSetDevice(ctx->device));
    ctx->ev_used = 0;
    int rc = fb_check_params(ctx, prm, 1);
    if (rc) return rc;
    struct {
        const fb_dfrags *df;
    } s{df};
    Engine e;
    e.ctx = ctx;
    e.df = s.df;
    std::vector<std::vector<uint32_t>> parts;
    if ((rc = fb_parts_engine(ctx, e, s.df, n_parts, part_ptr, part_reads, parts))) return rc;
    if ((rc = e.finalize_and_upload(prm->epsilon))) return rc;
    // read -> candidate parts (read_to_parts_map, part_block_manip.rs:185-193), canonical orders
    std::vector<std::pair<uint32_t, uint32_t>> rp;  // (read, part)
    for (uint64_t i = 0; i < n_parts; ++i)
        for (uint32_t r : parts[i]) rp.push_back(std::make_pair(r, (uint32_t)i));
    std::sort(rp.begin(), rp.end());
    std::vector<uint32_t> read_ids, cand;
    std::vector<uint64_t> cand_ptr;
    for (size_t k = 0; k < rp.size(); ++k) {
        if (k == 0 || rp[k].first != rp[k - 1].first) {
            read_ids.push_back(rp[k].first);
            cand_ptr.push_back(cand.size());
        }
        cand.push_back(rp[k].second);
    }
    cand_ptr.push_back(cand.size());
    for (size_t x = 0; x + 1 < cand_ptr.size(); ++x)
        if (cand_ptr[x + 1] - cand_ptr[x] > FB_FINAL_MAXCAND)
            FB_FAIL(FB_ERR_LIMIT, "a read belongs to more than %d haplosets", FB_FINAL_MAXCAND);
    std::vector<uint32_t> chosen(read_ids.size());
    if (!read_ids.empty()) {
        uint32_t *d_ids = nullptr, *d_cand = nullptr, *d_chosen = nullptr;
        uint64_t *d_cptr = nullptr;
        auto cleanup = [&]() {
            fb_cache_free(d_ids);
            fb_cache_free(d_cand);
            fb_cache_free(d_chosen);
            fb_cache_free(d_cptr);
        };
        if ((rc = fb_upload(ctx, &d_ids, read_ids)) || (rc = fb_upload(ctx, &d_cand, cand)) ||
            (rc = fb_upload(ctx, &d_cptr, cand_ptr)) || (rc = fb_dalloc(ctx, &d_chosen, read_ids.size()))) {
            cleanup();
            return rc;
        }
        // every read removed from every haploset: all tables start empty (part_block_manip.rs:195-200)
        cudaMemsetAsync(e.d_cnt[0], 0, std::max<uint64_t>(e.tot_cnt, 1) * 8, ctx->stream);
        cudaMemsetAsync(e.d_masks[0], 0, std::max<uint64_t>(e.tot_mask, 1) * 8, ctx->stream);
        FinalArgs a;
        a.fr = s.df->dev();
        a.inst = e.d_inst;
        a.cnt = e.d_cnt[0];
        a.masks = e.d_masks[0];
        a.lut = ctx->d_lut;
        a.n_active = (uint32_t)read_ids.size();
        a.read_ids = d_ids;
        a.cand_ptr = d_cptr;
        a.cand = d_cand;
        a.chosen = d_chosen;
        a.eps = prm->epsilon;
        a.eps_safe = fb_eps_is_safe(prm->epsilon);
        k_final_assign<<<1, FB_FINAL_THREADS, 0, ctx->stream>>>(a);
        ctx->tim.n_launches++;
        cudaMemcpyAsync(chosen.data(), d_chosen, chosen.size() * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        if (ce == cudaSuccess) ce = cudaGetLastError();
        cleanup();
        if (ce != cudaSuccess) {
            ctx->err = std::string("fb_process_reads_for_final_parts: ") + cudaGetErrorString(ce);
            return FB_ERR_CUDA;
        }
    }
    std::vector<std::vector<uint32_t>> np(n_parts);
    for (size_t x = 0; x < read_ids.size(); ++x) np[chosen[x]].push_back(read_ids[x]);  // ascending by construction
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    for (uint64_t i = 0; i < n_parts; ++i) ranges.push_back(std::make_pair(range_lo[i], range_hi[i]));
    fb_separate_broken_haplogroups(s.df, np, ranges);  // constants.rs:17 SEPARATE_BROKEN_HAPLOGROUPS = true
    // sort_parts (part_block_manip.rs:276-288): stable sort by range
    std::vector<size_t> idx(np.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return ranges[x] < ranges[y]; });
    fb_parts *r = (fb_parts *)calloc(1, sizeof(fb_parts));
    r->n_parts = np.size();
    r->part_ptr = (uint64_t *)calloc(np.size() + 1, sizeof(uint64_t));
    r->range_lo = (uint32_t *)calloc(np.size() + 1, sizeof(uint32_t));
    r->range_hi = (uint32_t *)calloc(np.size() + 1, sizeof(uint32_t));
    uint64_t tot = 0;
    for (size_t k = 0; k < idx.size(); ++k) {
        r->part_ptr[k] = tot;
        tot += np[idx[k]].size();
        r->range_lo[k] = ranges[idx[k]].first;
        r->range_hi[k] = ranges[idx[k]].second;
    }
    r->part_ptr[np.size()] = tot;
    r->read_ids = (uint32_t *)calloc(tot + 1, sizeof(uint32_t));
    uint64_t o = 0;
    for (size_t k = 0; k < idx.size(); ++k)
        for (uint32_t rd : np[idx[k]]) r->read_ids[o++] = rd;
    *out = r;
    return FB_OK;
}

int fb_process_reads_for_final_parts(fb_ctx *ctx, const fb_frags *fr, uint64_t n_parts, const uint64_t *part_ptr,
                                     const uint32_t *part_reads, const uint32_t *range_lo, const uint32_t *range_hi,
                                     const fb_params *prm, fb_parts **out) {
    if (!ctx) return FB_ERR_ARG;
    fb_dfrags *df = nullptr;
    int rc = fb_frags_upload(ctx, fr, &df);
    if (rc) return rc;
    rc = fb_process_reads_for_final_parts_resident(ctx, df, n_parts, part_ptr, part_reads, range_lo, range_hi, prm, out);
    fb_frags_free(ctx, df);
    return rc;
}

void fb_free_parts(fb_parts *r) {
    if (!r) return;
    free(r->part_ptr);
    free(r->read_ids);
    free(r->range_lo);
    free(r->range_hi);
    free(r);
}

int fb_get_hapq(fb_ctx *ctx, const fb_frags *fr, uint64_t n_parts, const uint64_t *part_ptr, const uint32_t *part_reads,
                const uint32_t *range_lo, const uint32_t *range_hi, const uint64_t *snp_to_genome_pos, uint64_t n_snps,
                const fb_params *prm, uint8_t *hapq, double *rel_err, double *avg_err) {
    if (!ctx) return FB_ERR_ARG;
    fb_dfrags *df = nullptr;
    int rc = fb_frags_upload(ctx, fr, &df);
    if (rc) return rc;
    rc = fb_get_hapq_resident(ctx, df, n_parts, part_ptr, part_reads, range_lo, range_hi, snp_to_genome_pos, n_snps, prm, hapq,
                              rel_err, avg_err);
    fb_frags_free(ctx, df);
    return rc;
}

int fb_get_hapq_resident(fb_ctx *ctx, const fb_dfrags *df, uint64_t n_parts, const uint64_t *part_ptr,
                         const uint32_t *part_reads, const uint32_t *range_lo, const uint32_t *range_hi,
                         const uint64_t *snp_to_genome_pos, uint64_t n_snps, const fb_params *prm, uint8_t *hapq,
                         double *rel_err, double *avg_err) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || (n_parts && (!part_ptr || !range_lo || !range_hi || !hapq || !rel_err)) || !snp_to_genome_pos || !avg_err)
        FB_FAIL(FB_ERR_ARG, "null argument");
    FB_CK(cudaSetDevice(ctx->device));
    ctx->ev_used = 0;
    int rc = fb_check_params(ctx, prm, 1);
    if (rc) return rc;
    for (uint64_t i = 0; i < n_parts; ++i)
        if (range_lo[i] < 1 || range_hi[i] > n_snps || range_lo[i] > range_hi[i])
            FB_FAIL(FB_ERR_ARG, "part %llu: snp range outside snp_to_genome_pos", (unsigned long long)i);
    struct {
        const fb_dfrags *df;
    } s{df};
    Engine e;
    e.ctx = ctx;
    e.df = s.df;
    std::vector<std::vector<uint32_t>> parts;
    if ((rc = fb_parts_engine(ctx, e, s.df, n_parts, part_ptr, part_reads, parts))) return rc;
    if ((rc = e.finalize_and_upload(prm->epsilon))) return rc;
    if (n_parts == 0) {
        *avg_err = 0.0 / 0.0;
        return FB_OK;
    }
    // unweighted tables -> buffer 0 (get_errors_cov_from_frags, :529-539); phred tables -> buffer 1 (:541)
    if ((rc = e.launch_hist(2, 0, 0, 0))) return rc;
    if ((rc = e.launch_hist(2, 1, 0, 1))) return rc;
    // find_overlapping_blocks(parts, 0.05, ranges) (part_block_manip.rs:454-515); rust-lapper semantics restated:
    // Lapper::new sorts by (start, stop) (stable); find(start, stop) yields iv.start < stop && iv.stop > start in order.
    struct Iv {
        uint32_t start, stop, val;
    };
    std::vector<Iv> sorted(n_parts);
    for (uint64_t i = 0; i < n_parts; ++i) sorted[i] = Iv{range_lo[i], range_hi[i], (uint32_t)i};
    std::stable_sort(sorted.begin(), sorted.end(), [](const Iv &x, const Iv &y) {
        return x.start != y.start ? x.start < y.start : x.stop < y.stop;
    });
    std::vector<uint32_t> prefmax(n_parts);
    for (uint64_t k = 0; k < n_parts; ++k) prefmax[k] = std::max(k ? prefmax[k - 1] : 0u, sorted[k].stop);
    std::vector<uint32_t> pair_i, pair_j;
    std::vector<double> pair_ol;
    std::vector<uint64_t> ov_ptr(n_parts + 1, 0);
    for (uint64_t i = 0; i < n_parts; ++i) {
        const uint32_t x1 = range_lo[i], x2 = range_hi[i];
        // candidates: start < x2 (upper bound by bisection) and stop > x1 (none before the first prefix max > x1)
        size_t ub = std::lower_bound(sorted.begin(), sorted.end(), x2,
                                     [](const Iv &iv, uint32_t v) { return iv.start < v; }) - sorted.begin();
        size_t lb = std::upper_bound(prefmax.begin(), prefmax.begin() + ub, x1) - prefmax.begin();
        for (size_t k = lb; k < ub; ++k) {
            const Iv &f = sorted[k];
            if (!(f.start < x2 && f.stop > x1)) continue;
            // overlap_percent (part_block_manip.rs:13-24), u32 arithmetic
            const uint32_t aa = x2 - f.start + 1, bb = f.stop - x1 + 1;
            const uint32_t intersect = std::min(aa, bb);
            const uint32_t min_length = x2 - x1 + 1;
            double p = (double)intersect / (double)min_length;
            if (p > 1.) p = 1.;
            if (p > 0.05 && f.val != (uint32_t)i) {
                pair_i.push_back((uint32_t)i);
                pair_j.push_back(f.val);
                pair_ol.push_back(p);
            }
        }
        ov_ptr[i + 1] = pair_i.size();
    }
    const int n_pairs = (int)pair_i.size();
    long long *d_ec = nullptr, *d_pd = nullptr;
    uint32_t *d_lo = nullptr, *d_hi = nullptr, *d_pi = nullptr, *d_pj = nullptr;
    auto cleanup = [&]() {
        fb_cache_free(d_ec);
        fb_cache_free(d_pd);
        fb_cache_free(d_lo);
        fb_cache_free(d_hi);
        fb_cache_free(d_pi);
        fb_cache_free(d_pj);
    };
    if ((rc = fb_dalloc(ctx, &d_ec, n_parts * 2)) || (rc = fb_dalloc(ctx, &d_pd, (size_t)n_pairs * 2)) ||
        (rc = fb_upload(ctx, &d_lo, range_lo, n_parts)) || (rc = fb_upload(ctx, &d_hi, range_hi, n_parts)) ||
        (rc = fb_upload(ctx, &d_pi, pair_i)) || (rc = fb_upload(ctx, &d_pj, pair_j))) {
        cleanup();
        return rc;
    }
    ErrCovArgs ea;
    ea.inst = e.d_inst;
    ea.n_parts = (int)n_parts;
    ea.cnt = e.d_cnt[0];
    ea.range_lo = d_lo;
    ea.range_hi = d_hi;
    ea.out = d_ec;
    k_errors_cov<<<(unsigned)((n_parts + 7) / 8), 256, 0, ctx->stream>>>(ea);
    ctx->tim.n_launches++;
    if (n_pairs) {
        HapDistArgs ha;
        ha.inst = e.d_inst;
        ha.cnt = e.d_cnt[1];
        ha.pair_i = d_pi;
        ha.pair_j = d_pj;
        ha.n_pairs = n_pairs;
        ha.out = d_pd;
        k_hap_distance<<<(unsigned)((n_pairs + 7) / 8), 256, 0, ctx->stream>>>(ha);
        ctx->tim.n_launches++;
    }
    std::vector<long long> ec(n_parts * 2), pd((size_t)n_pairs * 2 + 1);
    cudaMemcpyAsync(ec.data(), d_ec, n_parts * 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (n_pairs) cudaMemcpyAsync(pd.data(), d_pd, (size_t)n_pairs * 16, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    cleanup();
    if (ce != cudaSuccess) {
        ctx->err = std::string("fb_get_hapq: ") + cudaGetErrorString(ce);
        return FB_ERR_CUDA;
    }
    // part_block_manip.rs:523-540
    double weight = 0., error = 0.;
    std::vector<double> errs(n_parts);
    for (uint64_t i = 0; i < n_parts; ++i) {
        const double total_cov = (double)ec[i * 2], total_err = (double)ec[i * 2 + 1];
        weight += total_cov;
        error += total_err;
        errs[i] = total_err / total_cov;
    }
    const double avg = error / weight;
    const double HAPQ_CONSTANT = 40.;  // constants.rs:22
    for (uint64_t i = 0; i < n_parts; ++i) {
        double max_penalty = 0.;
        for (uint64_t k = ov_ptr[i]; k < ov_ptr[i + 1]; ++k) {
            const double same = (double)pd[k * 2], diff = (double)pd[k * 2 + 1];
            const double dist = (same + diff) == 0. ? 1. : diff / (same + diff);
            const double ol = pair_ol[k];
            if (ol * (1. - dist) > max_penalty) max_penalty = ol * (1. - dist);
        }
        uint32_t r0 = 0xFFFFFFFFu, r1 = 0;
        for (uint32_t rd : parts[i]) {
            if (s.df->h_first[rd] < r0) r0 = s.df->h_first[rd];
            if (s.df->h_last[rd] >= r1) r1 = s.df->h_last[rd];
        }
        uint64_t base_range = 0;
        if (!(r0 > r1)) base_range = snp_to_genome_pos[range_hi[i] - 1] - snp_to_genome_pos[range_lo[i] - 1];
        const double t1 = HAPQ_CONSTANT * (1. - max_penalty);
        const double t2 = std::min(1., (double)parts[i].size() / 3.);
        const double t3 = std::max(0.0, log(((double)base_range / (double)prm->block_length) + 1.));
        unsigned long long hq = fb_as_usize(t1 * t2 * t3);
        if (parts[i].size() == 1) hq = 0;
        hapq[i] = (uint8_t)std::min<unsigned long long>(hq, 60);
        rel_err[i] = errs[i] / avg;
    }
    *avg_err = avg;
    e.collect_timings();
    return FB_OK;
}

int fb_update_hap_graph(fb_ctx *ctx, const fb_frags *fr, uint64_t n_cols, const uint64_t *col_ptr,
                        const uint64_t *node_ptr, const uint32_t *node_reads, const uint32_t *node_lo,
                        const uint32_t *node_hi, const fb_params *prm, double *out_weights) {
    if (!ctx) return FB_ERR_ARG;
    fb_dfrags *df = nullptr;
    int rc = fb_frags_upload(ctx, fr, &df);
    if (rc) return rc;
    rc = fb_update_hap_graph_resident(ctx, df, n_cols, col_ptr, node_ptr, node_reads, node_lo, node_hi, prm, out_weights);
    fb_frags_free(ctx, df);
    return rc;
}

int fb_update_hap_graph_resident(fb_ctx *ctx, const fb_dfrags *df, uint64_t n_cols, const uint64_t *col_ptr,
                                 const uint64_t *node_ptr, const uint32_t *node_reads, const uint32_t *node_lo,
                                 const uint32_t *node_hi, const fb_params *prm, double *out_weights) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || !col_ptr || !node_ptr || !node_lo || !node_hi || !out_weights) FB_FAIL(FB_ERR_ARG, "null argument");
    FB_CK(cudaSetDevice(ctx->device));
    ctx->ev_used = 0;
    int rc = fb_check_params(ctx, prm, 1);
    if (rc) return rc;
    const uint64_t n_nodes = col_ptr[n_cols];
    struct {
        const fb_dfrags *df;
    } s{df};
    Engine e;
    e.ctx = ctx;
    e.df = s.df;
    std::vector<std::vector<uint32_t>> nodes;
    if ((rc = fb_parts_engine(ctx, e, s.df, n_nodes, node_ptr, node_reads, nodes))) return rc;
    // HapNode::new keeps only positions inside snp_endpoints (types_structs.rs:173)
    for (uint64_t v = 0; v < n_nodes; ++v) {
        InstDev &in = e.inst[v];
        const long long base0 = (long long)in.ag0 * 16;  // position0 of table row 0
        long long lo = (long long)node_lo[v] - 1 - base0, hi = (long long)node_hi[v] - 1 - base0;
        if (hi < 0 || lo > (long long)in.ng * 16) {  // nothing inside: an empty filter range
            in.flt_lo = 1;
            in.flt_hi = 0;
        } else {
            in.flt_lo = (uint32_t)std::max<long long>(lo, 0);
            in.flt_hi = (uint32_t)std::min<long long>(hi, 0xFFFFFFFELL);
        }
    }
    if ((rc = e.finalize_and_upload(prm->epsilon))) return rc;
    std::vector<uint64_t> out_off(n_nodes, 0), gprefix(n_nodes + 1, 0);
    std::vector<uint32_t> next_first(n_nodes, 0xFFFFFFFFu), next_count(n_nodes, 0), item_node, item_read;
    uint64_t tot_out = 0;
    for (uint64_t i = 0; i + 1 < n_cols; ++i)
        for (uint64_t v = col_ptr[i]; v < col_ptr[i + 1]; ++v) {
            next_first[v] = (uint32_t)col_ptr[i + 1];
            next_count[v] = (uint32_t)(col_ptr[i + 2] - col_ptr[i + 1]);
            out_off[v] = tot_out;
            tot_out += next_count[v];
            for (uint32_t r : nodes[v]) {
                item_node.push_back((uint32_t)v);
                item_read.push_back(r);
            }
        }
    for (uint64_t v = 0; v < n_nodes; ++v) gprefix[v + 1] = gprefix[v] + e.inst[v].ng;
    // sorted, de-duplicated node read lists for the membership test
    std::vector<uint64_t> nptr(n_nodes + 1, 0);
    std::vector<uint32_t> nreads;
    for (uint64_t v = 0; v < n_nodes; ++v) {
        nreads.insert(nreads.end(), nodes[v].begin(), nodes[v].end());
        nptr[v + 1] = nreads.size();
    }
    for (uint64_t k = 0; k < tot_out; ++k) out_weights[k] = 0.0;
    if (n_nodes == 0 || item_node.empty() || tot_out == 0) return FB_OK;
    if ((rc = e.launch_hist(2, 1, 0, 0))) return rc;  // phred-weighted hap_map of every node -> buffer 0
    uint64_t *d_gprefix = nullptr, *d_nptr = nullptr, *d_out_off = nullptr;
    uint32_t *d_next_first = nullptr, *d_next_count = nullptr, *d_item_node = nullptr, *d_item_read = nullptr,
             *d_nreads = nullptr;
    unsigned int *d_out = nullptr;
    uint4 *d_planes = nullptr;
    auto cleanup = [&]() {
        fb_cache_free(d_gprefix);
        fb_cache_free(d_nptr);
        fb_cache_free(d_out_off);
        fb_cache_free(d_next_first);
        fb_cache_free(d_next_count);
        fb_cache_free(d_item_node);
        fb_cache_free(d_item_read);
        fb_cache_free(d_nreads);
        fb_cache_free(d_out);
        fb_cache_free(d_planes);
    };
    if ((rc = fb_upload(ctx, &d_gprefix, gprefix)) || (rc = fb_upload(ctx, &d_nptr, nptr)) ||
        (rc = fb_upload(ctx, &d_out_off, out_off)) || (rc = fb_upload(ctx, &d_next_first, next_first)) ||
        (rc = fb_upload(ctx, &d_next_count, next_count)) || (rc = fb_upload(ctx, &d_item_node, item_node)) ||
        (rc = fb_upload(ctx, &d_item_read, item_read)) || (rc = fb_upload(ctx, &d_nreads, nreads)) ||
        (rc = fb_dalloc(ctx, &d_out, tot_out)) || (rc = fb_dalloc(ctx, &d_planes, gprefix[n_nodes]))) {
        cleanup();
        return rc;
    }
    cudaMemsetAsync(d_out, 0, tot_out * 4, ctx->stream);
    if (gprefix[n_nodes]) {
        k_planes_a4<<<(unsigned)((gprefix[n_nodes] + 127) / 128), 128, 0, ctx->stream>>>(e.d_inst, (int)n_nodes, d_gprefix,
                                                                                       e.d_cnt[0], d_planes);
        ctx->tim.n_launches++;
    }
    EdgeArgs a;
    a.fr = s.df->dev();
    a.inst = e.d_inst;
    a.group_prefix = d_gprefix;
    a.planes = d_planes;
    a.lut = ctx->d_lut;
    a.n_items = item_node.size();
    a.item_node = d_item_node;
    a.item_read = d_item_read;
    a.next_first = d_next_first;
    a.next_count = d_next_count;
    a.node_ptr = d_nptr;
    a.node_reads = d_nreads;
    a.out_off = d_out_off;
    a.out = d_out;
    k_edge_score<<<(unsigned)((a.n_items + 7) / 8), 256, 0, ctx->stream>>>(a);
    ctx->tim.n_launches++;
    std::vector<unsigned int> h_out(tot_out);
    cudaMemcpyAsync(h_out.data(), d_out, tot_out * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    cleanup();
    if (ce != cudaSuccess) {
        ctx->err = std::string("fb_update_hap_graph: ") + cudaGetErrorString(ce);
        return FB_ERR_CUDA;
    }
    for (uint64_t k = 0; k < tot_out; ++k) out_weights[k] = (double)h_out[k];
    e.collect_timings();
    return FB_OK;
}

}  // extern "C"

#include "fb_multi.cuh"
#include "fb_bench.cuh"
