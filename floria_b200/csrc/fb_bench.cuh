// fb_bench.cuh — measurement helpers (include/floria_b200_bench.h): device-side synthetic dense block + timed
// sweep/hist loop.  Not part of the reference-facing boundary.
#pragma once
#include "../../include/floria_b200_bench.h"
#include "fb_engine.cuh"

// counter-based PRNG of floria_b200/synth.py
__device__ __forceinline__ uint64_t fbs_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t fbs_u64(uint64_t seed, uint64_t stream, uint64_t idx) {
    return fbs_mix(seed + stream * 0xD1B54A32D192ED03ULL + (idx + 1) * 0x9E3779B97F4A7C15ULL);
}
__device__ __forceinline__ double fbs_u01(uint64_t seed, uint64_t stream, uint64_t idx) {
    return (double)(fbs_u64(seed, stream, idx) >> 11) * (1.0 / 9007199254740992.0);
}

// one thread per (read, group)
__global__ void k_synth_dense(uint64_t n_reads, uint32_t n_snps, uint32_t ng_per_read, uint32_t stride, uint64_t seed,
                              double present,
                              double flip, const uint8_t *__restrict__ truth, const uint8_t *__restrict__ nall,
                              const uint8_t *__restrict__ src, uint4 *__restrict__ qual, uint32_t *__restrict__ allele,
                              uint16_t *__restrict__ pres, uint32_t *__restrict__ nnz) {
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_reads * ng_per_read) return;
    const uint64_t r = x / ng_per_read;
    const uint32_t gl = (uint32_t)(x % ng_per_read);
    const uint32_t h = src[r];
    uint32_t qq[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};  // absent cells: 0xFF
    uint32_t al = 0, pr = 0;
    for (uint32_t k = 0; k < 16; ++k) {
        const uint32_t off = gl * 16 + k;  // position0 == offset inside the (full-span) read
        if (off >= n_snps) break;
        const uint64_t key = r * (1ULL << 20) + off;
        bool keep = fbs_u01(seed, 200, key) < present;
        keep |= (off == 0) | (off == n_snps - 1);
        if (!keep) continue;
        const uint32_t t = truth[(uint64_t)h * n_snps + off];
        const uint32_t na = nall[off];
        const bool do_flip = fbs_u01(seed, 201, key) < flip;
        const uint32_t shift = 1 + (uint32_t)(fbs_u64(seed, 202, key) % (uint64_t)(na - 1));
        const uint32_t a = do_flip ? (t + shift) % na : t;
        const uint32_t q = 5 + (uint32_t)(fbs_u64(seed, 203, key) % 36ULL);
        pr |= 1u << k;
        al |= ((a & 1u) << k) | (((a >> 1) & 1u) << (16 + k));
        qq[k >> 2] = (qq[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | (q << (8 * (k & 3)));
    }
    const uint64_t go = r * stride + gl;
    qual[go] = make_uint4(qq[0], qq[1], qq[2], qq[3]);
    allele[go] = al;
    pres[go] = (uint16_t)pr;
    atomicAdd(&nnz[r], (uint32_t)__popc(pr));
}

extern "C" {

int fb_bench_synth_dense(fb_ctx *ctx, uint64_t n_reads, uint32_t n_snps, uint32_t ploidy, uint64_t seed, double present,
                         double flip, const uint8_t *truth, const uint8_t *nall, const uint8_t *src, fb_dfrags **out) {
    if (!ctx) return FB_ERR_ARG;
    if (!out || !truth || !nall || !src || n_reads == 0 || n_snps == 0 || n_snps >= (1u << 20))
        FB_FAIL(FB_ERR_ARG, "bad argument");
    *out = nullptr;
    FB_CK(cudaSetDevice(ctx->device));
    const uint32_t ngr = (n_snps + 15) / 16;
    const uint32_t stride = (ngr + 7) & ~7u;  // reads start on 8-group boundaries (16-byte aligned planes)
    const uint64_t ng = n_reads * (uint64_t)stride;
    if (ng >= (1ull << 32) - 64) FB_FAIL(FB_ERR_LIMIT, "more than 2^32 groups");
    fb_dfrags *df = new fb_dfrags();
    df->ctx = ctx;
    df->n_reads = n_reads;
    df->n_groups = ng;
    df->h_first.assign(n_reads, 1);
    df->h_last.assign(n_reads, n_snps);
    df->h_gstart.assign(n_reads, 0);
    df->h_gptr.resize(n_reads + 1);
    df->h_prefmax_last.assign(n_reads, n_snps);
    df->h_nnz.resize(n_reads);
    for (uint64_t i = 0; i <= n_reads; ++i) df->h_gptr[i] = (uint32_t)(i * stride);
    df->h_gnum.assign(n_reads, ngr);
    uint8_t *d_truth = nullptr, *d_nall = nullptr, *d_src = nullptr;
    int rc;
    if ((rc = fb_upload(ctx, &df->d_first, df->h_first)) || (rc = fb_upload(ctx, &df->d_last, df->h_last)) ||
        (rc = fb_upload(ctx, &df->d_gstart, df->h_gstart)) || (rc = fb_upload(ctx, &df->d_gptr, df->h_gptr)) ||
        (rc = fb_upload(ctx, &df->d_gnum, df->h_gnum)) || (rc = fb_dalloc(ctx, &df->d_nnz, n_reads)) || (rc = fb_dalloc(ctx, &df->d_qual, ng + 1)) ||
        (rc = fb_dalloc(ctx, &df->d_allele, ng + 1)) || (rc = fb_dalloc(ctx, &df->d_present, ng + 2)) ||
        (rc = fb_upload(ctx, &d_truth, truth, (size_t)ploidy * n_snps)) || (rc = fb_upload(ctx, &d_nall, nall, n_snps)) ||
        (rc = fb_upload(ctx, &d_src, src, n_reads))) {
        fb_frags_free(ctx, df);
        return rc;
    }
    cudaMemsetAsync(df->d_nnz, 0, n_reads * 4, ctx->stream);
    cudaMemsetAsync(df->d_qual, 0xFF, (ng + 1) * sizeof(uint4), ctx->stream);
    cudaMemsetAsync(df->d_allele, 0, (ng + 1) * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(df->d_present, 0, (ng + 2) * sizeof(uint16_t), ctx->stream);
    k_synth_dense<<<(unsigned)((n_reads * (uint64_t)ngr + 255) / 256), 256, 0, ctx->stream>>>(n_reads, n_snps, ngr, stride, seed, present, flip, d_truth,
                                                                          d_nall, d_src, df->d_qual, df->d_allele,
                                                                          df->d_present, df->d_nnz);
    cudaMemcpyAsync(df->h_nnz.data(), df->d_nnz, n_reads * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    fb_cache_free(d_truth);
    fb_cache_free(d_nall);
    fb_cache_free(d_src);
    if (ce != cudaSuccess) {
        ctx->err = std::string("fb_bench_synth_dense: ") + cudaGetErrorString(ce);
        fb_frags_free(ctx, df);
        return FB_ERR_CUDA;
    }
    uint64_t nnz = 0;
    for (uint64_t i = 0; i < n_reads; ++i) nnz += df->h_nnz[i];
    df->nnz = nnz;
    df->bytes = ng * 22;
    *out = df;
    return FB_OK;
}

int fb_bench_sweep_hist(fb_ctx *ctx, const fb_dfrags *df, uint32_t ploidy, const uint8_t *hap, const fb_params *prm,
                        uint32_t iters, float *sweep_ms, float *hist_ms, uint64_t *cells) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || !hap || !prm) FB_FAIL(FB_ERR_ARG, "null argument");
    FB_CK(cudaSetDevice(ctx->device));
    ctx->ev_used = 0;
    int rc = fb_check_params(ctx, prm, ploidy);
    if (rc) return rc;
    Engine e;
    e.ctx = ctx;
    e.df = df;
    std::vector<uint32_t> reads(df->n_reads);
    for (uint64_t i = 0; i < df->n_reads; ++i) reads[i] = (uint32_t)i;
    int b = e.add_block(reads);
    e.add_instance(b, ploidy);
    if ((rc = e.finalize_and_upload(prm->epsilon))) return rc;
    FB_CK(cudaMemcpyAsync(e.d_assign[0], hap, df->n_reads, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = e.launch_sizes(0))) return rc;
    if ((rc = e.launch_hist(0, 1, 0))) return rc;  // warm-up + the table the sweep reads
    e.hist_ev.clear();
    for (uint32_t it = 0; it < iters; ++it) {
        if ((rc = e.launch_sweep(e.sweep_args(FB_SWEEP_MOVES)))) return rc;
        if ((rc = e.launch_hist(0, 1, 0))) return rc;
    }
    FB_CK(cudaStreamSynchronize(ctx->stream));
    FB_CK(cudaGetLastError());
    for (uint32_t it = 0; it < iters; ++it) {
        cudaEventElapsedTime(&sweep_ms[it], e.sweep_ev[it].first, e.sweep_ev[it].second);
        cudaEventElapsedTime(&hist_ms[it], e.hist_ev[it].first, e.hist_ev[it].second);
    }
    if (cells) *cells = e.blocks[0].nnz;
    return FB_OK;
}

int fb_bench_block_tables(fb_ctx *ctx, const fb_dfrags *df, uint32_t ploidy, const uint8_t *hap, const fb_params *prm,
                          uint64_t *n_pos, uint64_t *counts, int64_t *same_q26, int64_t *diff_q26, uint32_t *n_empty) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || !hap || !prm) FB_FAIL(FB_ERR_ARG, "null argument");
    FB_CK(cudaSetDevice(ctx->device));
    ctx->ev_used = 0;
    int rc = fb_check_params(ctx, prm, ploidy);
    if (rc) return rc;
    Engine e;
    e.ctx = ctx;
    e.df = df;
    std::vector<uint32_t> reads(df->n_reads);
    for (uint64_t i = 0; i < df->n_reads; ++i) reads[i] = (uint32_t)i;
    int b = e.add_block(reads);
    e.add_instance(b, ploidy);
    if ((rc = e.finalize_and_upload(prm->epsilon))) return rc;
    const uint64_t npos = (uint64_t)e.inst[0].ng * 16;
    if (n_pos) *n_pos = npos;
    if (!counts && !same_q26 && !diff_q26 && !n_empty) return FB_OK;  // size query
    FB_CK(cudaMemcpyAsync(e.d_assign[0], hap, df->n_reads, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = e.launch_sizes(0))) return rc;
    if ((rc = e.launch_hist(0, 1, 0))) return rc;
    if (counts) FB_CK(cudaMemcpyAsync(counts, e.d_cnt[0], e.tot_cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
    const uint64_t n = df->n_reads * (uint64_t)ploidy;
    long long *d_sq = nullptr, *d_dq = nullptr;
    uint32_t *d_ne = nullptr;
    if ((rc = fb_dalloc(ctx, &d_sq, n)) || (rc = fb_dalloc(ctx, &d_dq, n)) || (rc = fb_dalloc(ctx, &d_ne, n))) return rc;
    SweepArgs a = e.sweep_args(FB_SWEEP_SCORE);
    a.o_same_q26 = d_sq;
    a.o_diff_q26 = d_dq;
    a.o_nempty = d_ne;
    rc = e.launch_sweep(a);
    if (!rc) {
        if (same_q26) cudaMemcpyAsync(same_q26, d_sq, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (diff_q26) cudaMemcpyAsync(diff_q26, d_dq, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (n_empty) cudaMemcpyAsync(n_empty, d_ne, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        if (ce == cudaSuccess) ce = cudaGetLastError();
        if (ce != cudaSuccess) {
            ctx->err = std::string("fb_bench_block_tables: ") + cudaGetErrorString(ce);
            rc = FB_ERR_CUDA;
        }
    }
    fb_cache_free(d_sq);
    fb_cache_free(d_dq);
    fb_cache_free(d_ne);
    return rc;
}

// warp per read: the packed planes of a read back into CSR cells (ascending positions)
__global__ void k_unpack_csr(DFragsDev fr, uint64_t r0, uint64_t r1, const uint64_t *__restrict__ row_ptr /*[r1-r0+1], relative*/,
                             uint32_t *__restrict__ pos, uint8_t *__restrict__ allele, uint8_t *__restrict__ qual) {
    const uint64_t r = r0 + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= r1) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t g0 = fr.gptr[r], ng = fr.gnum[r], gs = fr.gstart[r];
    uint64_t o = row_ptr[r - r0];
    const uint8_t *q8 = reinterpret_cast<const uint8_t *>(fr.qual);
    for (uint32_t b = 0; b < ng; b += 32) {
        const uint32_t x = b + lane;
        const uint32_t pr = x < ng ? fr.present[g0 + x] : 0u;
        const uint32_t n = __popc(pr);
        uint32_t pre = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pre, d);
            if ((int)lane >= d) pre += t;
        }
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, pre, 31);
        if (n) {
            const uint32_t al = fr.allele[g0 + x];
            uint64_t w = o + pre - n;
            for (uint32_t bits = pr; bits;) {
                const int k = __ffs(bits) - 1;
                bits &= bits - 1;
                pos[w] = (gs + x) * 16u + (uint32_t)k + 1u;
                allele[w] = (uint8_t)(((al >> k) & 1u) | (((al >> (16 + k)) & 1u) << 1));
                qual[w] = q8[(uint64_t)(g0 + x) * 16 + k];
                ++w;
            }
        }
        o += tot;
    }
}

// Host CSR of a resident contig (the form a floria host would hand to fb_phase_block / fb_phase_blocks): row_ptr
// [n_reads+1], first / last [n_reads], pos / allele / qual [nnz]; the destination buffers may be pinned.  Used by bench.py
// to obtain HOST buffers of the 100k x 50k block for the end-to-end leg without a minutes-long host-side generator.
int fb_bench_export_csr(fb_ctx *ctx, const fb_dfrags *df, uint64_t *row_ptr, uint32_t *first, uint32_t *last,
                        uint32_t *pos, uint8_t *allele, uint8_t *qual) {
    if (!ctx) return FB_ERR_ARG;
    if (!df || !row_ptr || !first || !last || !pos || !allele || !qual) FB_FAIL(FB_ERR_ARG, "null argument");
    FB_CK(cudaSetDevice(ctx->device));
    const uint64_t R = df->n_reads;
    row_ptr[0] = 0;
    for (uint64_t i = 0; i < R; ++i) {
        row_ptr[i + 1] = row_ptr[i] + df->h_nnz[i];
        first[i] = df->h_first[i];
        last[i] = df->h_last[i];
    }
    const uint64_t chunk_cells = 256ull << 20;  // cells per device staging chunk (1.5 GB of temporaries)
    uint64_t *d_row = nullptr;
    uint32_t *d_pos = nullptr;
    uint8_t *d_al = nullptr, *d_q = nullptr;
    int rc = FB_OK;
    auto cleanup = [&]() {
        fb_cache_free(d_row);
        fb_cache_free(d_pos);
        fb_cache_free(d_al);
        fb_cache_free(d_q);
    };
    std::vector<uint64_t> rel;
    uint64_t cap_cells = 0, cap_rows = 0;
    for (uint64_t r0 = 0; r0 < R;) {
        uint64_t r1 = r0 + 1;
        while (r1 < R && row_ptr[r1 + 1] - row_ptr[r0] <= chunk_cells) ++r1;
        const uint64_t cells = row_ptr[r1] - row_ptr[r0], rows = r1 - r0;
        if (cells > cap_cells || rows > cap_rows) {
            cleanup();
            cap_cells = std::max(cells, chunk_cells);
            cap_rows = std::max<uint64_t>(rows, 1 << 16);
            if ((rc = fb_dalloc(ctx, &d_row, cap_rows + 1)) || (rc = fb_dalloc(ctx, &d_pos, cap_cells)) ||
                (rc = fb_dalloc(ctx, &d_al, cap_cells)) || (rc = fb_dalloc(ctx, &d_q, cap_cells))) {
                cleanup();
                return rc;
            }
        }
        rel.resize(rows + 1);
        for (uint64_t i = 0; i <= rows; ++i) rel[i] = row_ptr[r0 + i] - row_ptr[r0];
        cudaMemcpyAsync(d_row, rel.data(), (rows + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
        k_unpack_csr<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(df->dev(), r0, r1, d_row, d_pos, d_al, d_q);
        cudaMemcpyAsync(pos + row_ptr[r0], d_pos, cells * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(allele + row_ptr[r0], d_al, cells, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(qual + row_ptr[r0], d_q, cells, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        if (ce == cudaSuccess) ce = cudaGetLastError();
        if (ce != cudaSuccess) {
            ctx->err = std::string("fb_bench_export_csr: ") + cudaGetErrorString(ce);
            cleanup();
            return FB_ERR_CUDA;
        }
        r0 = r1;
    }
    cleanup();
    return FB_OK;
}

int fb_bench_download_planes(fb_ctx *ctx, const fb_dfrags *df, uint64_t *n_groups, uint8_t *qual, uint32_t *allele,
                             uint16_t *present) {
    if (!ctx) return FB_ERR_ARG;
    if (!df) FB_FAIL(FB_ERR_ARG, "null argument");
    FB_CK(cudaSetDevice(ctx->device));
    if (n_groups) *n_groups = df->n_groups;
    if (qual) FB_CK(cudaMemcpy(qual, df->d_qual, df->n_groups * 16, cudaMemcpyDeviceToHost));
    if (allele) FB_CK(cudaMemcpy(allele, df->d_allele, df->n_groups * 4, cudaMemcpyDeviceToHost));
    if (present) FB_CK(cudaMemcpy(present, df->d_present, df->n_groups * 2, cudaMemcpyDeviceToHost));
    return FB_OK;
}
}
