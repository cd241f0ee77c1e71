// fb_engine.cuh — host-side engine: context, HBM-resident fragments, the batched plan over (block, ploidy)
// instances and the launch sequences that replace optimize_clustering (local_clustering.rs:71-130) and
// get_mec_stats_epsilon_no_phred (local_clustering.rs:187-215).
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/floria_b200.h"
#include "fb_cache.cuh"
#include "fb_common.cuh"
#include "fb_kernels.cuh"

struct fb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // copy / pack stream of the pipelined upload (created on first use)
    std::string err;
    uint32_t *d_lut = nullptr;  // 256 x u32 (units of 2^-26)
    uint32_t h_lut[256];
    float h_lut_f[256];
    bool lut_valid = false;
    fb_timings tim;
    int *d_n_active = nullptr;
    int *h_n_active = nullptr;  // pinned
    int sm_count = 148;
    size_t mem_free_at_init = 0;  // cudaMemGetInfo at fb_init (free bytes of the device when the context was opened)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    FbCache cache;           // device memory of this context (free list is per context: see fb_cache.cuh)
    bool hist_attr = false;  // cudaFuncSetAttribute(k_hist) done on this context's device
};

struct fb_dfrags {
    fb_ctx *ctx = nullptr;
    uint64_t n_reads = 0, nnz = 0, n_groups = 0, bytes = 0;
    uint32_t *d_first = nullptr, *d_last = nullptr, *d_nnz = nullptr, *d_gstart = nullptr, *d_gptr = nullptr,
             *d_gnum = nullptr;
    uint4 *d_qual = nullptr;
    uint32_t *d_allele = nullptr;
    uint16_t *d_present = nullptr;
    std::vector<uint32_t> h_first, h_last, h_nnz, h_gstart, h_gptr, h_gnum, h_prefmax_last;
    // pipelined upload (fb_frags_upload_impl with pipelined = 1): the planes are filled chunk by chunk on ctx->stream2 while
    // the caller already computes on the leading reads; *d_ready = number of leading reads whose planes are complete
    bool pipelined = false;
    unsigned int *d_ready = nullptr;
    unsigned long long *d_pack_err = nullptr;
    cudaEvent_t ev_first = nullptr, ev_done = nullptr, ev_begin = nullptr;
    std::vector<void *> pipe_temps;  // device temporaries of the upload, released by fb_frags_finish
    DFragsDev dev() const {
        DFragsDev d;
        d.n_reads = n_reads;
        d.first = d_first;
        d.last = d_last;
        d.nnz = d_nnz;
        d.gstart = d_gstart;
        d.gptr = d_gptr;
        d.gnum = d_gnum;
        d.qual = d_qual;
        d.allele = d_allele;
        d.present = d_present;
        return d;
    }
};

static thread_local std::string g_init_err;

#define FB_CK(call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char b_[512];                                                                             \
            snprintf(b_, sizeof(b_), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = b_;                                                                            \
            return FB_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

#define FB_FAIL(code, ...)                        \
    do {                                          \
        char b_[512];                             \
        snprintf(b_, sizeof(b_), __VA_ARGS__);    \
        ctx->err = b_;                            \
        return code;                              \
    } while (0)

template <class T>
static int fb_dalloc(fb_ctx *ctx, T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    FB_CK(ctx->cache.alloc((void **)p, n * sizeof(T)));
    return FB_OK;
}
template <class T>
static int fb_upload(fb_ctx *ctx, T **p, const T *h, size_t n) {
    int rc = fb_dalloc(ctx, p, n);
    if (rc) return rc;
    if (n) FB_CK(cudaMemcpyAsync(*p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return FB_OK;
}
template <class T>
static int fb_upload(fb_ctx *ctx, T **p, const std::vector<T> &h) {
    return fb_upload(ctx, p, h.data(), h.size());
}

static cudaEvent_t fb_event(fb_ctx *ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    cudaEvent_t e = ctx->ev_pool[ctx->ev_used++];
    cudaEventRecord(e, ctx->stream);
    return e;
}

// local_clustering.rs:12-59 find_reads_in_interval on the sorted contig (same result as the reference's linear scan:
// reads are sorted by first_position, so the scan's `break` is an upper bound found by bisection, and the prefix maximum
// of last_position bounds where `last >= start` can first hold).
static void fb_find_reads(const fb_dfrags *df, uint32_t start, uint32_t end, std::vector<uint32_t> &out) {
    out.clear();
    const std::vector<uint32_t> &first = df->h_first, &last = df->h_last, &pm = df->h_prefmax_last;
    size_t n = first.size();
    size_t hi = std::upper_bound(first.begin(), first.end(), end) - first.begin();
    size_t lo = std::lower_bound(pm.begin(), pm.begin() + hi, start) - pm.begin();
    (void)n;
    for (size_t i = lo; i < hi; ++i) {
        if (last[i] < start) continue;
        if (last[i] - first[i] > 10000) continue;
        out.push_back((uint32_t)i);
    }
}

// ---- the batched plan -------------------------------------------------------------------------------------------------
struct BlockPlan {
    std::vector<uint32_t> reads;  // counter_ids ascending (may be left empty when the caller keeps the list: n_reads counts)
    uint32_t n_reads = 0;
    uint32_t ag0 = 0, ng = 0;
    uint64_t nnz = 0;
    uint64_t groups = 0;  // sum over the reads of their 16-SNP groups (the beam kernels are chosen by the mean per read)
    uint32_t read_off = 0;
};

// a block planned once (read selection, extents, per-read descriptors) and appended to the engine of every ploidy wave
struct PlannedBlock {
    BlockPlan plan;
    std::vector<RInfo> ri;
    std::vector<RExtra> rx;
};
static void fb_plan_block(const fb_dfrags *df, std::vector<uint32_t> &&reads, PlannedBlock &pb) {
    BlockPlan &b = pb.plan;
    b.reads = std::move(reads);
    b.n_reads = (uint32_t)b.reads.size();
    b.nnz = 0;
    b.groups = 0;
    uint32_t gmin = 0xFFFFFFFFu, gmax = 0;
    for (uint32_t r : b.reads) {
        uint32_t g0 = df->h_gstart[r];
        uint32_t g1 = g0 + df->h_gnum[r];
        gmin = std::min(gmin, g0);
        gmax = std::max(gmax, g1);
        b.nnz += df->h_nnz[r];
        b.groups += df->h_gnum[r];
    }
    if (b.reads.empty()) {
        gmin = 0;
        gmax = 0;
    }
    b.ag0 = gmin;
    b.ng = gmax - gmin;
    b.read_off = 0;
    pb.ri.resize(b.reads.size());
    pb.rx.resize(b.reads.size());
    for (size_t k = 0; k < b.reads.size(); ++k) {
        const uint32_t r = b.reads[k];
        RInfo &ri = pb.ri[k];
        ri.rid = r;
        ri.lg0 = df->h_gstart[r] - gmin;
        ri.lg1 = ri.lg0 + df->h_gnum[r];
        ri.gbase = df->h_gptr[r] - ri.lg0;
        pb.rx[k].first0 = (df->h_first[r] - 1u) - gmin * 16u;
        pb.rx[k].nnz = df->h_nnz[r];
    }
}

// per-read descriptors of a whole contig's blocks, built and uploaded once and shared by the engines of all ploidy waves
struct SharedReads {
    std::vector<RInfo> rinfo;
    std::vector<RExtra> rextra;
    RInfo *d_rinfo = nullptr;
    RExtra *d_rextra = nullptr;
    ~SharedReads() {
        fb_cache_free(d_rinfo);
        fb_cache_free(d_rextra);
    }
};

struct Engine {
    fb_ctx *ctx = nullptr;
    const SharedReads *shared = nullptr;  // when set: blocks reference its descriptors (add_shared) and nothing is copied
    const std::vector<RInfo> &host_rinfo() const { return shared ? shared->rinfo : rinfo; }
    const fb_dfrags *df = nullptr;
    std::vector<BlockPlan> blocks;
    std::vector<InstDev> inst;
    std::vector<InstState> st;
    std::vector<uint64_t> assign_prefix, tile_prefix, hap_prefix, moves_off, done_off;
    std::vector<uint64_t> mec_chunk_prefix;  // per (instance, haplotype): 32-position chunks of its count table
    uint64_t *d_mec_chunk_prefix = nullptr;
    MecChunk *d_mec_chunks = nullptr;
    std::vector<uint32_t> hist_splits;
    uint64_t tot_done = 0;
    uint32_t *d_hist_splits = nullptr, *d_done = nullptr;
    uint64_t *d_done_off = nullptr;
    bool any_split = false;
    std::vector<uint32_t> moves_cap;
    std::vector<RInfo> rinfo;
    std::vector<RExtra> rextra;
    uint64_t tot_assign = 0, tot_gain = 0, tot_cnt = 0, tot_mask = 0, tot_mec = 0, tot_moves = 0;
    // device
    InstDev *d_inst = nullptr;
    InstState *d_st = nullptr;
    uint64_t *d_assign_prefix = nullptr, *d_tile_prefix = nullptr, *d_hap_prefix = nullptr, *d_moves_off = nullptr;
    uint32_t *d_moves_cap = nullptr;
    RInfo *d_rinfo = nullptr;
    RExtra *d_rextra = nullptr;
    uint8_t *d_assign[2] = {nullptr, nullptr};
    uint64_t *d_cnt[2] = {nullptr, nullptr};
    uint2 *d_masks[2] = {nullptr, nullptr};
    int sweep_team = 0;  // lanes per read of k_sweep (0 = not chosen yet)
    double *d_mec[2] = {nullptr, nullptr};
    double *d_gain = nullptr;
    MoveRec *d_moves = nullptr;
    double eps = 0;
    int eps_safe = 0;

    ~Engine() { release(); }
    void release() {
        fb_cache_free(d_inst);
        fb_cache_free(d_st);
        fb_cache_free(d_assign_prefix);
        fb_cache_free(d_tile_prefix);
        fb_cache_free(d_hap_prefix);
        fb_cache_free(d_mec_chunk_prefix);
        fb_cache_free(d_mec_chunks);
        d_mec_chunk_prefix = nullptr;
        d_mec_chunks = nullptr;
        fb_cache_free(d_moves_off);
        fb_cache_free(d_moves_cap);
        if (!shared) fb_cache_free(d_rinfo);
        fb_cache_free(d_hist_splits);
        fb_cache_free(d_done_off);
        fb_cache_free(d_done);
        if (!shared) fb_cache_free(d_rextra);
        for (int b = 0; b < 2; ++b) {
            fb_cache_free(d_assign[b]);
            fb_cache_free(d_cnt[b]);
            fb_cache_free(d_masks[b]);
            fb_cache_free(d_mec[b]);
        }
        fb_cache_free(d_gain);
        fb_cache_free(d_moves);
        d_inst = nullptr;
        d_st = nullptr;
        d_assign_prefix = d_tile_prefix = d_hap_prefix = d_moves_off = nullptr;
        d_moves_cap = nullptr;
        d_rinfo = nullptr;
        d_hist_splits = nullptr;
        d_done_off = nullptr;
        d_done = nullptr;
        d_rextra = nullptr;
        for (int b = 0; b < 2; ++b) {
            d_assign[b] = nullptr;
            d_cnt[b] = nullptr;
            d_masks[b] = nullptr;
            d_mec[b] = nullptr;
        }
        d_gain = nullptr;
        d_moves = nullptr;
    }

    // add a block given its (ascending) read list; returns block index
    int add_block(const std::vector<uint32_t> &reads) {
        PlannedBlock pb;
        fb_plan_block(df, std::vector<uint32_t>(reads), pb);
        return add_planned(pb);
    }
    // a block whose descriptors sit at `read_off` of the shared arrays (the read list stays with the caller)
    int add_shared(const BlockPlan &plan, uint32_t read_off) {
        BlockPlan b;
        b.n_reads = plan.n_reads;
        b.ag0 = plan.ag0;
        b.ng = plan.ng;
        b.nnz = plan.nnz;
        b.groups = plan.groups;
        b.read_off = read_off;
        blocks.push_back(std::move(b));
        return (int)blocks.size() - 1;
    }
    int add_planned(const PlannedBlock &pb) {
        BlockPlan b = pb.plan;
        b.read_off = (uint32_t)rinfo.size();
        rinfo.insert(rinfo.end(), pb.ri.begin(), pb.ri.end());
        rextra.insert(rextra.end(), pb.rx.begin(), pb.rx.end());
        blocks.push_back(std::move(b));
        return (int)blocks.size() - 1;
    }
    int add_instance(int block, uint32_t ploidy) {
        const BlockPlan &b = blocks[block];
        InstDev in;
        memset(&in, 0, sizeof(in));
        in.block = (uint32_t)block;
        in.ploidy = ploidy;
        in.n_reads = b.n_reads;
        in.ng = b.ng;
        in.ag0 = b.ag0;
        in.read_off = b.read_off;
        in.mec_off = (uint32_t)tot_mec;
        in.flt_lo = 0;
        in.flt_hi = 0xFFFFFFFFu;
        in.assign_off = tot_assign;
        in.gain_off = tot_gain;
        in.cnt_off = tot_cnt;
        in.mask_off = tot_mask;
        assign_prefix.push_back(tot_assign);
        hap_prefix.push_back(tot_mec);
        uint32_t tiles = (b.ng + FB_HIST_TILE_GROUPS - 1) / FB_HIST_TILE_GROUPS;
        tot_assign += in.n_reads;
        tot_gain += (uint64_t)in.n_reads * ploidy;
        tot_cnt += (uint64_t)ploidy * b.ng * 64;
        tot_mask += (uint64_t)ploidy * b.ng;
        tot_mec += ploidy;
        // read splits per (tile, haplotype): enough CTAs to saturate HBM on big blocks, 1 (no atomics) on small ones
        uint32_t splits = std::min<uint32_t>(8, std::max<uint32_t>(1, in.n_reads / 8192));
        if (splits > 1) any_split = true;
        hist_splits.push_back(splits);
        done_off.push_back(tot_done);
        tot_done += (uint64_t)tiles * ploidy;
        tile_counts.push_back((uint64_t)tiles * ploidy * splits);
        // k_select packs a move as read | j << 24 | i << 28 and sorts over a power-of-two capacity held in 32 bits:
        // blocks beyond 2^24 reads are refused by the callers (fb_limits_ok), so neither can overflow here
        uint64_t maxm = (uint64_t)in.n_reads * (ploidy > 1 ? ploidy - 1 : 1);
        uint32_t cap = 1;
        while (cap < maxm && cap < (1u << 31)) cap <<= 1;
        moves_off.push_back(tot_moves);
        moves_cap.push_back(cap);
        tot_moves += cap;
        InstState s;
        memset(&s, 0, sizeof(s));
        s.cur = 0;
        s.active = 1;
        st.push_back(s);
        inst.push_back(in);
        return (int)inst.size() - 1;
    }
    std::vector<uint64_t> tile_counts;

    int finalize_and_upload(double eps_) {
        {
            std::string why;
            if (!limits_ok(why)) FB_FAIL(FB_ERR_LIMIT, "%s", why.c_str());
        }
        eps = eps_;
        eps_safe = fb_eps_is_safe(eps_);
        int n = (int)inst.size();
        assign_prefix.push_back(tot_assign);
        hap_prefix.push_back(tot_mec);
        tile_prefix.assign(n + 1, 0);
        for (int i = 0; i < n; ++i) tile_prefix[i + 1] = tile_prefix[i] + tile_counts[i];
        int rc;
        if ((rc = fb_upload(ctx, &d_inst, inst))) return rc;
        if ((rc = fb_upload(ctx, &d_st, st))) return rc;
        if ((rc = fb_upload(ctx, &d_assign_prefix, assign_prefix))) return rc;
        if ((rc = fb_upload(ctx, &d_tile_prefix, tile_prefix))) return rc;
        if ((rc = fb_upload(ctx, &d_hap_prefix, hap_prefix))) return rc;
        mec_chunk_prefix.assign(1, 0);
        for (const InstDev &in : inst)
            for (uint32_t h = 0; h < in.ploidy; ++h) mec_chunk_prefix.push_back(mec_chunk_prefix.back() + (in.ng + 1) / 2);
        if ((rc = fb_upload(ctx, &d_mec_chunk_prefix, mec_chunk_prefix))) return rc;
        if ((rc = fb_dalloc(ctx, &d_mec_chunks, (size_t)std::max<uint64_t>(mec_chunk_prefix.back(), 1)))) return rc;
        if ((rc = fb_upload(ctx, &d_moves_off, moves_off))) return rc;
        if ((rc = fb_upload(ctx, &d_moves_cap, moves_cap))) return rc;
        if (shared)
            d_rinfo = shared->d_rinfo;
        else if ((rc = fb_upload(ctx, &d_rinfo, rinfo)))
            return rc;
        if ((rc = fb_upload(ctx, &d_hist_splits, hist_splits))) return rc;
        if ((rc = fb_upload(ctx, &d_done_off, done_off))) return rc;
        if ((rc = fb_dalloc(ctx, &d_done, tot_done))) return rc;
        if (shared)
            d_rextra = shared->d_rextra;
        else if ((rc = fb_upload(ctx, &d_rextra, rextra)))
            return rc;
        for (int b = 0; b < 2; ++b) {
            if ((rc = fb_dalloc(ctx, &d_assign[b], tot_assign))) return rc;
            if ((rc = fb_dalloc(ctx, &d_cnt[b], tot_cnt))) return rc;
            if ((rc = fb_dalloc(ctx, &d_masks[b], tot_mask))) return rc;
            if ((rc = fb_dalloc(ctx, &d_mec[b], tot_mec * 2))) return rc;
            FB_CK(cudaMemsetAsync(d_assign[b], 0, std::max<uint64_t>(tot_assign, 1), ctx->stream));
        }
        if ((rc = fb_dalloc(ctx, &d_gain, tot_gain))) return rc;
        if ((rc = fb_dalloc(ctx, &d_moves, tot_moves))) return rc;
        return FB_OK;
    }

    int n_inst() const { return (int)inst.size(); }
    // capacity limits of the instance kernels: k_select's 24-bit read index, 4-bit haplotype fields
    bool limits_ok(std::string &why) const {
        for (const InstDev &in : inst) {
            if (in.n_reads >= (1u << 24)) {
                why = "a block holds " + std::to_string(in.n_reads) + " reads; the move selection supports fewer than 2^24 per block";
                return false;
            }
            if (in.ploidy > 15) {
                why = "ploidy above 15";
                return false;
            }
        }
        return true;
    }

    int launch_sizes(int which) {
        int n = n_inst();
        if (!n) return FB_OK;
        k_sizes<<<(n + 7) / 8, 256, 0, ctx->stream>>>(d_inst, d_st, n, d_assign[0], d_assign[1], which);
        ctx->tim.n_launches++;
        FB_CK(cudaGetLastError());
        return FB_OK;
    }
    int launch_hist(int which, int use_phred, int only_active, int buf = 0, int assign_cur = 0) {
        int n = n_inst();
        uint64_t ctas = tile_prefix[n];
        if (!ctas) return FB_OK;
        HistArgs a;
        a.fr = df->dev();
        a.inst = d_inst;
        a.st = d_st;
        a.n_inst = n;
        a.tile_prefix = d_tile_prefix;
        a.splits = d_hist_splits;
        a.done_off = d_done_off;
        a.done = d_done;
        a.rinfo = d_rinfo;
        a.assign[0] = d_assign[0];
        a.assign[1] = d_assign[1];
        a.cnt[0] = d_cnt[0];
        a.cnt[1] = d_cnt[1];
        a.masks[0] = d_masks[0];
        a.masks[1] = d_masks[1];
        a.lut = ctx->d_lut;
        a.use_phred = use_phred;
        a.which = which;
        a.buf = buf;
        a.only_active = only_active;
        a.assign_cur = assign_cur;
        cudaEvent_t e0 = fb_event(ctx);
        if (any_split) {
            cudaMemsetAsync(d_done, 0, std::max<uint64_t>(tot_done, 1) * 4, ctx->stream);
            k_hist_zero<<<dim3(64, (unsigned)n), 256, 0, ctx->stream>>>(a);
            ctx->tim.n_launches++;
        }
        if (!ctx->hist_attr) {  // function attributes are per device
            FB_CK(cudaFuncSetAttribute(k_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_HIST_SMEM));
            ctx->hist_attr = true;
        }
        k_hist<<<(unsigned)ctas, FB_HIST_THREADS, FB_HIST_SMEM, ctx->stream>>>(a);
        cudaEvent_t e1 = fb_event(ctx);
        hist_ev.push_back(std::make_pair(e0, e1));
        ctx->tim.n_launches++;
        ctx->tim.n_hist_launches++;
        FB_CK(cudaGetLastError());
        return FB_OK;
    }
    int launch_mec(int which, int only_active, int buf = 0) {
        int n = n_inst();
        uint64_t warps = hap_prefix[n];
        if (!warps) return FB_OK;
        MecArgs a;
        a.inst = d_inst;
        a.st = d_st;
        a.n_inst = n;
        a.hap_prefix = d_hap_prefix;
        a.chunk_prefix = d_mec_chunk_prefix;
        a.chunks = d_mec_chunks;
        a.cnt[0] = d_cnt[0];
        a.cnt[1] = d_cnt[1];
        a.mec[0] = d_mec[0];
        a.mec[1] = d_mec[1];
        a.eps = eps;
        a.eps_safe = eps_safe;
        a.which = which;
        a.buf = buf;
        a.only_active = only_active;
        cudaEvent_t e0 = fb_event(ctx);
        const uint64_t n_chunks = mec_chunk_prefix.back();
        if (n_chunks > 32 * warps) {  // long tables: chunk summaries first, then an ordered fold of the summaries
            k_mec_chunks<<<(unsigned)((n_chunks + 7) / 8), 256, 0, ctx->stream>>>(a, warps);
            k_mec<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(a);
            ctx->tim.n_launches++;
        } else {
            k_mec_flat<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(a);
        }
        cudaEvent_t e1 = fb_event(ctx);
        mec_ev.push_back(std::make_pair(e0, e1));
        ctx->tim.n_launches++;
        FB_CK(cudaGetLastError());
        return FB_OK;
    }
    SweepArgs sweep_args(int mode) {
        SweepArgs a;
        memset(&a, 0, sizeof(a));
        a.fr = df->dev();
        a.inst = d_inst;
        a.st = d_st;
        a.n_inst = n_inst();
        a.assign_prefix = d_assign_prefix;
        a.rinfo = d_rinfo;
        a.assign[0] = d_assign[0];
        a.assign[1] = d_assign[1];
        a.masks[0] = d_masks[0];
        a.masks[1] = d_masks[1];
        a.lut = ctx->d_lut;
        a.eps = eps;
        a.eps_safe = eps_safe;
        a.mode = mode;
        a.gain = d_gain;
        return a;
    }
    int launch_sweep(const SweepArgs &a) {
        if (!tot_assign) return FB_OK;
        cudaEvent_t e0 = fb_event(ctx);
        uint32_t pmax = 1;
        for (const InstDev &in : inst) pmax = std::max(pmax, in.ploidy);
        // lanes per read: a whole warp for long reads, teams of 8 / 2 lanes when the reads span few 16-SNP groups
        if (sweep_team == 0) {
            uint64_t g = 0;
            const std::vector<RInfo> &hr = host_rinfo();
            for (const RInfo &r : hr) g += r.lg1 - r.lg0;
            const double avg = hr.empty() ? 32.0 : (double)g / (double)hr.size();
            sweep_team = avg > 12.0 ? 32 : (avg > 2.5 ? 8 : 2);
            if (getenv("FB_SWEEP_TEAM")) sweep_team = atoi(getenv("FB_SWEEP_TEAM"));
        }
        // FB_SWEEP_TMA=1 selects the cp.async.bulk (1-D TMA) staged variant.  Measured on configs[2] (profiles/): the sweep
        // is instruction-issue bound, so the staging does not pay (p=4: 5.0 vs 4.7 ms, p=2: 3.0 vs 2.5 ms); default off.
        static const bool use_tma = getenv("FB_SWEEP_TMA") && atoi(getenv("FB_SWEEP_TMA")) != 0;
        const unsigned blk = FB_SWEEP_WARPS * 32;
        const unsigned teams = FB_SWEEP_WARPS * (32 / (sweep_team == 8 ? 8 : (sweep_team == 2 ? 2 : 32)));
        const unsigned grid = (unsigned)((tot_assign + teams - 1) / teams);
#define FB_LAUNCH_SWEEP(TMA_, L_)                                            \
    do {                                                                     \
        if (pmax <= 2)                                                       \
            k_sweep<2, TMA_, L_><<<grid, blk, 0, ctx->stream>>>(a);          \
        else if (pmax <= 4)                                                  \
            k_sweep<4, TMA_, L_><<<grid, blk, 0, ctx->stream>>>(a);          \
        else                                                                 \
            k_sweep<8, TMA_, L_><<<grid, blk, 0, ctx->stream>>>(a);          \
    } while (0)
        if (sweep_team == 8)
            FB_LAUNCH_SWEEP(false, 8);
        else if (sweep_team == 2)
            FB_LAUNCH_SWEEP(false, 2);
        else if (use_tma)
            FB_LAUNCH_SWEEP(true, 32);
        else
            FB_LAUNCH_SWEEP(false, 32);
#undef FB_LAUNCH_SWEEP
        cudaEvent_t e1 = fb_event(ctx);
        sweep_ev.push_back(std::make_pair(e0, e1));
        ctx->tim.n_launches++;
        ctx->tim.n_sweep_launches++;
        FB_CK(cudaGetLastError());
        return FB_OK;
    }
    int launch_select() {
        int n = n_inst();
        if (!n) return FB_OK;
        SelectArgs a;
        a.inst = d_inst;
        a.st = d_st;
        a.n_inst = n;
        a.gain = d_gain;
        a.assign[0] = d_assign[0];
        a.assign[1] = d_assign[1];
        a.moves = d_moves;
        a.moves_off = d_moves_off;
        a.moves_cap = d_moves_cap;
        cudaEvent_t e0 = fb_event(ctx);
        k_select<<<n, FB_SELECT_THREADS, 0, ctx->stream>>>(a);
        cudaEvent_t e1 = fb_event(ctx);
        select_ev.push_back(std::make_pair(e0, e1));
        ctx->tim.n_launches++;
        FB_CK(cudaGetLastError());
        return FB_OK;
    }
    int launch_accept(int init, uint32_t iter, uint32_t max_iters) {
        int n = n_inst();
        if (!n) return FB_OK;
        AcceptArgs a;
        a.inst = d_inst;
        a.st = d_st;
        a.n_inst = n;
        a.mec[0] = d_mec[0];
        a.mec[1] = d_mec[1];
        a.init = init;
        a.iter = iter;
        a.max_iters = max_iters;
        a.n_active = ctx->d_n_active;
        k_accept<<<(n + 127) / 128, 128, 0, ctx->stream>>>(a);
        ctx->tim.n_launches++;
        FB_CK(cudaGetLastError());
        return FB_OK;
    }

    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> hist_ev, mec_ev, sweep_ev, select_ev;
    uint64_t active_assign = 0;  // for cell accounting of launches (all instances are streamed by every launch)

    // optimize_clustering (local_clustering.rs:71-130) for every instance, starting from assign[0].
    int run_optimize(uint32_t max_iters) {
        int rc;
        if ((rc = launch_sizes(0))) return rc;
        if ((rc = launch_hist(0, 1, 0))) return rc;
        if ((rc = launch_mec(0, 0))) return rc;
        if ((rc = launch_accept(1, 0, max_iters))) return rc;
        // The host learns one round LATE whether any instance is still iterating: round `it` is queued before the count of
        // round it - 1 is read, so the device never idles between rounds (local_clustering.rs:85-129 runs up to
        // NUM_ITER_OPTIMIZE rounds per instance; an instance that has stopped is skipped by every kernel, so the one
        // surplus round after the last instance stops changes nothing).
        cudaEvent_t ev_prev = nullptr;
        for (uint32_t it = 0; it < max_iters; ++it) {
            if ((rc = launch_sweep(sweep_args(FB_SWEEP_MOVES)))) return rc;
            if ((rc = launch_select())) return rc;
            if ((rc = launch_hist(1, 1, 1))) return rc;
            if ((rc = launch_mec(1, 1))) return rc;
            FB_CK(cudaMemsetAsync(ctx->d_n_active, 0, sizeof(int), ctx->stream));
            if ((rc = launch_accept(0, it, max_iters))) return rc;
            FB_CK(cudaMemcpyAsync(ctx->h_n_active + (it & 1u), ctx->d_n_active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            cudaEvent_t ev = fb_event(ctx);
            if (ev_prev) {
                FB_CK(cudaEventSynchronize(ev_prev));
                if (ctx->h_n_active[(it - 1) & 1u] == 0) break;
            }
            ev_prev = ev;
        }
        return FB_OK;
    }

    void collect_timings() {
        auto sum = [](std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &v) {
            float t = 0;
            for (auto &p : v) {
                float ms = 0;
                cudaEventElapsedTime(&ms, p.first, p.second);
                t += ms;
            }
            return t;
        };
        ctx->tim.sweep_ms += sum(sweep_ev);
        ctx->tim.hist_ms += sum(hist_ev);
        ctx->tim.mec_ms += sum(mec_ev);
        ctx->tim.select_ms += sum(select_ev);
    }
};
