// fb_beam_wide.cuh — beam_search_phasing (global_clustering.rs:10-179) for ONE large instance spread over the whole GPU.
//
// k_beam (fb_beam.cuh) gives an instance one CTA: right for thousands of small blocks, hopeless for a block whose reads
// span tens of thousands of SNPs (the 100k-read x 50k-SNP block of BASELINE.json configs[2]: 3125 groups per read, 1.6 MB
// of counts per haplotype state).  Here every CTA of a cooperative grid owns an interleaved slice of the SNP axis
// (chunks of FB_BW_CH groups, chunk c belongs to CTA c mod gridDim.x) of EVERY haplotype state, and a step (one read) is
//   phase A  every CTA scores its slice of the read against every live state, sums its slice of delta(read) and of
//            the hash terms that leave the window; the partial sums (exact integers: order-free) are reduced across
//            the grid with RED atomics into a per-step slot of global memory
//   -------- ONE grid barrier per step --------
//   phase B  every CTA forms the p-values of all live states and runs the warp-0 decision section
//            (fb_beam_decide.inc: pruning, child scores, equality classes, BinaryHeap, next generation) REDUNDANTLY on
//            identical inputs, so all CTAs hold the same node tables / job list without a broadcast
//   phase C  every CTA materialises its slice of the surviving generation's new states (copy or in place) and the is-max
//            planes; the next step's phase A reads only the CTA's own slices, so no barrier is needed here
// Optional (FB_BW_EARLY, off: measured slower): while warp 0 forms the p-values and runs the decision section of step t,
// the other warps score read t + 1 against every state that is live NOW (early scoring, into the slot of step t + 1).  A
// state that survives unchanged is then done; the state that step t updates in place gets the difference between its
// updated and its old planes added by the materialisation itself (phase C holds both in registers); only states created
// by a copy in step t are scored by the regular phase A at the top of step t + 1.
// Full-state reads (the ordered epsilon replay of a non-dyadic epsilon, the word-by-word equality check behind a hash
// match) touch slices that other CTAs update in place in phase C of the same step: steps that execute one add a second
// grid barrier before phase C.  Both are rare (first reads of a haplotype; states that became equal after the window
// moved).  Results are bit-identical to k_beam and to the oracle (tests/test_gpu_beam_wide.py).
#pragma once
#include "fb_beam.cuh"

#define FB_BW_THREADS 256
#define FB_BW_WARPS (FB_BW_THREADS / 32)
#define FB_BW_CH 2  // groups per ownership chunk: 32 positions = 1 KB of counts = one warp pass of the materialisation
#define FB_BW_SLOTS 4  // step slots of the grid reduction: read t uses slot t mod 4 (written in steps t - 1 and t, read in step t,
                       // zeroed in step t - 2 after that step's grid barrier)
#ifndef FB_BW_EARLY
// 1: score the next read one step ahead, while warp 0 runs the p-values and the decision section, and let the
// materialisation add the difference for the state it updates in place.  Bit-identical results (tests/test_gpu_beam_wide.py
// passes either way), but measured SLOWER on the 100k x 50k block (964 ms against 917 ms per search): the early pass and the
// deltas cost the other warps more than the regular scoring pass they remove.  Kept for A/B runs (profiles/README.md).
#define FB_BW_EARLY 0
#endif

struct __align__(8) BeamWideAcc {  // one per (step slot, state)
    unsigned long long same, emptyw, sub;
    unsigned int ne_cnt;
    // Last diff / first empty position of the read on the state.  A read is scored one step EARLY, against the planes as
    // they are before that step's in-place update; sums are corrected by the update (signed deltas), extremes cannot be,
    // so they are kept apart: groups outside the updating read's range (`last_diff`, `first_empty`: never touched by the
    // update; also everything the regular scoring adds), groups inside it under the old planes (`_in`) and under the
    // updated planes (`_new`, written by the materialisation).  The p-value step takes outside + new for a state that
    // was updated in place, outside + in otherwise.
    int last_diff, first_empty;
    int last_diff_in, first_empty_in;
    int last_diff_new, first_empty_new;
    unsigned int _pad;
};
struct __align__(8) BeamWideStep {  // one per step slot
    unsigned long long total, delta;
};

int fb_beam_wide_max_grid(size_t smem_bytes, int sm_count, int *grid);
int fb_beam_wide_launch(unsigned grid, size_t smem_bytes, cudaStream_t stream, const struct BeamParams &bp);

__device__ __forceinline__ unsigned long long fb_ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Every spin of this kernel is bounded: a wait that outlasts FB_BW_SPIN_LIMIT polls (tens of seconds; a step takes
// microseconds, an upload chunk milliseconds) can only be a lost producer, and the kernel traps (the launch fails with an
// error) instead of hanging the device.
#define FB_BW_SPIN_LIMIT (1ull << 25)
// all threads of all CTAs; `target` = arrivals expected so far (monotonic counter, never reset during a launch)
__device__ __forceinline__ void fb_grid_barrier(unsigned long long *ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned long long spins = 0;
        while (fb_ld_acquire_u64(ctr) < target)
            if (++spins > FB_BW_SPIN_LIMIT) __trap();
    }
    __syncthreads();
}
// reads [0, n) of the contig are packed in HBM once *ready >= n (pipelined upload, fb_lib.cu): thread 0 of the CTA polls, the
// CTA-wide barrier that follows every call site publishes the result to the other threads
__device__ __forceinline__ void fb_wait_reads(const unsigned int *ready, unsigned int need, unsigned int &seen) {
    if (ready == nullptr || seen >= need) return;
    unsigned int v;
    unsigned long long spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ready) : "memory");
        if (v >= need) break;
        if (++spins > FB_BW_SPIN_LIMIT) __trap();
    }
    seen = v;
}

// stable_binom_cdf_p_rev (utils_frags.rs:211-248) with its two log terms on the two lanes of a pair (sub = 0 / 1); the
// operations and their order are those of fb_stable_binom_cdf_p_rev.  Every lane of the warp must call it.
__device__ __forceinline__ double fb_pvalue_pair(double same_f, double diff_f, uint32_t sub, double eps, double div_factor,
                                                 bool div_pow2, double inv_div) {
    const unsigned long long nn = fb_as_usize(same_f + diff_f), kk = fb_as_usize(diff_f);
    const double n64 = (double)nn, k64 = (double)kk;
    double a = nn ? k64 / n64 : 0.5;
    if (a == 1.0) a = 0.9999999;
    if (a == 0.0) a = 0.0000001;
    const double x = sub == 0 ? a : (1.0 - a);
    const double y = sub == 0 ? eps : (1.0 - eps);
    const double t = x * log(x / y);
    const double t1 = __shfl_xor_sync(0xFFFFFFFFu, t, 1);
    double rel_ent = sub == 0 ? t + t1 : t1 + t;  // lane `sub == 0` holds a*ln(a/p), the sum is a*ln(..) + (1-a)*ln(..)
    if (a < eps) rel_ent = -rel_ent;
    double pvs = 0.0;
    if (nn != 0) pvs = (div_pow2 ? -1.0 * n64 * inv_div : -1.0 * n64 / div_factor) * rel_ent;
    return 1.0 * pvs;
}

template <int P>
__device__ __noinline__ void fb_beam_wide_instance(const BeamParams &bp, const int ii, uint8_t *smem, uint8_t *slot,
                                                   unsigned long long &bar_target) {
    constexpr int NT = FB_BW_THREADS;
    constexpr int NW = FB_BW_WARPS;
    constexpr uint32_t CH = FB_BW_CH;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    const bool writer = cta == 0;
    const BeamSmem &L = bp.L;
    double *nd_score = reinterpret_cast<double *>(smem + L.off_nd_score);     // [2][W]
    double *nd_err = reinterpret_cast<double *>(smem + L.off_nd_err);         // [2][W][P]
    uint16_t *nd_ref = reinterpret_cast<uint16_t *>(smem + L.off_nd_ref);     // [2][W][P]
    unsigned long long *st_hash = reinterpret_cast<unsigned long long *>(smem + L.off_st_hash);
    double *sc_same = reinterpret_cast<double *>(smem + L.off_sc_same);
    double *sc_diff = reinterpret_cast<double *>(smem + L.off_sc_diff);
    double *sc_pv = reinterpret_cast<double *>(smem + L.off_sc_pv);
    int *st_hi = reinterpret_cast<int *>(smem + L.off_st_hi);
    int *st_mark = reinterpret_cast<int *>(smem + L.off_st_mark);
    int *st_free = reinterpret_cast<int *>(smem + L.off_free);
    int *live = reinterpret_cast<int *>(smem + L.off_live);
    double *ch_score = reinterpret_cast<double *>(smem + L.off_ch_score);
    unsigned long long *ch_fold = reinterpret_cast<unsigned long long *>(smem + L.off_ch_fold);
    int *ch_m = reinterpret_cast<int *>(smem + L.off_ch_m);
    uint16_t *ch_parent = reinterpret_cast<uint16_t *>(smem + L.off_ch_parent);
    uint16_t *ch_part = reinterpret_cast<uint16_t *>(smem + L.off_ch_part);
    uint16_t *ch_class = reinterpret_cast<uint16_t *>(smem + L.off_ch_class);
    double *hp_score = reinterpret_cast<double *>(smem + L.off_hp_score);
    int *hp_item = reinterpret_cast<int *>(smem + L.off_hp_item);
    uint32_t *lut_s = reinterpret_cast<uint32_t *>(smem + L.off_lut);
    uint32_t *wscr = reinterpret_cast<uint32_t *>(smem + L.off_wscr) + warp * 16;
    BeamJob *jobs = reinterpret_cast<BeamJob *>(smem + L.off_job);
    int *addnew = reinterpret_cast<int *>(smem + L.off_addnew);
    int *plain = reinterpret_cast<int *>(smem + L.off_plain);
    int *replay = reinterpret_cast<int *>(smem + L.off_replay);
    uint32_t *a_done = reinterpret_cast<uint32_t *>(smem + L.off_adone);  // read index whose sums are complete for the state
    uint32_t *inpl = reinterpret_cast<uint32_t *>(smem + L.off_inpl);     // step that must read the `_new` extremes
    struct Misc {
        unsigned long long delta[2];  // delta(read) of the current step (index = step parity, as in k_beam)
        int n_nodes[2];
        int n_live, n_free, n_jobs_copy, n_jobs_inplace;
        int hw;  // high-water mark of the state ids handed out so far
        int n_replay, full_reads;  // full_reads: this step read whole states (replay / exact comparison)
    };
    Misc *ms = reinterpret_cast<Misc *>(smem + L.off_misc);
    uint4 *rq = reinterpret_cast<uint4 *>(smem + L.off_rq);          // [2 * FB_BEAM_RG] this CTA's groups of the staged read
    uint32_t *ral = reinterpret_cast<uint32_t *>(smem + L.off_ral);
    uint16_t *rpr = reinterpret_cast<uint16_t *>(smem + L.off_rpr);
    uint32_t *btbuf = reinterpret_cast<uint32_t *>(smem + L.off_rq);  // backtrack staging (the read staging area of k_beam)

    uint32_t *hist = reinterpret_cast<uint32_t *>(slot + bp.hist_off);
    const uint32_t *__restrict__ qual32 = reinterpret_cast<const uint32_t *>(bp.fr.qual);
    const uint8_t *__restrict__ qual8 = reinterpret_cast<const uint8_t *>(bp.fr.qual);
    const InstDev in = bp.inst[ii];
    const uint32_t Wmax = P * bp.B;
    const uint32_t NS = P * bp.B * (P + 1) + 1;
    const uint32_t npos = in.ng * 16;
    const uint64_t state_words = ((uint64_t)npos * 4 + in.ng + 1) & ~1ULL;
    unsigned long long *pool = reinterpret_cast<unsigned long long *>(slot);
#define ST_CNT(s) (pool + (uint64_t)(s) * state_words)
#define ST_MASK(s) (reinterpret_cast<uint2 *>(pool + (uint64_t)(s) * state_words + (uint64_t)npos * 4))
    const uint32_t Wm = bp.maxW;  // smem strides
    const uint32_t Pm = bp.maxP;
#define ND_SCORE(g, n) nd_score[(g) * Wm + (n)]
#define ND_ERR(g, n, h) nd_err[((g) * Wm + (n)) * Pm + (h)]
#define ND_REF(g, n, h) nd_ref[((g) * Wm + (n)) * Pm + (h)]
    const RInfo *__restrict__ rinfo = bp.rinfo + in.read_off;
    const RExtra *__restrict__ rextra = bp.rextra + in.read_off;

    // chunks of CH groups are dealt round-robin to the CTAs: the first chunk >= cA that this CTA owns
    // (divisions by the grid size are on every warp's path several times per step: multiply by a 42-bit reciprocal instead,
    //  exact for numerators below 2^28 and grids of up to 1024 CTAs)
    const unsigned long long g_recip = ((1ULL << 42) + G - 1) / G;
    auto div_g = [&](uint32_t n) { return (uint32_t)(((unsigned long long)n * g_recip) >> 42); };
    auto my_first_chunk = [&](uint32_t cA) {
        uint32_t t = cta + G - (cA - div_g(cA) * G);  // in [1, 2 G)
        if (t >= G) t -= G;
        return cA + t;
    };

    // ---- init: one root node over the empty state (global_clustering.rs:29-47) ---------------------------------------
    fb_grid_barrier(bp.wbar, bar_target += G);  // the previous instance's readers of the step slots are done
    for (uint32_t s = tid; s < NS; s += NT) {
        st_hash[s] = 0;
        st_hi[s] = -1;
        st_mark[s] = 0;
        plain[s] = 0;
        addnew[s] = -1;
        a_done[s] = ~0u;
        inpl[s] = ~0u;
    }
    auto zero_slot = [&](uint32_t sl, int t0, int nt) {  // CTA 0, threads [t0, t0 + nt)
        BeamWideAcc z;
        z.same = z.emptyw = z.sub = 0;
        z.ne_cnt = 0;
        z.last_diff = z.last_diff_in = z.last_diff_new = -1;
        z.first_empty = z.first_empty_in = z.first_empty_new = INT_MAX;
        z._pad = 0;
        for (int s = tid - t0; s >= 0 && s < (int)NS; s += nt) bp.wacc[(uint64_t)sl * bp.maxNS + s] = z;
        if (tid == t0) {
            bp.wstep[sl].total = 0;
            bp.wstep[sl].delta = 0;
        }
    };
    if (writer)
        for (uint32_t sl = 0; sl < FB_BW_SLOTS; ++sl) zero_slot(sl, 0, NT);
    __syncthreads();
    if (tid == 0) {
        ms->n_free = 0;  // explicit free stack (ids below hw); state 0 is the root's empty state
        ms->hw = 1;
        ms->n_live = 1;
        live[0] = 0;
        ms->n_nodes[0] = 1;
        ms->n_replay = 0;
        ms->full_reads = 0;
        ND_SCORE(0, 0) = 0.0;
        for (uint32_t h = 0; h < P; ++h) {
            ND_ERR(0, 0, h) = 0.0;
            ND_REF(0, 0, h) = 0;
        }
    }
    fb_grid_barrier(bp.wbar, bar_target += G);

    // profile counters of thread 0 (FB_BEAM_PROF=1) live in shared memory: 24 64-bit registers would cost the kernel spills
    __shared__ long long pt[24];
    if (tid < 24) pt[tid] = 0;
    __syncthreads();
    long long tc = fb_clock();
#define PROF(i)                                \
    if (bp.prof && writer && tid == 0) {       \
        long long n_ = fb_clock();              \
        pt[i] += n_ - tc;                      \
        tc = n_;                               \
    }
    long long pa = 0, pc = 0, pw = 0;  // CTA 0, warp 1: cycles in phase A, phase C, and waiting for warp 0
    const bool prof1 = bp.prof && writer && tid == 32;
    const double ln_p = log((double)P);
    const bool div_pow2 = (fb_f64_bits(bp.div_factor) & 0xFFFFFFFFFFFFFULL) == 0 && bp.div_factor > 1e-300 && bp.div_factor < 1e300;
    const double inv_div = 1.0 / bp.div_factor;
    int gen = 0;
    uint32_t prev_start = 0;  // block-local position0 from which the hashes are valid
    int gmax = -1;            // last block-local group touched so far
    unsigned long long cells = 0, tapn = 0;
    // The reads' descriptors go through a shared-memory ring filled 32 steps ahead (a descriptor fetched from global
    // memory in the step that needs it costs a DRAM round trip on the critical path).
    __shared__ RInfo rd_i[64];
    __shared__ RExtra rd_x[64];
    if (tid < 64 && (uint32_t)tid < in.n_reads) {
        rd_i[tid] = rinfo[tid];
        rd_x[tid] = rextra[tid];
    }
    __syncthreads();
    RInfo ri_next = rd_i[0];
    RExtra rx_next = rd_x[0];

    // This CTA's groups of a read, staged in shared memory one step ahead (by the warps that idle while warp 0 runs the
    // decision section) together with the read-only sums of the step: total weight and delta(read) = sum of
    // G(pos, allele) * weight (see fb_beam.cuh), both reduced over the grid into the step's slot.
    constexpr uint32_t STG = 2 * FB_BEAM_RG;  // staging capacity in groups
    auto my_groups = [&](const RInfo &r, uint32_t &c0) {
        const uint32_t cA = r.lg0 / CH, cB = (r.lg1 - 1) / CH;
        c0 = my_first_chunk(cA);
        return c0 <= cB ? (div_g(cB - c0) + 1) * CH : 0u;  // group slots of my chunks (the ends may fall outside the read)
    };
    // The planes of read t + 1 are fetched while warp 0 runs the p-values and the decision section of step t (the DRAM
    // latency falls into the time the other warps would wait for the job list anyway), two groups per thread in flight.
#define FB_BW_PREFETCH_ONE(r_, gi_, q_, al_, pr_, c0_, nmg_)                          \
    {                                                                                \
        q_ = make_uint4(~0u, ~0u, ~0u, ~0u);                                         \
        al_ = 0;                                                                     \
        pr_ = 0;                                                                     \
        if ((gi_) < (nmg_)) {                                                        \
            const uint32_t lg_ = ((c0_) + ((gi_) / CH) * G) * CH + (gi_) % CH;       \
            if (lg_ >= (r_).lg0 && lg_ < (r_).lg1) {                                 \
                const uint32_t g_ = (r_).gbase + lg_;                                \
                q_ = bp.fr.qual[g_];                                                 \
                al_ = bp.fr.allele[g_];                                              \
                pr_ = bp.fr.present[g_];                                             \
            }                                                                        \
        }                                                                            \
    }
    auto prefetch_commit = [&](const RInfo &r, uint32_t slot_idx) {  // warps 1..NW-1
        uint32_t c0;
        const uint32_t nmg = my_groups(r, c0), nst = min(nmg, STG);
        {
            const uint32_t gi0 = (uint32_t)tid - 32, gi1 = gi0 + (uint32_t)(NT - 32);
            uint4 pf_q0, pf_q1;
            uint32_t pf_al0, pf_al1, pf_pr0, pf_pr1;
            FB_BW_PREFETCH_ONE(r, gi0, pf_q0, pf_al0, pf_pr0, c0, nst)
            FB_BW_PREFETCH_ONE(r, gi1, pf_q1, pf_al1, pf_pr1, c0, nst)
            if (gi0 < nst) {
                rq[gi0] = pf_q0;
                ral[gi0] = pf_al0;
                rpr[gi0] = (uint16_t)pf_pr0;
            }
            if (gi1 < nst) {
                rq[gi1] = pf_q1;
                ral[gi1] = pf_al1;
                rpr[gi1] = (uint16_t)pf_pr1;
            }
        }
        asm volatile("bar.sync 3, %0;" ::"n"(NT - 32) : "memory");  // warps 1..NW-1 only
        // the per-read sums, one thread per cell (the position hash fb_G is the expensive part)
        unsigned long long total = 0, dl = 0;
        const uint8_t *rq8 = reinterpret_cast<const uint8_t *>(rq);
        for (uint32_t x = (uint32_t)tid - 32; x < nmg * 16u; x += NT - 32) {
            const uint32_t gi = x >> 4, k = x & 15u;
            const uint32_t lg = (c0 + (gi / CH) * G) * CH + gi % CH;
            if (lg >= r.lg0 && lg < r.lg1) {
                uint32_t pr, al, qb;
                if (gi < nst) {
                    pr = rpr[gi];
                    al = ral[gi];
                    qb = rq8[gi * 16 + k];
                } else {  // beyond the staging capacity (reads of more than 256 x gridDim groups): straight from global memory
                    const uint32_t g = r.gbase + lg;
                    pr = bp.fr.present[g];
                    al = bp.fr.allele[g];
                    qb = qual8[(uint64_t)g * 16 + k];
                }
                if ((pr >> k) & 1u) {
                    const unsigned long long w = lut_s[qb];
                    const uint32_t av = ((al >> k) & 1u) | (((al >> (16 + k)) & 1u) << 1);
                    total += w;
                    dl += fb_G((in.ag0 + lg) * 16u + k, av) * w;
                }
            }
        }
        total = fb_warp_sum_u64(total);
        dl = fb_warp_sum_u64(dl);
        if (lane == 0) {
            if (total) atomicAdd(&bp.wstep[slot_idx].total, total);
            if (dl) atomicAdd(&bp.wstep[slot_idx].delta, dl);
        }
    };
    // One warp pass per (live state, tile of 32 of this CTA's groups of read `r`): the slice of the read against the state's
    // is-max planes, partial sums RED-added into the slot `accp` of step `stepno`.  The read's planes are the staged ones.
    //   early = false  regular scoring at the top of step `stepno`: states not yet scored for it (created by a copy)
    //   early = true   read `stepno` one step ahead, every state that is live now; [in_lo, in_hi) = groups of the read of
    //                  the running step, whose planes that step may still update in place (extremes kept apart)
    auto score_states = [&](const RInfo &r, BeamWideAcc *accp, uint32_t stepno, bool early, uint32_t in_lo, uint32_t in_hi) {
        const int n_live = ms->n_live;
        uint32_t c0;
        const uint32_t nmg = my_groups(r, c0);
        const int tiles_g = (int)((nmg + 31) / 32);
        const int n_tiles = n_live * tiles_g;
        for (int t = (int)warp - 1; t < n_tiles; t += NW - 1) {
            const int si = t % n_live, gt = t / n_live;
            const int s = live[si];
            if (!early && a_done[s] == stepno) continue;  // scored one step ahead (warp-uniform)
            const int hi = st_hi[s];
            const uint32_t gi = (uint32_t)gt * 32 + lane;
            const uint32_t lg = (c0 + (gi / CH) * G) * CH + gi % CH;
            const bool valid = gi < nmg && lg >= r.lg0 && lg < r.lg1;
            const bool in_rng = early && lg >= in_lo && lg < in_hi;
            uint32_t same = 0, emptyw = 0, ne_cnt = 0;  // per lane: at most 16 weights of 2^26
            int last_diff = -1, first_empty = INT_MAX;
            if (valid) {
                const uint2 m = ((int)lg <= hi) ? __ldcg(ST_MASK(s) + lg) : make_uint2(0u, 0u);
                uint4 q;
                uint32_t al, pr;
                if (gi < STG) {
                    q = rq[gi];
                    al = ral[gi];
                    pr = rpr[gi];
                } else {
                    const uint32_t g = r.gbase + lg;
                    q = bp.fr.qual[g];
                    al = bp.fr.allele[g];
                    pr = bp.fr.present[g];
                }
                uint32_t w[16];
                fb_group_weights(q, pr, lut_s, w);
                uint32_t sb, ne;
                fb_group_masks(al, m, sb, ne);
                same = fb_masked_sum(w, sb);
                const uint32_t eb = pr & ~ne & 0xFFFFu;
                if (eb) {
                    emptyw = fb_masked_sum(w, eb);
                    ne_cnt = __popc(eb);
                    first_empty = (int)(lg * 16u) + __ffs(eb) - 1;
                }
                const uint32_t db = pr & ne & ~sb & 0xFFFFu;
                if (db) last_diff = (int)(lg * 16u) + 31 - __clz(db);
            }
            // warp sums of values below 2^30 as two 16-bit halves (each REDUX result fits 32 bits)
            const unsigned long long same_w = (unsigned long long)__reduce_add_sync(0xFFFFFFFFu, same & 0xFFFFu) +
                                              ((unsigned long long)__reduce_add_sync(0xFFFFFFFFu, same >> 16) << 16);
            ne_cnt = __reduce_add_sync(0xFFFFFFFFu, ne_cnt);
            unsigned long long emptyw_w = 0;
            int fe_out = INT_MAX, fe_in = INT_MAX, ld_out = -1, ld_in = -1;
            if (ne_cnt) {
                emptyw_w = (unsigned long long)__reduce_add_sync(0xFFFFFFFFu, emptyw & 0xFFFFu) +
                           ((unsigned long long)__reduce_add_sync(0xFFFFFFFFu, emptyw >> 16) << 16);
                fe_out = __reduce_min_sync(0xFFFFFFFFu, in_rng ? INT_MAX : first_empty);
                if (early) fe_in = __reduce_min_sync(0xFFFFFFFFu, in_rng ? first_empty : INT_MAX);
            }
            if (!bp.eps_safe) {  // only the epsilon order needs it
                ld_out = __reduce_max_sync(0xFFFFFFFFu, in_rng ? -1 : last_diff);
                if (early) ld_in = __reduce_max_sync(0xFFFFFFFFu, in_rng ? last_diff : -1);
            }
            if (lane == 0) {
                if (same_w) atomicAdd(&accp[s].same, same_w);
                if (ne_cnt) {
                    atomicAdd(&accp[s].ne_cnt, ne_cnt);
                    if (emptyw_w) atomicAdd(&accp[s].emptyw, emptyw_w);
                    if (fe_out != INT_MAX) atomicMin(&accp[s].first_empty, fe_out);
                    if (fe_in != INT_MAX) atomicMin(&accp[s].first_empty_in, fe_in);
                }
                if (ld_out >= 0) atomicMax(&accp[s].last_diff, ld_out);
                if (ld_in >= 0) atomicMax(&accp[s].last_diff_in, ld_in);
            }
        }
        if (early)
            for (int i = tid - 32; i < n_live; i += NT - 32) a_done[live[i]] = stepno;
    };
    // pipelined upload: the planes of read t are valid once *bp.ready > t.  Thread 0 waits (bounded) before the CTA touches
    // a read; step t touches read t (phase C) and read t + 1 (prefetch).
    unsigned int ready_seen = 0;
    if (bp.ready) {
        if (tid == 0) fb_wait_reads(bp.ready, min(2u, in.n_reads), ready_seen);
        __syncthreads();
    }
    if (warp != 0) prefetch_commit(ri_next, 0);
    __syncthreads();

    for (uint32_t step = 0; step < in.n_reads; ++step) {
        const uint32_t width = step < 25 ? Wmax : bp.B;  // global_clustering.rs:50-53
        const RInfo ri = ri_next;
        const RExtra rx = rx_next;
        if (step + 1 < in.n_reads) {
            ri_next = rd_i[(step + 1) & 63u];
            rx_next = rd_x[(step + 1) & 63u];
        }
        if ((step & 31u) == 0 && step >= 32 && warp == NW - 1) {  // ring slots of steps [step - 32, step) are dead: refill
            const uint32_t t = step + 32 + lane;
            if (t < in.n_reads) {
                rd_i[t & 63u] = rinfo[t];
                rd_x[t & 63u] = rextra[t];
            }
        }
        const uint32_t cur_start = rx.first0;
        const int par = (int)(step & 1u);
        const int gmax_new = max(gmax, (int)ri.lg1 - 1);
        const uint32_t wend = (uint32_t)(gmax_new + 1) * 16u;  // one past the last live window position
        const uint32_t sl = step % FB_BW_SLOTS;
        BeamWideAcc *acc = bp.wacc + (uint64_t)sl * bp.maxNS;
        BeamWideStep *sacc = bp.wstep + sl;

        // ---- phase A (warps 1..; warp 0 is still closing the previous step's bookkeeping): this CTA's slice of the
        //      read against every live state ------------------------------------------------------------------------------
        if (warp != 0) {
            long long q0 = 0;
            if (prof1) q0 = fb_clock();
            const int n_live = ms->n_live;
            score_states(ri, acc, step, false, 0u, 0u);  // states created by a copy in the previous step (all of them at step 0)
            // hash terms of the positions [prev_start, cur_start) that leave the window, my slice of every live state
            if (cur_start > prev_start) {
                const uint32_t dA = (prev_start >> 4) / CH, dB = ((cur_start - 1) >> 4) / CH;
                const uint32_t d0 = my_first_chunk(dA);
                if (d0 <= dB) {
                    for (int si = (int)warp - 1; si < n_live; si += NW - 1) {
                        const int s = live[si];
                        const uint32_t pend = min(cur_start, (uint32_t)(st_hi[s] + 1) * 16u);
                        unsigned long long sub = 0;
                        const unsigned long long *c = ST_CNT(s);
                        for (uint32_t ch = d0; ch <= dB; ch += G) {
                            const uint32_t pos = ch * CH * 16u + lane;  // the chunk's 32 positions, lane = position
                            if (pos >= prev_start && pos < pend) {
#pragma unroll
                                for (uint32_t a = 0; a < 4; ++a)
                                    sub += fb_G(in.ag0 * 16u + pos, a) * (__ldcg(c + (uint64_t)pos * 4 + a) & FB_CNT_MASK);
                            }
                        }
                        sub = fb_warp_sum_u64(sub);
                        if (lane == 0 && sub) atomicAdd(&acc[s].sub, sub);
                    }
                }
            }
            if (prof1) pa += fb_clock() - q0;
        }
        PROF(1)  // warp 0: the bookkeeping that follows the previous step's live list (overlaps phase A)
        fb_grid_barrier(bp.wbar, bar_target += G);
        PROF(3)
        const int n_nodes = ms->n_nodes[gen];
        const int n_live = ms->n_live;

        // ---- phase B.1 (warp 0 of every CTA, redundantly): scores and p-values of all live states, a pair of lanes per state;
        //      the other warps are already staging / scoring the next read --------------------------------------------------
        if (warp == 0) {
            const unsigned long long tot_q = __ldcg(&sacc->total);
            if (lane == 0) ms->delta[par] = __ldcg(&sacc->delta);
            for (int base = 0; base < n_live; base += 16) {
                const int si = base + (int)(lane >> 1);
                const uint32_t sub = lane & 1u;
                const bool v = si < n_live;
                const int s = v ? live[si] : 0;
                const unsigned long long same_q = v ? __ldcg(&acc[s].same) : 0, emptyw = v ? __ldcg(&acc[s].emptyw) : 0;
                const unsigned long long hsub = v ? __ldcg(&acc[s].sub) : 0;
                const uint32_t ne_cnt = v ? __ldcg(&acc[s].ne_cnt) : 0;
                const long long diff_q = v ? (long long)(tot_q - same_q - emptyw) : 0;
                if (bp.prof && writer && tid == 0) pt[16] += fb_clock_after((unsigned long long)diff_q + hsub + ne_cnt) - tc;  // sums arrived from L2
                double diff_f = 0.0;
                bool need_replay = false;
                if (ne_cnt == 0)
                    diff_f = fb_q26_to_f64(diff_q);
                else if (bp.eps_safe)
                    diff_f = fb_q26_to_f64(diff_q + (long long)ne_cnt * (long long)(bp.eps * FB_Q26));
                else {
                    // extremes: groups the previous step's read does not cover + those it covers, under the planes that apply
                    const bool upd = inpl[s] == step;  // updated in place by the previous step
                    const int ld = max(__ldcg(&acc[s].last_diff), upd ? __ldcg(&acc[s].last_diff_new) : __ldcg(&acc[s].last_diff_in));
                    const int fe = min(__ldcg(&acc[s].first_empty), upd ? __ldcg(&acc[s].first_empty_new) : __ldcg(&acc[s].first_empty_in));
                    if (ld < fe)
                        // every empty position lies right of every diff position: the exact dyadic part followed by ne_cnt
                        // consecutive `+= epsilon`, evaluated in closed form per binade (fb_add_eps_n)
                        diff_f = fb_add_eps_n(fb_q26_to_f64(diff_q), bp.eps, ne_cnt);
                    else
                        need_replay = true;
                }
                const double same_f = fb_q26_to_f64((long long)same_q);
                const double pv = fb_pvalue_pair(same_f, diff_f, sub, bp.eps, bp.div_factor, div_pow2, inv_div);
                if (bp.prof && writer && tid == 0) pt[17] += fb_clock_after((unsigned long long)__double_as_longlong(pv)) - tc;  // p-value formed
                if (v && sub == 0) {
                    sc_same[s] = same_f;
                    sc_diff[s] = diff_f;
                    sc_pv[s] = pv;
                    if (hsub) st_hash[s] -= hsub;
                    if (need_replay) replay[atomicAdd(&ms->n_replay, 1)] = s;
                }
            }
            __syncwarp();
            const int n_replay = ms->n_replay;
            if (n_replay) {
                // ordered replay over the whole read (utils_frags.rs:33-72 in canonical order); reads planes that other
                // CTAs own, hence the second grid barrier of this step before phase C
                const uint32_t g0 = ri.gbase + ri.lg0, g1 = ri.gbase + ri.lg1;
                for (int x = 0; x < n_replay; ++x) {
                    const int s = replay[x];
                    const double diff_f =
                        fb_replay_diff_t<32, true>(bp.fr, g0, g1, ST_MASK(s), ri.lg0, st_hi[s], lut_s, bp.eps, wscr);
                    const double pv = fb_pvalue_pair(sc_same[s], diff_f, lane & 1u, bp.eps, bp.div_factor, div_pow2, inv_div);
                    if (lane == 0) {
                        sc_diff[s] = diff_f;
                        sc_pv[s] = pv;
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    ms->n_replay = 0;
                    ms->full_reads = 1;
                }
                __syncwarp();
            }
        }
        if (bp.ready) {
            // the next step prefetches read step + 2 at the top of its phase A: make sure it has landed (thread 0 waits, the
            // named barriers of phase B.2 / C order the other warps behind it: warp 0 arrives on them only afterwards)
            if (tid == 0) fb_wait_reads(bp.ready, min(step + 3u, in.n_reads), ready_seen);
        }
        PROF(4)

        // ---- phase B.2 / C ---------------------------------------------------------------------------------------------------
        // warp 0 of every CTA runs the decision section (redundantly, identical inputs); the other warps stage the next
        // read, wait for the job list (named barrier 1), materialise this CTA's slices of the new states, wait for the
        // next step's live list (named barrier 2) and go straight into the next step's phase A.
        // Steps that read whole states (replay / exact comparison: full_reads, same value in every CTA) must not
        // overlap another CTA's in-place update of the same step: all CTAs meet at a second grid barrier first.
        if (warp != 0) {
            long long q0 = 0;
            if (prof1) q0 = fb_clock();
            const bool fuse = FB_BW_EARLY && step + 1 < in.n_reads;  // there is a next read: it is scored now, one step ahead
            if (writer) zero_slot((step + 2) % FB_BW_SLOTS, 32, NT - 32);  // last read two steps ago, first written after the next barrier
            if (step + 1 < in.n_reads) prefetch_commit(ri_next, (step + 1) % FB_BW_SLOTS);
            if (fuse) {
                score_states(ri_next, bp.wacc + (uint64_t)((step + 1) % FB_BW_SLOTS) * bp.maxNS, step + 1, true, ri.lg0, ri.lg1);
            }
            if (prof1) {
                const long long n_ = fb_clock();
                pc += n_ - q0;
                q0 = n_;
            }
            // warp 0 joins this barrier with the job list (it must not touch the live list / extents the early scoring reads)
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
            if (ms->full_reads) fb_grid_barrier(bp.wbar, bar_target += G);
            if (fuse)  // states updated in place: the next step's p-values take the extremes written by phase C
                for (int j = tid - 32; j < ms->n_jobs_inplace; j += NT - 32) inpl[jobs[(int)Wm + 1 - j].dst] = step + 1;
            if (prof1) {
                const long long n_ = fb_clock();
                pw += n_ - q0;
                q0 = n_;
            }
            const int nj_copy = ms->n_jobs_copy, nj_inpl = ms->n_jobs_inplace;
            const int gs = (int)(cur_start >> 4);
            for (int pass = 0; pass < 2; ++pass) {
                const int nj = pass == 0 ? nj_copy : nj_inpl;
                if (nj == 0) continue;
                // copies: the whole window [gs, gmax_new]; in place: the read's groups only
                const int glo = pass == 0 ? gs : (int)ri.lg0;
                const int ghi = pass == 0 ? gmax_new : (int)ri.lg1 - 1;
                const uint32_t cA = (uint32_t)glo / CH, cB = (uint32_t)ghi / CH;
                const uint32_t c0 = my_first_chunk(cA);
                const int nmc = c0 <= cB ? (int)(div_g(cB - c0) + 1) : 0;
                const int total = nj * nmc;  // one warp pass per (job, chunk)
                // two passes in flight per warp: the loads of both are issued before either is finished
                struct CItem {
                    uint32_t src, dst;
                    int src_hi, lg;
                    bool act, in_read;
                    ulonglong2 v0, v1;
                    uint32_t pr, al, qb;
                    uint2 om;  // in place: the group's is-max planes before the update
                };
                const uint32_t k = lane & 15u;
                const bool fuse_c = fuse && pass == 1;
                BeamWideAcc *accn = bp.wacc + (uint64_t)((step + 1) % FB_BW_SLOTS) * bp.maxNS;
                const uint32_t c0n = my_first_chunk(ri_next.lg0 / CH);  // this CTA's first chunk of the next read
                auto c_load = [&](int x, CItem &it) {
                    const int jn = nj == 1 ? 0 : x / nmc, kc = x - jn * nmc;  // (one job per step is the steady state)
                    const BeamJob jb = pass == 0 ? jobs[jn] : jobs[(int)Wm + 1 - jn];
                    it.src = jb.src;
                    it.dst = jb.dst;
                    it.src_hi = jb.src_hi;
                    it.lg = (int)((c0 + (uint32_t)kc * G) * CH + (lane >> 4));
                    it.act = it.lg >= glo && it.lg <= ghi;
                    it.in_read = it.act && it.lg >= (int)ri.lg0 && it.lg < (int)ri.lg1;
                    it.v0 = make_ulonglong2(0ULL, 0ULL);
                    it.v1 = it.v0;
                    it.pr = it.al = it.qb = 0;
                    it.om = make_uint2(0u, 0u);
                    if (it.act && it.lg <= it.src_hi) {
                        const ulonglong2 *src =
                            reinterpret_cast<const ulonglong2 *>(ST_CNT(it.src) + ((uint64_t)it.lg * 16 + k) * 4);
                        it.v0 = __ldcg(src);
                        it.v1 = __ldcg(src + 1);
                        if (fuse_c) it.om = __ldcg(ST_MASK(it.src) + it.lg);
                    }
                    if (it.in_read) {
                        const uint32_t g = ri.gbase + (uint32_t)it.lg;
                        it.pr = bp.fr.present[g];
                        it.al = bp.fr.allele[g];
                        it.qb = qual8[(uint64_t)g * 16 + k];
                    }
                };
                auto c_finish = [&](const CItem &it) {
                    bool im0 = false, im1 = false, im2 = false, im3 = false;
                    if (it.act) {
                        unsigned long long c0w = it.v0.x, c1w = it.v0.y, c2w = it.v1.x, c3w = it.v1.y;
                        if (it.in_read && ((it.pr >> k) & 1u)) {
                            const uint32_t av = ((it.al >> k) & 1u) | (((it.al >> (16 + k)) & 1u) << 1);
                            const unsigned long long w = lut_s[it.qb];
                            if (av == 0) c0w = (c0w + w) | FB_PRESENT;
                            if (av == 1) c1w = (c1w + w) | FB_PRESENT;
                            if (av == 2) c2w = (c2w + w) | FB_PRESENT;
                            if (av == 3) c3w = (c3w + w) | FB_PRESENT;
                        }
                        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(ST_CNT(it.dst) + ((uint64_t)it.lg * 16 + k) * 4);
                        dst[0] = make_ulonglong2(c0w, c1w);
                        dst[1] = make_ulonglong2(c2w, c3w);
                        const unsigned long long m0 = c0w & FB_CNT_MASK, m1 = c1w & FB_CNT_MASK, m2 = c2w & FB_CNT_MASK,
                                                 m3 = c3w & FB_CNT_MASK;
                        unsigned long long mx = m0 > m1 ? m0 : m1;
                        const unsigned long long my = m2 > m3 ? m2 : m3;
                        mx = mx > my ? mx : my;
                        if (mx > 0) {
                            im0 = m0 == mx;
                            im1 = m1 == mx;
                            im2 = m2 == mx;
                            im3 = m3 == mx;
                        }
                    }
                    const uint32_t sh = lane & 16u;
                    const uint32_t b0 = (__ballot_sync(0xFFFFFFFFu, im0) >> sh) & 0xFFFFu;
                    const uint32_t b1 = (__ballot_sync(0xFFFFFFFFu, im1) >> sh) & 0xFFFFu;
                    const uint32_t b2 = (__ballot_sync(0xFFFFFFFFu, im2) >> sh) & 0xFFFFu;
                    const uint32_t b3 = (__ballot_sync(0xFFFFFFFFu, im3) >> sh) & 0xFFFFu;
                    const uint32_t nmx = b0 | (b1 << 16), nmy = b2 | (b3 << 16);
                    if (it.act && k == 0) ST_MASK(it.dst)[it.lg] = make_uint2(nmx, nmy);
                    if (fuse_c) {
                        // The next read was scored against the OLD planes of this state (early scoring): add what the update
                        // changes, one lane per cell of the next read in this item's two groups.  same / empty weight /
                        // empty count take signed deltas (wrapping adds); the extremes under the new planes go to `_new`.
                        uint32_t w = 0;
                        int ds = 0, de = 0;  // change of "same" (+1 / -1) and of "empty" (-1: a position can only fill up)
                        int pe = INT_MAX, pd = -1;  // position if the cell is empty / a diff under the new planes
                        if (it.act && it.lg >= (int)ri_next.lg0 && it.lg < (int)ri_next.lg1) {
                            const uint32_t gi = div_g((uint32_t)it.lg / CH - c0n) * CH + (uint32_t)it.lg % CH;  // staging slot
                            uint32_t pr, al, qb;
                            if (gi < STG) {
                                pr = rpr[gi];
                                al = ral[gi];
                                qb = reinterpret_cast<const uint8_t *>(rq)[gi * 16 + k];
                            } else {
                                const uint32_t g = ri_next.gbase + (uint32_t)it.lg;
                                pr = bp.fr.present[g];
                                al = bp.fr.allele[g];
                                qb = qual8[(uint64_t)g * 16 + k];
                            }
                            if ((pr >> k) & 1u) {
                                const uint32_t av = ((al >> k) & 1u) | (((al >> (16 + k)) & 1u) << 1);
                                const uint32_t sh_a = k + 16u * (av & 1u);
                                const uint32_t sb_o = ((av < 2 ? it.om.x : it.om.y) >> sh_a) & 1u;
                                const uint32_t sb_n = ((av < 2 ? nmx : nmy) >> sh_a) & 1u;
                                const uint32_t ao = it.om.x | it.om.y, an = nmx | nmy;
                                const uint32_t ne_o = ((ao | (ao >> 16)) >> k) & 1u, ne_n = ((an | (an >> 16)) >> k) & 1u;
                                w = lut_s[qb];
                                ds = (int)sb_n - (int)sb_o;
                                de = (int)ne_o - (int)ne_n;
                                const int pos = it.lg * 16 + (int)k;
                                if (!ne_n) pe = pos;
                                if (ne_n && !sb_n) pd = pos;
                            }
                        }
                        if (__any_sync(0xFFFFFFFFu, ds != 0 || de != 0)) {  // weights are at most 2^26: 32 of them fit 32 bits
                            const uint32_t ps = __reduce_add_sync(0xFFFFFFFFu, ds > 0 ? w : 0u);
                            const uint32_t ns = __reduce_add_sync(0xFFFFFFFFu, ds < 0 ? w : 0u);
                            const uint32_t nE = __reduce_add_sync(0xFFFFFFFFu, de < 0 ? w : 0u);
                            const uint32_t nC = __reduce_add_sync(0xFFFFFFFFu, de < 0 ? 1u : 0u);
                            if (lane == 0) {
                                if (ps != ns) atomicAdd(&accn[it.dst].same, (unsigned long long)ps - (unsigned long long)ns);
                                if (nC) {
                                    if (nE) atomicAdd(&accn[it.dst].emptyw, 0ULL - (unsigned long long)nE);
                                    atomicSub(&accn[it.dst].ne_cnt, nC);
                                }
                            }
                        }
                        if (!bp.eps_safe) {
                            pe = __reduce_min_sync(0xFFFFFFFFu, pe);
                            pd = __reduce_max_sync(0xFFFFFFFFu, pd);
                            if (lane == 0) {
                                if (pe != INT_MAX) atomicMin(&accn[it.dst].first_empty_new, pe);
                                if (pd >= 0) atomicMax(&accn[it.dst].last_diff_new, pd);
                            }
                        }
                    }
                };
                for (int x = (int)warp - 1; x < total; x += 2 * (NW - 1)) {
                    CItem ia, ib;
                    c_load(x, ia);
                    const bool two = x + (NW - 1) < total;  // warp-uniform
                    if (two) c_load(x + (NW - 1), ib);
                    c_finish(ia);
                    if (two) c_finish(ib);
                }
            }
            if (prof1) {
                const long long n_ = fb_clock();
                pc += n_ - q0;
                q0 = n_;
            }
            asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");  // the next step's live list
            if (prof1) pw += fb_clock() - q0;
        } else {
#define FB_BEAM_POOL_LD(p) __ldcg(p)
#define FB_BEAM_WRITER writer
#define FB_BEAM_VERIFIED(x) \
    if (lane == 0) ms->full_reads = 1
#define FB_BEAM_ARRIVE()                                              \
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");             \
    if (ms->full_reads) {                                             \
        fb_grid_barrier(bp.wbar, bar_target += G);                    \
        if (lane == 0) ms->full_reads = 0;                            \
    }
#define FB_BEAM_ARRIVE2() asm volatile("bar.arrive 2, %0;" ::"n"(NT) : "memory")
#include "fb_beam_decide.inc"
#undef FB_BEAM_POOL_LD
#undef FB_BEAM_WRITER
#undef FB_BEAM_VERIFIED
#undef FB_BEAM_ARRIVE
#undef FB_BEAM_ARRIVE2
        }
        cells += (unsigned long long)n_nodes * rx.nnz;
        tapn += (unsigned long long)n_nodes * P;
        gen ^= 1;
        prev_start = cur_start;
        gmax = gmax_new;
        // no closing barrier: the next step's grid barrier orders warp 0's bookkeeping before anything that reads it
    }
    __syncthreads();
    if (prof1) {
        atomicAdd(bp.prof + 0, (unsigned long long)pa);
        atomicAdd(bp.prof + 2, (unsigned long long)pc);
        atomicAdd(bp.prof + 20, (unsigned long long)pw);
    }

    // ---- global_clustering.rs:149-176: best = into_sorted_vec()[0]; walk the parent pointers (CTA 0, warp 0) ----------
    if (writer && warp == 0) {
        const int len = ms->n_nodes[gen];
        int e = 0;
        if (lane == 0) {
            HeapRef hp;
            hp.score = hp_score;
            hp.item = hp_item;
            hp.len = len;
            for (int x = 0; x < len; ++x) {
                hp_score[x] = ND_SCORE(gen, x);
                hp_item[x] = x;
            }
            hp.into_sorted();
            e = hp_item[0];
            bp.best_out[ii] = hp_score[0];
            bp.cells_out[ii] = cells;
            bp.tapn_out[ii] = tapn;
        }
        // the history rows are staged in shared memory a batch of steps at a time, so that the dependent walk runs at
        // shared-memory latency (one L2 round trip per batch instead of one per read)
        uint8_t *as = bp.assign_out + in.assign_off;
        const int rows = max(1, (int)(2u * FB_BEAM_RG * 16u / 4u / Wm));
        for (int hi_step = (int)in.n_reads; hi_step > 0; hi_step -= rows) {
            const int lo_step = max(0, hi_step - rows);
            const int n = (hi_step - lo_step) * (int)Wm;
            __syncwarp();
            for (int x = (int)lane; x < n; x += 32) btbuf[x] = hist[(uint64_t)lo_step * Wm + x];
            __syncwarp();
            if (lane == 0)
                for (int st = hi_step - 1; st >= lo_step; --st) {
                    const uint32_t v = btbuf[(st - lo_step) * (int)Wm + e];
                    as[st] = (uint8_t)(v >> 16);
                    e = (int)(v & 0xFFFFu);
                }
        }
        if (lane == 0 && bp.prof) {
            PROF(5)
            for (int i = 0; i < 12; ++i) atomicAdd(bp.prof + i, (unsigned long long)pt[i]);
            for (int i = 13; i < 24; ++i) atomicAdd(bp.prof + i, (unsigned long long)pt[i]);
            atomicAdd(bp.prof + 12, (unsigned long long)in.n_reads);
        }
    }
#undef PROF
#undef FB_BW_PREFETCH_ONE
#undef ST_CNT
#undef ST_MASK
#undef ND_SCORE
#undef ND_ERR
#undef ND_REF
}

#ifdef FB_BEAM_WIDE_IMPL  // defined by fb_beam_wide_tu.cu only: the kernel is not a template
// Cooperative launch: every CTA must be resident (the grid barrier spins).  bp.order lists the instances, run one after
// the other by the whole grid; the scratch is ONE slot (pool + history) sized for the largest of them.
__global__ void __launch_bounds__(FB_BW_THREADS, 1) k_beam_wide(BeamParams bp) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x;
    {
        const BeamSmem &L = bp.L;
        uint32_t *lut_s = reinterpret_cast<uint32_t *>(smem + L.off_lut);
        for (int i = tid; i < 256; i += FB_BW_THREADS) lut_s[i] = bp.lut[i];
    }
    unsigned long long bar_target = 0;
    for (int wk = 0; wk < bp.n_work; ++wk) {
        const int ii = bp.order[wk];
        __syncthreads();
        switch (bp.inst[ii].ploidy) {
            case 2: fb_beam_wide_instance<2>(bp, ii, smem, bp.scratch, bar_target); break;
            case 3: fb_beam_wide_instance<3>(bp, ii, smem, bp.scratch, bar_target); break;
            case 4: fb_beam_wide_instance<4>(bp, ii, smem, bp.scratch, bar_target); break;
            case 5: fb_beam_wide_instance<5>(bp, ii, smem, bp.scratch, bar_target); break;
            case 6: fb_beam_wide_instance<6>(bp, ii, smem, bp.scratch, bar_target); break;
            case 7: fb_beam_wide_instance<7>(bp, ii, smem, bp.scratch, bar_target); break;
            case 8: fb_beam_wide_instance<8>(bp, ii, smem, bp.scratch, bar_target); break;
            default: break;
        }
    }
}
#endif  // FB_BEAM_WIDE_IMPL
