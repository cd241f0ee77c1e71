// fb_beam.cuh — K5/K6: beam_search_phasing (global_clustering.rs:10-179) as a persistent kernel.
//
// One CTA runs one (block, ploidy) instance from first read to backtrack; CTAs pull instances from a work queue
// (largest first).  The reference keeps, per search node, a deep copy of `ploidy` nested hash maps
// (types_structs.rs:326-376 build_truncated_hap_block).  Here a node is `ploidy` references into a pool of dense
// HAPLOTYPE STATES (counts [pos][4] + is-max planes), shared copy-on-write between nodes:
//   * a child (node n, read -> hap j) differs from its parent in ONE state: ref[n][j] + read;
//   * children are scored, de-duplicated and pushed through an exact emulation of Rust's BinaryHeap BEFORE any state is
//     materialised; only the states of the surviving generation are built (in place when nothing else still needs the
//     parent state);
//   * HapBlock equality (global_clustering.rs:123-127) is decided exactly: a linear hash of the in-window counts
//     (H = sum G(pos,allele) * count mod 2^64, so H(state + read) = H(state) + delta(read)) filters candidates, and
//     every hash match is verified word by word on the virtual states;
//   * the sliding window (positions < current read's first SNP are dropped, types_structs.rs:346-360) never needs
//     a physical truncation: reads arrive sorted by first position, so dropped positions are never scored again;
//     they are removed from the hashes and excluded from the equality test.
// Heap order, `<=` tie behaviour, the width rule (ploidy*B for the first 25 reads), pruning by p - lse > ln(0.01) and
// the final into_sorted_vec()[0] + parent walk follow the reference line by line (see fb_seq.h for the heap).
//
// Anatomy of a step (one read), three CTA-wide synchronisation points:
//   phase 1  all warps   one warp per live state scores the read (planes from the state, the read's cells from the
//                        shared-memory staging buffer) and forms its p-value; other warps advance the states' window hashes
//   phase 2  warp 0      pruning (no exp/log outside the decision band), child scores, compaction, hash fold, duplicate
//                        classes (exact verification only on a fold match), BinaryHeap pushes/pops, job list
//            warps 1..   stage the NEXT read's planes in shared memory and sum its delta(read), then wait on named barrier 1
//   phase 3  warps 1..   materialise the surviving generation's new states, one thread per position, is-max planes by
//                        ballot -- overlapped with warp 0 writing the next generation's node tables, history, hashes,
//                        live / free lists (warp 0 only ARRIVES on barrier 1 after publishing the job list)
// FB_BEAM_PROF=1 accumulates per-phase cycle counts (printed by fb_run_beam); profiles/README.md has the numbers.
#pragma once
#include <limits.h>

#include <vector>

#include "fb_common.cuh"
#include "fb_replay.cuh"

// CTA sizes the kernel is instantiated for: 256 threads (one CTA per SM: lowest latency per read step, used when there are
// fewer instances than SMs can hold) and 128 threads (three CTAs per SM at 168 registers: the dependent chains of
// several instances interleave; measured 172 vs 186 ms with two CTAs on the configs[4]-shaped batch, 45 vs 56 ms on short reads,
// used for many-instance work queues such as a batched metagenome)
#define FB_BEAM_THREADS 256
#define FB_BEAM_THREADS_SMALL 128
#define FB_BEAM_WARPS (FB_BEAM_THREADS / 32)
#define FB_BEAM_RG 128  // reads of up to this many groups (2048 positions) are staged in shared memory

struct BeamTapDev {
    double *same, *diff, *logp;
    unsigned long long cap;
};

// shared-memory carve-up (same arithmetic on host and device)
struct BeamSmem {
    uint32_t off_nd_score, off_nd_err, off_nd_ref, off_st_hash, off_sc_same, off_sc_diff, off_st_hi, off_st_mark,
        off_free, off_live, off_ch_score, off_ch_parent, off_ch_part, off_ch_class, off_ch_diff, off_hp_score,
        off_hp_item, off_lut, off_wscr, off_misc, off_job, off_addnew, off_plain, off_sc_pv, off_ch_fold, off_ch_m, off_rq,
        off_ral, off_rpr, off_replay, off_adone, off_inpl, total;
    __host__ __device__ void layout(uint32_t P, uint32_t W, uint32_t NS) {
        uint32_t o = 0;
        auto take = [&](uint32_t bytes) {
            uint32_t r = o;
            o += (bytes + 15u) & ~15u;
            return r;
        };
        off_nd_score = take(2 * W * 8);
        off_nd_err = take(2 * W * P * 8);
        off_nd_ref = take(2 * W * P * 2);
        off_st_hash = take(NS * 8);
        off_sc_same = take(NS * 8);
        off_sc_diff = take(NS * 8);
        off_st_hi = take(NS * 4);
        off_st_mark = take(NS * 4);
        off_free = take(NS * 4);
        off_live = take(NS * 4);
        off_ch_score = take(W * P * 8);
        off_ch_diff = take(W * P * 8);
        off_ch_parent = take(W * P * 2);
        off_ch_part = take(W * P * 2);
        off_ch_class = take(W * P * 2);
        off_hp_score = take((W + 2) * 8);
        off_hp_item = take((W + 2) * 4);
        off_lut = take(256 * 4);
        off_wscr = take(FB_BEAM_WARPS * 16 * 4);
        off_misc = take(256);
        off_job = take((W + 2) * 16);
        off_addnew = take(NS * 4);
        off_plain = take(NS * 4);
        off_sc_pv = take(NS * 8);
        off_ch_fold = take(W * P * 8);
        off_ch_m = take(W * P * 4);
        off_rq = take(2 * FB_BEAM_RG * 16);   // staged planes of the current / next read (double buffered)
        off_ral = take(2 * FB_BEAM_RG * 4);
        off_rpr = take(2 * FB_BEAM_RG * 2);
        off_replay = take(NS * 4);  // k_beam_wide: states whose epsilon sum needs the ordered replay
        off_adone = take(NS * 4);   // k_beam_wide: read index up to which a state has been scored (early scoring)
        off_inpl = take(NS * 4);    // k_beam_wide: step whose p-values see the state as updated in place one step before
        total = o;
    }
};

struct BeamParams {
    DFragsDev fr;
    const InstDev *inst;
    const RInfo *rinfo;
    const RExtra *rextra;
    const uint32_t *lut;
    const int *order;  // instance indices to run, largest first
    int n_work;
    int *work_counter;
    uint8_t *assign_out;  // engine assign buffer 0
    double eps, div_factor, cutoff;
    int eps_safe;
    uint32_t B;           // max_number_solns
    uint32_t maxP, maxW, maxNS;
    BeamSmem L;                      // the shared-memory carve-up for (maxP, maxW, maxNS), computed once on the host
    uint8_t *scratch;     // per-CTA slots
    uint64_t slot_bytes;  // pool + history
    uint64_t hist_off;    // offset of the history array inside a slot
    unsigned long long *cells_out;  // [n_inst]
    double *best_out;               // [n_inst]
    unsigned long long *tapn_out;   // [n_inst]
    BeamTapDev tap;                 // only meaningful for single-instance calls
    unsigned long long *prof;       // optional [8] cycle counters per phase (FB_BEAM_PROF=1), else NULL
    // k_beam_wide only (fb_beam_wide.cuh): global workspace of the grid-wide reductions and the grid barrier
    struct BeamWideAcc *wacc;        // [3][maxNS] per-state partial sums of a step (three step slots)
    struct BeamWideStep *wstep;      // [3] per-read sums of a step
    unsigned long long *wbar;        // monotonic arrival counter of the grid barrier
    const unsigned int *ready;       // pipelined upload: number of leading reads whose planes are packed (NULL: all)
};


// host entry points of the beam translation unit (fb_beam_tu.cu): the kernel is compiled on its own, in parallel with
// the rest of the library
int fb_beam_occupancy(int threads, size_t smem_bytes, int *blocks_per_sm);
int fb_beam_launch(int threads, unsigned grid, size_t smem_bytes, cudaStream_t stream, const struct BeamParams &bp);

struct BeamJob {
    uint32_t src, dst;
    int src_hi;      // st_hi of the source state before this step (groups beyond it are empty)
    uint32_t _pad;
};

// cycle counter of the FB_BEAM_PROF probes: the memory clobber keeps it on its side of barriers and named barriers
__device__ __forceinline__ long long fb_clock() {
    long long c;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
    return c;
}
// the same, read only once `dep` has been produced (an operand of the instruction)
__device__ __forceinline__ long long fb_clock_after(unsigned long long dep) {
    long long c;
    asm volatile("{ .reg .u64 t; mov.u64 t, %1; mov.u64 %0, %%clock64; }" : "=l"(c) : "l"(dep) : "memory");
    return c;
}
__device__ __forceinline__ void fb_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int P, int NT>
__device__ __noinline__ void fb_beam_instance(const BeamParams &bp, const int ii, uint8_t *smem, uint8_t *slot) {
    constexpr int NW = NT / 32;  // warps of the CTA
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
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
    struct Misc {
        unsigned long long delta[2];  // delta(read) of the current / next read (step parity)
        int n_nodes[2];
        int n_live, n_free, n_jobs_copy, n_jobs_inplace;
        int hw;  // high-water mark of the state ids handed out so far (ids >= hw are free and untouched)
    };
    Misc *ms = reinterpret_cast<Misc *>(smem + L.off_misc);
    uint4 *rq = reinterpret_cast<uint4 *>(smem + L.off_rq);            // [2][FB_BEAM_RG]
    uint32_t *ral = reinterpret_cast<uint32_t *>(smem + L.off_ral);    // [2][FB_BEAM_RG]
    uint16_t *rpr = reinterpret_cast<uint16_t *>(smem + L.off_rpr);    // [2][FB_BEAM_RG]

    uint32_t *hist = reinterpret_cast<uint32_t *>(slot + bp.hist_off);
    const uint32_t *__restrict__ qual32 = reinterpret_cast<const uint32_t *>(bp.fr.qual);
    {
        const InstDev in = bp.inst[ii];
        const uint32_t Wmax = P * bp.B;
        const uint32_t NS = P * bp.B * (P + 1) + 1;
        const uint32_t npos = in.ng * 16;
        // counts + planes (uint2 == one 64-bit word), rounded to an even word count to keep 16-byte alignment
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

        // ---- init: one root node over the empty state (global_clustering.rs:29-47) ---------------------------------
        for (uint32_t s = tid; s < NS; s += NT) {
            st_hash[s] = 0;
            st_hi[s] = -1;
            st_mark[s] = 0;
            plain[s] = 0;
            addnew[s] = -1;
        }
        __syncthreads();
        if (tid == 0) {
            ms->n_free = 0;  // explicit free stack (ids below hw); state 0 is the root's empty state
            ms->hw = 1;
            ms->n_live = 1;
            live[0] = 0;
            ms->n_nodes[0] = 1;
            ND_SCORE(0, 0) = 0.0;
            for (uint32_t h = 0; h < P; ++h) {
                ND_ERR(0, 0, h) = 0.0;
                ND_REF(0, 0, h) = 0;
            }
        }
        long long pt[24] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tc = clock64();
#define PROF(i)                         \
    if (bp.prof && tid == 0) {          \
        long long n_ = clock64();       \
        pt[i] += n_ - tc;               \
        tc = n_;                        \
    }
        const double ln_p = log((double)P);
        const bool div_pow2 = (fb_f64_bits(bp.div_factor) & 0xFFFFFFFFFFFFFULL) == 0 && bp.div_factor > 1e-300 && bp.div_factor < 1e300;
        const double inv_div = 1.0 / bp.div_factor;
        int gen = 0;
        uint32_t prev_start = 0;  // block-local position0 from which the hashes are valid
        int gmax = -1;            // last block-local group touched so far
        unsigned long long cells = 0, tapn = 0;
        RInfo ri_next = rinfo[0];
        RExtra rx_next = rextra[0];
        // Everything of a step that depends on the read only is prepared one step ahead, off the critical path, by the
        // warps that idle while warp 0 runs phase 2: the read's planes are staged in shared memory (reads of up to
        // FB_BEAM_RG groups; longer ones are read from global memory) and delta(read) is summed.
        auto stage_read = [&](const RInfo &rn, int par, int w0, int nw) {  // warps [w0, w0 + nw) cooperate
            const uint32_t ngr = rn.lg1 - rn.lg0, gb = rn.gbase + rn.lg0;
            if (ngr <= FB_BEAM_RG && (int)warp >= w0 && (int)warp < w0 + nw)
                for (uint32_t x = (warp - w0) * 32 + lane; x < ngr; x += nw * 32) {
                    rq[par * FB_BEAM_RG + x] = bp.fr.qual[gb + x];
                    ral[par * FB_BEAM_RG + x] = bp.fr.allele[gb + x];
                    rpr[par * FB_BEAM_RG + x] = bp.fr.present[gb + x];
                }
        };
        auto read_delta = [&](const RInfo &rn, int par) {  // one warp: delta(read) = sum of G(pos, allele) * weight
            unsigned long long d = 0;
            for (uint32_t x = lane; x < (rn.lg1 - rn.lg0) * 4; x += 32) {
                const uint32_t lg = rn.lg0 + (x >> 2), sub = x & 3;
                const uint32_t g = rn.gbase + lg;
                const uint32_t q = qual32[(uint64_t)g * 4 + sub];
                const uint32_t al = bp.fr.allele[g], pr = bp.fr.present[g];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t c = sub * 4 + k;
                    if ((pr >> c) & 1u) {
                        const uint32_t av = ((al >> c) & 1u) | (((al >> (16 + c)) & 1u) << 1);
                        d += fb_G((in.ag0 + lg) * 16u + c, av) * (unsigned long long)lut_s[(q >> (8 * k)) & 0xFFu];
                    }
                }
            }
            d = fb_warp_sum_u64(d);
            if (lane == 0) ms->delta[par] = d;
        };
        stage_read(ri_next, 0, 0, NW - 1);
        if (warp == NW - 1) read_delta(ri_next, 0);
        __syncthreads();

        for (uint32_t step = 0; step < in.n_reads; ++step) {
            const uint32_t width = step < 25 ? Wmax : bp.B;  // global_clustering.rs:50-53
            const RInfo ri = ri_next;
            const RExtra rx = rx_next;
            if (step + 1 < in.n_reads) {  // prefetch the next read's descriptor (consumed next iteration)
                ri_next = rinfo[step + 1];
                rx_next = rextra[step + 1];
            }
            const uint32_t cur_start = rx.first0;
            const uint32_t g0 = ri.gbase + ri.lg0, g1 = ri.gbase + ri.lg1;  // global groups of the read
            const int par = (int)(step & 1u);
            const bool staged = (ri.lg1 - ri.lg0) <= FB_BEAM_RG;
            const uint4 *__restrict__ rq_c = rq + par * FB_BEAM_RG;
            const uint32_t *__restrict__ ral_c = ral + par * FB_BEAM_RG;
            const uint16_t *__restrict__ rpr_c = rpr + par * FB_BEAM_RG;
            const int n_nodes = ms->n_nodes[gen];
            const int n_live = ms->n_live;
            const int gmax_new = max(gmax, (int)ri.lg1 - 1);
            const uint32_t wend = (uint32_t)(gmax_new + 1) * 16u;  // one past the last live window position

            // ---- phase 1 (all warps): per live state: window advance of the hash, score of the read, p-value --------
            // tasks: [0, n_live) score a live state; [n_live, 2 n_live) advance a live state's hash window (only when the
            // window start moved).  The two kinds are independent, so they run on different warps.
            const int n_tasks = cur_start > prev_start ? 2 * n_live : n_live;
            for (int task = warp; task < n_tasks; task += NW) {
                const int s = live[task < n_live ? task : task - n_live];
                const int hi = st_hi[s];
                if (task >= n_live) {
                    // drop positions [prev_start, cur_start) from the state's hash
                    unsigned long long sub = 0;
                    const uint32_t pend = min(cur_start, (uint32_t)(hi + 1) * 16u);
                    if (pend > prev_start) {
                        const unsigned long long *c = ST_CNT(s);
                        for (uint32_t x = prev_start * 4 + lane; x < pend * 4; x += 32) {
                            const uint32_t pos = x >> 2, a = x & 3;
                            sub += fb_G(in.ag0 * 16u + pos, a) * (c[x] & FB_CNT_MASK);
                        }
                    }
                    sub = fb_warp_sum_u64(sub);
                    if (lane == 0) st_hash[s] -= sub;
                    continue;
                }
                long long q0 = 0;
                if (bp.prof && tid == 0) q0 = clock64();
                const uint2 *mk = ST_MASK(s);
                unsigned long long total = 0, same = 0, emptyw = 0;
                uint32_t ne_cnt = 0;
                int last_diff = -1, first_empty = INT_MAX;  // block-local positions of the last diff / first empty cell
                for (uint32_t g = g0 + lane; g < g1; g += 32) {
                    uint4 q;
                    uint32_t al, pr;
                    if (staged) {
                        q = rq_c[g - g0];
                        al = ral_c[g - g0];
                        pr = rpr_c[g - g0];
                    } else {
                        q = bp.fr.qual[g];
                        al = bp.fr.allele[g];
                        pr = bp.fr.present[g];
                    }
                    uint32_t w[16];
                    fb_group_weights(q, pr, lut_s, w);
                    uint32_t t = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) t += w[k];
                    total += t;
                    const uint32_t lg = ri.lg0 + (g - g0);
                    uint2 m = ((int)lg <= hi) ? mk[lg] : make_uint2(0u, 0u);
                    uint32_t sb, ne;
                    fb_group_masks(al, m, sb, ne);
                    same += fb_masked_sum(w, sb);
                    uint32_t eb = pr & ~ne & 0xFFFFu;
                    if (eb) {
                        emptyw += fb_masked_sum(w, eb);
                        ne_cnt += __popc(eb);
                        first_empty = min(first_empty, (int)(lg * 16u) + __ffs(eb) - 1);
                    }
                    const uint32_t db = pr & ne & ~sb & 0xFFFFu;
                    if (db) last_diff = max(last_diff, (int)(lg * 16u) + 31 - __clz(db));
                }
                if (bp.prof && tid == 0) { long long n_ = clock64(); pt[16] += n_ - q0; q0 = n_; }
                total = fb_warp_sum_u64(total);
                same = fb_warp_sum_u64(same);
                emptyw = fb_warp_sum_u64(emptyw);
                ne_cnt = fb_warp_sum_u32(ne_cnt);
                const long long diff_q = (long long)(total - same - emptyw);
                if (bp.prof && tid == 0) { long long n_ = clock64(); pt[17] += n_ - q0; q0 = n_; }
                double diff_f;
                if (ne_cnt == 0)
                    diff_f = fb_q26_to_f64(diff_q);
                else if (bp.eps_safe)
                    diff_f = fb_q26_to_f64(diff_q + (long long)ne_cnt * (long long)(bp.eps * FB_Q26));
                else if (__reduce_max_sync(0xFFFFFFFFu, last_diff) < __reduce_min_sync(0xFFFFFFFFu, first_empty))
                    // every empty position lies right of every diff position (the usual case: the read's tail runs past
                    // the haplotype's coverage): the reference's sum is the exact dyadic part followed by ne_cnt
                    // consecutive `+= epsilon`, which fb_add_eps_n evaluates in closed form per binade
                    diff_f = fb_add_eps_n(fb_q26_to_f64(diff_q), bp.eps, ne_cnt);
                else
                    diff_f = fb_replay_diff(bp.fr, g0, g1, mk, ri.lg0, hi, lut_s, bp.eps, wscr);
                if (bp.prof && tid == 0) { long long n_ = clock64(); pt[18] += n_ - q0 + (long long)(diff_f * 0.0); q0 = n_; }
                {
                    // stable_binom_cdf_p_rev (utils_frags.rs:211-248) with its two log terms evaluated on two lanes; the
                    // operations and their order are those of fb_stable_binom_cdf_p_rev (global_clustering.rs:81-88).
                    const double same_f = fb_q26_to_f64((long long)same);
                    const unsigned long long nn = fb_as_usize(same_f + diff_f), kk = fb_as_usize(diff_f);
                    double pvs = 0.0;
                    if (nn != 0) {
                        const double n64 = (double)nn, k64 = (double)kk;
                        double a = k64 / n64;
                        if (a == 1.0) a = 0.9999999;
                        if (a == 0.0) a = 0.0000001;
                        const double x = lane == 0 ? a : (1.0 - a);
                        const double y = lane == 0 ? bp.eps : (1.0 - bp.eps);
                        const double t = x * log(x / y);
                        const double t1 = __shfl_sync(0xFFFFFFFFu, t, 1);
                        double rel_ent = t + t1;
                        if (a < bp.eps) rel_ent = -rel_ent;
                        // x / div_factor == x * (1 / div_factor) bit for bit when div_factor is a power of two (the
                        // reference's DIV_FACTOR is 0.25): skip the software division on the critical path
                        pvs = (div_pow2 ? -1.0 * n64 * inv_div : -1.0 * n64 / bp.div_factor) * rel_ent;
                    }
                    if (lane == 0) {
                        sc_same[s] = same_f;
                        sc_diff[s] = diff_f;
                        sc_pv[s] = 1.0 * pvs;
                    }
                    if (bp.prof && tid == 0) { long long n_ = clock64(); pt[19] += n_ - q0 + (long long)(pvs * 0.0); q0 = n_; }
                }
            }
            __syncthreads();
            PROF(0)

            // ---- phase 2 (warp 0): pruning, child scores, equality classes, exact BinaryHeap emulation, next generation;
            //      the other warps stage the next read (planes -> shared memory, delta) ------------------------------------------------
            if (warp != 0) {
                if (step + 1 < in.n_reads) {
                    stage_read(ri_next, par ^ 1, 1, NW - 2);
                    if (warp == NW - 1) read_delta(ri_next, par ^ 1);
                }
                // wait until warp 0 has published the job list of this step (named barrier 1: warp 0 only arrives and goes
                // on with the next generation's node tables, so that bookkeeping overlaps the materialisation)
                asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
                    // ---- phase 3 (warps 1..7): materialise the new states (types_structs.rs:368-373 on the dense layout) ---------
                    {
                    const int nj_copy = ms->n_jobs_copy, nj_inpl = ms->n_jobs_inplace;
                    const int gs = (int)(cur_start >> 4);
                    const uint8_t *__restrict__ qual8 = reinterpret_cast<const uint8_t *>(bp.fr.qual);
                    // copies: groups [gs, gmax_new]; in place: the read's groups only.  One thread per (job, group, position):
                    // a warp touches 1 KB of contiguous counts (two 16-byte accesses per lane), and the is-max planes of a
                    // group are assembled with one ballot per allele.
                    for (int pass = 0; pass < 2; ++pass) {
                        const int nj = pass == 0 ? nj_copy : nj_inpl;
                        const int glo = pass == 0 ? gs : (int)ri.lg0;
                        const int ghi = pass == 0 ? gmax_new : (int)ri.lg1 - 1;
                        const int npj = (ghi - glo + 1) * 16;  // positions per job
                        const int total = nj * npj;
                        for (int base = 0; base < total; base += NT - 32) {
                            const int idx = base + tid - 32;
                            const bool act = idx < total;
                            const uint32_t k = (uint32_t)tid & 15u;
                            bool im0 = false, im1 = false, im2 = false, im3 = false;
                            int lg = 0;
                            uint32_t dst_state = 0;
                            if (act) {
                                const int jn = idx / npj, rem = idx - jn * npj;
                                const BeamJob jb = pass == 0 ? jobs[jn] : jobs[(int)Wm + 1 - jn];
                                lg = glo + (rem >> 4);
                                dst_state = jb.dst;
                                const uint64_t pos = (uint64_t)lg * 16 + k;
                                unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                                if (lg <= jb.src_hi) {
                                    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(ST_CNT(jb.src) + pos * 4);
                                    const ulonglong2 v0 = src[0], v1 = src[1];
                                    c0 = v0.x;
                                    c1 = v0.y;
                                    c2 = v1.x;
                                    c3 = v1.y;
                                }
                                if (lg >= (int)ri.lg0 && lg < (int)ri.lg1) {
                                    const uint32_t g = ri.gbase + (uint32_t)lg;
                                    // the three loads are independent (absent cells carry a valid dummy quality byte)
                                    uint32_t pr, al, qb;
                                    if (staged) {
                                        const uint32_t x = (uint32_t)lg - ri.lg0;
                                        pr = rpr_c[x];
                                        al = ral_c[x];
                                        qb = reinterpret_cast<const uint8_t *>(rq_c)[x * 16 + k];
                                    } else {
                                        pr = bp.fr.present[g];
                                        al = bp.fr.allele[g];
                                        qb = qual8[(uint64_t)g * 16 + k];
                                    }
                                    if ((pr >> k) & 1u) {
                                        const uint32_t av = ((al >> k) & 1u) | (((al >> (16 + k)) & 1u) << 1);
                                        const unsigned long long w = lut_s[qb];
                                        if (av == 0) c0 = (c0 + w) | FB_PRESENT;
                                        if (av == 1) c1 = (c1 + w) | FB_PRESENT;
                                        if (av == 2) c2 = (c2 + w) | FB_PRESENT;
                                        if (av == 3) c3 = (c3 + w) | FB_PRESENT;
                                    }
                                }
                                ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(ST_CNT(jb.dst) + pos * 4);
                                dst[0] = make_ulonglong2(c0, c1);
                                dst[1] = make_ulonglong2(c2, c3);
                                const unsigned long long m0 = c0 & FB_CNT_MASK, m1 = c1 & FB_CNT_MASK, m2 = c2 & FB_CNT_MASK,
                                                         m3 = c3 & FB_CNT_MASK;
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
                            if (act && k == 0) ST_MASK(dst_state)[lg] = make_uint2(b0 | (b1 << 16), b2 | (b3 << 16));
                        }
                    }
                }
            } else {
#define FB_BEAM_POOL_LD(p) (*(p))
#define FB_BEAM_WRITER true
#define FB_BEAM_VERIFIED(x)
#define FB_BEAM_ARRIVE() asm volatile("bar.arrive 1, %0;" ::"n"(NT) : "memory")
#define FB_BEAM_ARRIVE2()
#include "fb_beam_decide.inc"
#undef FB_BEAM_POOL_LD
#undef FB_BEAM_WRITER
#undef FB_BEAM_VERIFIED
#undef FB_BEAM_ARRIVE
#undef FB_BEAM_ARRIVE2
            }
            PROF(1)
            cells += (unsigned long long)n_nodes * rx.nnz;
            tapn += (unsigned long long)n_nodes * P;
            gen ^= 1;
            prev_start = cur_start;
            gmax = gmax_new;
            __syncthreads();
            PROF(2)
        }

        // ---- global_clustering.rs:149-176: best = into_sorted_vec()[0]; walk the parent pointers ------------------------------
        if (tid == 0) {
            const int len = ms->n_nodes[gen];
            HeapRef hp;
            hp.score = hp_score;
            hp.item = hp_item;
            hp.len = len;
            for (int e = 0; e < len; ++e) {
                hp_score[e] = ND_SCORE(gen, e);
                hp_item[e] = e;
            }
            hp.into_sorted();
            int e = hp_item[0];
            bp.best_out[ii] = hp_score[0];
            uint8_t *as = bp.assign_out + in.assign_off;
            for (int step = (int)in.n_reads - 1; step >= 0; --step) {
                const uint32_t v = hist[(uint64_t)step * Wm + e];
                as[step] = (uint8_t)(v >> 16);
                e = (int)(v & 0xFFFFu);
            }
            bp.cells_out[ii] = cells;
            bp.tapn_out[ii] = tapn;
            if (bp.prof) {
                PROF(5)
                for (int i = 0; i < 12; ++i) atomicAdd(bp.prof + i, (unsigned long long)pt[i]);
                for (int i = 13; i < 24; ++i) atomicAdd(bp.prof + i, (unsigned long long)pt[i]);
                atomicAdd(bp.prof + 12, (unsigned long long)in.n_reads);
            }
        }
#undef PROF
#undef ST_CNT
#undef ST_MASK
#undef ND_SCORE
#undef ND_ERR
#undef ND_REF
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == FB_BEAM_THREADS ? 1 : 3) k_beam(BeamParams bp) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_work;
    const int tid = threadIdx.x;
    {
        const BeamSmem &L = bp.L;
        uint32_t *lut_s = reinterpret_cast<uint32_t *>(smem + L.off_lut);
        for (int i = tid; i < 256; i += NT) lut_s[i] = bp.lut[i];
    }
    uint8_t *slot = bp.scratch + (uint64_t)blockIdx.x * bp.slot_bytes;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_work = atomicAdd(bp.work_counter, 1);
        __syncthreads();
        const int wk = s_work;
        if (wk >= bp.n_work) break;
        const int ii = bp.order[wk];
        switch (bp.inst[ii].ploidy) {
            case 2: fb_beam_instance<2, NT>(bp, ii, smem, slot); break;
            case 3: fb_beam_instance<3, NT>(bp, ii, smem, slot); break;
            case 4: fb_beam_instance<4, NT>(bp, ii, smem, slot); break;
            case 5: fb_beam_instance<5, NT>(bp, ii, smem, slot); break;
            case 6: fb_beam_instance<6, NT>(bp, ii, smem, slot); break;
            case 7: fb_beam_instance<7, NT>(bp, ii, smem, slot); break;
            case 8: fb_beam_instance<8, NT>(bp, ii, smem, slot); break;
            default: break;
        }
    }
}
