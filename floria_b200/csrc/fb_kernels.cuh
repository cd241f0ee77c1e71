// fb_kernels.cuh — packing, scoring sweep (K1), haplotype histogram (K2), MEC reduction (K3), move selection (K4).
// Reference functions each kernel reproduces are cited at the kernel.  All paths relative to /root/reference.
#pragma once
#include "fb_common.cuh"
#include "fb_replay.cuh"

// ---------------------------------------------------------------------------------------------------------------------
// pack: CSR cells -> banded planes.  One thread per stored cell.
// ---------------------------------------------------------------------------------------------------------------------
// cell_base: index of pos[0] / allele[0] / qual[0] in the contig's cell numbering (chunked uploads hand in one chunk of
// whole reads at a time; row_ptr always is the contig's full array)
__global__ void k_pack(uint64_t nnz, uint64_t cell_base, uint64_t n_reads, const uint64_t *__restrict__ row_ptr,
                       const uint32_t *__restrict__ pos, const uint8_t *__restrict__ allele,
                       const uint8_t *__restrict__ qual, const uint32_t *__restrict__ gstart,
                       const uint32_t *__restrict__ gptr, const uint32_t *__restrict__ first,
                       const uint32_t *__restrict__ last, uint8_t *__restrict__ qual_out,
                       uint32_t *__restrict__ allele_out, uint32_t *__restrict__ present_out32,
                       unsigned long long *__restrict__ err /* first bad cell + 1, 0 = none */,
                       const uint32_t *__restrict__ rshift /* per read: added to its positions (batched contigs), or NULL */) {
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // index into this launch's arrays
    const bool in = x < nnz;
    const uint32_t lane = threadIdx.x & 31u;
    // read of a cell: last r with row_ptr[r] <= c.  The CTA's cells are consecutive: two threads search the reads of its
    // first and last cell over the whole array, everybody else searches between those two (no search at all inside a long read).
    __shared__ uint64_t r_ends[2];
    auto find_read = [&](uint64_t c, uint64_t lo, uint64_t hi) {  // invariant: row_ptr[lo] <= c < row_ptr[hi]
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (row_ptr[mid] <= c)
                lo = mid;
            else
                hi = mid;
        }
        return lo;
    };
    if (threadIdx.x == 0 || threadIdx.x == 32) {
        const uint64_t x0 = (uint64_t)blockIdx.x * blockDim.x;
        const uint64_t xe = threadIdx.x == 0 ? x0 : min(x0 + blockDim.x, nnz) - 1;
        r_ends[threadIdx.x >> 5] = find_read(cell_base + xe, 0, n_reads);
    }
    __syncthreads();
    const uint64_t c = cell_base + (in ? x : 0);
    uint64_t lo = r_ends[0];
    if (in && r_ends[1] != lo) lo = find_read(c, lo, r_ends[1] + 1);
    bool ok = in;
    uint32_t g = 0xFFFFFFFFu - lane, k = 0, abits = 0, q = 0;  // a key of its own for the threads that store nothing
    if (in) {
        const uint32_t sh = rshift ? rshift[lo] : 0u;
        const uint32_t p = pos[x] + sh;
        const uint32_t a = allele[x];
        // per-cell validation (the per-read checks are done on the host): allele index fits 2 bits, positions strictly
        // ascending inside [first, last] with the end points present
        bool bad = a > 3 || p < first[lo] || p > last[lo];
        if (c > row_ptr[lo] && pos[x - 1] + sh >= p) bad = true;  // (chunks start on read boundaries: x >= 1 here)
        if (c == row_ptr[lo] && p != first[lo]) bad = true;
        if (c + 1 == row_ptr[lo + 1] && p != last[lo]) bad = true;
        if (bad) {
            atomicMin(err, c + 1);
            ok = false;
        } else {
            const uint32_t p0 = p - 1u;
            g = gptr[lo] + ((p0 >> 4) - gstart[lo]);
            k = p0 & 15u;
            abits = ((a & 1u) << k) | (((a >> 1) & 1u) << (16 + k));
            q = qual[x];
        }
    }
    // the cells of a group sit on neighbouring lanes (positions ascend inside a read): one atomic per group and warp
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, g);
    const uint32_t al_or = __reduce_or_sync(peers, abits);
    const uint32_t pr_or = __reduce_or_sync(peers, ok ? (1u << k) : 0u);
    if (ok) {
        if (lane == (uint32_t)__ffs(peers) - 1u) {
            atomicOr(&allele_out[g], al_or);
            atomicOr(&present_out32[g >> 1], pr_or << (16u * (g & 1u)));
        }
        qual_out[(uint64_t)g * 16 + k] = (uint8_t)q;
    }
}

// pipelined upload: the planes of reads [0, n) are complete
__global__ void k_set_ready(unsigned int *ready, unsigned int n) {
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ready), "r"(n) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 score_sweep.  One warp per (instance, read); lanes stride over the read's 16-cell groups.
// Reproduces utils_frags.rs:32-75 distance_read_haplo_epsilon_empty for the read against EVERY haplotype of the
// instance's current table, and (mode MOVES) the candidate-move generation of opt_iterate, local_clustering.rs:300-325.
// Exactness: same/diff weight sums are integer (units of 2^-26) warp reductions; the f64 `diff` is formed exactly as the
// reference's left-to-right sum in canonical position order (closed form when no epsilon term can round, otherwise the
// warp-cooperative replay below).
// ---------------------------------------------------------------------------------------------------------------------
#define FB_SWEEP_MOVES 0
#define FB_SWEEP_SCORE 1
#define FB_SWEEP_WARPS 8
#define FB_SWEEP_STAGES 4

// one pipeline stage of a warp: 32 groups of a read, staged by three 1-D TMA bulk copies
struct __align__(128) SweepStage {
    uint4 qual[32];       // 512 B
    uint32_t allele[32];  // 128 B
    uint16_t present[32]; //  64 B
    uint8_t _pad[64];
};

struct SweepArgs {
    DFragsDev fr;
    const InstDev *inst;
    InstState *st;
    int n_inst;
    const uint64_t *assign_prefix;  // [n_inst+1] == inst[i].assign_off, last = total
    const RInfo *rinfo;
    const uint8_t *assign[2];
    const uint2 *masks[2];
    const uint32_t *lut;  // 256 weights in units of 2^-26
    double eps;
    int eps_safe;
    int mode;
    // MOVES
    double *gain;
    // SCORE (any may be null)
    double *o_same, *o_diff;
    long long *o_same_q26, *o_diff_q26;
    uint32_t *o_nempty;
};

template <int P, int MODE, bool TMA, int L>
__device__ void fb_sweep_body(const SweepArgs &a, const InstDev &in, int ii, uint64_t slot, const RInfo ri,
                              const uint32_t *lut_s, uint32_t *wscratch, SweepStage *stg, uint64_t *bars) {
    const uint32_t lane = fb_lane() % L;  // lane inside the team of L lanes that owns this read
    const uint32_t tmask = fb_team_mask<L>();
    const int cur = a.st[ii].cur;
    const uint2 *__restrict__ masks = a.masks[cur] + in.mask_off;
    const uint32_t g0 = a.fr.gptr[ri.rid], g1 = g0 + (ri.lg1 - ri.lg0);
    // MOVES needs only `diff` per haplotype (opt_iterate); SCORE also reports `same`
    unsigned long long acc[P], emptyw[P], total = 0;  // acc = diff (MOVES) or same (SCORE) weight sums
    uint32_t ne_cnt[P];
#pragma unroll
    for (int h = 0; h < P; ++h) {
        acc[h] = 0;
        emptyw[h] = 0;
        ne_cnt[h] = 0;
    }
    // scoring of one 16-cell group against every haplotype: the 16 LUT weights first, then haplotype by haplotype a
    // masked sum over one bit word (fb_masked_sum_p: two R2P + 16 predicated adds per haplotype)
    auto score_group = [&](const uint4 q, const uint32_t al, const uint32_t pr, const uint32_t g) {
            const uint32_t lg = ri.lg0 + (g - g0);
            uint32_t w[16];
            fb_group_weights_raw(q, lut_s, w);
            if (MODE == FB_SWEEP_SCORE) total += fb_masked_sum_p(w, pr);
#pragma unroll
            for (int h = 0; h < P; ++h) {
                const uint2 m = masks[(uint32_t)h * in.ng + lg];
                uint32_t sb, ne;
                fb_group_masks(al, m, sb, ne);
                const uint32_t sel = MODE == FB_SWEEP_SCORE ? (sb & pr) : (pr & ne & ~sb);
                const uint32_t eb = pr & ~ne & 0xFFFFu;
                acc[h] += fb_masked_sum_p(w, sel);
                if (eb) {  // cells on positions the haplotype does not cover: rare inside a block
                    ne_cnt[h] += __popc(eb);
                    if (MODE == FB_SWEEP_SCORE) emptyw[h] += fb_masked_sum(w, eb);
                }
            }
    };
    if (TMA) {
        // TMA pipeline: lane 0 keeps FB_SWEEP_STAGES chunks of 32 groups in flight (three cp.async.bulk per chunk,
        // completion on the stage's mbarrier); every lane then scores one group out of shared memory.  Reads start on
        // 8-group boundaries and their planes are padded to 8 groups, so every copy is 16-byte aligned and a multiple
        // of 16 bytes.
        const uint32_t ng_pad = (g1 - g0 + 7u) & ~7u;
        const uint32_t n_chunks = (g1 - g0 + 31u) >> 5;
        auto issue = [&](uint32_t c) {
            const uint32_t st = c % FB_SWEEP_STAGES;
            const uint32_t n = min(32u, ng_pad - c * 32u);
            const uint32_t gsrc = g0 + c * 32u;
            fb_mbar_expect_tx(&bars[st], n * 22u);
            fb_bulk_g2s(stg[st].qual, a.fr.qual + gsrc, n * 16u, &bars[st]);
            fb_bulk_g2s(stg[st].allele, a.fr.allele + gsrc, n * 4u, &bars[st]);
            fb_bulk_g2s(stg[st].present, a.fr.present + gsrc, n * 2u, &bars[st]);
        };
        if (lane == 0)
            for (uint32_t c = 0; c < min(n_chunks, (uint32_t)FB_SWEEP_STAGES); ++c) issue(c);
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const uint32_t st = c % FB_SWEEP_STAGES;
            fb_mbar_wait(&bars[st], (c / FB_SWEEP_STAGES) & 1u);
            const uint32_t g = g0 + c * 32u + lane;
            const bool live = g < g1;
            const uint4 q = stg[st].qual[lane];
            const uint32_t al = stg[st].allele[lane];
            const uint32_t pr = live ? (uint32_t)stg[st].present[lane] : 0u;
            __syncwarp();
            if (lane == 0 && c + FB_SWEEP_STAGES < n_chunks) issue(c + FB_SWEEP_STAGES);
            if (live) score_group(q, al, pr, g);
        }
    } else {
        // register pipeline: the next group's loads are in flight while the current one is scored
        uint32_t g = g0 + lane;
        uint4 q_n = make_uint4(0, 0, 0, 0);
        uint32_t al_n = 0, pr_n = 0;
        if (g < g1) {
            q_n = a.fr.qual[g];
            al_n = a.fr.allele[g];
            pr_n = a.fr.present[g];
        }
        for (; g < g1; g += L) {
            const uint4 q = q_n;
            const uint32_t al = al_n, pr = pr_n;
            if (g + L < g1) {
                q_n = a.fr.qual[g + L];
                al_n = a.fr.allele[g + L];
                pr_n = a.fr.present[g + L];
            }
            score_group(q, al, pr, g);
        }
    }
    double diff_f[P];
    long long same_q[P], diff_q[P];
    if (MODE == FB_SWEEP_SCORE) total = fb_team_sum_u64<L>(total, tmask);
#pragma unroll
    for (int h = 0; h < P; ++h) {
        acc[h] = fb_team_sum_u64<L>(acc[h], tmask);
        ne_cnt[h] = fb_team_sum_u32<L>(ne_cnt[h], tmask);
        if (MODE == FB_SWEEP_SCORE) {
            emptyw[h] = fb_team_sum_u64<L>(emptyw[h], tmask);
            same_q[h] = (long long)acc[h];
            diff_q[h] = (long long)(total - acc[h] - emptyw[h]);
        } else {
            same_q[h] = 0;
            diff_q[h] = (long long)acc[h];
        }
    }
#pragma unroll
    for (int h = 0; h < P; ++h) {
        if (ne_cnt[h] == 0) {
            diff_f[h] = fb_q26_to_f64(diff_q[h]);
        } else if (a.eps_safe) {
            // every term is a multiple of 2^-26: the sum is exact in any order
            diff_f[h] = fb_q26_to_f64(diff_q[h] + (long long)ne_cnt[h] * (long long)(a.eps * FB_Q26));
        } else {
            diff_f[h] = fb_replay_diff_t<L>(a.fr, g0, g1, masks + (uint32_t)h * in.ng, ri.lg0, 0x7FFFFFFF, lut_s, a.eps, wscratch);
        }
    }
    if (lane != 0) return;
    if (MODE == FB_SWEEP_SCORE) {
#pragma unroll
        for (int h = 0; h < P; ++h) {
            uint64_t o = slot * P + h;
            if (a.o_same) a.o_same[o] = fb_q26_to_f64(same_q[h]);
            if (a.o_diff) a.o_diff[o] = diff_f[h];
            if (a.o_same_q26) a.o_same_q26[o] = same_q[h];
            if (a.o_diff_q26) a.o_diff_q26[o] = diff_q[h];
            if (a.o_nempty) a.o_nempty[o] = ne_cnt[h];
        }
    } else {
        // opt_iterate (local_clustering.rs:300-325): gain of moving the read from its haplotype i to j
        const uint32_t r_local = (uint32_t)(slot - in.assign_off);
        const int i = a.assign[cur][slot];
        double *gout = a.gain + in.gain_off + (uint64_t)r_local * P;
        const bool skip = (i >= P) || a.st[ii].sizes[cur][i] <= 1;  // `if partition[i].len() <= 1 { continue; }`
        double errors_read = 0.0;
#pragma unroll
        for (int h = 0; h < P; ++h)
            if (h == i) errors_read = diff_f[h];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            double gn = 0.0;
            if (!skip && j != i) {
                const double diff_score = errors_read - diff_f[j];
                if (diff_score > 0.0) gn = diff_score;
            }
            gout[j] = gn;
        }
    }
}

template <int P, bool TMA, int L>
__device__ __forceinline__ void fb_sweep_dispatch(const SweepArgs &a, const InstDev &in, int ii, uint64_t slot,
                                                  const RInfo ri, const uint32_t *lut_s, uint32_t *ws, SweepStage *stg,
                                                  uint64_t *bars) {
    if (a.mode == FB_SWEEP_SCORE)
        fb_sweep_body<P, FB_SWEEP_SCORE, TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars);
    else
        fb_sweep_body<P, FB_SWEEP_MOVES, TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars);
}

// PMAX = largest ploidy this instantiation handles: the register budget of a kernel is that of its widest path, so the
// host launches the narrowest variant that covers the batch (2, 4 or 8).  L = lanes per read: a whole warp for long
// reads, teams of 8 or 2 lanes for short ones (a paired short read has 1-2 groups; a 100-SNP read 7), so that the lanes
// of a warp are not idle.  Teams diverge freely (fb_team_mask), e.g. across an instance boundary with another ploidy.
template <int PMAX, bool TMA, int L>
__global__ void __launch_bounds__(FB_SWEEP_WARPS * 32) k_sweep(SweepArgs a) {
    static_assert(!TMA || L == 32, "the TMA staging is per warp");
    constexpr int TEAMS = 32 / L;
    __shared__ uint32_t lut_s[256];
    __shared__ uint32_t wscr[FB_SWEEP_WARPS * TEAMS][16];
    __shared__ SweepStage stages[TMA ? FB_SWEEP_WARPS : 1][TMA ? FB_SWEEP_STAGES : 1];
    __shared__ __align__(8) uint64_t mbars[TMA ? FB_SWEEP_WARPS : 1][FB_SWEEP_STAGES];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut_s[i] = a.lut[i];
    if (TMA && (threadIdx.x & 31) == 0) {
        for (int s2 = 0; s2 < FB_SWEEP_STAGES; ++s2) fb_mbar_init(&mbars[threadIdx.x >> 5][s2], 1);
        fb_mbar_fence_init();
    }
    __syncthreads();
    SweepStage *stg = stages[TMA ? (threadIdx.x >> 5) : 0];
    uint64_t *bars = mbars[TMA ? (threadIdx.x >> 5) : 0];
    const uint64_t total = a.assign_prefix[a.n_inst];
    const uint32_t team = threadIdx.x / L;  // team index inside the CTA
    const uint64_t slot = (uint64_t)blockIdx.x * (FB_SWEEP_WARPS * TEAMS) + team;
    if (slot >= total) return;
    const int ii = fb_upper_seg(a.assign_prefix, a.n_inst, slot);
    const InstDev in = a.inst[ii];
    if (a.mode == FB_SWEEP_MOVES && !a.st[ii].active) return;
    const RInfo ri = a.rinfo[in.read_off + (uint32_t)(slot - in.assign_off)];
    uint32_t *ws = wscr[team];
    switch (in.ploidy) {
        case 1: fb_sweep_dispatch<1, TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        case 2: fb_sweep_dispatch<2, TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        case 3: if (PMAX >= 3) fb_sweep_dispatch<(PMAX >= 3 ? 3 : 1), TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        case 4: if (PMAX >= 4) fb_sweep_dispatch<(PMAX >= 4 ? 4 : 1), TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        case 5: if (PMAX >= 5) fb_sweep_dispatch<(PMAX >= 5 ? 5 : 1), TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        case 6: if (PMAX >= 6) fb_sweep_dispatch<(PMAX >= 6 ? 6 : 1), TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        case 7: if (PMAX >= 7) fb_sweep_dispatch<(PMAX >= 7 ? 7 : 1), TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        case 8: if (PMAX >= 8) fb_sweep_dispatch<(PMAX >= 8 ? 8 : 1), TMA, L>(a, in, ii, slot, ri, lut_s, ws, stg, bars); break;
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2 hap_histogram.  One CTA per (instance, tile of 512 positions, haplotype, read split).  Lane l of every warp owns
// group l of the tile (16 positions); the 8 warps take the matching reads round-robin.  A read's slice of the tile
// is staged into a per-warp ring of shared-memory stages by cp.async (LDGSTS: every lane copies its own group's 16 B of
// quals, 4 B of alleles and the 4-byte word holding its 2 B of presence into a lane-private slot) several rows ahead,
// so the row loads cost no registers and stay in flight while earlier rows are accumulated.  (A 1-D TMA bulk-copy
// variant of the same ring was measured slower: three UBLKCP per row cost ~60 issue slots of address set-up, see
// profiles/.)  Per position a lane keeps 32-bit partial sums T (all alleles), C1 (allele bit 0 set) and C2 (allele
// bit 1 set) in registers (fb_padd: R2P-unpacked predicates) and flushes them every 60 rows into 64-bit tables in
// shared memory held as two 32-bit halves (native 32-bit shared atomics, carry into the high half); allele 3 and
// zero-weight keys are rare and go straight to shared-memory atomics.  n3 = C3, n1 = C1-C3, n2 = C2-C3, n0 = T-C1-C2+C3.
// With several read splits per (tile, haplotype) the partial tables are merged with 64-bit global atomics (exact integer
// adds: deterministic) and the last CTA to finish derives the is-max planes.
// Reproduces utils_frags.rs:160-184 set_to_seq_dict / hap_block_from_partition; the planes are the consensus test of
// utils_frags.rs:53-69.
// ---------------------------------------------------------------------------------------------------------------------
#define FB_HIST_THREADS 256
#define FB_HIST_WARPS 8
#define FB_HIST_TILE_GROUPS 32  // 512 positions
#define FB_HIST_LIST 512
#define FB_HIST_STAGES 4

struct __align__(16) HistStage {
    uint4 qual[32];       // lane-private slots
    uint32_t allele[32];
    uint32_t present[32];  // the aligned 4-byte word holding the lane's 16-bit presence mask
};
// dynamic shared memory: T, C1, C2, C3 tables as [table][lo|hi][512] 32-bit words, then the per-warp stage rings
#define FB_HIST_TAB_BYTES (4 * 2 * 512 * 4)
#define FB_HIST_SMEM (FB_HIST_TAB_BYTES + FB_HIST_WARPS * FB_HIST_STAGES * (int)sizeof(HistStage))

__device__ __forceinline__ void fb_cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fb_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fb_cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(fb_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fb_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct HistArgs {
    DFragsDev fr;
    const InstDev *inst;
    const InstState *st;
    int n_inst;
    const uint64_t *tile_prefix;  // [n_inst+1] prefix of tiles(i) * ploidy(i) * splits(i)
    const uint32_t *splits;       // [n_inst] read splits per (tile, haplotype)
    const uint64_t *done_off;     // [n_inst] offset into done[] ([tiles][ploidy] counters)
    uint32_t *done;               // zeroed before the launch
    const RInfo *rinfo;
    const uint8_t *assign[2];
    uint64_t *cnt[2];
    uint2 *masks[2];
    const uint32_t *lut;
    int use_phred;    // 0: every cell weighs 1.0 (get_mec_stats_epsilon_no_phred / get_errors_cov_from_frags)
    int which;        // 0: the instance's current buffer, 1: the other ("new") one, 2: explicit buffer `buf`
    int buf;
    int only_active;  // skip instances whose optimize loop has finished
    int assign_cur;   // 1: read the partition from the CURRENT buffer whatever buffer the table is written to
};

// zero the count tables of the instances that merge with atomics
__global__ void k_hist_zero(HistArgs a) {
    const int ii = blockIdx.y;
    if (a.splits[ii] <= 1) return;
    if (a.only_active && !a.st[ii].active) return;
    const InstDev in = a.inst[ii];
    const int cur = a.st[ii].cur;
    const int buf = a.which == 0 ? cur : (a.which == 1 ? (cur ^ 1) : a.buf);
    ulonglong2 *p = reinterpret_cast<ulonglong2 *>(a.cnt[buf] + in.cnt_off);
    const uint64_t n2 = (uint64_t)in.ploidy * in.ng * 32;  // 64 words per group = 32 ulonglong2
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x)
        p[x] = make_ulonglong2(0ULL, 0ULL);
}

// 64-bit add into a (lo, hi) pair of 32-bit shared-memory words
__device__ __forceinline__ void fb_tab_add(uint32_t *lo, uint32_t *hi, uint32_t idx, uint32_t v) {
    if (v) {
        const uint32_t old = atomicAdd(&lo[idx], v);
        if (old > ~v) atomicAdd(&hi[idx], 1u);
    }
}

#ifndef FB_HIST_MIN_CTAS
#define FB_HIST_MIN_CTAS 2  // three resident CTAs (85 registers) were measured: profiles/README.md
#endif
__global__ void __launch_bounds__(FB_HIST_THREADS, FB_HIST_MIN_CTAS) k_hist(HistArgs a) {
    __shared__ uint32_t lut_s[256];
    __shared__ uint4 s_list[FB_HIST_LIST];
    extern __shared__ __align__(128) uint8_t hist_dyn[];
    uint32_t(*tab)[2][512] = reinterpret_cast<uint32_t(*)[2][512]>(hist_dyn);  // [T|C1|C2|C3][lo|hi][k*32+lane]
    __shared__ uint32_t zf_tab[512];  // bit a: a zero-weight cell inserted allele key a
    __shared__ int s_n, s_last, s_zero_any, s_zero_other;
    const int t = threadIdx.x;
    const uint32_t lane = t & 31, warp = t >> 5;
    HistStage *stg = reinterpret_cast<HistStage *>(hist_dyn + FB_HIST_TAB_BYTES) + warp * FB_HIST_STAGES;
    if (t == 0) {
        s_n = 0;
        s_zero_any = 0;
        s_zero_other = 0;
    }
    for (int i = t; i < FB_HIST_TAB_BYTES / 4; i += FB_HIST_THREADS) reinterpret_cast<uint32_t *>(hist_dyn)[i] = 0u;
    for (int i = t; i < 512; i += FB_HIST_THREADS) zf_tab[i] = 0;
    __syncthreads();
    for (int i = t; i < 256; i += FB_HIST_THREADS) {
        const uint32_t v = a.use_phred ? a.lut[i] : (1u << 26);
        lut_s[i] = v;
        if (v == 0) {
            s_zero_any = 1;
            if (i != 0) s_zero_other = 1;
        }
    }
    const uint64_t cta = blockIdx.x;
    const int ii = fb_upper_seg(a.tile_prefix, a.n_inst, cta);
    const InstDev in = a.inst[ii];
    if (a.only_active && !a.st[ii].active) return;
    const uint32_t S = a.splits[ii];
    uint32_t rel = (uint32_t)(cta - a.tile_prefix[ii]);
    const uint32_t split = rel % S;
    rel /= S;
    const uint32_t h = rel % in.ploidy, tile = rel / in.ploidy;
    const int cur = a.st[ii].cur;
    const int buf = a.which == 0 ? cur : (a.which == 1 ? (cur ^ 1) : a.buf);
    const uint8_t *__restrict__ assign = a.assign[a.assign_cur ? cur : buf] + in.assign_off;
    const RInfo *__restrict__ rinfo = a.rinfo + in.read_off;
    const uint32_t tg0 = tile * FB_HIST_TILE_GROUPS;  // first block-local group of the tile
    const uint32_t G = tg0 + lane;                    // this lane's block-local group
    const uint32_t r_begin = (uint32_t)((uint64_t)in.n_reads * split / S);
    const uint32_t r_end = (uint32_t)((uint64_t)in.n_reads * (split + 1) / S);
    // optional position filter (types_structs.rs:173: HapNode::new keeps positions inside snp_endpoints only)
    uint32_t flt16 = 0xFFFFu;
    if (in.flt_lo != 0 || in.flt_hi != 0xFFFFFFFFu) {
        flt16 = 0;
        for (uint32_t k = 0; k < 16; ++k) {
            const uint32_t pos = G * 16 + k;
            if (pos >= in.flt_lo && pos <= in.flt_hi) flt16 |= 1u << k;
        }
    }
    uint32_t t32[16], c1[16], c2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        t32[k] = 0;
        c1[k] = 0;
        c2[k] = 0;
    }
    int since_flush = 0;
    uint32_t it0 = 0;  // rows this warp has pushed through its stage ring so far (stage = it % STAGES, parity = it / STAGES)
    __syncthreads();
    const bool lut_zero = s_zero_any != 0, lut_zero_only_q0 = s_zero_other == 0;
    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t idx = k * 32 + lane;
            fb_tab_add(tab[0][0], tab[0][1], idx, t32[k]);
            fb_tab_add(tab[1][0], tab[1][1], idx, c1[k]);
            fb_tab_add(tab[2][0], tab[2][1], idx, c2[k]);
            t32[k] = 0;
            c1[k] = 0;
            c2[k] = 0;
        }
    };
    // every lane: start the copies of its group of list entry e into its slot of stage `it % STAGES` (one commit group
    // per row, also when the lane has nothing to copy, so that the group counts of all lanes agree)
    auto issue = [&](uint32_t e, uint32_t it) {
        const uint4 en = s_list[e];
        if (G >= en.y && G < en.z) {
            HistStage &sg = stg[it % FB_HIST_STAGES];
            const uint32_t g = en.x + G;
            fb_cp_async16(&sg.qual[lane], a.fr.qual + g);
            fb_cp_async4(&sg.allele[lane], a.fr.allele + g);
            fb_cp_async4(&sg.present[lane], a.fr.present + (g & ~1u));
        }
        fb_cp_async_commit();
    };
    for (uint32_t base = r_begin; base < r_end; base += FB_HIST_LIST) {
        // reads are sorted by first position: once a chunk starts right of the tile, nothing later overlaps it
        if (rinfo[base].lg0 >= tg0 + FB_HIST_TILE_GROUPS) break;
#pragma unroll
        for (int x = 0; x < FB_HIST_LIST / FB_HIST_THREADS; ++x) {
            const uint32_t rl = base + x * FB_HIST_THREADS + t;
            if (rl < r_end) {
                const RInfo ri = rinfo[rl];
                if (assign[rl] == h && ri.lg1 > tg0 && ri.lg0 < tg0 + FB_HIST_TILE_GROUPS) {
                    const int o = atomicAdd(&s_n, 1);  // order is irrelevant: integer adds commute
                    s_list[o] = make_uint4(ri.gbase, ri.lg0, ri.lg1, 0u);
                }
            }
        }
        __syncthreads();
        const uint32_t n = (uint32_t)s_n;
        const uint32_t cnt_w = n > warp ? (n - warp + FB_HIST_WARPS - 1) / FB_HIST_WARPS : 0u;  // entries warp, warp+8, ...
        // prologue: FB_HIST_STAGES - 1 rows in flight (empty commit groups pad a short list)
#pragma unroll
        for (uint32_t i = 0; i < FB_HIST_STAGES - 1; ++i) {
            if (i < cnt_w)
                issue(warp + i * FB_HIST_WARPS, it0 + i);
            else
                fb_cp_async_commit();
        }
        for (uint32_t i = 0; i < cnt_w; ++i) {
            const uint32_t it = it0 + i;
            // keep the ring full: row i + STAGES - 1 goes into the stage consumed in the previous iteration
            if (i + FB_HIST_STAGES - 1 < cnt_w)
                issue(warp + (i + FB_HIST_STAGES - 1) * FB_HIST_WARPS, it + FB_HIST_STAGES - 1);
            else
                fb_cp_async_commit();
            fb_cp_async_wait<FB_HIST_STAGES - 1>();  // this lane's copies of row i have landed (slots are lane-private)
            const uint4 en = s_list[warp + i * FB_HIST_WARPS];
            const HistStage &sg = stg[it % FB_HIST_STAGES];
            uint4 q = make_uint4(~0u, ~0u, ~0u, ~0u);  // uncovered group: every bit word below is 0, the weights are unused
            uint32_t al = 0, pr = 0;
            if (G >= en.y && G < en.z) {
                q = sg.qual[lane];
                al = sg.allele[lane];
                pr = (sg.present[lane] >> (16u * ((en.x + G) & 1u))) & 0xFFFFu;
            }
            const uint32_t P16 = pr & flt16;
            const uint32_t A0 = al & P16, A1 = (al >> 16) & P16;
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
            {
                uint32_t w[16];
                fb_group_weights_raw(q, lut_s, w);
#pragma unroll
                for (int k = 0; k < 16; ++k) fb_padd(t32[k], P16, 1u << k, w[k]);
#pragma unroll
                for (int k = 0; k < 16; ++k) fb_padd(c1[k], A0, 1u << k, w[k]);
                if (__any_sync(0xFFFFFFFFu, A1 != 0)) {  // allele bit 1 (alleles 2/3): warp-uniform branch
#pragma unroll
                    for (int k = 0; k < 16; ++k) fb_padd(c2[k], A1, 1u << k, w[k]);
                    for (uint32_t bits = A1 & A0; bits;) {  // allele 3 is rare: straight to the shared table
                        const int k = __ffs(bits) - 1;
                        bits &= bits - 1;
                        fb_tab_add(tab[3][0], tab[3][1], k * 32 + lane, lut_s[(qq[k >> 2] >> (8 * (k & 3))) & 0xFFu]);
                    }
                }
            }
            // zero-weight keys: a cell whose weight is 0 still inserts its allele key (utils_frags.rs:165-166).
            // When only q = 0 maps to weight 0 a zero-byte test of the quality words filters the common case.
            if (lut_zero) {
                bool maybe = true;
                if (lut_zero_only_q0) {
                    uint32_t z = 0;
#pragma unroll
                    for (int x = 0; x < 4; ++x) z |= (qq[x] - 0x01010101u) & ~qq[x] & 0x80808080u;
                    maybe = z != 0;
                }
                if (maybe) {
                    for (uint32_t bits = P16; bits;) {
                        const int k = __ffs(bits) - 1;
                        bits &= bits - 1;
                        if (lut_s[(qq[k >> 2] >> (8 * (k & 3))) & 0xFFu] == 0)
                            atomicOr(&zf_tab[k * 32 + lane], 1u << (((A0 >> k) & 1u) | (((A1 >> k) & 1u) << 1)));
                    }
                }
            }
            if (++since_flush >= 60) {  // 63 rows of weights <= 2^26 fit 32 bits
                flush();
                since_flush = 0;
            }
        }
        it0 += cnt_w;
        fb_cp_async_wait<0>();
        __syncthreads();
        if (t == 0) s_n = 0;
        __syncthreads();
    }
    flush();
    __syncthreads();
    // epilogue: thread t finalises positions 2t and 2t+1 of the tile
    const uint32_t eg = (uint32_t)t >> 3;            // group of the tile
    const uint32_t ek = ((uint32_t)t & 7u) * 2;      // first of the two positions inside the group
    const uint32_t EG = tg0 + eg;
    const bool in_range = EG < in.ng;
    unsigned long long *out = reinterpret_cast<unsigned long long *>(a.cnt[buf] + in.cnt_off) +
                              (((uint64_t)h * in.ng + EG) * 16 + ek) * 4;
    unsigned long long c4[2][4];
    uint32_t zf[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        const uint32_t idx = (ek + x) * 32 + eg;
        const unsigned long long T = tab[0][0][idx] | ((unsigned long long)tab[0][1][idx] << 32);
        const unsigned long long C1 = tab[1][0][idx] | ((unsigned long long)tab[1][1][idx] << 32);
        const unsigned long long C2 = tab[2][0][idx] | ((unsigned long long)tab[2][1][idx] << 32);
        const unsigned long long C3 = tab[3][0][idx] | ((unsigned long long)tab[3][1][idx] << 32);
        c4[x][3] = C3;
        c4[x][1] = C1 - C3;
        c4[x][2] = C2 - C3;
        c4[x][0] = T - C1 - C2 + C3;
        zf[x] = zf_tab[idx];
    }
    if (S > 1) {
        if (in_range) {
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
                for (int av = 0; av < 4; ++av) {
                    if (c4[x][av]) atomicAdd(out + x * 4 + av, c4[x][av]);
                    if (zf[x] & (1u << av)) atomicOr(out + x * 4 + av, FB_PRESENT);
                }
        }
        __threadfence();
        __syncthreads();
        if (t == 0) {
            const uint32_t prev = atomicAdd(a.done + a.done_off[ii] + (uint64_t)tile * in.ploidy + h, 1u);
            s_last = prev == S - 1;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (in_range) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const ulonglong2 u = __ldcg(reinterpret_cast<const ulonglong2 *>(out + x * 4));
                const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(out + x * 4) + 1);
                c4[x][0] = u.x;
                c4[x][1] = u.y;
                c4[x][2] = v.x;
                c4[x][3] = v.y;
                zf[x] = 0;  // presence of zero-weight keys already sits in bit 62
            }
        }
    }
    uint32_t pl[4] = {0, 0, 0, 0};
    if (in_range) {
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            unsigned long long mx = 0;
#pragma unroll
            for (int av = 0; av < 4; ++av) {
                const unsigned long long v = c4[x][av] & FB_CNT_MASK;
                if (v > 0 || (zf[x] & (1u << av))) c4[x][av] |= FB_PRESENT;
                mx = v > mx ? v : mx;
            }
            reinterpret_cast<ulonglong2 *>(out + x * 4)[0] = make_ulonglong2(c4[x][0], c4[x][1]);
            reinterpret_cast<ulonglong2 *>(out + x * 4)[1] = make_ulonglong2(c4[x][2], c4[x][3]);
            if (mx > 0) {
#pragma unroll
                for (int av = 0; av < 4; ++av)
                    if ((c4[x][av] & FB_CNT_MASK) == mx) pl[av] |= 1u << (ek + x);
            }
        }
    }
    // combine the 8 threads of a group
#pragma unroll
    for (int av = 0; av < 4; ++av) {
        pl[av] |= __shfl_xor_sync(0xFFFFFFFFu, pl[av], 1);
        pl[av] |= __shfl_xor_sync(0xFFFFFFFFu, pl[av], 2);
        pl[av] |= __shfl_xor_sync(0xFFFFFFFFu, pl[av], 4);
    }
    if (in_range && (t & 7) == 0)
        a.masks[buf][in.mask_off + (uint64_t)h * in.ng + EG] = make_uint2(pl[0] | (pl[1] << 16), pl[2] | (pl[3] << 16));
}

// ---------------------------------------------------------------------------------------------------------------------
// K3 mec_reduce.  One warp per (instance, haplotype).  Reproduces local_clustering.rs:218-260 get_mec_stats_epsilon
// (and :187-215 on the unweighted table): per position in ascending order: bases += max count; errors += the other
// counts in ascending-count order; errors += epsilon when max <= 1.0.  The two f64 accumulators are emulated exactly
// (SeqSum) in canonical position order.
// ---------------------------------------------------------------------------------------------------------------------
struct MecChunk {  // summary of 32 consecutive positions of one haplotype table
    long long bases;   // sum of the consensus counts (units of 2^-26)
    long long others;  // sum of the non-consensus counts
    uint32_t eps;      // bit l: position l adds an epsilon (consensus count <= 1.0)
    uint32_t _pad[3];
};

struct MecArgs {
    const InstDev *inst;
    const InstState *st;
    int n_inst;
    const uint64_t *hap_prefix;    // [n_inst+1] prefix of ploidy(i)
    const uint64_t *chunk_prefix;  // [n_haps+1] prefix of 32-position chunks per (instance, haplotype)
    MecChunk *chunks;              // scratch, [chunk_prefix[n_haps]]
    const uint64_t *cnt[2];
    double *mec[2];  // [mec_off + h][2] = (bases, errors)
    double eps;
    int eps_safe;
    int which, buf, only_active;
};

__device__ __forceinline__ void fb_cswap(unsigned long long &x, unsigned long long &y) {
    const unsigned long long lo = x < y ? x : y, hi = x < y ? y : x;
    x = lo;
    y = hi;
}

// one position of a haplotype table: allele_counts (keys present) sorted ascending by count
// (local_clustering.rs:229-236).  An absent key holds 0 and adding 0.0 changes nothing, so all four words go through a
// sorting network as items: v0 <= v1 <= v2 are the non-consensus counts, v3 the consensus count.
struct MecPos {
    unsigned long long v0, v1, v2, v3;
    bool anyp, has_eps;
};
__device__ __forceinline__ MecPos fb_mec_pos(const ulonglong2 *__restrict__ c2, uint32_t p, uint32_t npos) {
    MecPos r;
    ulonglong2 x = make_ulonglong2(0ULL, 0ULL), y = x;
    if (p < npos) {
        x = c2[(uint64_t)p * 2];
        y = c2[(uint64_t)p * 2 + 1];
    }
    r.anyp = ((x.x | x.y | y.x | y.y) & FB_PRESENT) != 0;
    r.v0 = x.x & FB_CNT_MASK;
    r.v1 = x.y & FB_CNT_MASK;
    r.v2 = y.x & FB_CNT_MASK;
    r.v3 = y.y & FB_CNT_MASK;
    fb_cswap(r.v0, r.v1);
    fb_cswap(r.v2, r.v3);
    fb_cswap(r.v0, r.v2);
    fb_cswap(r.v1, r.v3);
    fb_cswap(r.v1, r.v2);
    r.has_eps = r.anyp && r.v3 <= (1ULL << 26);  // cons_bases <= 1.0
    return r;
}

__device__ __forceinline__ bool fb_mec_locate(const MecArgs &a, uint64_t hapidx, InstDev &in, uint32_t &h, int &buf) {
    const int ii = fb_upper_seg(a.hap_prefix, a.n_inst, hapidx);
    in = a.inst[ii];
    if (a.only_active && !a.st[ii].active) return false;
    h = (uint32_t)(hapidx - a.hap_prefix[ii]);
    const int cur = a.st[ii].cur;
    buf = a.which == 0 ? cur : (a.which == 1 ? (cur ^ 1) : a.buf);
    return true;
}

// level 1: one warp per (instance, haplotype, chunk of 32 positions): exact integer sums + the epsilon mask
__global__ void __launch_bounds__(256) k_mec_chunks(MecArgs a, uint64_t n_haps) {
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= a.chunk_prefix[n_haps]) return;
    const uint32_t lane = fb_lane();
    const uint64_t hapidx = (uint64_t)fb_upper_seg(a.chunk_prefix, (int)n_haps, wid);
    InstDev in;
    uint32_t h;
    int buf;
    if (!fb_mec_locate(a, hapidx, in, h, buf)) return;
    const ulonglong2 *__restrict__ c2 =
        reinterpret_cast<const ulonglong2 *>(a.cnt[buf] + in.cnt_off + (uint64_t)h * in.ng * 64);
    const uint32_t p0 = (uint32_t)(wid - a.chunk_prefix[hapidx]) * 32u;
    const MecPos m = fb_mec_pos(c2, p0 + lane, in.ng * 16);
    const unsigned long long sb = fb_warp_sum_u64(m.v3), so = fb_warp_sum_u64(m.v0 + m.v1 + m.v2);
    const unsigned e = __ballot_sync(0xFFFFFFFFu, m.has_eps);
    if (lane == 0) {
        MecChunk c;
        c.bases = (long long)sb;
        c.others = (long long)so;
        c.eps = e;
        c._pad[0] = c._pad[1] = c._pad[2] = 0;
        a.chunks[wid] = c;
    }
}

// errors of one chunk, position by position: per position up to 3 dyadic items then an optional epsilon.  Runs of lanes
// between two epsilon items are added as one exact lump (prefix sums) whenever SeqSum proves that identical to
// item-by-item addition.
__device__ void fb_mec_chunk_detail(const ulonglong2 *__restrict__ c2, uint32_t npos, uint32_t p0, SeqSum &errors,
                                    double eps, int eps_safe) {
    const uint32_t lane = fb_lane();
    const MecPos m = fb_mec_pos(c2, p0 + lane, npos);
    const long long others = (long long)(m.v0 + m.v1 + m.v2);
    long long pre = others;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xFFFFFFFFu, pre, o);
        if ((int)lane >= o) pre += t;
    }
    const unsigned E = __ballot_sync(0xFFFFFFFFu, m.has_eps);
    int cur = 0;
    while (cur < 32) {
        const unsigned rest = E >> cur;
        const int e = rest ? cur + __ffs(rest) - 1 : 31;  // last lane of the run (its epsilon, if any, follows)
        const long long hi_sum = __shfl_sync(0xFFFFFFFFu, pre, e);
        const long long lo_sum = cur ? __shfl_sync(0xFFFFFFFFu, pre, cur - 1) : 0;
        if (!errors.add_dyadic_run(hi_sum - lo_sum)) {
            for (int l = cur; l <= e; ++l) {
                const long long o_l = __shfl_sync(0xFFFFFFFFu, others, l);
                if (errors.add_dyadic_run(o_l)) continue;
                errors.add_dyadic((long long)__shfl_sync(0xFFFFFFFFu, m.v0, l));
                errors.add_dyadic((long long)__shfl_sync(0xFFFFFFFFu, m.v1, l));
                errors.add_dyadic((long long)__shfl_sync(0xFFFFFFFFu, m.v2, l));
            }
        }
        if (!rest) break;
        errors.add_eps(eps, eps_safe);
        cur = e + 1;
    }
}

// single-level variant for small tables (a few chunks per haplotype: the summaries would not pay): one warp per
// (instance, haplotype) walks the chunks in order
__global__ void __launch_bounds__(256) k_mec_flat(MecArgs a) {
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= a.hap_prefix[a.n_inst]) return;
    InstDev in;
    uint32_t h;
    int buf;
    if (!fb_mec_locate(a, wid, in, h, buf)) return;
    const ulonglong2 *__restrict__ c2 =
        reinterpret_cast<const ulonglong2 *>(a.cnt[buf] + in.cnt_off + (uint64_t)h * in.ng * 64);
    const uint32_t npos = in.ng * 16;
    SeqSum bases, errors;
    bases.init();
    errors.init();
    for (uint32_t p0 = 0; p0 < npos; p0 += 32) {
        const MecPos m = fb_mec_pos(c2, p0 + fb_lane(), npos);
        if (!__ballot_sync(0xFFFFFFFFu, m.anyp)) continue;
        const long long tot = (long long)fb_warp_sum_u64(m.v3);
        if (!bases.add_dyadic_run(tot))
            for (int l = 0; l < 32; ++l) bases.add_dyadic((long long)__shfl_sync(0xFFFFFFFFu, m.v3, l));
        fb_mec_chunk_detail(c2, npos, p0, errors, a.eps, a.eps_safe);
    }
    if (fb_lane() == 0) {
        a.mec[buf][((uint64_t)in.mec_off + h) * 2 + 0] = bases.S;
        a.mec[buf][((uint64_t)in.mec_off + h) * 2 + 1] = errors.S;
    }
}

// level 2: one warp per (instance, haplotype) folds the chunk summaries in position order: a lane holds one chunk, runs
// of epsilon-free chunks are added as one exact lump, chunks with epsilon items (the thin ends of a table) or runs that
// SeqSum cannot prove exact are expanded position by position.
__global__ void __launch_bounds__(256) k_mec(MecArgs a) {
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= a.hap_prefix[a.n_inst]) return;
    const uint32_t lane = fb_lane();
    InstDev in;
    uint32_t h;
    int buf;
    if (!fb_mec_locate(a, wid, in, h, buf)) return;
    const ulonglong2 *__restrict__ c2 =
        reinterpret_cast<const ulonglong2 *>(a.cnt[buf] + in.cnt_off + (uint64_t)h * in.ng * 64);
    const uint32_t npos = in.ng * 16;
    const MecChunk *__restrict__ ch = a.chunks + a.chunk_prefix[wid];
    const uint32_t n_chunks = (uint32_t)(a.chunk_prefix[wid + 1] - a.chunk_prefix[wid]);
    SeqSum bases, errors;
    bases.init();
    errors.init();
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 32) {
        const uint32_t c = c0 + lane;
        long long sb = 0, so = 0;
        uint32_t em = 0;
        if (c < n_chunks) {
            const MecChunk k = ch[c];
            sb = k.bases;
            so = k.others;
            em = k.eps;
        }
        // bases: dyadic items only (one per position)
        {
            const long long tot = (long long)fb_warp_sum_u64((unsigned long long)sb);
            if (!bases.add_dyadic_run(tot)) {
                for (int l = 0; l < 32; ++l) {
                    const long long sb_l = __shfl_sync(0xFFFFFFFFu, sb, l);
                    if (bases.add_dyadic_run(sb_l)) continue;
                    const MecPos m = fb_mec_pos(c2, (c0 + l) * 32u + lane, npos);
                    for (int q = 0; q < 32; ++q) bases.add_dyadic((long long)__shfl_sync(0xFFFFFFFFu, m.v3, q));
                }
            }
        }
        // errors
        {
            long long pre = so;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xFFFFFFFFu, pre, o);
                if ((int)lane >= o) pre += t;
            }
            const unsigned X = __ballot_sync(0xFFFFFFFFu, em != 0);  // chunks that hold epsilon items
            // chunks whose only items are epsilons (a single read covers those positions: no non-consensus counts)
            // contribute popc(mask) consecutive `+= epsilon`; a run of such chunks is one fb_add_eps_n
            const unsigned Y = __ballot_sync(0xFFFFFFFFu, em != 0 && so == 0);
            uint32_t pe = (em != 0 && so == 0) ? (uint32_t)__popc(em) : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pe, o);
                if ((int)lane >= o) pe += t;
            }
            int cur = 0;
            while (cur < 32) {
                const unsigned rest = X >> cur;
                const int e = rest ? cur + __ffs(rest) - 1 : 32;  // next chunk to expand
                if (e > cur) {
                    const long long hi_sum = __shfl_sync(0xFFFFFFFFu, pre, e - 1);
                    const long long lo_sum = cur ? __shfl_sync(0xFFFFFFFFu, pre, cur - 1) : 0;
                    if (!errors.add_dyadic_run(hi_sum - lo_sum)) {
                        for (int l = cur; l < e; ++l) {
                            const long long so_l = __shfl_sync(0xFFFFFFFFu, so, l);
                            if (errors.add_dyadic_run(so_l)) continue;
                            fb_mec_chunk_detail(c2, npos, (c0 + l) * 32u, errors, a.eps, a.eps_safe);
                        }
                    }
                }
                if (e == 32) break;
                if ((Y >> e) & 1u) {
                    const unsigned inv = ~(Y >> e);
                    const int r = inv ? __ffs(inv) - 1 : 32 - e;  // consecutive epsilon-only chunks from e on
                    const uint32_t hi_n = __shfl_sync(0xFFFFFFFFu, pe, e + r - 1);
                    const uint32_t lo_n = e ? __shfl_sync(0xFFFFFFFFu, pe, e - 1) : 0u;
                    fb_seqsum_add_eps_n(errors, a.eps, a.eps_safe, (unsigned long long)(hi_n - lo_n));
                    cur = e + r;
                } else {
                    fb_mec_chunk_detail(c2, npos, (c0 + e) * 32u, errors, a.eps, a.eps_safe);
                    cur = e + 1;
                }
            }
        }
    }
    if (lane == 0) {
        a.mec[buf][((uint64_t)in.mec_off + h) * 2 + 0] = bases.S;
        a.mec[buf][((uint64_t)in.mec_off + h) * 2 + 1] = errors.S;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// reads per haplotype of a buffer (thread per instance-hap would be enough; one warp per instance keeps it simple)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_sizes(const InstDev *inst, InstState *st, int n_inst, const uint8_t *assign0, const uint8_t *assign1,
                        int which /*0 cur, 1 other*/) {
    const int ii = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ii >= n_inst) return;
    const uint32_t lane = fb_lane();
    const InstDev in = inst[ii];
    const int buf = which == 0 ? st[ii].cur : (st[ii].cur ^ 1);
    const uint8_t *as = (buf == 0 ? assign0 : assign1) + in.assign_off;
    uint32_t cnt[FB_MAXP];
#pragma unroll
    for (int h = 0; h < FB_MAXP; ++h) cnt[h] = 0;
    for (uint32_t r = lane; r < in.n_reads; r += 32) {
        const uint32_t hv = as[r];
#pragma unroll
        for (int h = 0; h < FB_MAXP; ++h) cnt[h] += (hv == (uint32_t)h);
    }
#pragma unroll
    for (int h = 0; h < FB_MAXP; ++h) {
        cnt[h] = fb_warp_sum_u32(cnt[h]);
        if (lane == 0) st[ii].sizes[buf][h] = cnt[h];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K4 move selection + application.  One CTA per instance.  Reproduces opt_iterate, local_clustering.rs:327-357:
// stable sort of the candidate moves by gain (descending; ties keep generation order = haplotype i asc, read asc,
// j asc under the canonical set order), number_of_moves = len/10 (len/3+1 when that is 0), then the sequential walk
// with the "already moved" / "source haplotype would become empty" guards and the `mv_num > number_of_moves` break.
// ---------------------------------------------------------------------------------------------------------------------
struct MoveRec {
    unsigned long long key;  // ~gain bits: ascending key == descending gain
    uint32_t gen;            // generation index (tie-break, ascending)
    uint32_t info;           // read_local | j << 24 | i << 28
};

struct SelectArgs {
    const InstDev *inst;
    InstState *st;
    int n_inst;
    const double *gain;
    uint8_t *assign[2];
    MoveRec *moves;                // scratch, [gain_off ... ) rounded to pow2 per instance by the host
    const uint64_t *moves_off;     // [n_inst] offset into moves
    const uint32_t *moves_cap;     // [n_inst] pow2 capacity
};

__device__ __forceinline__ bool fb_move_less(const MoveRec &x, const MoveRec &y) {
    return x.key < y.key || (x.key == y.key && x.gen < y.gen);
}

#define FB_SELECT_THREADS 256

__global__ void __launch_bounds__(FB_SELECT_THREADS) k_select(SelectArgs a) {
    const int ii = blockIdx.x;
    InstState &st = a.st[ii];
    if (!st.active) return;
    const InstDev in = a.inst[ii];
    const int cur = st.cur;
    const uint32_t P = in.ploidy;
    const uint8_t *__restrict__ as_cur = a.assign[cur] + in.assign_off;
    uint8_t *__restrict__ as_new = a.assign[cur ^ 1] + in.assign_off;
    const double *__restrict__ gain = a.gain + in.gain_off;
    MoveRec *mv = a.moves + a.moves_off[ii];
    const uint32_t cap = a.moves_cap[ii];
    const int t = threadIdx.x;
    __shared__ uint32_t s_scan[FB_SELECT_THREADS];
    __shared__ uint32_t s_base;
    if (t == 0) s_base = 0;
    __syncthreads();
    // 1. compaction in generation order: for i in 0..P, reads ascending, j ascending
    for (uint32_t i = 0; i < P; ++i) {
        for (uint32_t r0 = 0; r0 < in.n_reads; r0 += FB_SELECT_THREADS) {
            const uint32_t r = r0 + t;
            uint32_t cnt = 0;
            if (r < in.n_reads && as_cur[r] == i)
                for (uint32_t j = 0; j < P; ++j) cnt += gain[(uint64_t)r * P + j] > 0.0;
            s_scan[t] = cnt;
            __syncthreads();
            // inclusive scan (Hillis-Steele)
            for (int o = 1; o < FB_SELECT_THREADS; o <<= 1) {
                uint32_t v = t >= o ? s_scan[t - o] : 0;
                __syncthreads();
                s_scan[t] += v;
                __syncthreads();
            }
            uint32_t off = s_base + s_scan[t] - cnt;
            if (cnt) {
                for (uint32_t j = 0; j < P; ++j) {
                    double g = gain[(uint64_t)r * P + j];
                    if (g > 0.0) {
                        MoveRec m;
                        m.key = ~(unsigned long long)__double_as_longlong(g);
                        m.gen = off;
                        m.info = r | (j << 24) | (i << 28);
                        mv[off++] = m;
                    }
                }
            }
            __syncthreads();
            if (t == FB_SELECT_THREADS - 1) s_base += s_scan[t];
            __syncthreads();
        }
    }
    const uint32_t M = s_base;
    if (t == 0) {
        st.n_moves = (int)M;
    }
    // new_part = partition.clone()
    for (uint32_t r = t; r < in.n_reads; r += FB_SELECT_THREADS) as_new[r] = as_cur[r];
    if (M == 0) {
        if (t == 0)
            for (uint32_t h = 0; h < FB_MAXP; ++h) st.sizes[cur ^ 1][h] = st.sizes[cur][h];
        return;
    }
    // 2. bitonic sort of (key, gen) ascending over the next power of two >= M (padding sorts last)
    uint32_t n2 = 1;
    while (n2 < M) n2 <<= 1;
    if (n2 > cap) n2 = cap;  // host guarantees cap >= next_pow2(max moves)
    for (uint32_t x = M + t; x < n2; x += FB_SELECT_THREADS) {
        MoveRec m;
        m.key = ~0ULL;
        m.gen = 0xFFFFFFFFu;
        m.info = 0xFFFFFFFFu;
        mv[x] = m;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t x = t; x < n2; x += FB_SELECT_THREADS) {
                uint32_t y = x ^ j;
                if (y > x) {
                    MoveRec mx = mv[x], my = mv[y];
                    bool up = (x & k) == 0;
                    bool sw = up ? fb_move_less(my, mx) : fb_move_less(mx, my);
                    if (sw) {
                        mv[x] = my;
                        mv[y] = mx;
                    }
                }
            }
            __syncthreads();
        }
    }
    // 3. sequential application
    if (t == 0) {
        uint32_t sizes[FB_MAXP];
        for (uint32_t h = 0; h < FB_MAXP; ++h) sizes[h] = st.sizes[cur][h];
        uint32_t number_of_moves = M / 10;
        if (number_of_moves == 0 && M > 0) number_of_moves = M / 3 + 1;
        for (uint32_t mv_num = 0; mv_num < M; ++mv_num) {
            const uint32_t info = mv[mv_num].info;
            const uint32_t r = info & 0xFFFFFFu, j = (info >> 24) & 0xFu, i = info >> 28;
            if (as_new[r] != as_cur[r]) continue;  // moved_reads.contains(read): a moved read always changes haplotype
            if (sizes[i] == 1) continue;
            as_new[r] = (uint8_t)j;
            sizes[j] += 1;
            sizes[i] -= 1;
            if (mv_num > number_of_moves) break;
        }
        for (uint32_t h = 0; h < FB_MAXP; ++h) st.sizes[cur ^ 1][h] = sizes[h];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// accept / reject of optimize_clustering (local_clustering.rs:97-99, 105-127).  One thread per instance.
// ---------------------------------------------------------------------------------------------------------------------
struct AcceptArgs {
    const InstDev *inst;
    InstState *st;
    int n_inst;
    const double *mec[2];
    int init;  // 1: set prev_score from the current buffer (lines 97-99)
    uint32_t iter, max_iters;
    int *n_active;  // device counter of instances still iterating
};

__global__ void k_accept(AcceptArgs a) {
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= a.n_inst) return;
    InstState &st = a.st[ii];
    const InstDev in = a.inst[ii];
    if (a.init) {
        double s = 0.0;  // binom_vec.iter().map(|x| x.1).sum()
        for (uint32_t h = 0; h < in.ploidy; ++h) s += a.mec[st.cur][((uint64_t)in.mec_off + h) * 2 + 1];
        st.prev_score = s * -1.;
        st.n_hist += 1;
        return;
    }
    if (!st.active) return;
    const int nb = st.cur ^ 1;
    double s = 0.0;
    for (uint32_t h = 0; h < in.ploidy; ++h) s += a.mec[nb][((uint64_t)in.mec_off + h) * 2 + 1];
    const double new_score = s * -1.;
    st.new_score = new_score;
    st.n_opt_iterate += 1;
    st.n_hist += 1;
    if (new_score > st.prev_score) {
        st.prev_score = new_score;
        st.cur = nb;
        st.accepted += 1;
        if (a.iter + 1 >= a.max_iters)
            st.active = 0;
        else
            atomicAdd(a.n_active, 1);
    } else {
        st.active = 0;
    }
}
