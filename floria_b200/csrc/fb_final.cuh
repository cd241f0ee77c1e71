// fb_final.cuh — rows a14/a15 of SURVEY.md §8: final read refinement and HAPQ.
//   k_final_assign   part_block_manip.rs:195-222  online greedy re-insertion of every read into its single best haploset
//   k_errors_cov     utils_frags.rs:596-657       get_errors_cov_from_frags on the unweighted per-part histogram
//   k_hap_distance   utils_frags.rs:659-700       distance_between_haplotypes for overlapping part pairs
// Parts are engine instances of ploidy 1 (one dense count table + is-max planes per part).
#pragma once
#include "fb_beam.cuh"
#include "fb_common.cuh"

#define FB_FINAL_THREADS 256
#define FB_FINAL_WARPS (FB_FINAL_THREADS / 32)
#define FB_FINAL_MAXCAND 64

struct FinalArgs {
    DFragsDev fr;
    const InstDev *inst;      // one per part
    uint64_t *cnt;            // zero-initialised tables (all reads removed, part_block_manip.rs:195-200)
    uint2 *masks;
    const uint32_t *lut;
    uint32_t n_active;        // reads that belong to at least one part
    const uint32_t *read_ids; // ascending counter_id (canonical FxHashMap<&Frag,_> order)
    const uint64_t *cand_ptr; // [n_active+1]
    const uint32_t *cand;     // part ids, ascending within a read (canonical FxHashSet<usize> order)
    uint32_t *chosen;         // [n_active] out
    double eps;
    int eps_safe;
};

__global__ void __launch_bounds__(FB_FINAL_THREADS) k_final_assign(FinalArgs a) {
    __shared__ uint32_t lut_s[256];
    __shared__ uint32_t wscr[FB_FINAL_WARPS][16];
    __shared__ double s_same[FB_FINAL_MAXCAND], s_diff[FB_FINAL_MAXCAND];
    __shared__ uint32_t s_best;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += FB_FINAL_THREADS) lut_s[i] = a.lut[i];
    __syncthreads();
    const uint32_t *__restrict__ qual32 = reinterpret_cast<const uint32_t *>(a.fr.qual);
    for (uint32_t x = 0; x < a.n_active; ++x) {
        const uint32_t rid = a.read_ids[x];
        const uint64_t c0 = a.cand_ptr[x], c1 = a.cand_ptr[x + 1];
        const uint32_t nc = (uint32_t)(c1 - c0);
        const uint32_t g0 = a.fr.gptr[rid], g1 = g0 + a.fr.gnum[rid];
        const uint32_t gs = a.fr.gstart[rid];
        uint32_t best = a.cand[c0];
        if (nc > 1) {
            // distance_read_haplo_epsilon_empty against every candidate haploset (part_block_manip.rs:205-212)
            for (uint32_t k = warp; k < nc; k += FB_FINAL_WARPS) {
                const InstDev in = a.inst[a.cand[c0 + k]];
                const uint2 *mk = a.masks + in.mask_off;
                const int hi = (int)in.ng - 1;
                const uint32_t lg0 = gs - in.ag0;  // the part's extent covers all of its reads
                unsigned long long total = 0, same = 0, emptyw = 0;
                uint32_t ne_cnt = 0;
                for (uint32_t g = g0 + lane; g < g1; g += 32) {
                    uint4 q = a.fr.qual[g];
                    uint32_t al = a.fr.allele[g];
                    uint32_t pr = a.fr.present[g];
                    uint32_t w[16];
                    fb_group_weights(q, pr, lut_s, w);
                    uint32_t t = 0;
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) t += w[kk];
                    total += t;
                    uint2 m = mk[lg0 + (g - g0)];
                    uint32_t sb, ne;
                    fb_group_masks(al, m, sb, ne);
                    same += fb_masked_sum(w, sb);
                    uint32_t eb = pr & ~ne & 0xFFFFu;
                    if (eb) {
                        emptyw += fb_masked_sum(w, eb);
                        ne_cnt += __popc(eb);
                    }
                }
                total = fb_warp_sum_u64(total);
                same = fb_warp_sum_u64(same);
                emptyw = fb_warp_sum_u64(emptyw);
                ne_cnt = fb_warp_sum_u32(ne_cnt);
                const long long diff_q = (long long)(total - same - emptyw);
                double diff_f;
                if (ne_cnt == 0)
                    diff_f = fb_q26_to_f64(diff_q);
                else if (a.eps_safe)
                    diff_f = fb_q26_to_f64(diff_q + (long long)ne_cnt * (long long)(a.eps * FB_Q26));
                else
                    diff_f = fb_replay_diff(a.fr, g0, g1, mk, lg0, hi, lut_s, a.eps, wscr[warp]);
                if (lane == 0) {
                    s_same[k] = fb_q26_to_f64((long long)same);
                    s_diff[k] = diff_f;
                }
            }
            __syncthreads();
            if (tid == 0) {
                // min_by over (diff + 1., id, same), lexicographic, first minimum wins (part_block_manip.rs:214-218)
                uint32_t bk = 0;
                for (uint32_t k = 1; k < nc; ++k) {
                    const double d1 = s_diff[k] + 1., db = s_diff[bk] + 1.;
                    bool less;
                    if (d1 != db)
                        less = d1 < db;
                    else if (a.cand[c0 + k] != a.cand[c0 + bk])
                        less = a.cand[c0 + k] < a.cand[c0 + bk];
                    else
                        less = s_same[k] < s_same[bk];
                    if (less) bk = k;
                }
                s_best = a.cand[c0 + bk];
            }
            __syncthreads();
            best = s_best;
        }
        if (tid == 0) a.chosen[x] = best;
        // add_read_to_block (utils_frags.rs:465-474) on the dense table, and refresh the is-max planes it touches
        {
            const InstDev in = a.inst[best];
            unsigned long long *cnt = reinterpret_cast<unsigned long long *>(a.cnt + in.cnt_off);
            uint2 *mk = a.masks + in.mask_off;
            const uint32_t lg0 = gs - in.ag0;
            const int total = (int)(g1 - g0) * 4;
            for (int base = 0; base < total; base += FB_FINAL_THREADS) {
                const int idx = base + tid;
                const bool act = idx < total;
                uint32_t pl[4] = {0, 0, 0, 0};
                uint32_t lg = 0, sub = 0;
                if (act) {
                    const uint32_t g = g0 + (uint32_t)(idx >> 2);
                    sub = (uint32_t)idx & 3u;
                    lg = lg0 + (g - g0);
                    unsigned long long wv[16];
                    ulonglong2 *p = reinterpret_cast<ulonglong2 *>(cnt + ((uint64_t)lg * 16 + sub * 4) * 4);
#pragma unroll
                    for (int y = 0; y < 8; ++y) {
                        ulonglong2 v = p[y];
                        wv[2 * y] = v.x;
                        wv[2 * y + 1] = v.y;
                    }
                    const uint32_t q = qual32[(uint64_t)g * 4 + sub];
                    const uint32_t al = a.fr.allele[g], pr = a.fr.present[g];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t c = sub * 4 + k;
                        if ((pr >> c) & 1u) {
                            const uint32_t av = ((al >> c) & 1u) | (((al >> (16 + c)) & 1u) << 1);
                            const unsigned long long w = lut_s[(q >> (8 * k)) & 0xFFu];
#pragma unroll
                            for (int aa = 0; aa < 4; ++aa)
                                if ((uint32_t)aa == av) wv[k * 4 + aa] = (wv[k * 4 + aa] + w) | FB_PRESENT;
                        }
                    }
#pragma unroll
                    for (int y = 0; y < 8; ++y) p[y] = make_ulonglong2(wv[2 * y], wv[2 * y + 1]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        unsigned long long mx = 0;
#pragma unroll
                        for (int aa = 0; aa < 4; ++aa) {
                            unsigned long long v = wv[k * 4 + aa] & FB_CNT_MASK;
                            mx = v > mx ? v : mx;
                        }
                        if (mx > 0) {
#pragma unroll
                            for (int aa = 0; aa < 4; ++aa)
                                if ((wv[k * 4 + aa] & FB_CNT_MASK) == mx) pl[aa] |= 1u << (sub * 4 + k);
                        }
                    }
                }
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) {
                    pl[aa] |= __shfl_xor_sync(0xFFFFFFFFu, pl[aa], 1);
                    pl[aa] |= __shfl_xor_sync(0xFFFFFFFFu, pl[aa], 2);
                }
                if (act && sub == 0) mk[lg] = make_uint2(pl[0] | (pl[1] << 16), pl[2] | (pl[3] << 16));
            }
        }
        __syncthreads();
        __threadfence_block();
    }
}

// ---- get_errors_cov_from_frags (utils_frags.rs:596-657) on the UNWEIGHTED table of each part ---------------------------
// One warp per part.  All quantities are integer read counts, so the f64 sums of the reference are exact and order free.
// out[part] = (total_support, errors) as integers.
struct ErrCovArgs {
    const InstDev *inst;
    int n_parts;
    const uint64_t *cnt;  // no-phred tables (units of 2^-26: one read = 2^26)
    const uint32_t *range_lo, *range_hi;  // 1-based inclusive SNP ranges
    long long *out;       // [n_parts][2]
};

__global__ void k_errors_cov(ErrCovArgs a) {
    const int part = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (part >= a.n_parts) return;
    const uint32_t lane = fb_lane();
    const InstDev in = a.inst[part];
    const unsigned long long *cnt = reinterpret_cast<const unsigned long long *>(a.cnt + in.cnt_off);
    const long long base0 = (long long)in.ag0 * 16;  // position0 of table row 0
    const long long lo0 = (long long)a.range_lo[part] - 1, hi0 = (long long)a.range_hi[part] - 1;
    long long p_begin = lo0 - base0, p_end = hi0 - base0 + 1;
    if (p_begin < 0) p_begin = 0;
    if (p_end > (long long)in.ng * 16) p_end = (long long)in.ng * 16;
    long long support = 0, errors = 0;
    for (long long p = p_begin + lane; p < p_end; p += 32) {
        long long snp_support = 0, max_count_pos = 0;
#pragma unroll
        for (int al = 0; al < 4; ++al) {
            const unsigned long long w = cnt[(uint64_t)p * 4 + al];
            if (!(w & FB_PRESENT)) continue;
            const long long c = (long long)((w & FB_CNT_MASK) >> 26);
            if (c > snp_support) max_count_pos = c;  // sic: compared with the running SUM (utils_frags.rs:620-623)
            snp_support += c;
        }
        support += snp_support;
        errors += snp_support - max_count_pos;
    }
    support = (long long)fb_warp_sum_u64((unsigned long long)support);
    errors = (long long)fb_warp_sum_u64((unsigned long long)errors);
    if (lane == 0) {
        a.out[part * 2 + 0] = support;
        a.out[part * 2 + 1] = errors;
    }
}

// ---- distance_between_haplotypes (utils_frags.rs:659-700) with range = (MIN, MAX): every shared position counts --------
// One warp per (i, j) pair.  consensus = max_by_key = the LAST maximum in ascending-allele (canonical) order among the
// allele keys present.
struct HapDistArgs {
    const InstDev *inst;
    const uint64_t *cnt;  // phred tables
    const uint32_t *pair_i, *pair_j;
    int n_pairs;
    long long *out;  // [n_pairs][2] = (same, diff)
};

__device__ __forceinline__ int fb_consensus_last_max(const unsigned long long *w4) {
    int best = -1;
    unsigned long long bv = 0;
#pragma unroll
    for (int al = 0; al < 4; ++al) {
        if (!(w4[al] & FB_PRESENT)) continue;
        const unsigned long long v = w4[al] & FB_CNT_MASK;
        if (best < 0 || v >= bv) {
            best = al;
            bv = v;
        }
    }
    return best;
}

__global__ void k_hap_distance(HapDistArgs a) {
    const int pr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pr >= a.n_pairs) return;
    const uint32_t lane = fb_lane();
    const InstDev A = a.inst[a.pair_i[pr]], B = a.inst[a.pair_j[pr]];
    const unsigned long long *ca = reinterpret_cast<const unsigned long long *>(a.cnt + A.cnt_off);
    const unsigned long long *cb = reinterpret_cast<const unsigned long long *>(a.cnt + B.cnt_off);
    const long long a0 = (long long)A.ag0 * 16, a1 = a0 + (long long)A.ng * 16;
    const long long b0 = (long long)B.ag0 * 16, b1 = b0 + (long long)B.ng * 16;
    const long long lo = a0 > b0 ? a0 : b0, hi = a1 < b1 ? a1 : b1;
    long long same = 0, diff = 0;
    for (long long p = lo + lane; p < hi; p += 32) {
        unsigned long long wa[4], wb[4];
#pragma unroll
        for (int al = 0; al < 4; ++al) {
            wa[al] = ca[(uint64_t)(p - a0) * 4 + al];
            wb[al] = cb[(uint64_t)(p - b0) * 4 + al];
        }
        const int c1 = fb_consensus_last_max(wa), c2 = fb_consensus_last_max(wb);
        if (c1 < 0 || c2 < 0) continue;  // position key absent in one of the haplotypes
        if (c1 == c2)
            same += 1;
        else
            diff += 1;
    }
    same = (long long)fb_warp_sum_u64((unsigned long long)same);
    diff = (long long)fb_warp_sum_u64((unsigned long long)diff);
    if (lane == 0) {
        a.out[pr * 2 + 0] = same;
        a.out[pr * 2 + 1] = diff;
    }
}
