// fb_replay.cuh — exact replay of the reference's left-to-right f64 `diff` accumulation (utils_frags.rs:33-72) by a team
// of lanes; shared by the scoring sweep (fb_kernels.cuh) and the beam search (fb_beam.cuh).
#pragma once
#include "fb_common.cuh"

// Exact left-to-right f64 sum of the diff/epsilon items of one read vs one haplotype table (canonical position order),
// i.e. the `diff` accumulator of utils_frags.rs:33-72.  Warp-cooperative; every lane returns the same value.
// Planes of groups beyond `hi` are empty.  Per chunk of 32 groups the lanes are split by ballot into runs of
// epsilon-free lanes (added as ONE exact integer lump when SeqSum proves that identical to item-by-item addition)
// and lanes holding epsilon items (replayed item by item).
// CG: the planes are read through L2 only (__ldcg), for tables that other CTAs of the running kernel write.
template <int L, bool CG = false>
__device__ double fb_replay_diff_t(const DFragsDev &fr, uint32_t g0, uint32_t g1, const uint2 *__restrict__ mh,
                                   uint32_t lg0, int hi, const uint32_t *lut, double eps, uint32_t *wscratch /*16 per team*/) {
    const uint32_t lane = fb_lane() % L;               // lane inside the team
    const uint32_t tmask = fb_team_mask<L>();          // the team's lanes
    const uint32_t tshift = (fb_lane() / L) * L;
    const uint32_t lowmask = L == 32 ? 0xFFFFFFFFu : ((1u << (L & 31)) - 1u);
    SeqSum ss;
    ss.init();
    for (uint32_t base = g0; base < g1; base += L) {
        uint32_t g = base + lane;
        bool valid = g < g1;
        uint32_t w[16];
        uint32_t diffbits = 0, emptybits = 0;
        if (valid) {
            uint4 q = fr.qual[g];
            uint32_t al = fr.allele[g];
            uint32_t pr = fr.present[g];
            fb_group_weights(q, pr, lut, w);
            const uint32_t lg = lg0 + (g - g0);
            uint2 m = ((int)lg <= hi) ? (CG ? __ldcg(mh + lg) : mh[lg]) : make_uint2(0u, 0u);
            uint32_t same, ne;
            fb_group_masks(al, m, same, ne);
            diffbits = pr & ne & ~same;
            emptybits = pr & ~ne & 0xFFFFu;
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) w[k] = 0;
        }
        const long long Wl = (long long)fb_masked_sum(w, diffbits);
        const unsigned E = (__ballot_sync(tmask, emptybits != 0) >> tshift) & lowmask;
        // inclusive prefix sums of the per-lane dyadic sums
        long long pre = Wl;
#pragma unroll
        for (int o = 1; o < L; o <<= 1) {
            long long v = __shfl_up_sync(tmask, pre, o, L);
            if ((int)lane >= o) pre += v;
        }
        int cur = 0;
        while (cur < L) {
            const unsigned rest = E >> cur;
            const int e = rest ? cur + __ffs(rest) - 1 : L;  // next lane holding epsilon items
            if (e > cur) {
                const long long hi_sum = __shfl_sync(tmask, pre, e - 1, L);
                const long long lo_sum = cur ? __shfl_sync(tmask, pre, cur - 1, L) : 0;
                if (!ss.add_dyadic_run(hi_sum - lo_sum)) {
                    for (int l = cur; l < e; ++l) {
                        const long long Wl_l = __shfl_sync(tmask, Wl, l, L);
                        if (ss.add_dyadic_run(Wl_l)) continue;
                        const uint32_t db = __shfl_sync(tmask, diffbits, l, L);
                        __syncwarp(tmask);
                        if ((int)lane == l) {
#pragma unroll
                            for (int k = 0; k < 16; ++k) wscratch[k] = w[k];
                        }
                        __syncwarp(tmask);
                        uint32_t bits = db;
                        while (bits) {
                            const int k = __ffs(bits) - 1;
                            bits &= bits - 1;
                            ss.add_dyadic((long long)wscratch[k]);
                        }
                    }
                }
            }
            if (e == L) break;
            const uint32_t eb = __shfl_sync(tmask, emptybits, e, L);
            const uint32_t db = __shfl_sync(tmask, diffbits, e, L);
            if (db == 0) {
                ss.S = fb_add_eps_n(ss.S, eps, (unsigned long long)__popc(eb));  // an all-epsilon lane needs no weights
                ss.tail = 1;
            } else {
                __syncwarp(tmask);
                if ((int)lane == e) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) wscratch[k] = w[k];
                }
                __syncwarp(tmask);
                uint32_t bits = eb | db;
                while (bits) {
                    const int k = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if ((eb >> k) & 1u)
                        ss.add_eps(eps, 0);
                    else
                        ss.add_dyadic((long long)wscratch[k]);
                }
            }
            cur = e + 1;
        }
    }
    return ss.S;
}
__device__ __forceinline__ double fb_replay_diff(const DFragsDev &fr, uint32_t g0, uint32_t g1, const uint2 *__restrict__ mh,
                                                 uint32_t lg0, int hi, const uint32_t *lut, double eps, uint32_t *wscratch) {
    return fb_replay_diff_t<32>(fr, g0, g1, mh, lg0, hi, lut, eps, wscratch);
}

