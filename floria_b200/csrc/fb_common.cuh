// fb_common.cuh — device-side data model of the B200 hot path.
//
// HBM layout of a contig's reads (replaces Frag::{seq_dict, qual_dict, positions}, types_structs.rs:68-85):
//   reads are rows of a banded read x SNP matrix cut into GROUPS of 16 consecutive SNP positions aligned to the
//   absolute position grid (position0 = pos-1; group = position0 >> 4), so a read's group k lines up with every
//   haplotype table's group (gstart + k).  Per group (22 B = 1.375 B per stored cell):
//     qual    uint4    16 x 8-bit phred bytes (cell c in byte c)
//     allele  uint32   2-bit alleles, planar: bit c = allele bit 0 of cell c, bit 16+c = allele bit 1
//     present uint16   bit c = the read has a cell at that position
//   gptr[r] is the first group of read r (aligned to 8 groups so that every plane of a read starts on a 16-byte
//   boundary: the 1-D TMA bulk copies of the sweep need that), gnum[r] its number of groups, gstart[r] the absolute
//   index of its first group.
//
// Haplotype tables (replace Haplotype = FxHashMap<pos, FxHashMap<allele, f64>>, types_structs.rs:15):
//   counts  uint64 [ploidy][n_pos][4]   weight sums in units of 2^-26 (exact), bit 62 = allele key present
//   masks   uint2  [ploidy][n_groups]   per group 4 planes x 16 bit: plane a bit c = "allele a holds the maximum
//                                       (non-zero) count at position c" (x = plane0 | plane1<<16, y = plane2 | plane3<<16)
//   A cell with allele a at a position scores `same` iff plane a has its bit (utils_frags.rs:60-69: equal to the
//   consensus or tied with it), `empty` (+epsilon) iff no plane has it (utils_frags.rs:36-48), else `diff`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fb_seq.h"

#define FB_GROUP 16
#define FB_MAXP 8  // largest ploidy the kernels are instantiated for

struct DFragsDev {
    uint64_t n_reads;
    const uint32_t *first, *last, *nnz, *gstart, *gptr, *gnum;
    const uint4 *qual;
    const uint32_t *allele;
    const uint16_t *present;
};

// one (block, ploidy) unit of work
struct InstDev {
    uint32_t block, ploidy, n_reads, ng;  // ng = groups spanned by the block's reads
    uint32_t ag0;                         // absolute index of the block's first group
    uint32_t read_off;                    // into blk_reads / blk_rinfo
    uint32_t mec_off;                     // into mec arrays: [ploidy] pairs
    uint32_t _pad;
    uint32_t flt_lo, flt_hi;              // block-local position0 range counted by k_hist (HapNode::new endpoints)
    uint64_t assign_off;  // into assign_cur/new; prefix over instances of n_reads
    uint64_t gain_off;    // into gain slots [n_reads][ploidy]
    uint64_t cnt_off;     // into count buffers, words: [ploidy][ng*16][4]
    uint64_t mask_off;    // into mask buffers, uint2: [ploidy][ng]
};

// per (block, read): where the read's groups sit relative to the block's table
struct RInfo {
    uint32_t rid;    // counter_id
    uint32_t gbase;  // global group index of the read's group that aligns with block-local group 0 (wraps)
    uint32_t lg0;    // block-local index of the read's first group
    uint32_t lg1;    // one past the last
};

// per (block, read) extras for the beam search
struct RExtra {
    uint32_t first0;  // block-local position0 of the read's first SNP
    uint32_t nnz;     // stored cells of the read
};

struct InstState {
    int cur;       // which of the two count/mask/assign buffers holds the accepted partition
    int active;    // still iterating optimize_clustering
    int n_moves;   // candidate moves produced by the last sweep
    uint32_t accepted, n_opt_iterate, n_hist;
    double prev_score, new_score;
    uint32_t sizes[2][FB_MAXP];  // reads per haplotype for both buffers
};

__device__ __forceinline__ uint32_t fb_lane() { return threadIdx.x & 31; }

// ---- 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers --------------------------------------------------
__device__ __forceinline__ uint32_t fb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fb_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fb_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fb_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fb_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     fb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(fb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fb_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(fb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// isSame / nonEmpty bit words of one group for one haplotype (16 low bits significant)
__device__ __forceinline__ void fb_group_masks(uint32_t al, uint2 m, uint32_t &same, uint32_t &nonempty) {
    uint32_t a0 = al & 0xFFFFu, a1 = al >> 16;
    uint32_t m0 = m.x & 0xFFFFu, m1 = m.x >> 16, m2 = m.y & 0xFFFFu, m3 = m.y >> 16;
    uint32_t t01 = (a0 & m1) | (~a0 & m0);
    uint32_t t23 = (a0 & m3) | (~a0 & m2);
    same = ((a1 & t23) | (~a1 & t01)) & 0xFFFFu;
    nonempty = m0 | m1 | m2 | m3;
}

// weights (units of 2^-26) of the 16 cells of a group; absent cells get 0
__device__ __forceinline__ void fb_group_weights(uint4 q, uint32_t pres, const uint32_t *__restrict__ lut,
                                                 uint32_t (&w)[16]) {
    const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        uint32_t b = (qq[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
        uint32_t v = lut[b];
        w[k] = ((pres >> k) & 1u) ? v : 0u;
    }
}

__device__ __forceinline__ uint32_t fb_masked_sum(const uint32_t (&w)[16], uint32_t bits) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (bits & (1u << k)) s += w[k];
    return s;
}

// acc += w when (bits & m) != 0, as `and` + `setp` + a predicated add.  With the 16 tests of one bit word issued back to
// back ptxas turns the tests into two R2P (7 predicates each) + two LOP3.P instead of 16 LOP3.P, and picks
// IMAD.IADD / IADD3 for the adds to balance the FMA and ALU pipes (profiles/: the scoring kernels are bound by the ALU
// pipe and by issue slots, not by HBM).
__device__ __forceinline__ void fb_padd(uint32_t &acc, uint32_t bits, uint32_t m, uint32_t w) {
    asm("{\n.reg .pred p;\n.reg .u32 t;\nand.b32 t, %1, %2;\nsetp.ne.u32 p, t, 0;\n@p add.u32 %0, %0, %3;\n}"
        : "+r"(acc)
        : "r"(bits), "r"(m), "r"(w));
}
// sum of w[k] over the set bits of `bits` (16 low bits), haplotype-outer / cell-inner order (see fb_padd)
__device__ __forceinline__ uint32_t fb_masked_sum_p(const uint32_t (&w)[16], uint32_t bits) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) fb_padd(s, bits, 1u << k, w[k]);
    return s;
}

// LUT weights of the 16 cells of a group WITHOUT zeroing absent cells (callers mask their bit sets with `present`)
__device__ __forceinline__ void fb_group_weights_raw(uint4 q, const uint32_t *__restrict__ lut, uint32_t (&w)[16]) {
    const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = lut[__byte_perm(qq[k >> 2], 0, 0x4440 | (k & 3))];
}

__device__ __forceinline__ unsigned long long fb_warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t fb_warp_sum_u32(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// sub-warp teams of L consecutive lanes (L = 32, 8 or 2): short reads get a team each instead of a whole warp.  Every
// shuffle of a team names only the team's lanes in its mask, so the teams of a warp may diverge freely.
template <int L>
__device__ __forceinline__ uint32_t fb_team_mask() {
    return L == 32 ? 0xFFFFFFFFu : (((1u << (L & 31)) - 1u) << ((fb_lane() / L) * L));
}
template <int L>
__device__ __forceinline__ unsigned long long fb_team_sum_u64(unsigned long long v, uint32_t mask) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <int L>
__device__ __forceinline__ uint32_t fb_team_sum_u32(uint32_t v, uint32_t mask) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

// first index i in [0,n] with prefix[i+1] > x, for an ascending prefix array of n+1 entries (prefix[0] == 0)
__device__ __forceinline__ int fb_upper_seg(const uint64_t *__restrict__ prefix, int n, uint64_t x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid + 1] > x)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}
