// fb_graph.cuh — row f1 of SURVEY.md §8: update_hap_graph edge weights (graph_processing.rs:22-76).
//   k_planes_a4   consensus / tie planes of a node's hap_map for distance_read_haplo (utils_frags.rs:77-108)
//   k_edge_score  every read of a node of column i scored against every node of column i+1, unambiguity test,
//                 out_weights[hap_id_in] += 1
// Nodes are engine instances of ploidy 1 whose histogram is restricted to the node's snp_endpoints
// (types_structs.rs:169-180).
#pragma once
#include "fb_common.cuh"

// per group: x = consensus planes 0|1<<16, y = consensus planes 2|3<<16, z = tie planes 0|1<<16, w = tie planes 2|3<<16
//   consensus (utils_frags.rs:86-92): max_by_key over the allele keys present = the LAST maximum in ascending allele
//   order (canonical), even when every count is 0;  tie plane a: key a present and count[a] == max.
__global__ void k_planes_a4(const InstDev *inst, int n_inst, const uint64_t *group_prefix /*[n_inst+1]*/,
                            const uint64_t *cnt, uint4 *planes) {
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= group_prefix[n_inst]) return;
    const int ii = fb_upper_seg(group_prefix, n_inst, x);
    const InstDev in = inst[ii];
    const uint32_t g = (uint32_t)(x - group_prefix[ii]);
    const unsigned long long *c = reinterpret_cast<const unsigned long long *>(cnt + in.cnt_off) + (uint64_t)g * 64;
    uint32_t cons[4] = {0, 0, 0, 0}, tie[4] = {0, 0, 0, 0};
    for (int k = 0; k < 16; ++k) {
        unsigned long long w4[4];
        int best = -1;
        unsigned long long bv = 0;
        for (int a = 0; a < 4; ++a) {
            w4[a] = c[k * 4 + a];
            if (!(w4[a] & FB_PRESENT)) continue;
            const unsigned long long v = w4[a] & FB_CNT_MASK;
            if (best < 0 || v >= bv) {
                best = a;
                bv = v;
            }
        }
        if (best < 0) continue;
        cons[best] |= 1u << k;
        for (int a = 0; a < 4; ++a)
            if ((w4[a] & FB_PRESENT) && (w4[a] & FB_CNT_MASK) == bv) tie[a] |= 1u << k;
    }
    planes[x] = make_uint4(cons[0] | (cons[1] << 16), cons[2] | (cons[3] << 16), tie[0] | (tie[1] << 16),
                           tie[2] | (tie[3] << 16));
}

struct EdgeArgs {
    DFragsDev fr;
    const InstDev *inst;              // one per node
    const uint64_t *group_prefix;     // planes offset of node v = group_prefix[v]
    const uint4 *planes;
    const uint32_t *lut;
    // work items: one warp per (node1, read)
    uint64_t n_items;
    const uint32_t *item_node;        // node1 of the item
    const uint32_t *item_read;        // counter_id
    const uint32_t *next_first;       // [n_nodes] first node id of the next column (or 0xFFFFFFFF for the last column)
    const uint32_t *next_count;       // [n_nodes] nodes in the next column
    const uint64_t *node_ptr;         // node read lists (ascending counter_id) for the membership test
    const uint32_t *node_reads;
    const uint64_t *out_off;          // [n_nodes] offset of node1's row in out
    unsigned int *out;                // counts
};

__device__ __forceinline__ void fb_mux4(uint32_t al, uint32_t lo, uint32_t hi, uint32_t &sel, uint32_t &any) {
    const uint32_t a0 = al & 0xFFFFu, a1 = al >> 16;
    const uint32_t m0 = lo & 0xFFFFu, m1 = lo >> 16, m2 = hi & 0xFFFFu, m3 = hi >> 16;
    const uint32_t t01 = (a0 & m1) | (~a0 & m0);
    const uint32_t t23 = (a0 & m3) | (~a0 & m2);
    sel = ((a1 & t23) | (~a1 & t01)) & 0xFFFFu;
    any = m0 | m1 | m2 | m3;
}

__global__ void __launch_bounds__(256) k_edge_score(EdgeArgs a) {
    __shared__ uint32_t lut_s[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut_s[i] = a.lut[i];
    __syncthreads();
    const uint64_t item = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= a.n_items) return;
    const uint32_t lane = fb_lane();
    const uint32_t v1 = a.item_node[item], rid = a.item_read[item];
    const uint32_t b0 = a.next_first[v1], nb = a.next_count[v1];
    if (b0 == 0xFFFFFFFFu || nb == 0) return;
    const uint32_t g0 = a.fr.gptr[rid], g1 = g0 + a.fr.gnum[rid], gs = a.fr.gstart[rid];
    unsigned long long best = ~0ULL, second = ~0ULL;  // two smallest rounded diffs (read_to_hap_sim.sort(), :43-45)
    uint32_t hap_id_in = 0xFFFFFFFFu;
    for (uint32_t l = 0; l < nb; ++l) {
        const uint32_t v2 = b0 + l;
        const InstDev in = a.inst[v2];
        const uint4 *pl = a.planes + a.group_prefix[v2];
        const long long ag0 = in.ag0, ag1 = (long long)in.ag0 + in.ng;
        unsigned long long diff = 0;
        for (uint32_t g = g0 + lane; g < g1; g += 32) {
            const long long ag = (long long)gs + (g - g0);  // absolute group of this read group
            if (ag < ag0 || ag >= ag1) continue;            // node2's table has no key there: positions are skipped (:81-83)
            const uint4 q = a.fr.qual[g];
            const uint32_t al = a.fr.allele[g], pr = a.fr.present[g];
            const uint4 p4 = pl[ag - ag0];
            uint32_t tiesel, anyc, conssel, anyk;
            fb_mux4(al, p4.x, p4.y, conssel, anyk);  // anyk: a key exists at the position
            fb_mux4(al, p4.z, p4.w, tiesel, anyc);
            (void)conssel;
            (void)anyc;
            const uint32_t db = pr & anyk & ~tiesel & 0xFFFFu;  // neither the consensus nor tied with it (:93-103)
            uint32_t w[16];
            fb_group_weights_raw(q, lut_s, w);
            diff += fb_masked_sum(w, db);
        }
        diff = fb_warp_sum_u64(diff);
        const unsigned long long dr = (diff + (1ULL << 25)) >> 26;  // diff.round() as usize (:107)
        if (dr < best) {
            second = best;
            best = dr;
        } else if (dr < second) {
            second = dr;
        }
        // hap_node2.frag_set.contains(read) (:37-39): the last such l wins
        const uint32_t *rd = a.node_reads + a.node_ptr[v2];
        int lo = 0, hi = (int)(a.node_ptr[v2 + 1] - a.node_ptr[v2]);
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rd[mid] < rid)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo < (int)(a.node_ptr[v2 + 1] - a.node_ptr[v2]) && rd[lo] == rid) hap_id_in = l;
    }
    if (lane == 0 && hap_id_in != 0xFFFFFFFFu) {
        const bool unambiguous = nb > 1 ? (best != second) : true;  // :44-56
        if (unambiguous) atomicAdd(a.out + a.out_off[v1] + hap_id_in, 1u);
    }
}
