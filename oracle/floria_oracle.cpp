// floria_oracle.cpp — CPU ORACLE for the floria hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C++17 restatement of the reference's algorithm (bluenote-1577/floria @ ee27af85, Rust), function by
// function, with the same data model (nested maps position -> allele -> f64 count, deep clones per beam child,
// std BinaryHeap semantics).  Every function cites the reference lines it follows (paths relative to
// /root/reference).  Nothing under floria_b200/ may call into this file; only tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs load the shared object built from it.
//
// PARITY UNPINNED (see floria_oracle.h): the reference has no tests / golden vectors for this path.
//
// Iteration-order model ("order_model = 0", SURVEY.md §7 hard part 2): the reference iterates FxHashSet /
// FxHashMap containers in hashbrown bucket order, which is platform dependent and cannot be validated here
// (fxhash/hashbrown sources are not under /root/reference).  This oracle pins the CANONICAL order instead:
// sets of reads iterate in ascending counter_id, position maps in ascending SNP position, allele maps in
// ascending allele.  That is a declared deviation from the real binary only where the reference's result
// depends on hash iteration order (exact ties, and ulp-level f64 summation order when epsilon is not dyadic).
//
// Third-party semantics restated from their published algorithms (crates not vendored in /root/reference):
//   std::collections::BinaryHeap (Rust std): push = sift_up stopping on `<=`; pop = swap-remove + sift_down_to_bottom
//   + sift_up; into_sorted_vec = in-place heapsort with sift_down_range.  ordered-float 2.10.1: total order on f64.
//   rust-lapper 1.1.0: Lapper::new sorts by (start, stop); find(start, stop) yields iv.start < stop && iv.stop > start.

#include "floria_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace orc {

typedef uint32_t SnpPosition;  // types_structs.rs:12
typedef uint8_t Genotype;      // types_structs.rs:13
static const Genotype GAP_CHAR = 9;  // types_structs.rs:16

// constants.rs:3-22
static const double MIN_SHARED_READS_UNAMBIG = 2.;
static const double HAPQ_CONSTANT = 40.;
static const double DIST_COV_CUTOFF = 0.5;
static const bool SEPARATE_BROKEN_HAPLOGROUPS = true;

static thread_local std::string g_err;

// ---- data model (types_structs.rs:68-85, 11-16, 253-256) ----------------------------------------------------
struct Frag {
    size_t counter_id;
    SnpPosition first_position, last_position;
    // `positions` (FxHashSet) / seq_dict / qual_dict keys, in canonical (ascending) iteration order.
    std::vector<SnpPosition> positions;
    std::vector<Genotype> seq;
    std::vector<uint8_t> qual;
};

// FxHashMap<Genotype, GenotypeCount>, iterated in ascending allele order.
struct AlleleMap {
    std::vector<std::pair<Genotype, double>> v;
    const double *get(Genotype a) const {
        for (auto &kv : v)
            if (kv.first == a) return &kv.second;
        return nullptr;
    }
    double &entry_or_zero(Genotype a) {  // .entry(a).or_insert(OrderedFloat(0.))
        size_t i = 0;
        for (; i < v.size(); ++i) {
            if (v[i].first == a) return v[i].second;
            if (v[i].first > a) break;
        }
        v.insert(v.begin() + i, std::make_pair(a, 0.0));
        return v[i].second;
    }
    void remove(Genotype a) {
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].first == a) {
                v.erase(v.begin() + i);
                return;
            }
    }
    bool empty() const { return v.empty(); }
    bool operator==(const AlleleMap &o) const { return v == o.v; }
    bool operator!=(const AlleleMap &o) const { return !(v == o.v); }
    // .iter().max_by_key(|entry| entry.1): the LAST maximal element in iteration order.
    size_t max_by_count() const {
        size_t best = 0;
        for (size_t i = 1; i < v.size(); ++i)
            if (v[i].second >= v[best].second) best = i;
        return best;
    }
};
typedef std::map<SnpPosition, AlleleMap> Haplotype;  // types_structs.rs:15
struct HapBlock {                                    // types_structs.rs:253-256
    std::vector<Haplotype> blocks;
    bool operator==(const HapBlock &o) const { return blocks == o.blocks; }
};

struct FragIdLess {
    bool operator()(const Frag *a, const Frag *b) const { return a->counter_id < b->counter_id; }
};
typedef std::set<const Frag *, FragIdLess> FragSet;  // FxHashSet<&Frag>, canonical order = ascending counter_id

// ---- utils_frags.rs:702-711 phred_scale ---------------------------------------------------------------------
struct PhredLut {
    double w[256];
    PhredLut() {
        for (int q = 0; q < 256; ++q) {
            float prob = 1.0f - powf(10.0f, (float)q / -10.0f);  // 1. - 10_f32.powf(q as f32 / -10.)
            w[q] = (double)prob;
        }
    }
};
static const PhredLut &default_lut() {
    static PhredLut lut;
    return lut;
}
struct Weights {  // the table in effect for one call (fb_params.phred_lut or the default)
    double w[256];
    explicit Weights(const fb_params *p) {
        if (p && p->phred_lut)
            for (int q = 0; q < 256; ++q) w[q] = (double)p->phred_lut[q];
        else
            memcpy(w, default_lut().w, sizeof(w));
    }
};
static thread_local const Weights *g_w = nullptr;
static inline double phred_scale(const Frag &f, size_t k) { return g_w->w[f.qual[k]]; }

// ---- utils_frags.rs:32-75 distance_read_haplo_epsilon_empty ---------------------------------------------------
static std::pair<double, double> distance_read_haplo_epsilon_empty(const Frag &r, const Haplotype &hap,
                                                                   double epsilon) {
    double diff = 0.0;
    double same = 0.0;
    for (size_t k = 0; k < r.positions.size(); ++k) {  // for pos in r.positions.iter()
        SnpPosition pos = r.positions[k];
        bool empty_pos = true;
        auto it = hap.find(pos);
        if (it != hap.end()) {
            for (auto &kv : it->second.v) {
                if (kv.second != 0.) {
                    empty_pos = false;
                    break;
                }
            }
        }
        if (empty_pos) {
            diff += epsilon;
            continue;
        }
        Genotype frag_var = r.seq[k];
        const AlleleMap &m = it->second;
        size_t cons = m.max_by_count();
        Genotype consensus_var = m.v[cons].first;
        if (frag_var == consensus_var) {
            same += phred_scale(r, k);
        } else {
            const double *count = m.get(frag_var);
            if (count) {
                if (*count == m.v[cons].second) {
                    same += phred_scale(r, k);
                    continue;
                }
            }
            diff += phred_scale(r, k);
        }
    }
    return std::make_pair(same, diff);
}

// Rust's `f64::round() as usize`: round half away from zero, saturating cast.
static inline uint64_t round_as_usize(double x) {
    double r = std::round(x);
    if (!(r > 0.0)) return 0;
    if (r >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)r;
}
static inline uint64_t trunc_as_usize(double x) {  // `x as usize`
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}

// ---- utils_frags.rs:77-108 distance_read_haplo ------------------------------------------------------------------
static std::pair<uint64_t, uint64_t> distance_read_haplo(const Frag &r1, const Haplotype &hap) {
    double diff = 0.;
    double same = 0.;
    for (size_t k = 0; k < r1.positions.size(); ++k) {
        SnpPosition pos = r1.positions[k];
        auto it = hap.find(pos);
        if (it == hap.end()) continue;
        Genotype frag_var = r1.seq[k];
        const AlleleMap &m = it->second;
        // NOTE: the reference unwraps max_by_key here; an existing position key always has >= 1 allele entry on
        // every path that reaches this function (HapNode::new inserts both together, types_structs.rs:175-177).
        if (m.empty()) continue;
        size_t cons = m.max_by_count();
        Genotype consensus_var = m.v[cons].first;
        if (frag_var == consensus_var) {
            same += phred_scale(r1, k);
        } else {
            const double *count = m.get(frag_var);
            if (count) {
                if (*count == m.v[cons].second) {
                    continue;  // a tie contributes to neither (lines 97-101)
                }
            }
            diff += phred_scale(r1, k);
        }
    }
    return std::make_pair(round_as_usize(same), round_as_usize(diff));
}

// ---- utils_frags.rs:160-184 set_to_seq_dict / hap_block_from_partition ----------------------------------------------
static Haplotype set_to_seq_dict(const FragSet &frag_set, bool use_phred) {
    Haplotype hap_map;
    for (const Frag *frag : frag_set) {
        for (size_t k = 0; k < frag->positions.size(); ++k) {
            Genotype var_at_pos = frag->seq[k];
            AlleleMap &sites = hap_map[frag->positions[k]];
            double &site_counter = sites.entry_or_zero(var_at_pos);
            if (use_phred)
                site_counter += phred_scale(*frag, k);
            else
                site_counter += 1.;
        }
    }
    return hap_map;
}
static HapBlock hap_block_from_partition(const std::vector<FragSet> &part, bool use_qual) {
    HapBlock b;
    for (auto &set : part) b.blocks.push_back(set_to_seq_dict(set, use_qual));
    return b;
}

// ---- utils_frags.rs:211-248 stable_binom_cdf_p_rev -------------------------------------------------------------------
static double stable_binom_cdf_p_rev(uint64_t n, uint64_t k, double p, double div_factor) {
    if (n == 0) return 0.0;
    double n64 = (double)n;
    double k64 = (double)k;
    double a = k64 / n64;
    if (a == 1.0) a = 0.9999999;
    if (a == 0.0) a = 0.0000001;
    double rel_ent = a * std::log(a / p) + (1.0 - a) * std::log((1.0 - a) / (1.0 - p));
    if (a < p) rel_ent = -rel_ent;
    double large_dev_val = -1.0 * n64 / div_factor * rel_ent;
    return large_dev_val;
}

// ---- utils_frags.rs:250-258 log_sum_exp ---------------------------------------------------------------------------------
static inline double rust_f64_max(double a, double b) {  // f64::max ignores a NaN operand
    if (std::isnan(a)) return b;
    if (std::isnan(b)) return a;
    return a > b ? a : b;
}
static double log_sum_exp(const std::vector<double> &probs) {
    double max = std::numeric_limits<double>::quiet_NaN();
    for (double x : probs) max = rust_f64_max(max, x);
    double sum = 0.0;
    for (double logpval : probs) sum += std::exp(logpval - max);
    return max + std::log(sum);
}

// ---- utils_frags.rs:465-490 add_read_to_block / remove_read_from_block -------------------------------------------------
static void add_read_to_block(HapBlock &block, const Frag &frag, size_t part) {
    for (size_t k = 0; k < frag.positions.size(); ++k) {
        AlleleMap &sites = block.blocks[part][frag.positions[k]];
        double &site_counter = sites.entry_or_zero(frag.seq[k]);
        site_counter += phred_scale(frag, k);
    }
}
static void remove_read_from_block(HapBlock &block, const Frag &frag, size_t part) {
    for (size_t k = 0; k < frag.positions.size(); ++k) {
        Genotype var_at_pos = frag.seq[k];
        AlleleMap &sites = block.blocks[part][frag.positions[k]];
        double &site_counter = sites.entry_or_zero(var_at_pos);
        if (site_counter != 0.) site_counter -= phred_scale(frag, k);
        if (site_counter <= 0.) sites.remove(var_at_pos);
    }
}

// ---- Rust std BinaryHeap<T> (max-heap over a Vec) ------------------------------------------------------------------------
// T must provide le(a,b) [a <= b], ge(a,b), lt(a,b).  Here every comparison is on the node score only
// (types_structs.rs:127-131 for SearchNode; HapBlock::cmp compares blocks.len(), always equal, :258-262).
template <class T>
struct BinaryHeap {
    std::vector<T> data;
    static bool le(const T &a, const T &b) { return a.score() <= b.score(); }
    static bool ge(const T &a, const T &b) { return a.score() >= b.score(); }
    static bool lt(const T &a, const T &b) { return a.score() < b.score(); }
    size_t len() const { return data.size(); }
    void sift_up(size_t start, size_t pos) {
        T elt = std::move(data[pos]);
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (le(elt, data[parent])) break;
            data[pos] = std::move(data[parent]);
            pos = parent;
        }
        data[pos] = std::move(elt);
    }
    void push(T item) {
        size_t old_len = data.size();
        data.push_back(std::move(item));
        sift_up(0, old_len);
    }
    void sift_down_to_bottom(size_t pos) {
        size_t end = data.size();
        size_t start = pos;
        T elt = std::move(data[pos]);
        size_t child = 2 * pos + 1;
        while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {
            if (le(data[child], data[child + 1])) child += 1;
            data[pos] = std::move(data[child]);
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            data[pos] = std::move(data[child]);
            pos = child;
        }
        data[pos] = std::move(elt);
        sift_up(start, pos);
    }
    void pop() {
        if (data.empty()) return;
        T item = std::move(data.back());
        data.pop_back();
        if (!data.empty()) {
            std::swap(item, data[0]);
            sift_down_to_bottom(0);
        }
    }
    void sift_down_range(size_t pos, size_t end) {
        T elt = std::move(data[pos]);
        size_t child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            if (le(data[child], data[child + 1])) child += 1;
            if (ge(elt, data[child])) {
                data[pos] = std::move(elt);
                return;
            }
            data[pos] = std::move(data[child]);
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1 && lt(elt, data[child])) {
            data[pos] = std::move(data[child]);
            pos = child;
        }
        data[pos] = std::move(elt);
    }
    std::vector<T> into_sorted_vec() {
        size_t end = data.size();
        while (end > 1) {
            end -= 1;
            std::swap(data[0], data[end]);
            sift_down_range(0, end);
        }
        return std::move(data);
    }
};

// ---- types_structs.rs:114-153 SearchNode, 216-251 build_child_node ------------------------------------------------------
struct SearchNode {
    const Frag *read;
    size_t part;
    double score;
    std::vector<size_t> freqs;
    std::vector<std::pair<double, double>> error_vec;
    std::shared_ptr<SearchNode> parent_node;
    SnpPosition current_pos;
    std::set<size_t> broken_blocks;
    ~SearchNode() {  // iterative drop like types_structs.rs:141-153 (avoids deep recursion on long chains)
        std::shared_ptr<SearchNode> prev = std::move(parent_node);
        while (prev && prev.use_count() == 1) {
            std::shared_ptr<SearchNode> next = std::move(prev->parent_node);
            prev = std::move(next);
        }
    }
};
static SearchNode *build_child_node(const Frag *read, size_t part, std::shared_ptr<SearchNode> parent,
                                    std::vector<std::pair<double, double>> error_vec, double score,
                                    SnpPosition current_pos) {
    SearchNode *n = new SearchNode();
    if (parent) {
        n->freqs = parent->freqs;
        n->freqs[part] += 1;
    } else {
        n->freqs.assign(error_vec.size(), 1);
    }
    n->read = read;
    n->part = part;
    n->score = score;
    n->error_vec = std::move(error_vec);
    n->parent_node = std::move(parent);
    n->current_pos = current_pos;
    return n;
}

// ---- types_structs.rs:326-376 build_truncated_hap_block ----------------------------------------------------------------
static std::pair<std::set<size_t>, HapBlock> build_truncated_hap_block(const HapBlock &block, const Frag &frag,
                                                                       size_t part, SnpPosition current_startpos) {
    size_t ploidy = block.blocks.size();
    std::set<size_t> blocks_broken;
    std::vector<Haplotype> block_vec = block.blocks;  // manual deepcopy
    std::vector<int> num_after(ploidy, 0), num_before(ploidy, 0);
    SnpPosition boundary_end = current_startpos + 50;
    for (size_t i = 0; i < ploidy; ++i) {
        for (auto &kv : block.blocks[i]) {
            SnpPosition pos = kv.first;
            SnpPosition boundary_start = pos + 50;
            if (pos >= current_startpos && pos < boundary_end) num_after[i] += 1;
            if (pos < current_startpos && boundary_start > current_startpos) num_before[i] += 1;
            if (pos < current_startpos) block_vec[i].erase(pos);
        }
    }
    for (size_t i = 0; i < ploidy; ++i)
        if (num_after[i] == 0 && num_before[i] != 0) blocks_broken.insert(i);
    for (size_t k = 0; k < frag.positions.size(); ++k) {
        AlleleMap &sites = block_vec[part][frag.positions[k]];
        double &site_counter = sites.entry_or_zero(frag.seq[k]);
        site_counter += phred_scale(frag, k);
    }
    HapBlock nb;
    nb.blocks = std::move(block_vec);
    return std::make_pair(std::move(blocks_broken), std::move(nb));
}

// ---- global_clustering.rs:181-208 read_to_node_value ------------------------------------------------------------------------
static std::pair<double, std::vector<std::pair<double, double>>> read_to_node_value(const SearchNode &node,
                                                                                   const Frag &frag,
                                                                                   const HapBlock &block,
                                                                                   size_t part_index,
                                                                                   double epsilon) {
    size_t ploidy = block.blocks.size();
    auto sd = distance_read_haplo_epsilon_empty(frag, block.blocks[part_index], epsilon);
    std::vector<std::pair<double, double>> new_error_vec;
    for (size_t i = 0; i < ploidy; ++i) {
        if (i == part_index)
            new_error_vec.push_back(
                std::make_pair(node.error_vec[i].first + sd.first, node.error_vec[i].second + sd.second));
        else
            new_error_vec.push_back(node.error_vec[i]);
    }
    double mec = 0.0;  // .iter().map(|x| x.1).sum()
    for (auto &x : new_error_vec) mec += x.second;
    return std::make_pair(-1.0 * mec, std::move(new_error_vec));
}

struct HeapEntry {  // (Rc<SearchNode>, HapBlock)
    std::shared_ptr<SearchNode> node;
    HapBlock block;
    double score() const { return node->score; }
};

struct BeamTap {
    double *same, *diff, *logp;
    uint64_t cap, n;
};
struct BeamCounters {
    uint64_t cells_beam = 0;
};

// ---- global_clustering.rs:10-179 beam_search_phasing ------------------------------------------------------------------------
static std::pair<std::map<SnpPosition, std::set<size_t>>, std::vector<FragSet>> beam_search_phasing(
    const std::vector<FragSet> &clique, const std::vector<const Frag *> &all_reads, double epsilon,
    double div_factor, double cutoff_value, size_t max_number_solns, double *best_score_out, BeamTap *tap,
    BeamCounters *ctr) {
    if (all_reads.size() == 0) return std::make_pair(std::map<SnpPosition, std::set<size_t>>(), std::vector<FragSet>());
    std::vector<FragSet> partition = clique;
    size_t ploidy = clique.size();
    HapBlock first_block = hap_block_from_partition(clique, true);
    const Frag *random_frag = all_reads[0];

    std::shared_ptr<SearchNode> first_node(new SearchNode());
    first_node->read = random_frag;
    first_node->part = SIZE_MAX;
    first_node->score = 0.0;
    first_node->freqs.assign(clique.size(), 1);
    first_node->error_vec.assign(ploidy, std::make_pair(0.0, 0.0));
    first_node->current_pos = 0;

    BinaryHeap<HeapEntry> search_node_heap;
    search_node_heap.push(HeapEntry{first_node, std::move(first_block)});

    for (size_t i = 0; i < all_reads.size(); ++i) {
        size_t max_num_soln_mut = max_number_solns;
        if (i < 25) max_num_soln_mut = ploidy * max_number_solns;
        BinaryHeap<HeapEntry> search_node_heap_next;
        const Frag *frag = all_reads[i];
        bool frag_in_clique = false;
        for (size_t j = 0; j < ploidy; ++j)
            if (clique[j].count(frag)) frag_in_clique = true;
        if (frag_in_clique) continue;
        SnpPosition current_startpos = frag->first_position;
        if (ctr) ctr->cells_beam += (uint64_t)search_node_heap.len() * frag->positions.size();
        for (const HeapEntry &he : search_node_heap.data) {  // search_node_heap.iter(): backing-Vec order
            const std::shared_ptr<SearchNode> &node = he.node;
            const HapBlock &block = he.block;
            std::vector<double> p_value_list;
            for (size_t part_index = 0; part_index < ploidy; ++part_index) {
                auto sd = distance_read_haplo_epsilon_empty(*frag, block.blocks[part_index], epsilon);
                double dist = 1.0 * stable_binom_cdf_p_rev(trunc_as_usize(sd.first + sd.second),
                                                           trunc_as_usize(sd.second), epsilon, div_factor);
                if (tap && tap->n < tap->cap) {
                    if (tap->same) tap->same[tap->n] = sd.first;
                    if (tap->diff) tap->diff[tap->n] = sd.second;
                    if (tap->logp) tap->logp[tap->n] = dist;
                }
                if (tap) tap->n++;
                p_value_list.push_back(dist);
            }
            double lse = log_sum_exp(p_value_list);
            for (size_t j = 0; j < ploidy; ++j) {
                if (p_value_list[j] - lse > cutoff_value) {
                    auto sv = read_to_node_value(*node, *frag, block, j, epsilon);
                    double new_node_score = -sv.first;
                    std::shared_ptr<SearchNode> new_node(build_child_node(frag, j, node, std::move(sv.second),
                                                                          new_node_score, current_startpos));
                    auto tb = build_truncated_hap_block(block, *frag, j, current_startpos);
                    for (size_t index : tb.first) new_node->broken_blocks.insert(index);
                    HapBlock &new_block = tb.second;
                    bool project_exists = false;
                    for (const HeapEntry &e : search_node_heap_next.data) {
                        if (e.block == new_block && e.node->score >= new_node->score) project_exists = true;
                    }
                    if (!project_exists) {
                        search_node_heap_next.push(HeapEntry{new_node, std::move(new_block)});
                        if (search_node_heap_next.len() > max_num_soln_mut) search_node_heap_next.pop();
                    }
                }
            }
        }
        search_node_heap = std::move(search_node_heap_next);
    }

    std::vector<HeapEntry> search_node_heap_to_list = search_node_heap.into_sorted_vec();
    const SearchNode *node_pointer = search_node_heap_to_list[0].node.get();
    if (best_score_out) *best_score_out = node_pointer->score;
    std::map<SnpPosition, std::set<size_t>> break_positions;
    while (true) {
        SnpPosition current_pos = node_pointer->current_pos;
        if (node_pointer->broken_blocks.size() > 0) {
            std::set<size_t> &haps_to_break = break_positions[current_pos];
            for (size_t index : node_pointer->broken_blocks) haps_to_break.insert(index);
        }
        if (!node_pointer->parent_node) {
            break;
        } else {
            partition[node_pointer->part].insert(node_pointer->read);
            node_pointer = node_pointer->parent_node.get();
        }
    }
    return std::make_pair(std::move(break_positions), std::move(partition));
}

// ---- local_clustering.rs:12-59 find_reads_in_interval ----------------------------------------------------------------------
static std::vector<size_t> find_reads_in_interval_idx(SnpPosition start, SnpPosition end, uint64_t n_reads,
                                                      const uint32_t *first, const uint32_t *last,
                                                      size_t max_num_reads) {
    std::vector<size_t> final_set;
    for (size_t i = 0; i < n_reads; ++i) {
        if (final_set.size() > max_num_reads) break;
        if (last[i] < start) continue;
        if (first[i] > end) break;
        if (last[i] - first[i] > 10000) continue;
        final_set.push_back(i);
    }
    return final_set;
}

// ---- local_clustering.rs:218-260 get_mec_stats_epsilon ----------------------------------------------------------------------
static std::vector<std::pair<double, double>> get_mec_stats_epsilon(const HapBlock &hap_block, double epsilon,
                                                                    bool use_gaps) {
    std::vector<std::pair<double, double>> binom_vec;
    for (const Haplotype &hap : hap_block.blocks) {
        double errors = 0.;
        double bases = 0.;
        for (auto &pk : hap) {  // for seq_dict in hap.values()
            std::vector<std::pair<Genotype, double>> allele_counts = pk.second.v;
            if (!use_gaps) {
                long index_to_remove = -1;
                for (size_t i = 0; i < allele_counts.size(); ++i)
                    if (allele_counts[i].first == GAP_CHAR) index_to_remove = (long)i;
                if (index_to_remove >= 0) allele_counts.erase(allele_counts.begin() + index_to_remove);
            }
            if (allele_counts.empty()) continue;
            std::stable_sort(allele_counts.begin(), allele_counts.end(),
                             [](const std::pair<Genotype, double> &x, const std::pair<Genotype, double> &y) {
                                 return x.second < y.second;
                             });
            double cons_bases = allele_counts.back().second;
            bases += cons_bases;
            for (size_t i = 0; i + 1 < allele_counts.size(); ++i) errors += allele_counts[i].second;
            if (cons_bases <= 1.) errors += epsilon;
        }
        binom_vec.push_back(std::make_pair(bases, errors));
    }
    return binom_vec;
}
// ---- local_clustering.rs:187-215 get_mec_stats_epsilon_no_phred (same loop on the unweighted histogram) -------------------
static std::vector<std::pair<double, double>> get_mec_stats_epsilon_no_phred(const std::vector<FragSet> &read_part,
                                                                             double epsilon) {
    HapBlock hap_block_no_phred = hap_block_from_partition(read_part, false);
    return get_mec_stats_epsilon(hap_block_no_phred, epsilon, true);
}

// ---- local_clustering.rs:292-358 opt_iterate -----------------------------------------------------------------------------------
struct Move {
    double gain;
    size_t i;
    const Frag *read;
    size_t j;
};
static std::vector<FragSet> opt_iterate(const std::vector<FragSet> &partition, const HapBlock &hap_block,
                                        double epsilon) {
    size_t ploidy = partition.size();
    std::vector<Move> best_moves;
    for (size_t i = 0; i < ploidy; ++i) {
        if (partition[i].size() <= 1) continue;
        for (const Frag *read : partition[i]) {
            const Haplotype &haplo_i = hap_block.blocks[i];
            double errors_read = distance_read_haplo_epsilon_empty(*read, haplo_i, epsilon).second;
            for (size_t j = 0; j < ploidy; ++j) {
                if (j == i) continue;
                const Haplotype &haplo_j = hap_block.blocks[j];
                double read_errors_movej = distance_read_haplo_epsilon_empty(*read, haplo_j, epsilon).second;
                double diff_score = errors_read - read_errors_movej;
                if (diff_score > 0.0) best_moves.push_back(Move{diff_score, i, read, j});
            }
        }
    }
    std::set<const Frag *> moved_reads;
    std::vector<FragSet> new_part = partition;
    // best_moves.sort_by(|a, b| b.0.partial_cmp(&a.0).unwrap()) — stable, descending gain
    std::stable_sort(best_moves.begin(), best_moves.end(),
                     [](const Move &a, const Move &b) { return b.gain < a.gain; });
    size_t number_of_moves = best_moves.size() / 10;
    if (number_of_moves == 0 && best_moves.size() > 0) number_of_moves = best_moves.size() / 3 + 1;
    for (size_t mv_num = 0; mv_num < best_moves.size(); ++mv_num) {
        const Move &mv = best_moves[mv_num];
        if (moved_reads.count(mv.read)) continue;
        if (new_part[mv.i].size() == 1) continue;
        new_part[mv.j].insert(mv.read);
        new_part[mv.i].erase(mv.read);
        moved_reads.insert(mv.read);
        if (mv_num > number_of_moves) break;
    }
    return new_part;
}

struct OptCounters {
    uint64_t n_opt_iterate = 0;  // opt_iterate calls (scoring sweeps)
    uint64_t n_hist = 0;         // hap_block_from_partition calls
    uint32_t n_accepted = 0;
};

// ---- local_clustering.rs:71-130 optimize_clustering --------------------------------------------------------------------------
static double optimize_clustering(std::vector<FragSet> partition, double epsilon, size_t max_iters,
                                  std::vector<FragSet> *out_part, HapBlock *out_block, OptCounters *ctr) {
    bool not_empty = false;
    for (auto &part : partition)
        if (part.size() > 0) not_empty = true;
    if (!not_empty) {
        HapBlock prev = hap_block_from_partition(partition, true);
        if (ctr) ctr->n_hist++;
        *out_part = std::move(partition);
        if (out_block) *out_block = std::move(prev);
        return 0.0;
    }
    HapBlock prev_hap_block = hap_block_from_partition(partition, true);
    if (ctr) ctr->n_hist++;
    auto binom_vec = get_mec_stats_epsilon(prev_hap_block, epsilon, true);
    double prev_score = 0.0;
    for (auto &x : binom_vec) prev_score += x.second;
    prev_score *= -1.;
    std::vector<FragSet> best_part = std::move(partition);
    for (size_t i = 0; i < max_iters; ++i) {
        std::vector<FragSet> new_part = opt_iterate(best_part, prev_hap_block, epsilon);
        HapBlock new_block = hap_block_from_partition(new_part, true);
        if (ctr) {
            ctr->n_opt_iterate++;
            ctr->n_hist++;
        }
        auto new_binom_vec = get_mec_stats_epsilon(new_block, epsilon, true);
        double s = 0.0;
        for (auto &x : new_binom_vec) s += x.second;
        double new_score = s * -1.;
        if (new_score > prev_score) {
            prev_score = new_score;
            best_part = std::move(new_part);
            prev_hap_block = std::move(new_block);
            if (ctr) ctr->n_accepted++;
        } else {
            break;
        }
    }
    *out_part = std::move(best_part);
    if (out_block) *out_block = std::move(prev_hap_block);
    return prev_score;
}

// ---- graph_processing.rs:205-222 MEC-ratio threshold -----------------------------------------------------------------------------
static double mec_threshold(size_t ploidy, double epsilon, uint32_t sensitivity) {
    if (sensitivity == 1)
        return 1.0 / (1.0 - epsilon) / (1.0 + 1.0 / (std::pow((double)ploidy, 0.50) + 1.00));
    else if (sensitivity == 2)
        return 1.0 / (1.0 - epsilon) / (1.0 + 1.0 / (std::pow((double)ploidy, 1.00) + 1. / 3.));
    else
        return 1.0 / (1.0 - epsilon) / (1.0 + 1.0 / (std::pow((double)ploidy, 1.00) + 1.00));
}

struct BlockResult {
    uint32_t best_ploidy = 0;
    uint32_t ploidies_run = 0;
    std::vector<double> mec_vector, expected_errors;
    std::vector<uint32_t> read_ids;
    std::vector<uint8_t> hap;
    uint64_t cells_sweep = 0, cells_hist = 0, cells_beam = 0;
};

// ---- graph_processing.rs:103-304 get_local_hap_blocks (up to, not including, HapNode construction) ------------------------------
static BlockResult get_local_hap_blocks(const std::vector<Frag> &all_frags, const uint32_t *first,
                                        const uint32_t *last, SnpPosition lo, SnpPosition hi,
                                        const fb_params &options) {
    BlockResult res;
    size_t max_ploidy = options.max_ploidy;
    double epsilon = options.epsilon;
    size_t max_number_solns = options.max_number_solns;
    size_t ploidy_start = 1;
    size_t ploidy_end = max_ploidy + 1;
    size_t num_ploidies = ploidy_end - ploidy_start;
    std::vector<double> mec_vector(num_ploidies, 0.);
    std::vector<std::vector<FragSet>> parts_vector;
    std::vector<double> expected_errors_ref;
    std::vector<size_t> reads = find_reads_in_interval_idx(lo, hi, all_frags.size(), first, last, SIZE_MAX);
    res.mec_vector.assign(num_ploidies, 0.);
    res.expected_errors.assign(num_ploidies, 0.);
    size_t best_ploidy = ploidy_start;
    if (reads.empty()) return res;  // None
    uint64_t nnz_block = 0;
    for (size_t r : reads) nnz_block += all_frags[r].positions.size();
    for (size_t ploidy = ploidy_start; ploidy < ploidy_end; ++ploidy) {
        best_ploidy = ploidy;
        res.ploidies_run++;
        double num_alleles = 0.0;
        std::vector<const Frag *> vec_reads_own;
        for (size_t r : reads) vec_reads_own.push_back(&all_frags[r]);
        // vec_reads_own.sort(): Frag::cmp == ascending counter_id for a contig sorted by Frag::cmp (floria.rs:289-293)
        std::sort(vec_reads_own.begin(), vec_reads_own.end(), [](const Frag *a, const Frag *b) {
            if (a->first_position != b->first_position) return a->first_position < b->first_position;
            if (a->last_position != b->last_position) return a->last_position > b->last_position;
            return a->counter_id < b->counter_id;
        });
        BeamCounters bc;
        auto bs = beam_search_phasing(std::vector<FragSet>(ploidy), vec_reads_own, epsilon, options.div_factor,
                                      options.prob_cutoff_ln, max_number_solns, nullptr, nullptr, &bc);
        res.cells_beam += bc.cells_beam;
        std::vector<FragSet> optimized_part;
        OptCounters oc;
        optimize_clustering(std::move(bs.second), epsilon, options.num_iter_optimize, &optimized_part, nullptr, &oc);
        auto binom_vec = get_mec_stats_epsilon_no_phred(optimized_part, epsilon);
        oc.n_hist++;
        res.cells_sweep += oc.n_opt_iterate * nnz_block;
        res.cells_hist += oc.n_hist * nnz_block;
        for (auto &gb : binom_vec) {
            mec_vector[ploidy - ploidy_start] += gb.second;
            num_alleles += gb.first;
            num_alleles += gb.second;
        }
        parts_vector.push_back(std::move(optimized_part));
        expected_errors_ref.push_back(num_alleles * epsilon);
        if (ploidy > ploidy_start) {
            double thr = mec_threshold(ploidy, epsilon, options.ploidy_sensitivity);
            if ((mec_vector[ploidy - ploidy_start] / mec_vector[ploidy - ploidy_start - 1]) < thr) {
                // do nothing
            } else {
                if (options.stopping_heuristic) {
                    best_ploidy -= 1;
                    break;
                }
            }
            if (mec_vector[ploidy - ploidy_start] < expected_errors_ref[ploidy - ploidy_start]) break;
        } else {
            if (mec_vector[ploidy - ploidy_start] < expected_errors_ref[ploidy - ploidy_start]) break;
        }
    }
    res.best_ploidy = (uint32_t)best_ploidy;
    for (size_t k = 0; k < expected_errors_ref.size(); ++k) res.expected_errors[k] = expected_errors_ref[k];
    res.mec_vector = mec_vector;
    const std::vector<FragSet> &best_part = parts_vector[best_ploidy - ploidy_start];
    std::vector<std::pair<uint32_t, uint8_t>> rows;
    for (size_t h = 0; h < best_part.size(); ++h)
        for (const Frag *f : best_part[h]) rows.push_back(std::make_pair((uint32_t)f->counter_id, (uint8_t)h));
    std::sort(rows.begin(), rows.end());
    for (auto &r : rows) {
        res.read_ids.push_back(r.first);
        res.hap.push_back(r.second);
    }
    return res;
}

// ---- utils_frags.rs:405-463 get_range_with_lengths ----------------------------------------------------------------------------------
static std::vector<std::pair<SnpPosition, SnpPosition>> get_range_with_lengths(const uint64_t *snp_to_genome_pos,
                                                                               uint64_t n, uint64_t block_length,
                                                                               uint64_t overlap_len,
                                                                               double minimal_density) {
    std::vector<std::pair<SnpPosition, SnpPosition>> return_vec;
    if (n == 0) return return_vec;
    uint64_t cum_pos = 0;
    uint64_t last_pos = snp_to_genome_pos[0];
    SnpPosition left_endpoint = 0;
    SnpPosition new_left_end = 0;
    bool hit_new_left = false;
    for (uint64_t ii = 0; ii < n; ++ii) {
        uint64_t pos = snp_to_genome_pos[ii];
        SnpPosition i = (SnpPosition)ii;
        if (i == (SnpPosition)(n - 1)) {
            return_vec.push_back(std::make_pair(left_endpoint, i));
            break;
        }
        if (pos < last_pos) {
            g_err = "VCF malformed. Positions are not increasing";
            return std::vector<std::pair<SnpPosition, SnpPosition>>();
        }
        cum_pos += pos - last_pos;
        last_pos = pos;
        if (cum_pos > block_length - overlap_len && hit_new_left == false) {
            new_left_end = i;
            hit_new_left = true;
        }
        if (cum_pos > block_length) {
            cum_pos = 0;
            double snp_density = (double)(i - left_endpoint) / (double)block_length;
            if (snp_density > minimal_density) return_vec.push_back(std::make_pair(left_endpoint, i - 1));
            if (snp_to_genome_pos[new_left_end] + block_length < snp_to_genome_pos[new_left_end + 1])
                left_endpoint = new_left_end;
            else
                left_endpoint = new_left_end + 1;
            last_pos = snp_to_genome_pos[left_endpoint];
            hit_new_left = false;
        }
    }
    for (auto &x : return_vec) {
        x.first += 1;
        x.second += 1;
    }
    return return_vec;
}

// ---- part_block_manip.rs:13-24 overlap_percent -----------------------------------------------------------------------------------------
static double overlap_percent(SnpPosition x1, SnpPosition x2, SnpPosition y1, SnpPosition y2) {
    SnpPosition a = x2 - y1 + 1, b = y2 - x1 + 1;  // u32 arithmetic (wrapping in release builds)
    SnpPosition intersect = std::max(std::min(a, b), (SnpPosition)0);
    SnpPosition min_length = x2 - x1 + 1;
    double p = (double)intersect / (double)min_length;
    if (p > 1.) return 1.;
    return p;
}

// ---- part_block_manip.rs:27-98 separate_broken_haplogroups -----------------------------------------------------------------------------
static void separate_broken_haplogroups(std::vector<FragSet> &all_joined_path_parts,
                                        std::vector<std::pair<SnpPosition, SnpPosition>> &snp_range_parts_vec) {
    std::vector<std::pair<size_t, std::vector<SnpPosition>>> all_breaks;
    auto sorted_by_first = [](const FragSet &part) {
        std::vector<const Frag *> v(part.begin(), part.end());
        std::stable_sort(v.begin(), v.end(),
                         [](const Frag *x, const Frag *y) { return x->first_position < y->first_position; });
        return v;
    };
    for (size_t i = 0; i < snp_range_parts_vec.size(); ++i) {
        std::vector<const Frag *> vec_of_frags = sorted_by_first(all_joined_path_parts[i]);
        SnpPosition current_lastest_pos = 0;
        std::vector<SnpPosition> breaks;
        for (const Frag *frag : vec_of_frags) {
            if (current_lastest_pos != 0 && frag->first_position > current_lastest_pos) {
                if (current_lastest_pos >= snp_range_parts_vec[i].first &&
                    current_lastest_pos < snp_range_parts_vec[i].second)
                    breaks.push_back(current_lastest_pos);
            }
            if (frag->last_position > current_lastest_pos) current_lastest_pos = frag->last_position;
        }
        if (!breaks.empty()) all_breaks.push_back(std::make_pair(i, breaks));
    }
    std::vector<FragSet> new_parts;
    std::vector<std::pair<SnpPosition, SnpPosition>> new_ranges;
    for (auto &break_info : all_breaks) {
        size_t spot_index = 0;
        const std::vector<SnpPosition> &break_spots = break_info.second;
        SnpPosition break_start = snp_range_parts_vec[break_info.first].first;
        std::vector<const Frag *> vec_of_frags = sorted_by_first(all_joined_path_parts[break_info.first]);
        SnpPosition end_spot = break_spots[spot_index];
        FragSet new_part;
        for (const Frag *frag : vec_of_frags) {
            if (frag->last_position <= end_spot) {
                new_part.insert(frag);
            } else {
                // NOTE (faithful to :71-84): the frag that triggers the switch is NOT inserted anywhere.
                new_ranges.push_back(std::make_pair(break_start, end_spot));
                new_parts.push_back(std::move(new_part));
                break_start = end_spot + 1;
                spot_index += 1;
                if (spot_index != break_spots.size())
                    end_spot = break_spots[spot_index];
                else
                    end_spot = std::numeric_limits<SnpPosition>::max();
                new_part = FragSet();
            }
        }
        new_ranges.push_back(std::make_pair(break_start, snp_range_parts_vec[break_info.first].second));
        new_parts.push_back(std::move(new_part));
    }
    for (auto &break_info : all_breaks) all_joined_path_parts[break_info.first].clear();
    for (size_t i = 0; i < new_parts.size(); ++i) {
        all_joined_path_parts.push_back(std::move(new_parts[i]));
        snp_range_parts_vec.push_back(new_ranges[i]);
    }
}

// ---- part_block_manip.rs:174-274 process_reads_for_final_parts (+ sort_parts 276-288) -----------------------------------------------
static void process_reads_for_final_parts(std::vector<FragSet> &all_joined_path_parts,
                                          std::vector<std::pair<SnpPosition, SnpPosition>> &snp_range_parts_vec,
                                          double epsilon) {
    HapBlock all_parts_block = hap_block_from_partition(all_joined_path_parts, true);
    std::map<const Frag *, std::set<size_t>, FragIdLess> read_to_parts_map;
    for (size_t i = 0; i < all_joined_path_parts.size(); ++i)
        for (const Frag *frag : all_joined_path_parts[i]) read_to_parts_map[frag].insert(i);
    for (auto &kv : read_to_parts_map) {
        for (size_t id : kv.second) {
            all_joined_path_parts[id].erase(kv.first);
            remove_read_from_block(all_parts_block, *kv.first, id);
        }
    }
    for (auto &kv : read_to_parts_map) {
        const Frag *frag = kv.first;
        // min_by over (diff + 1., id, same) with lexicographic partial_cmp; first minimum wins
        bool have = false;
        double bd = 0, bs = 0;
        size_t bid = 0;
        for (size_t id : kv.second) {
            auto sd = distance_read_haplo_epsilon_empty(*frag, all_parts_block.blocks[id], epsilon);
            double d1 = sd.second + 1.;
            bool less;
            if (!have)
                less = true;
            else if (d1 != bd)
                less = d1 < bd;
            else if (id != bid)
                less = id < bid;
            else
                less = sd.first < bs;
            if (less) {
                have = true;
                bd = d1;
                bid = id;
                bs = sd.first;
            }
        }
        all_joined_path_parts[bid].insert(frag);
        add_read_to_block(all_parts_block, *frag, bid);
    }
    if (SEPARATE_BROKEN_HAPLOGROUPS) separate_broken_haplogroups(all_joined_path_parts, snp_range_parts_vec);
    // reassign_short (hidden --reassign-short, :235-270) is not restated: out of scope.
    // sort_parts: stable sort of (part, range) by range
    std::vector<size_t> idx(all_joined_path_parts.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(),
                     [&](size_t a, size_t b) { return snp_range_parts_vec[a] < snp_range_parts_vec[b]; });
    std::vector<FragSet> np;
    std::vector<std::pair<SnpPosition, SnpPosition>> nr;
    for (size_t i : idx) {
        np.push_back(std::move(all_joined_path_parts[i]));
        nr.push_back(snp_range_parts_vec[i]);
    }
    all_joined_path_parts = std::move(np);
    snp_range_parts_vec = std::move(nr);
}

// ---- utils_frags.rs:596-657 get_errors_cov_from_frags ---------------------------------------------------------------------------------
static void get_errors_cov_from_frags(const FragSet &frags, SnpPosition left_snp_pos, SnpPosition right_snp_pos,
                                      double *cov_out, double *err_out, double *errors_out, double *support_out) {
    Haplotype hap_map = set_to_seq_dict(frags, false);
    std::vector<double> snp_counter_list;
    double errors = 0.;
    double total_support = 0.;
    size_t snp_nonzero = 0;
    for (uint64_t pos = left_snp_pos; pos <= (uint64_t)right_snp_pos; ++pos) {
        double snp_support = 0.;
        double max_count_pos = 0.;
        auto it = hap_map.find((SnpPosition)pos);
        if (it != hap_map.end() && !it->second.empty()) {
            snp_nonzero += 1;
            for (auto &sc : it->second.v) {
                if (sc.first == GAP_CHAR) continue;
                if (sc.second > snp_support) max_count_pos = sc.second;  // sic: compares with the running SUM
                snp_support += sc.second;
            }
        }
        total_support += snp_support;
        errors += snp_support - max_count_pos;
        snp_counter_list.push_back(snp_support);
    }
    std::sort(snp_counter_list.begin(), snp_counter_list.end());
    double cov;
    if (snp_counter_list.empty()) {
        cov = 0.;
    } else {
        if (snp_nonzero > 0) {
            double s = 0.;
            for (double c : snp_counter_list) s += c;
            cov = s / (double)snp_nonzero;
        } else
            cov = 0.;
    }
    *cov_out = cov;
    *err_out = errors / total_support;
    *errors_out = errors;
    *support_out = total_support;
}

// ---- utils_frags.rs:659-700 distance_between_haplotypes ---------------------------------------------------------------------------------
static std::pair<double, double> distance_between_haplotypes(const Haplotype &hap1, const Haplotype &hap2,
                                                             SnpPosition r0, SnpPosition r1) {
    double same = 0., diff = 0.;
    for (auto &kv : hap1) {
        SnpPosition pos = kv.first;
        double cov_pos_1 = 0.;
        for (auto &x : kv.second.v) cov_pos_1 += x.second;
        auto it2 = hap2.find(pos);
        if (it2 != hap2.end()) {
            double cov_pos_2 = 0.;
            for (auto &x : it2->second.v) cov_pos_2 += x.second;
            if ((cov_pos_1 > DIST_COV_CUTOFF && cov_pos_2 > DIST_COV_CUTOFF) || (pos >= r0 && pos <= r1)) {
                // the reference unwraps max_by_key; on the get_hapq path every key has >= 1 allele entry
                if (kv.second.empty() || it2->second.empty()) continue;
                Genotype c1 = kv.second.v[kv.second.max_by_count()].first;
                Genotype c2 = it2->second.v[it2->second.max_by_count()].first;
                if (c1 == c2)
                    same += 1.;
                else
                    diff += 1.;
            }
        }
    }
    return std::make_pair(same, diff);
}

// ---- part_block_manip.rs:454-515 find_overlapping_blocks (rust-lapper semantics restated) ------------------------------------------------
struct Iv {
    SnpPosition start, stop;
    size_t val;
};
static void find_overlapping_blocks(size_t n_parts, double ol_cutoff,
                                    const std::vector<std::pair<SnpPosition, SnpPosition>> &ranges,
                                    std::map<size_t, std::vector<Iv>> &all_overlaps,
                                    std::map<size_t, std::vector<double>> &all_overlaps_percentage) {
    std::vector<Iv> interval_vec;
    for (size_t i = 0; i < n_parts; ++i) interval_vec.push_back(Iv{ranges[i].first, ranges[i].second, i});
    std::vector<Iv> sorted = interval_vec;
    std::stable_sort(sorted.begin(), sorted.end(), [](const Iv &a, const Iv &b) {
        if (a.start != b.start) return a.start < b.start;
        return a.stop < b.stop;
    });
    for (size_t i = 0; i < interval_vec.size(); ++i) {
        const Iv &range = interval_vec[i];
        for (const Iv &found : sorted) {
            if (!(found.start < range.stop && found.stop > range.start)) continue;  // Lapper::find
            double overlap_p = overlap_percent(range.start, range.stop, found.start, found.stop);
            if (overlap_p > ol_cutoff && found.val != i) {
                all_overlaps[i].push_back(found);
                all_overlaps_percentage[i].push_back(overlap_p);
            }
        }
    }
}

// ---- part_block_manip.rs:517-620 get_hapq ---------------------------------------------------------------------------------------------------
static void get_hapq(const std::vector<FragSet> &parts, const uint64_t *snp_to_genome_pos,
                     const std::vector<std::pair<SnpPosition, SnpPosition>> &ranges, uint64_t block_length,
                     std::vector<uint8_t> &hapqs, std::vector<double> &purities, double &avg_err_out) {
    double weight = 0., error = 0.;
    std::vector<double> total_covs, errs;
    for (size_t i = 0; i < parts.size(); ++i) {
        double cov, err, total_err, total_cov;
        get_errors_cov_from_frags(parts[i], ranges[i].first, ranges[i].second, &cov, &err, &total_err, &total_cov);
        weight += total_cov;
        error += total_err;
        total_covs.push_back(total_cov);
        errs.push_back(err);
    }
    double avg_err = error / weight;
    HapBlock all_parts_block = hap_block_from_partition(parts, true);
    std::map<size_t, std::vector<Iv>> all_ol;
    std::map<size_t, std::vector<double>> all_overlaps_p;
    find_overlapping_blocks(parts.size(), 0.05, ranges, all_ol, all_overlaps_p);
    for (size_t i = 0; i < parts.size(); ++i) {
        double max_penalty = 0.;
        auto itp = all_overlaps_p.find(i);
        if (itp != all_overlaps_p.end()) {
            for (size_t _j = 0; _j < itp->second.size(); ++_j) {
                double ol = itp->second[_j];
                size_t j = all_ol[i][_j].val;
                auto sd = distance_between_haplotypes(all_parts_block.blocks[i], all_parts_block.blocks[j], 0,
                                                      std::numeric_limits<SnpPosition>::max());
                double dist;
                if ((sd.first + sd.second) == 0.)
                    dist = 1.;
                else
                    dist = sd.second / (sd.first + sd.second);
                if (ol * (1. - dist) > max_penalty) max_penalty = ol * (1. - dist);
            }
        }
        SnpPosition r0 = std::numeric_limits<SnpPosition>::max(), r1 = 0;
        for (const Frag *read : parts[i]) {
            if (read->first_position < r0) r0 = read->first_position;
            if (read->last_position >= r1) r1 = read->last_position;
        }
        uint64_t base_range;
        if (r0 > r1)
            base_range = 0;
        else
            base_range = snp_to_genome_pos[ranges[i].second - 1] - snp_to_genome_pos[ranges[i].first - 1];
        double t1 = HAPQ_CONSTANT * (1. - max_penalty);
        double t2 = std::min(1., (double)parts[i].size() / 3.);
        double t3 = std::max(0.0, std::log(((double)base_range / (double)block_length) + 1.));
        uint64_t hapq = trunc_as_usize(t1 * t2 * t3);
        if (parts[i].size() == 1) hapq = 0;
        hapqs.push_back((uint8_t)std::min<uint64_t>(hapq, 60));
        purities.push_back(errs[i] / avg_err);
    }
    avg_err_out = avg_err;
}

// ---- types_structs.rs:169-180 HapNode::new hap_map; graph_processing.rs:22-76 update_hap_graph out_weights ---------------------------------
static Haplotype hap_node_map(const FragSet &frag_set, SnpPosition e0, SnpPosition e1) {
    Haplotype hap_map;
    for (const Frag *frag : frag_set)
        for (size_t k = 0; k < frag->positions.size(); ++k) {
            SnpPosition pos = frag->positions[k];
            if (pos <= e1 && pos >= e0) {
                AlleleMap &sites = hap_map[pos];
                sites.entry_or_zero(frag->seq[k]) += phred_scale(*frag, k);
            }
        }
    return hap_map;
}

// ---- helpers for the C API ---------------------------------------------------------------------------------------------------------------------
static bool build_frags(const fb_frags *in, std::vector<Frag> &out) {
    if (!in) {
        g_err = "null frags";
        return false;
    }
    out.resize(in->n_reads);
    for (uint64_t i = 0; i < in->n_reads; ++i) {
        Frag &f = out[i];
        f.counter_id = i;
        f.first_position = in->first[i];
        f.last_position = in->last[i];
        uint64_t a = in->row_ptr[i], b = in->row_ptr[i + 1];
        f.positions.assign(in->pos + a, in->pos + b);
        f.seq.assign(in->allele + a, in->allele + b);
        f.qual.assign(in->qual + a, in->qual + b);
        for (size_t k = 1; k < f.positions.size(); ++k)
            if (f.positions[k] <= f.positions[k - 1]) {
                g_err = "positions of a read must be strictly ascending";
                return false;
            }
    }
    return true;
}
static std::vector<FragSet> build_partition(const std::vector<Frag> &frags, uint64_t n_sel, const uint32_t *sel,
                                            const uint8_t *hap, uint32_t ploidy) {
    std::vector<FragSet> part(ploidy);
    for (uint64_t i = 0; i < n_sel; ++i)
        if (hap[i] < ploidy) part[hap[i]].insert(&frags[sel[i]]);
    return part;
}
static std::vector<FragSet> build_parts_csr(const std::vector<Frag> &frags, uint64_t n_parts, const uint64_t *ptr,
                                            const uint32_t *reads) {
    std::vector<FragSet> parts(n_parts);
    for (uint64_t i = 0; i < n_parts; ++i)
        for (uint64_t k = ptr[i]; k < ptr[i + 1]; ++k) parts[i].insert(&frags[reads[k]]);
    return parts;
}
struct WeightScope {
    Weights w;
    const Weights *prev;
    explicit WeightScope(const fb_params *p) : w(p), prev(g_w) { g_w = &w; }
    ~WeightScope() { g_w = prev; }
};

}  // namespace orc

using namespace orc;

extern "C" {

const char *orc_last_error(void) { return g_err.c_str(); }

// test hook: drive the oracle's BinaryHeap like global_clustering.rs:128-135 / 149 do
struct HeapProbe {
    double s;
    int id;
    double score() const { return s; }
};
int orc_heap_trace(const double *scores, int n, int width, int *data_out, int *sorted_out) {
    BinaryHeap<HeapProbe> h;
    for (int c = 0; c < n; ++c) {
        h.push(HeapProbe{scores[c], c});
        if ((int)h.len() > width) h.pop();
    }
    int len = (int)h.len();
    for (int e = 0; e < len; ++e) data_out[e] = h.data[e].id;
    std::vector<HeapProbe> v = h.into_sorted_vec();
    for (int e = 0; e < len; ++e) sorted_out[e] = v[e].id;
    return len;
}

void orc_phred_lut(float *out256) {
    for (int q = 0; q < 256; ++q) out256[q] = (float)default_lut().w[q];
}
double orc_stable_binom_cdf_p_rev(uint64_t n, uint64_t k, double p, double div_factor) {
    return stable_binom_cdf_p_rev(n, k, p, div_factor);
}
double orc_log_sum_exp(const double *probs, uint64_t n) {
    return log_sum_exp(std::vector<double>(probs, probs + n));
}
double orc_mec_threshold(uint32_t ploidy, double epsilon, uint32_t sensitivity) {
    return mec_threshold(ploidy, epsilon, sensitivity);
}

int64_t orc_get_range_with_lengths(const uint64_t *snp_to_genome_pos, uint64_t n_snps, uint64_t block_length,
                                   uint64_t overlap_len, double minimal_density, uint32_t *lo, uint32_t *hi,
                                   uint64_t cap) {
    g_err.clear();
    auto v = get_range_with_lengths(snp_to_genome_pos, n_snps, block_length, overlap_len, minimal_density);
    if (!g_err.empty()) return -1;
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
        lo[i] = v[i].first;
        hi[i] = v[i].second;
    }
    return (int64_t)v.size();
}

int64_t orc_find_reads_in_interval(uint32_t start, uint32_t end, uint64_t n_reads, const uint32_t *first,
                                   const uint32_t *last, uint32_t *out_ids, uint64_t cap) {
    auto v = find_reads_in_interval_idx(start, end, n_reads, first, last, SIZE_MAX);
    for (size_t i = 0; i < v.size() && i < cap; ++i) out_ids[i] = (uint32_t)v[i];
    return (int64_t)v.size();
}

int orc_score_reads(const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap, uint32_t ploidy,
                    const fb_params *prm, double *same, double *diff) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    auto part = build_partition(frags, n_sel, sel, hap, ploidy);
    HapBlock block = hap_block_from_partition(part, true);
    for (uint64_t i = 0; i < n_sel; ++i)
        for (uint32_t h = 0; h < ploidy; ++h) {
            auto sd = distance_read_haplo_epsilon_empty(frags[sel[i]], block.blocks[h], prm->epsilon);
            if (same) same[i * ploidy + h] = sd.first;
            if (diff) diff[i * ploidy + h] = sd.second;
        }
    return 0;
}

int orc_score_reads_noeps(const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                          uint32_t ploidy, const fb_params *prm, uint64_t *same, uint64_t *diff) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    auto part = build_partition(frags, n_sel, sel, hap, ploidy);
    HapBlock block = hap_block_from_partition(part, true);
    for (uint64_t i = 0; i < n_sel; ++i)
        for (uint32_t h = 0; h < ploidy; ++h) {
            auto sd = distance_read_haplo(frags[sel[i]], block.blocks[h]);
            if (same) same[i * ploidy + h] = sd.first;
            if (diff) diff[i * ploidy + h] = sd.second;
        }
    return 0;
}

int orc_hap_block_from_partition(const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                                 uint32_t ploidy, int use_qual, const fb_params *prm, uint32_t pos_lo,
                                 uint32_t n_pos, double *counts, uint8_t *key_mask) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    auto part = build_partition(frags, n_sel, sel, hap, ploidy);
    HapBlock block = hap_block_from_partition(part, use_qual != 0);
    if (counts) memset(counts, 0, sizeof(double) * (size_t)ploidy * n_pos * 4);
    if (key_mask) memset(key_mask, 0, (size_t)ploidy * n_pos);
    for (uint32_t h = 0; h < ploidy; ++h)
        for (auto &kv : block.blocks[h]) {
            if (kv.first < pos_lo || kv.first >= pos_lo + n_pos) continue;
            size_t p = kv.first - pos_lo;
            for (auto &ac : kv.second.v) {
                if (ac.first > 3) {
                    g_err = "allele > 3";
                    return 1;
                }
                if (counts) counts[((size_t)h * n_pos + p) * 4 + ac.first] = ac.second;
                if (key_mask) key_mask[(size_t)h * n_pos + p] |= (uint8_t)(1u << ac.first);
            }
        }
    return 0;
}

int orc_get_mec_stats_epsilon(const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap,
                              uint32_t ploidy, int use_phred, const fb_params *prm, double *bases,
                              double *errors) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    auto part = build_partition(frags, n_sel, sel, hap, ploidy);
    std::vector<std::pair<double, double>> bv;
    if (use_phred) {
        HapBlock block = hap_block_from_partition(part, true);
        bv = get_mec_stats_epsilon(block, prm->epsilon, true);
    } else {
        bv = get_mec_stats_epsilon_no_phred(part, prm->epsilon);
    }
    for (uint32_t h = 0; h < ploidy; ++h) {
        bases[h] = bv[h].first;
        errors[h] = bv[h].second;
    }
    return 0;
}

int orc_beam_search_phasing(const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy,
                            const fb_params *prm, uint8_t *hap_out, double *best_score, double *tap_same,
                            double *tap_diff, double *tap_logp, uint64_t tap_cap, uint64_t *tap_n) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    std::vector<const Frag *> reads;
    for (uint64_t i = 0; i < n_sel; ++i) reads.push_back(&frags[sel[i]]);
    BeamTap tap{tap_same, tap_diff, tap_logp, tap_cap, 0};
    auto res = beam_search_phasing(std::vector<FragSet>(ploidy), reads, prm->epsilon, prm->div_factor,
                                   prm->prob_cutoff_ln, prm->max_number_solns, best_score, &tap, nullptr);
    if (tap_n) *tap_n = tap.n;
    for (uint64_t i = 0; i < n_sel; ++i) hap_out[i] = 255;
    for (size_t h = 0; h < res.second.size(); ++h)
        for (const Frag *f : res.second[h]) {
            // sel is ascending: locate by binary search
            const uint32_t *it = std::lower_bound(sel, sel + n_sel, (uint32_t)f->counter_id);
            hap_out[it - sel] = (uint8_t)h;
        }
    return 0;
}

int orc_optimize_clustering(const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, const uint8_t *hap_in,
                            uint32_t ploidy, const fb_params *prm, uint8_t *hap_out, double *score,
                            uint32_t *n_rounds) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    auto part = build_partition(frags, n_sel, sel, hap_in, ploidy);
    std::vector<FragSet> out;
    OptCounters oc;
    double s = optimize_clustering(std::move(part), prm->epsilon, prm->num_iter_optimize, &out, nullptr, &oc);
    if (score) *score = s;
    if (n_rounds) *n_rounds = oc.n_accepted;
    for (uint64_t i = 0; i < n_sel; ++i) hap_out[i] = 255;
    for (size_t h = 0; h < out.size(); ++h)
        for (const Frag *f : out[h]) {
            const uint32_t *it = std::lower_bound(sel, sel + n_sel, (uint32_t)f->counter_id);
            hap_out[it - sel] = (uint8_t)h;
        }
    return 0;
}

// graph_processing.rs:140-162 for one block at a fixed ploidy on an explicit read list (mirror of fb_phase_block):
// beam_search_phasing -> optimize_clustering -> get_mec_stats_epsilon_no_phred, with the work counters of SURVEY.md 8d.
int orc_phase_block(const fb_frags *fr, uint64_t n_sel, const uint32_t *sel, uint32_t ploidy, const fb_params *prm,
                    uint8_t *hap_out, double *mec_bases, double *mec_errors, fb_block_phase *out) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    std::vector<uint32_t> all;
    if (!sel) {
        n_sel = frags.size();
        all.resize(n_sel);
        for (uint64_t i = 0; i < n_sel; ++i) all[i] = (uint32_t)i;
        sel = all.data();
    }
    std::vector<const Frag *> reads;
    uint64_t nnz_block = 0;
    for (uint64_t i = 0; i < n_sel; ++i) {
        reads.push_back(&frags[sel[i]]);
        nnz_block += frags[sel[i]].positions.size();
    }
    BeamCounters bc;
    double best = 0.0;
    auto bs = beam_search_phasing(std::vector<FragSet>(ploidy), reads, prm->epsilon, prm->div_factor,
                                  prm->prob_cutoff_ln, prm->max_number_solns, &best, nullptr, &bc);
    std::vector<FragSet> optimized_part;
    OptCounters oc;
    const double s = optimize_clustering(std::move(bs.second), prm->epsilon, prm->num_iter_optimize, &optimized_part,
                                         nullptr, &oc);
    auto binom_vec = get_mec_stats_epsilon_no_phred(optimized_part, prm->epsilon);
    oc.n_hist++;
    for (uint32_t h = 0; h < ploidy; ++h) {
        if (mec_bases) mec_bases[h] = binom_vec[h].first;
        if (mec_errors) mec_errors[h] = binom_vec[h].second;
    }
    if (hap_out) {
        for (uint64_t i = 0; i < n_sel; ++i) hap_out[i] = 255;
        for (size_t h = 0; h < optimized_part.size(); ++h)
            for (const Frag *f : optimized_part[h]) {
                const uint32_t *it = std::lower_bound(sel, sel + n_sel, (uint32_t)f->counter_id);
                hap_out[it - sel] = (uint8_t)h;
            }
    }
    if (out) {
        memset(out, 0, sizeof(*out));
        out->beam_score = ploidy > 1 ? best : 0.0;
        out->opt_score = s;
        out->n_rounds = (uint32_t)oc.n_accepted;
        out->ploidy = ploidy;
        out->cells_sweep = oc.n_opt_iterate * nnz_block;
        out->cells_hist = oc.n_hist * nnz_block;
        out->cells_beam = bc.cells_beam;
    }
    return 0;
}

int orc_phase_blocks(const fb_frags *fr, uint64_t n_blocks, const uint32_t *blk_lo, const uint32_t *blk_hi,
                     const fb_params *prm, uint32_t n_threads, fb_block_results **out) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    Weights w(prm);
    std::vector<BlockResult> results(n_blocks);
    std::atomic<uint64_t> next(0);
    if (n_threads == 0) n_threads = 1;
    auto worker = [&]() {
        g_w = &w;
        for (;;) {
            uint64_t j = next.fetch_add(1);
            if (j >= n_blocks) break;
            results[j] = get_local_hap_blocks(frags, fr->first, fr->last, blk_lo[j], blk_hi[j], *prm);
        }
        g_w = nullptr;
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < n_threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();

    fb_block_results *r = (fb_block_results *)calloc(1, sizeof(fb_block_results));
    uint32_t mp = prm->max_ploidy;
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
        tot += results[j].read_ids.size();
    }
    r->read_ptr[n_blocks] = tot;
    r->read_ids = (uint32_t *)calloc(tot + 1, sizeof(uint32_t));
    r->hap = (uint8_t *)calloc(tot + 1, 1);
    for (uint64_t j = 0; j < n_blocks; ++j) {
        const BlockResult &b = results[j];
        r->best_ploidy[j] = b.best_ploidy;
        r->ploidies_run[j] = b.ploidies_run;
        for (uint32_t k = 0; k < mp; ++k) {
            r->mec_vector[j * mp + k] = b.mec_vector[k];
            r->expected_errors[j * mp + k] = b.expected_errors[k];
        }
        if (!b.read_ids.empty()) {
            memcpy(r->read_ids + r->read_ptr[j], b.read_ids.data(), b.read_ids.size() * sizeof(uint32_t));
            memcpy(r->hap + r->read_ptr[j], b.hap.data(), b.hap.size());
        }
        r->cells_sweep += b.cells_sweep;
        r->cells_hist += b.cells_hist;
        r->cells_beam += b.cells_beam;
        r->block_cells[j] = b.cells_sweep + b.cells_hist + b.cells_beam;
    }
    *out = r;
    return 0;
}

void orc_free_block_results(fb_block_results *r) {
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

int orc_process_reads_for_final_parts(const fb_frags *fr, uint64_t n_parts, const uint64_t *part_ptr,
                                      const uint32_t *part_reads, const uint32_t *range_lo,
                                      const uint32_t *range_hi, const fb_params *prm, fb_parts **out) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    auto parts = build_parts_csr(frags, n_parts, part_ptr, part_reads);
    std::vector<std::pair<SnpPosition, SnpPosition>> ranges;
    for (uint64_t i = 0; i < n_parts; ++i) ranges.push_back(std::make_pair(range_lo[i], range_hi[i]));
    process_reads_for_final_parts(parts, ranges, prm->epsilon);
    fb_parts *r = (fb_parts *)calloc(1, sizeof(fb_parts));
    r->n_parts = parts.size();
    r->part_ptr = (uint64_t *)calloc(parts.size() + 1, sizeof(uint64_t));
    r->range_lo = (uint32_t *)calloc(parts.size() + 1, sizeof(uint32_t));
    r->range_hi = (uint32_t *)calloc(parts.size() + 1, sizeof(uint32_t));
    uint64_t tot = 0;
    for (size_t i = 0; i < parts.size(); ++i) {
        r->part_ptr[i] = tot;
        tot += parts[i].size();
        r->range_lo[i] = ranges[i].first;
        r->range_hi[i] = ranges[i].second;
    }
    r->part_ptr[parts.size()] = tot;
    r->read_ids = (uint32_t *)calloc(tot + 1, sizeof(uint32_t));
    uint64_t k = 0;
    for (auto &p : parts)
        for (const Frag *f : p) r->read_ids[k++] = (uint32_t)f->counter_id;
    *out = r;
    return 0;
}
void orc_free_parts(fb_parts *r) {
    if (!r) return;
    free(r->part_ptr);
    free(r->read_ids);
    free(r->range_lo);
    free(r->range_hi);
    free(r);
}

int orc_get_hapq(const fb_frags *fr, uint64_t n_parts, const uint64_t *part_ptr, const uint32_t *part_reads,
                 const uint32_t *range_lo, const uint32_t *range_hi, const uint64_t *snp_to_genome_pos,
                 uint64_t n_snps, const fb_params *prm, uint8_t *hapq, double *rel_err, double *avg_err) {
    (void)n_snps;
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    auto parts = build_parts_csr(frags, n_parts, part_ptr, part_reads);
    std::vector<std::pair<SnpPosition, SnpPosition>> ranges;
    for (uint64_t i = 0; i < n_parts; ++i) ranges.push_back(std::make_pair(range_lo[i], range_hi[i]));
    std::vector<uint8_t> hq;
    std::vector<double> pur;
    double ae = 0;
    get_hapq(parts, snp_to_genome_pos, ranges, prm->block_length, hq, pur, ae);
    for (uint64_t i = 0; i < n_parts; ++i) {
        hapq[i] = hq[i];
        rel_err[i] = pur[i];
    }
    *avg_err = ae;
    return 0;
}

int orc_update_hap_graph(const fb_frags *fr, uint64_t n_cols, const uint64_t *col_ptr, const uint64_t *node_ptr,
                         const uint32_t *node_reads, const uint32_t *node_lo, const uint32_t *node_hi,
                         const fb_params *prm, double *out_weights) {
    std::vector<Frag> frags;
    if (!build_frags(fr, frags)) return 1;
    WeightScope ws(prm);
    uint64_t n_nodes = col_ptr[n_cols];
    std::vector<FragSet> sets(n_nodes);
    std::vector<Haplotype> maps(n_nodes);
    for (uint64_t v = 0; v < n_nodes; ++v) {
        for (uint64_t k = node_ptr[v]; k < node_ptr[v + 1]; ++k) sets[v].insert(&frags[node_reads[k]]);
        maps[v] = hap_node_map(sets[v], node_lo[v], node_hi[v]);
    }
    uint64_t w = 0;
    for (uint64_t i = 0; i + 1 < n_cols; ++i) {
        uint64_t a0 = col_ptr[i], a1 = col_ptr[i + 1], b0 = col_ptr[i + 1], b1 = col_ptr[i + 2];
        size_t nb2 = b1 - b0;
        for (uint64_t v1 = a0; v1 < a1; ++v1) {
            std::vector<double> out(nb2, 0.0);
            for (const Frag *read : sets[v1]) {
                std::vector<std::pair<uint64_t, size_t>> read_to_hap_sim;
                size_t hap_id_in = SIZE_MAX;
                for (size_t l = 0; l < nb2; ++l) {
                    if (sets[b0 + l].count(read)) hap_id_in = l;
                    auto sd = distance_read_haplo(*read, maps[b0 + l]);
                    read_to_hap_sim.push_back(std::make_pair(sd.second, l));
                }
                std::sort(read_to_hap_sim.begin(), read_to_hap_sim.end());
                if (read_to_hap_sim.size() > 1) {
                    if (read_to_hap_sim[0].first != read_to_hap_sim[1].first) {
                        if (hap_id_in != SIZE_MAX) out[hap_id_in] += 1.;
                    }
                } else {
                    if (hap_id_in != SIZE_MAX) out[hap_id_in] += 1.;
                }
            }
            for (size_t l = 0; l < nb2; ++l) out_weights[w++] = out[l];
        }
    }
    (void)MIN_SHARED_READS_UNAMBIG;
    return 0;
}

}  // extern "C"
