// fb_seq.h — sequential/exact-arithmetic building blocks shared by the CUDA kernels and (for CPU unit tests of
// the host-compilable logic, tests/test_host_logic.py) a host build.  Everything here is __host__ __device__.
//
// Reference semantics these helpers reproduce (paths relative to /root/reference):
//   * utils_frags.rs:211-248  stable_binom_cdf_p_rev
//   * utils_frags.rs:250-258  log_sum_exp
//   * Rust std BinaryHeap push / pop / into_sorted_vec as used at global_clustering.rs:46-57,130-133,149
//   * the f64 accumulation `diff += epsilon` / `diff += w` of utils_frags.rs:33-72 in canonical (ascending
//     position) order, see SeqSum below.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

// Weights are integers in units of 2^-26 (every phred weight (1f32 - 10f32.powf(-q/10)) as f64 is a multiple of
// 2^-26, SURVEY.md §8 exactness notes).  Count words carry a sticky "allele key present" flag in bit 62 so that
// map equality (key sets + values, types_structs.rs:253) is plain word equality.
#define FB_Q26 67108864.0
#define FB_INV_Q26 (1.0 / 67108864.0)
#define FB_PRESENT (1ULL << 62)
#define FB_CNT_MASK (FB_PRESENT - 1ULL)

FB_HD double fb_q26_to_f64(int64_t v) { return (double)v * FB_INV_Q26; }  // exact for |v| < 2^53

// ---- utils_frags.rs:211-248 ------------------------------------------------------------------------------------
FB_HD double fb_stable_binom_cdf_p_rev(unsigned long long n, unsigned long long k, double p, double div_factor) {
    if (n == 0) return 0.0;
    double n64 = (double)n;
    double k64 = (double)k;
    double a = k64 / n64;
    if (a == 1.0) a = 0.9999999;
    if (a == 0.0) a = 0.0000001;
    double rel_ent = a * log(a / p) + (1.0 - a) * log((1.0 - a) / (1.0 - p));
    if (a < p) rel_ent = -rel_ent;
    return -1.0 * n64 / div_factor * rel_ent;
}

// `x as usize` for the non-negative values that occur here
FB_HD unsigned long long fb_as_usize(double x) {
    if (!(x > 0.0)) return 0ULL;
    if (x >= 18446744073709551616.0) return 0xFFFFFFFFFFFFFFFFULL;
    return (unsigned long long)x;
}

// ---- utils_frags.rs:250-258 (probs.len() == n >= 1; no NaNs occur) ------------------------------------------------
FB_HD double fb_log_sum_exp(const double *probs, int n) {
    double mx = probs[0];
    for (int i = 1; i < n; ++i) mx = probs[i] > mx ? probs[i] : mx;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += exp(probs[i] - mx);
    return mx + log(sum);
}

// ---- exact emulation of a left-to-right f64 sum of {dyadic weight | epsilon} items ------------------------------------
// The reference adds, per position in iteration order, either a phred weight (a multiple of 2^-26) or epsilon to an
// f64 accumulator.  All-dyadic prefixes are exact in any order; once a non-dyadic epsilon has been added the
// accumulator has bits below 2^-26 ("tail") and later additions round.  add_dyadic_run() adds the exact integer sum
// of a RUN of consecutive dyadic items in one step when that is provably identical to adding them one by one:
//   - no tail yet: every partial sum is a multiple of 2^-26 below 2^27 -> exact;
//   - tail, but S and S+W lie in the same binade: every partial sum is a multiple of ulp(S) inside the binade ->
//     exact (requires S < 2^27 so that 2^-26 >= ulp(S)).
// Otherwise the caller must feed the items one at a time (add_dyadic / add_eps), which is always exact emulation.
struct SeqSum {
    double S;
    int tail;  // 0: S is an exact multiple of 2^-26
    FB_HD void init() {
        S = 0.0;
        tail = 0;
    }
    FB_HD static int same_binade(double a, double b) {
#if defined(__CUDA_ARCH__)
        return (__double2hiint(a) >> 20) == (__double2hiint(b) >> 20);
#else
        union {
            double d;
            uint64_t u;
        } x, y;
        x.d = a;
        y.d = b;
        return (x.u >> 52) == (y.u >> 52);
#endif
    }
    // returns 1 if the run was added, 0 if the caller has to add the items one by one
    FB_HD int add_dyadic_run(int64_t w_q26) {
        if (w_q26 == 0) return 1;
        double r = S + fb_q26_to_f64(w_q26);
        if (!tail) {
            if (r < 134217728.0) {  // 2^27: S, w and every partial sum are multiples of 2^-26 that fit 53 bits
                S = r;
                return 1;
            }
            return 0;
        }
        if (S > 0.0 && S < 134217728.0 && same_binade(S, r)) {
            S = r;
            return 1;
        }
        return 0;
    }
    FB_HD void add_dyadic(int64_t w_q26) {  // single item: one rounded add
        S += fb_q26_to_f64(w_q26);
        if (S >= 134217728.0) tail = 1;  // may have rounded: no longer a guaranteed multiple of 2^-26
    }
    FB_HD void add_eps(double eps, int eps_safe) {
        S += eps;
        if (!eps_safe || S >= 134217728.0) tail = 1;
    }
};

// ---- m consecutive `S += eps` in O(number of binades crossed) ---------------------------------------------------------
// The reference adds epsilon once per empty position (utils_frags.rs:45-48), so a read whose tail lies beyond a
// haplotype's coverage costs hundreds of dependent f64 adds.  While S stays inside one binade its ulp u is constant and
// S is a multiple of u, so fl(S + eps) = S + c with the same c = round-to-nearest(eps / u) * u at every step (unless
// eps / u ends in exactly .5, where ties-to-even alternates: those steps are done one by one).  k steps are therefore
// S + k*c, exactly, as long as S + k*c stays below the top of the binade; the step that crosses it is a real add.
// Bit-identical to the loop `for (i < m) S += eps` for S >= 0, eps > 0 (tests/test_abi_and_host_logic.py).
FB_HD uint64_t fb_f64_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    union {
        double d;
        uint64_t u;
    } v;
    v.d = x;
    return v.u;
#endif
}
FB_HD double fb_bits_f64(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union {
        double d;
        uint64_t u;
    } v;
    v.u = b;
    return v.d;
#endif
}
FB_HD double fb_add_eps_n(double S, double eps, unsigned long long m) {
    while (m > 0) {
        const uint64_t sb = fb_f64_bits(S), eb = fb_f64_bits(eps);
        const int es = (int)((sb >> 52) & 0x7FF), ee = (int)((eb >> 52) & 0x7FF);  // biased exponents
        // bulk steps need: S and eps normal, eps < 2^e(S) (so that eps / ulp(S) < 2^52 and has an exact fraction)
        if (es > 0 && es < 0x7FE && ee > 0 && ee < es) {
            const uint64_t sq = (sb & 0xFFFFFFFFFFFFFULL) | (1ULL << 52);  // S / u, in [2^52, 2^53)
            const uint64_t em = (eb & 0xFFFFFFFFFFFFFULL) | (1ULL << 52);  // eps = em * 2^(ee - 1075)
            const int sh = es - ee;                                        // eps / u = em >> sh, sh >= 1
            uint64_t q, rem2;  // q = floor(eps / u); rem2: 0 below one half, 1 exactly one half, 2 above
            if (sh > 53) {
                q = 0;
                rem2 = 0;
            } else {
                q = sh >= 64 ? 0 : (em >> sh);
                const uint64_t frac = em & ((1ULL << sh) - 1ULL), half = 1ULL << (sh - 1);
                rem2 = frac > half ? 2 : (frac == half ? 1 : 0);
            }
            if (rem2 != 1) {
                const uint64_t cq = q + (rem2 == 2 ? 1ULL : 0ULL);
                if (cq == 0) return S;  // eps < u / 2: every add rounds back to S
                const uint64_t room = ((1ULL << 53) - 1ULL) - sq;
                uint64_t k = room / cq;
                if (k > m) k = m;
                if (k > 0) {
                    const uint64_t nq = sq + k * cq;  // < 2^53: same binade
                    S = fb_bits_f64((sb & 0xFFF0000000000000ULL) | (nq & 0xFFFFFFFFFFFFFULL));
                    m -= k;
                    continue;
                }
            }
        }
        S = S + eps;  // a real add: binade crossing, tie, or S not yet above eps
        m -= 1;
    }
    return S;
}

// m consecutive SeqSum::add_eps in one step (same S and tail as the loop)
FB_HD void fb_seqsum_add_eps_n(SeqSum &ss, double eps, int eps_safe, unsigned long long m) {
    if (m == 0) return;
    ss.S = fb_add_eps_n(ss.S, eps, m);
    if (!eps_safe || ss.S >= 134217728.0) ss.tail = 1;
}

// epsilon is "safe" when it is itself a multiple of 2^-26: then every quantity on the path is exact and the sum is
// order independent (the dyadic-epsilon gate of BASELINE.md §4).
FB_HD int fb_eps_is_safe(double eps) {
    double x = eps * FB_Q26;
    return x == floor(x) && x >= 0.0 && x < 9007199254740992.0;
}

// ---- Rust std BinaryHeap over parallel arrays (score, payload) -----------------------------------------------------------
// Comparisons use the score only; `<=` is true on ties (types_structs.rs:127-131, 258-262).
struct HeapRef {
    double *score;
    int *item;
    int len;
    FB_HD void sift_up(int start, int pos) {
        double es = score[pos];
        int ei = item[pos];
        while (pos > start) {
            int parent = (pos - 1) / 2;
            if (es <= score[parent]) break;
            score[pos] = score[parent];
            item[pos] = item[parent];
            pos = parent;
        }
        score[pos] = es;
        item[pos] = ei;
    }
    FB_HD void push(double s, int it) {
        int old_len = len;
        score[len] = s;
        item[len] = it;
        len++;
        sift_up(0, old_len);
    }
    FB_HD void sift_down_to_bottom(int pos) {
        int end = len;
        int start = pos;
        double es = score[pos];
        int ei = item[pos];
        int child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            if (score[child] <= score[child + 1]) child += 1;
            score[pos] = score[child];
            item[pos] = item[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            score[pos] = score[child];
            item[pos] = item[child];
            pos = child;
        }
        score[pos] = es;
        item[pos] = ei;
        sift_up(start, pos);
    }
    // BinaryHeap::pop: remove the maximum
    FB_HD void pop() {
        if (len == 0) return;
        len--;
        double s = score[len];
        int it = item[len];
        if (len > 0) {
            score[0] = s;  // swap(item, data[0]); the old root is dropped
            item[0] = it;
            sift_down_to_bottom(0);
        }
    }
    FB_HD void sift_down_range(int pos, int end) {
        double es = score[pos];
        int ei = item[pos];
        int child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            if (score[child] <= score[child + 1]) child += 1;
            if (es >= score[child]) {
                score[pos] = es;
                item[pos] = ei;
                return;
            }
            score[pos] = score[child];
            item[pos] = item[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1 && es < score[child]) {
            score[pos] = score[child];
            item[pos] = item[child];
            pos = child;
        }
        score[pos] = es;
        item[pos] = ei;
    }
    // BinaryHeap::into_sorted_vec (ascending); afterwards item[0] is the minimum-score entry
    FB_HD void into_sorted() {
        int end = len;
        while (end > 1) {
            end -= 1;
            double ts = score[0];
            score[0] = score[end];
            score[end] = ts;
            int ti = item[0];
            item[0] = item[end];
            item[end] = ti;
            sift_down_range(0, end);
        }
    }
};

// ---- position/allele hash for the linear state hash of the beam search (fb_beam.cu) ----------------------------------------
FB_HD uint64_t fb_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
FB_HD uint64_t fb_G(uint32_t pos0, uint32_t allele) { return fb_mix64(((uint64_t)pos0 << 2) | allele) | 1ULL; }

// graph_processing.rs:205-222
FB_HD double fb_mec_threshold(unsigned ploidy, double epsilon, unsigned sensitivity) {
    if (sensitivity == 1)
        return 1.0 / (1.0 - epsilon) / (1.0 + 1.0 / (pow((double)ploidy, 0.50) + 1.00));
    else if (sensitivity == 2)
        return 1.0 / (1.0 - epsilon) / (1.0 + 1.0 / (pow((double)ploidy, 1.00) + 1. / 3.));
    else
        return 1.0 / (1.0 - epsilon) / (1.0 + 1.0 / (pow((double)ploidy, 1.00) + 1.00));
}
