// fb_hosttest.cpp — host build of the __host__ __device__ logic in fb_seq.h, for CPU unit tests only
// (tests/test_host_logic.py).  Not part of the product path.
#include "fb_seq.h"

extern "C" {

// items: w_q26[i] >= 0 is a dyadic weight, w_q26[i] < 0 is an epsilon add.  Adds them with the run-lumping rules the
// kernels use: consecutive dyadic items are offered as one run (add_dyadic_run), falling back to item-by-item.
double ht_seqsum(const long long *items, unsigned long long n, double eps, int chunk) {
    SeqSum ss;
    ss.init();
    const int eps_safe = fb_eps_is_safe(eps);
    unsigned long long i = 0;
    while (i < n) {
        // a chunk of up to `chunk` items, as a warp would see it
        unsigned long long e = i + (unsigned long long)chunk < n ? i + (unsigned long long)chunk : n;
        bool any_eps = false;
        long long tot = 0;
        for (unsigned long long k = i; k < e; ++k) {
            if (items[k] < 0) any_eps = true; else tot += items[k];
        }
        if (!any_eps && ss.add_dyadic_run(tot)) { i = e; continue; }
        for (unsigned long long k = i; k < e; ++k) {
            if (items[k] < 0) ss.add_eps(eps, eps_safe);
            else if (!ss.add_dyadic_run(items[k])) ss.add_dyadic(items[k]);
        }
        i = e;
    }
    return ss.S;
}

// pushes scores in order into a bounded max-heap (pop when len > width), returns the backing-array order of the
// item ids and, in sorted_out, the into_sorted_vec order.
int ht_heap(const double *scores, int n, int width, int *data_out, int *sorted_out) {
    double *hs = new double[n + 2];
    int *hi = new int[n + 2];
    HeapRef hp;
    hp.score = hs; hp.item = hi; hp.len = 0;
    for (int c = 0; c < n; ++c) {
        hp.push(scores[c], c);
        if (hp.len > width) hp.pop();
    }
    int len = hp.len;
    for (int e = 0; e < len; ++e) data_out[e] = hi[e];
    hp.into_sorted();
    for (int e = 0; e < len; ++e) sorted_out[e] = hi[e];
    delete[] hs; delete[] hi;
    return len;
}

double ht_binom(unsigned long long n, unsigned long long k, double p, double div) { return fb_stable_binom_cdf_p_rev(n, k, p, div); }
double ht_lse(const double *p, int n) { return fb_log_sum_exp(p, n); }
double ht_mec_threshold(unsigned ploidy, double eps, unsigned s) { return fb_mec_threshold(ploidy, eps, s); }
int ht_eps_safe(double eps) { return fb_eps_is_safe(eps); }
double ht_add_eps_n(double S, double eps, unsigned long long m) { return fb_add_eps_n(S, eps, m); }
}
