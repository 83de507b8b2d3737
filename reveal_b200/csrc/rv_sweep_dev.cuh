// rv_sweep_dev.cuh -- per-slot device predicates of the MUM sweeps, shared by the grid-wide sweep
// kernels (rv_sweep.cu) and the single-block recursion step (rv_split.cu).
#pragma once
#include "rv_sweep.h"

namespace rv {

__device__ __forceinline__ bool is_lower(unsigned char c) { return c >= 'a' && c <= 'z'; }

// reveal.c:81-85 / :145-149 / :246-256: the match cannot be extended to the left
__device__ __forceinline__ bool left_maximal(const unsigned char *T, i64 a, i64 b) {  // no __restrict__: the recursion step lower-cases T in the same kernel
    if (a == 0 || b == 0) return true;
    unsigned char ca = T[a - 1], cb = T[b - 1];
    return ca != cb || ca == 'N' || ca == '$' || is_lower(ca);
}

// ---- pair sweep ---------------------------------------------------------------
__device__ __forceinline__ bool pair_test(const SweepArgs &p, i64 i, i64 &l, i64 &a, i64 &b) {
    if (i < 1 || i >= p.n) return false;
    int li = p.LCP[i];
    if (li < p.minl) return false;
    if (p.LCP[i - 1] >= li) return false;                        // not unique (reveal.c:86-95)
    if (i + 1 < p.n && p.LCP[i + 1] >= li) return false;
    if (0 >= li) return false;                                   // la := 0 at the last slot
    i64 s1 = p.SA[i], s0 = p.SA[i - 1];
    if ((s1 > p.nsep0) == (s0 > p.nsep0)) return false;          // both in the same sample
    a = s1 < s0 ? s1 : s0;
    b = s1 < s0 ? s0 : s1;
    if (!left_maximal(p.T, a, b)) return false;
    l = li;
    if (p.rc == 1) b = p.nsep0 + ((p.flavour ? p.n : p.nT) - b - l);  // reveal.c:98-100 / :162-164
    return true;
}

// ---- multi sweep --------------------------------------------------------------
__device__ __forceinline__ int sample_of(const SweepArgs &p, i64 pos) { return p.SO ? (int)p.SO[pos] : (pos > p.nsep0 ? 1 : 0); }

// ismultimum (reveal.c:227-259) for the interval [lb,ub] of value l > 0
__device__ __forceinline__ bool multi_ok(const SweepArgs &p, i64 lb, i64 ub) {
    // a conjunction of two order-independent predicates: the left-maximality test (which rejects most candidates of
    // similar genomes) runs first, the per-member sample look-ups only for the survivors
    bool maximal = false;
    for (i64 j = lb; j < ub; j++)
        if (left_maximal(p.T, p.SA[j], p.SA[j + 1])) {
            maximal = true;
            break;
        }
    if (!maximal) return false;
    if (p.main_nsamples == 2) {
        if ((p.SA[ub] > p.nsep0) == (p.SA[lb] > p.nsep0)) return false;
    } else if (p.main_nsamples <= 64) {
        u64 seen = 0;
        for (i64 j = lb; j <= ub; j++) {
            u64 bit = 1ull << p.SO[p.SA[j]];
            if (seen & bit) return false;
            seen |= bit;
        }
    } else {
        for (i64 j = lb + 1; j <= ub; j++) {
            int s = p.SO[p.SA[j]];
            for (i64 q = lb; q < j; q++)
                if ((int)p.SO[p.SA[q]] == s) return false;
        }
    }
    return true;
}

// Visits every reportable interval that closes at slot ub, inner first.
template <class F> __device__ __forceinline__ void multi_visit(const SweepArgs &p, i64 ub, F emit) {
    if (ub < 1 || ub >= p.n) return;
    const i64 next = ub + 1 < p.n ? (i64)p.LCP[ub + 1] : -1;  // -1: the final flush closes everything (reveal.c:538)
    i64 m = p.LCP[ub];
    i64 lb = ub - 1;
    for (;;) {
        if (m <= next || m <= 0) break;
        i64 size = ub - lb + 1;
        if ((i64)p.LCP[lb] < m) {  // lb is the left boundary of an lcp-interval of value m
            if (m >= p.minl && size >= p.minn && size <= p.main_nsamples && multi_ok(p, lb, ub)) emit(m, lb, size);
        }
        if (lb == 0 || size >= p.main_nsamples) break;
        i64 v = p.LCP[lb];
        m = v < m ? v : m;
        lb--;
    }
}

}  // namespace rv
