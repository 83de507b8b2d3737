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
// the LCP half of the test: slot i holds a value >= minl that is larger than both neighbours (unique, reveal.c:86-95)
__device__ __forceinline__ bool pair_candidate(const SweepArgs &p, i64 i, int &li) {
    if (i < 1 || i >= p.n) return false;
    li = p.LCP[i];
    if (li < p.minl) return false;
    if (p.LCP[i - 1] >= li) return false;                        // not unique (reveal.c:86-95)
    if (i + 1 < p.n && p.LCP[i + 1] >= li) return false;
    if (0 >= li) return false;                                   // la := 0 at the last slot
    return true;
}
// the SA / text half: the two suffixes lie in different samples and the match cannot be extended to the left
__device__ __forceinline__ bool pair_confirm(const SweepArgs &p, i64 i, int li, i64 &l, i64 &a, i64 &b) {
    i64 s1 = p.SA[i], s0 = p.SA[i - 1];
    if ((s1 > p.nsep0) == (s0 > p.nsep0)) return false;          // both in the same sample
    a = s1 < s0 ? s1 : s0;
    b = s1 < s0 ? s0 : s1;
    if (!left_maximal(p.T, a, b)) return false;
    l = li;
    if (p.rc == 1) b = p.nsep0 + ((p.flavour ? p.n : p.nT) - b - l);  // reveal.c:98-100 / :162-164
    return true;
}
__device__ __forceinline__ bool pair_test(const SweepArgs &p, i64 i, i64 &l, i64 &a, i64 &b) {
    int li;
    return pair_candidate(p, i, li) && pair_confirm(p, i, li, l, a, b);
}

// ---- multi sweep --------------------------------------------------------------
__device__ __forceinline__ int sample_of(const SweepArgs &p, i64 pos) {
    if (p.nsep_n > 0) {
        int smp = 0;
        for (int k = 0; k < p.nsep_n; k++) smp += p.nsep_v[k] < pos ? 1 : 0;
        return smp;
    }
    return p.SO ? (int)p.SO[pos] : (pos > p.nsep0 ? 1 : 0);
}

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

// The members of the interval [lb, ub] while lb walks to the left: what ismultimum (reveal.c:227-259) asks about them, kept up to
// date one member at a time -- every member's suffix, left character and sample are fetched ONCE per closing slot, not once
// per nested interval and neighbour pair (five similar genomes: 15 gathers instead of 48).
//   left-maximal  <=>  some neighbouring pair differs in its left character, or has a special one ('N', '$', lower case), or
//                      sits at the start of the text  <=>  a member at position 0, or not all left characters equal, or the
//                      common left character is special
//   one suffix per sample: two samples -- the two ends lie on different sides of nsep0; up to 64 -- a bit per sample
struct MultiMembers {
    bool anyzero, allsame, havec, dup, side_first, side_last;
    unsigned char c0;
    u64 seen;
    int count;
    __device__ __forceinline__ void init() {
        anyzero = havec = dup = side_first = side_last = false;
        allsame = true;
        c0 = 0;
        seen = 0;
        count = 0;
    }
    __device__ __forceinline__ void add(const SweepArgs &p, i64 pos) {
        if (pos == 0) {
            anyzero = true;
        } else {
            const unsigned char c = p.T[pos - 1];
            if (!havec) {
                c0 = c;
                havec = true;
            } else if (c != c0) {
                allsame = false;
            }
        }
        if (p.main_nsamples == 2 || !p.SO) {   // (no SO array: one or two samples; one sample never reports anything)
            const bool side = pos > p.nsep0;
            if (count == 0) side_first = side;
            side_last = side;
        } else {
            int smp = 0;
            if (p.nsep_n > 0) {
                for (int k = 0; k < p.nsep_n; k++) smp += p.nsep_v[k] < pos ? 1 : 0;   // = SO[pos] (so_fill_kernel)
            } else {
                smp = (int)p.SO[pos];
            }
            const u64 bit = 1ull << smp;
            if (seen & bit) dup = true;
            seen |= bit;
        }
        count++;
    }
    __device__ __forceinline__ bool ok(const SweepArgs &p) const {
        const bool maximal = anyzero || !allsame || c0 == 'N' || c0 == '$' || is_lower(c0);
        if (!maximal) return false;
        return p.main_nsamples == 2 ? side_first != side_last : !dup;
    }
};

// Visits every reportable interval that closes at slot ub, inner first.
template <class F> __device__ __forceinline__ void multi_visit(const SweepArgs &p, i64 ub, F emit) {
    if (ub < 1 || ub >= p.n) return;
    const i64 next = ub + 1 < p.n ? (i64)p.LCP[ub + 1] : -1;  // -1: the final flush closes everything (reveal.c:538)
    i64 m = p.LCP[ub];
    if (m <= next || m <= 0 || m < p.minl) return;            // nothing closes here, or nothing long enough (m only shrinks below)
    const bool tracked = p.main_nsamples <= 64;
    MultiMembers mem;
    mem.init();
    if (tracked) mem.add(p, p.SA[ub]);
    i64 lb = ub - 1;
    for (;;) {
        if (m <= next || m <= 0 || m < p.minl) break;
        if (tracked) mem.add(p, p.SA[lb]);
        i64 size = ub - lb + 1;
        if ((i64)p.LCP[lb] < m) {  // lb is the left boundary of an lcp-interval of value m
            if (size >= p.minn && size <= p.main_nsamples && (tracked ? mem.ok(p) : multi_ok(p, lb, ub))) emit(m, lb, size);
        }
        if (lb == 0 || size >= p.main_nsamples) break;
        i64 v = p.LCP[lb];
        m = v < m ? v : m;
        lb--;
    }
}



// ---- multi-MEM walk (getmultimems, reveal.c:292-434 + ismultimem :261-290) --------------------------------
// The reference's lcp-interval stack walk reports intervals of ANY size whose members cover >= minn samples, and
// its `continue` at reveal.c:340-342 (taken when an interval is a multi-MEM of too few samples) skips the
// `lb = i_lb` hand-over, which changes the left boundary RECORDED for the interval pushed next -- so the output
// depends on the walk itself and cannot be restated per interval.  But the walk decomposes: a position with
// LCP < minl pops every entry that could ever be reported, and an entry of value >= minl only ever inherits its
// left boundary from entries of value >= minl popped at the same position.  So every maximal run of positions with
// LCP >= minl ("segment") is walked by its own thread with its own stack (slots [i0, ...) of two global arrays:
// the depth never exceeds the run length), exactly as the reference walks it.
__device__ __forceinline__ bool mem_ok(const SweepArgs &p, i64 l, i64 lb, i64 ub, int &c) {
    c = 0;
    if (l <= 0) return false;
    if (p.main_nsamples == 2) {
        c = 1;  // reveal.c:267: exactly one of the two flags is incremented
    } else {
        u64 seen = 0;
        for (i64 j = lb; j <= ub; j++) seen |= 1ull << p.SO[p.SA[j]];
        c = __popcll(seen);
    }
    for (i64 j = lb; j < ub; j++)
        if (left_maximal(p.T, p.SA[j], p.SA[j + 1])) return true;
    return false;
}

__device__ __forceinline__ bool mem_segment_start(const SweepArgs &p, i64 i) {
    if (i < 1 || i >= p.n) return false;
    if (p.LCP[i] < p.minl) return false;
    return i == 1 || p.LCP[i - 1] < p.minl;
}

// walks the segment that starts at i0; emit(l, c, lb, size) in the reference's pop order
template <class F> __device__ __forceinline__ void mem_walk(const SweepArgs &p, i64 i0, int *__restrict__ st_l, int *__restrict__ st_lb, F emit) {
    int *sl = st_l + i0, *slb = st_lb + i0;
    i64 depth = 0;
    for (i64 i = i0;; i++) {
        const i64 cur = i < p.n ? (i64)p.LCP[i] : -1;  // -1: the final flush (reveal.c:390-434) closes everything at n-1
        i64 lb = i - 1;
        while (depth > 0 && cur < (i64)sl[depth - 1]) {
            depth--;
            const i64 l = sl[depth], ilb = slb[depth], iub = i - 1, size = iub - ilb + 1;
            bool skip_handover = false;
            int c = 0;
            if (l >= p.minl && size >= p.minn && mem_ok(p, l, ilb, iub, c)) {
                if (c < p.minn) skip_handover = true;  // the reference's `continue` (reveal.c:340-342)
                else emit(l, (i64)c, ilb, size);
            }
            if (!skip_handover) lb = ilb;
        }
        if (i >= p.n || cur < p.minl) break;
        if (depth == 0 || cur > (i64)sl[depth - 1]) {
            sl[depth] = (int)cur;
            slb[depth] = (int)lb;
            depth++;
        }
    }
}

}  // namespace rv
