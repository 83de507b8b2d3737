// rv_split.cu -- deriving the child indexes of one recursion step on the device.
//
// Replaces, for one step of the reference's `aligner` (reveallib/reveal.c:731-1338):
//   label scatter   D[SAi[j]] = 1 lead / 2 trail / 4 parallel / 3 matched     reveal.c:1005-1117
//   split           3-way stable compaction of SA by label, child LCP = running
//                   minimum of the parent LCP over the skipped span, inverse SA
//                   rewritten to child-local ranks                            reveal.c:582-664
//   T lower-casing  of the matched bases                                      reveal.c:1230-1234
//   bubble_sort     boundary fix-up of the leading child                      reveal.c:666-727
// so that a child is exactly the index the reference would hand to the next step
// (same SA order, same LCP values, same inverse), not a rebuild.
//
// split on the GPU: the reference keeps, per class, a running minimum that is reset after
// every entry of that class; this is an associative scan over per-class states
// (count, seen, minimum since the last entry of the class), done tile-wise:
// reduce -> single-block scan of the tile states -> apply.  The reference's `continue`
// for entries with no label (reveal.c:616-620) skips folding LCP[i+1]; kept.
//
// bubble_sort is sequential by nature but touches only suffixes that start shortly before
// a matched interval: one block scans the child for those candidates (their LCP, which
// only ever decreases during the pass, must reach across `begin`), sorts them, and one
// thread replays the reference's loop body on exactly those entries, in order.
#include "rv_internal.h"
#include "rv_sweep.h"
#include "rv_sweep_dev.cuh"
#include <stdlib.h>
#include <vector>
#include <map>

namespace rv {

static const int SP_THREADS = 256;
static const int SP_IPT = 8;
static const int SP_TILE = SP_THREADS * SP_IPT;
static const int INF_LCP = 0x7fffffff;

// ---- label scatter ---------------------------------------------------------------------------
// intervals k = 0..m-1: text positions [ibeg[k], ibeg[k]+len) get label lab[k]; pre[k] = exclusive prefix of lengths
__global__ void __launch_bounds__(256) label_kernel(const i64 *__restrict__ ibeg, const i64 *__restrict__ pre, const unsigned char *__restrict__ lab,
                                                   int m, i64 total, const int *__restrict__ SAi, unsigned char *__restrict__ D, i64 n) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int lo = 0, hi = m - 1;  // last k with pre[k] <= g
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pre[mid] <= g) lo = mid; else hi = mid - 1;
    }
    i64 j = ibeg[lo] + (g - pre[lo]);
    // a position that is not a suffix of this parent carries some other sub-index's rank: never write outside D
    // (the class counts checked by the host afterwards then report the bad interval)
    const i64 r = SAi[j];
    if (r >= 0 && r < n) D[r] = lab[lo];
}

__global__ void __launch_bounds__(256) lower_kernel(const i64 *__restrict__ ibeg, const i64 *__restrict__ pre, int m, i64 total, unsigned char *__restrict__ T) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int lo = 0, hi = m - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pre[mid] <= g) lo = mid; else hi = mid - 1;
    }
    i64 j = ibeg[lo] + (g - pre[lo]);
    unsigned char c = T[j];
    if (c >= 'A' && c <= 'Z') T[j] = (unsigned char)(c + 32);  // tolower in the C locale (reveal.c:1232)
}

// ---- split -----------------------------------------------------------------------------------
struct SplitState {  // scan element over a range of parent slots, per class c = 0 lead, 1 trail, 2 parallel
    u32 cnt[3];      // entries of the class in the range
    int m[3];        // min of the folded LCP values after the last entry of the class (whole range if none)
    u32 seen;        // bit c: the class occurs in the range
};
__device__ __forceinline__ SplitState split_identity() {
    SplitState s;
    s.cnt[0] = s.cnt[1] = s.cnt[2] = 0;
    s.m[0] = s.m[1] = s.m[2] = INF_LCP;
    s.seen = 0;
    return s;
}
// a followed by b
__device__ __forceinline__ SplitState split_combine(const SplitState &a, const SplitState &b) {
    SplitState r;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        r.cnt[c] = a.cnt[c] + b.cnt[c];
        r.m[c] = ((b.seen >> c) & 1u) ? b.m[c] : (a.m[c] < b.m[c] ? a.m[c] : b.m[c]);
    }
    r.seen = a.seen | b.seen;
    return r;
}
__device__ __forceinline__ int class_of(unsigned char d) { return d == 1 ? 0 : (d == 2 ? 1 : (d == 4 ? 2 : -1)); }
// value folded into the running minima when the reference's loop goes from slot i-1 to slot i
__device__ __forceinline__ int folded(const int *__restrict__ LCP, const unsigned char *__restrict__ D, i64 i) {
    if (i == 0) return INF_LCP;
    unsigned char dp = D[i - 1];
    return (dp >= 1 && dp <= 4) ? LCP[i] : INF_LCP;  // `continue` on unlabeled entries skips the fold (reveal.c:616-620)
}
// state of the single slot i, and (through out_m) the class minimum BEFORE the reset at i is applied by the caller
__device__ __forceinline__ void split_step(SplitState &s, int v, int cls, int &out_m) {
#pragma unroll
    for (int c = 0; c < 3; c++) s.m[c] = s.m[c] < v ? s.m[c] : v;
    out_m = INF_LCP;
    if (cls >= 0) {
        out_m = s.m[cls];
        s.m[cls] = INF_LCP;
        s.cnt[cls]++;
        s.seen |= 1u << cls;
    }
}
__device__ __forceinline__ SplitState shfl_up_state(const SplitState &s, int d) {
    SplitState r;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        r.cnt[c] = __shfl_up_sync(FULL, s.cnt[c], d);
        r.m[c] = __shfl_up_sync(FULL, s.m[c], d);
    }
    r.seen = __shfl_up_sync(FULL, s.seen, d);
    return r;
}
// inclusive scan of per-thread states over the block; returns the EXCLUSIVE prefix of the calling thread
__device__ __forceinline__ SplitState block_excl_scan_state(const SplitState &mine, SplitState *s_warp /*[32]*/, SplitState *block_total) {
    SplitState inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        SplitState t = shfl_up_state(inc, d);
        if (lane_id() >= (unsigned)d) inc = split_combine(t, inc);
    }
    const unsigned w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane_id() == 31) s_warp[w] = inc;
    __syncthreads();
    SplitState pre = split_identity();
    for (unsigned ww = 0; ww < w; ww++) pre = split_combine(pre, s_warp[ww]);
    if (block_total) {
        SplitState tot = split_identity();
        for (unsigned ww = 0; ww < nw; ww++) tot = split_combine(tot, s_warp[ww]);
        *block_total = tot;
    }
    SplitState excl_in_warp = shfl_up_state(inc, 1);
    if (lane_id() == 0) excl_in_warp = split_identity();
    __syncthreads();
    return split_combine(pre, excl_in_warp);
}

__global__ void __launch_bounds__(SP_THREADS) split_reduce_kernel(const int *__restrict__ LCP, const unsigned char *__restrict__ D, i64 n,
                                                                  SplitState *__restrict__ tiles) {
    __shared__ SplitState s_warp[32];
    i64 base = (i64)blockIdx.x * SP_TILE + (i64)threadIdx.x * SP_IPT;
    SplitState st = split_identity();
    for (int k = 0; k < SP_IPT; k++) {
        i64 i = base + k;
        if (i < n) {
            int dummy;
            split_step(st, folded(LCP, D, i), class_of(D[i]), dummy);
        }
    }
    SplitState total;
    block_excl_scan_state(st, s_warp, &total);
    if (threadIdx.x == 0) tiles[blockIdx.x] = total;
}

// single thread block: exclusive scan of the tile states.  Every thread folds a contiguous run of tiles, the block scans the
// per-thread states (the combine is associative), every thread writes the exclusive prefixes of its run.
static const int TS_THREADS = 256;
__global__ void __launch_bounds__(TS_THREADS) split_tilescan_kernel(SplitState *__restrict__ tiles, i64 ntiles, u32 *__restrict__ counts) {
    __shared__ SplitState s_warp[32];
    const i64 chunk = (ntiles + TS_THREADS - 1) / TS_THREADS;
    const i64 lo = (i64)threadIdx.x * chunk, hi = lo + chunk < ntiles ? lo + chunk : ntiles;
    SplitState mine = split_identity();
    for (i64 t = lo; t < hi; t++) mine = split_combine(mine, tiles[t]);
    SplitState total;
    SplitState run = block_excl_scan_state(mine, s_warp, &total);
    for (i64 t = lo; t < hi; t++) {
        const SplitState x = tiles[t];
        tiles[t] = run;
        run = split_combine(run, x);
    }
    if (threadIdx.x == 0) {
        counts[0] = total.cnt[0];
        counts[1] = total.cnt[1];
        counts[2] = total.cnt[2];
    }
}

__global__ void __launch_bounds__(SP_THREADS)
split_apply_kernel(const int *__restrict__ SA, const int *__restrict__ LCP, const unsigned char *__restrict__ D, i64 n,
                   const SplitState *__restrict__ tiles, int *__restrict__ SAi, int *__restrict__ sa0, int *__restrict__ lcp0,
                   int *__restrict__ sa1, int *__restrict__ lcp1, int *__restrict__ sa2, int *__restrict__ lcp2, u32 cn0, u32 cn1, u32 cn2) {
    __shared__ SplitState s_warp[32];
    i64 base = (i64)blockIdx.x * SP_TILE + (i64)threadIdx.x * SP_IPT;
    SplitState st = split_identity();
    for (int k = 0; k < SP_IPT; k++) {
        i64 i = base + k;
        if (i < n) {
            int dummy;
            split_step(st, folded(LCP, D, i), class_of(D[i]), dummy);
        }
    }
    SplitState excl = block_excl_scan_state(st, s_warp, nullptr);
    SplitState run = split_combine(tiles[blockIdx.x], excl);  // state just before this thread's first slot
    for (int k = 0; k < SP_IPT; k++) {
        i64 i = base + k;
        if (i >= n) break;
        int cls = class_of(D[i]);
        int out_m;
        u32 rank = cls >= 0 ? run.cnt[cls] : 0u;
        split_step(run, folded(LCP, D, i), cls, out_m);
        if (cls >= 0 && rank < (cls == 0 ? cn0 : (cls == 1 ? cn1 : cn2))) {  // more entries than the intervals cover: reported by the count check
            int *csa = cls == 0 ? sa0 : (cls == 1 ? sa1 : sa2);
            int *clcp = cls == 0 ? lcp0 : (cls == 1 ? lcp1 : lcp2);
            int s = SA[i];
            csa[rank] = s;
            clcp[rank] = rank == 0 ? 0 : out_m;
            SAi[s] = (int)rank;  // update inverse (reveal.c:598, 611, 632)
        }
    }
}

// ---- bubble_sort -------------------------------------------------------------------------------
static const int BB_THREADS = 1024;
static const int BB_CAP = 4096;  // candidates per matched interval handled by the sorted fast path

// The reference's loop body for slot i (reveal.c:686-722): if the match of SA[i] with its upper neighbour runs across `begin`, the
// entry is moved down to the first slot whose LCP is below the truncated length begin - SA[i] (entries in between shift up by
// one), LCP[x + 1] becomes that length and LCP[i + 1] is capped by the old LCP[i]; else, if its match with the LOWER neighbour
// runs across `begin`, LCP[i + 1] is cut to begin - SA[i] when it exceeds LCP[i].
static inline int bubble_cap() {   // candidates per matched interval of the sorted fast path (test hook: RV_BUBBLE_CAP)
    if (const char *e = getenv("RV_BUBBLE_CAP")) {
        int x = atoi(e);
        if (x >= 32 && x <= BB_CAP) return x;
    }
    return BB_CAP;
}

// Replays the loop body for the sorted candidate slots cand[0..cnt), candidates one after the other (their
// effects depend on each other) but each one block-parallel: the insertion of reveal.c:688-700 is a backward
// search for the first LCP below the truncated length followed by a shift of the whole span by one slot, and
// for a suffix that starts right before `begin` that span is a large part of the child.
__device__ __forceinline__ void bubble_replay_block(int *SA, int *LCP, int *SAi, i64 n, i64 begin, const int *s_cand, int cnt, int *s_tmp /*[8]*/) {
    const int NT = (int)blockDim.x;
    const int tid = (int)threadIdx.x;
    for (int c = 0; c < cnt; c++) {
        const i64 i = s_cand[c];
        if (tid == 0) {
            int mode = 0;
            const int sa = SA[i], lcp = LCP[i];
            if ((sa < begin) && ((i64)sa + lcp > begin)) {
                mode = 1;
            } else if (i < n - 1 && (sa < begin) && ((i64)sa + LCP[i + 1] > begin)) {
                if (LCP[i + 1] > lcp) LCP[i + 1] = (int)(begin - sa);
            }
            s_tmp[0] = mode;
            s_tmp[1] = sa;
            s_tmp[2] = lcp;
            s_tmp[3] = -1;  // stop slot found by the search
        }
        __syncthreads();
        if (s_tmp[0] == 1) {
            const int tmpSA = s_tmp[1], tmpLCP = s_tmp[2];
            const i64 thr = begin - tmpSA;
            // ---- search: x = first slot going down from i with LCP[x] < thr, else 0 ----
            i64 hi = i;
            i64 xstop = 0;
            while (hi >= 1) {
                i64 x = hi - tid;
                if (x >= 1 && (i64)LCP[x] < thr) atomicMax(&s_tmp[3], (int)x);
                __syncthreads();
                int found = s_tmp[3];
                __syncthreads();
                if (found >= 0) {
                    xstop = found;
                    break;
                }
                hi -= NT;
            }
            // ---- shift [xstop, i-1] up by one slot, top window first ----
            for (i64 top = i; top > xstop; top -= NT) {
                i64 x = top - tid;  // destination slot
                int vs = 0, vl = 0;
                const bool mine = x > xstop;
                if (mine) {
                    vs = SA[x - 1];
                    vl = LCP[x - 1];
                }
                __syncthreads();
                if (mine) {
                    SA[x] = vs;
                    LCP[x] = vl;
                    SAi[vs] = (int)x;
                }
                __syncthreads();
            }
            if (tid == 0) {
                SAi[tmpSA] = (int)xstop;
                SA[xstop] = tmpSA;
                LCP[xstop + 1] = (int)thr;
                if (i < n - 1 && tmpLCP < LCP[i + 1]) LCP[i + 1] = tmpLCP;
            }
        }
        __syncthreads();
    }
}

// ascending bitonic sort of s_cand[0..cnt) by the whole block (padded with INT_MAX to a power of two <= BB_CAP)
__device__ __forceinline__ void bubble_sort_cands(int *s_cand, int cnt) {
    const int NT = (int)blockDim.x;
    int m = 1;
    while (m < cnt) m <<= 1;
    for (int i = cnt + threadIdx.x; i < m; i += NT) s_cand[i] = 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += NT) {
                int ixj = i ^ j;
                if (ixj > i) {
                    int a = s_cand[i], c = s_cand[ixj];
                    bool up = (i & k) == 0;
                    if ((a > c) == up) {
                        s_cand[i] = c;
                        s_cand[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ bool bubble_candidate(const int *SA, const int *LCP, i64 n, i64 i, i64 begin) {
    const i64 s = SA[i];
    return s < begin && (s + LCP[i] > begin || (i < n - 1 && s + LCP[i + 1] > begin));
}

// More candidates than the sorted fast path holds (a repeat of thousands of characters runs across `begin`): the child is taken
// in windows of `cap` consecutive slots, in slot order -- a window has at most `cap` candidates, they are collected, sorted and
// replayed like above.  Equivalent to the reference's loop over all slots: the body of slot i only moves entries among the slots
// <= i and lowers LCP[i + 1], so the slots of a later window still hold their own entries when their window is looked at, and
// LCP values only ever shrink, so no slot becomes a candidate after its window was collected (the replay re-checks each one).
// (Round 1 handed this case to ONE thread replaying all n slots.)
__device__ __forceinline__ void bubble_windows(int *SA, int *LCP, int *SAi, i64 n, i64 begin, int *s_cand, int *s_cnt /*[12]*/, int cap) {
    const int NT = (int)blockDim.x;
    for (i64 w0 = 0; w0 < n; w0 += cap) {
        if (threadIdx.x == 0) *s_cnt = 0;
        __syncthreads();
        const i64 w1 = w0 + cap < n ? w0 + cap : n;
        for (i64 i = w0 + threadIdx.x; i < w1; i += NT)
            if (bubble_candidate(SA, LCP, n, i, begin)) s_cand[atomicAdd(s_cnt, 1)] = (int)i;
        __syncthreads();
        const int cnt = *s_cnt;
        __syncthreads();
        if (cnt > 0) {
            bubble_sort_cands(s_cand, cnt);
            bubble_replay_block(SA, LCP, SAi, n, begin, s_cand, cnt, s_cnt + 1);
        }
        __syncthreads();
    }
}

// bubble_sort of one child by one thread block (any block size that is a power of two <= 1024); cap <= BB_CAP: candidates the
// sorted fast path takes at once (BB_CAP but for tests)
__device__ __forceinline__ void bubble_block(int *SA, int *LCP, int *SAi, i64 n, const i64 *begins, int nbegins, int *s_cand, int *s_cnt, int cap) {
    const int NT = (int)blockDim.x;
    for (int b = 0; b < nbegins; b++) {
        const i64 begin = begins[b];
        if (threadIdx.x == 0) *s_cnt = 0;
        __syncthreads();
        // candidates: LCP values only decrease during the pass, so a slot whose original values do not
        // reach across `begin` can never take either branch
        for (i64 i = threadIdx.x; i < n; i += NT) {
            if (bubble_candidate(SA, LCP, n, i, begin)) {
                int at = atomicAdd(s_cnt, 1);
                if (at < cap) s_cand[at] = (int)i;
            }
        }
        __syncthreads();
        const int cnt = *s_cnt;
        __syncthreads();
        if (cnt > cap) {
            bubble_windows(SA, LCP, SAi, n, begin, s_cand, s_cnt, cap);
        } else if (cnt > 0) {
            bubble_sort_cands(s_cand, cnt);
            bubble_replay_block(SA, LCP, SAi, n, begin, s_cand, cnt, s_cnt + 1);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BB_THREADS) bubble_kernel(int *SA, int *LCP, int *SAi, i64 n, const i64 *__restrict__ begins, int nbegins, int cap) {
    __shared__ int s_cand[BB_CAP];
    __shared__ int s_cnt[12];  // [0] candidate count, [1..] scratch of the replay
    bubble_block(SA, LCP, SAi, n, begins, nbegins, s_cand, s_cnt, cap);
}

// Large children: the candidate scan runs grid-wide, the sort + replay in one block.
__global__ void __launch_bounds__(256) bubble_detect_kernel(const int *__restrict__ SA, const int *__restrict__ LCP, i64 n, const i64 *__restrict__ begins,
                                                           int b, int *__restrict__ cand, int *__restrict__ cand_cnt, int cap) {
    const i64 begin = begins[b];
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (bubble_candidate(SA, LCP, n, i, begin)) {
            int at = atomicAdd(cand_cnt, 1);
            if (at < cap) cand[at] = (int)i;
        }
    }
}

__global__ void __launch_bounds__(BB_THREADS) bubble_apply_kernel(int *SA, int *LCP, int *SAi, i64 n, const i64 *__restrict__ begins, int b,
                                                                  const int *__restrict__ cand, int *cand_cnt, int cap) {
    __shared__ int s_cand[BB_CAP];
    __shared__ int s_cnt[12];
    const i64 begin = begins[b];
    const int cnt = *cand_cnt;
    __syncthreads();
    if (threadIdx.x == 0) *cand_cnt = 0;  // ready for the next matched interval
    if (cnt > cap) {  // rare: more candidates than the fast path holds
        bubble_windows(SA, LCP, SAi, n, begin, s_cand, s_cnt, cap);
        return;
    }
    if (cnt == 0) return;
    for (int i = threadIdx.x; i < cnt; i += BB_THREADS) s_cand[i] = cand[i];
    __syncthreads();
    bubble_sort_cands(s_cand, cnt);
    bubble_replay_block(SA, LCP, SAi, n, begin, s_cand, cnt, s_cnt + 1);
}

// ---- one whole recursion step of a SMALL sub-index in a single launch ---------------------------
// Deep in the recursion almost every sub-index has a few hundred to a few thousand suffixes; ten launches
// and three host synchronisations per step would be pure latency.  One 1024-thread block does the step:
// labels in shared memory, block-wide split scan, T lower-casing, bubble_sort of the leading child, and
// then already the MUM sweep of every child that will need one (it has no precomputed skipmums), written
// in list order into a host-mapped buffer.  The host sees one launch and one synchronisation per step.
static const int SM_THREADS = 1024;
static const int SM_MAXN = 16384;   // parent entries handled by the single-block path
static const int SM_MAXIV = 40;     // intervals that fit the kernel argument
static const int SM_MAXMUM = 16;

struct SmallStepArgs {
    unsigned char *T;
    int *SAi;
    const unsigned short *SO;
    i64 nT, nsep0;
    int main_nsamples, rc, minl, minn;
    const int *pSA, *pLCP;
    int n;
    int m1;
    i64 total1;
    i64 ibeg[SM_MAXIV], pre[SM_MAXIV];
    unsigned char lab[SM_MAXIV];
    int m2;
    i64 mum_l;
    i64 mbeg[SM_MAXMUM];
    int nb;
    i64 bbeg[SM_MAXMUM];
    int *cSA[3], *cLCP[3];
    int cn[3], do_sweep[3];
    int bb_cap;     // candidates per matched interval the sorted fast path of bubble_sort takes (BB_CAP but for tests)
    i64 *out;       // host-mapped: [0..15] header, then rows / members
    i64 out_words;
};

// ordered emission of one child's sweep by the whole block; returns through hdr[3*c..]: nrec, nmem, overflow
__device__ __forceinline__ void small_sweep(const SmallStepArgs &a, int c, u32 *s_scan_a, u32 *s_scan_b, i64 *s_cursor) {
    SweepArgs p;
    p.T = a.T;
    p.SA = a.cSA[c];
    p.LCP = a.cLCP[c];
    p.SO = a.SO;
    p.n = a.cn[c];
    p.nT = a.nT;
    p.nsep0 = a.nsep0;
    p.rc = a.rc;
    p.flavour = 1;
    p.minl = a.minl;
    p.minn = a.minn;
    p.main_nsamples = a.main_nsamples;
    const bool multi = a.main_nsamples > 2;
    const int n = a.cn[c];
    const int chunk = (n + SM_THREADS - 1) / SM_THREADS;
    const int lo = (int)threadIdx.x * chunk;
    const int hi = lo + chunk < n ? lo + chunk : n;
    u32 nr = 0, nm = 0;
    for (int i = lo; i < hi; i++) {
        if (multi) {
            multi_visit(p, (i64)i, [&](i64, i64, i64 size) { nr++; nm += (u32)size; });
        } else {
            i64 l, x, y;
            nr += pair_test(p, (i64)i, l, x, y) ? 1u : 0u;
        }
    }
    u32 tr, tm;
    u32 ir = block_incl_sum<SM_THREADS, u32>(nr, s_scan_a, &tr);
    u32 im = block_incl_sum<SM_THREADS, u32>(nm, s_scan_b, &tm);
    const i64 cur = *s_cursor;
    const i64 need = (i64)tr * 3 + (i64)tm * 2;
    const bool fits = cur + need <= a.out_words;
    if (fits) {
        i64 *rows = a.out + cur;
        i64 *mem = rows + (i64)tr * 3;
        u32 at_r = ir - nr, at_m = im - nm;
        for (int i = lo; i < hi && nr; i++) {
            if (multi) {
                multi_visit(p, (i64)i, [&](i64 l, i64 lb, i64 size) {
                    rows[3 * (i64)at_r + 0] = l;
                    rows[3 * (i64)at_r + 1] = size;
                    rows[3 * (i64)at_r + 2] = (i64)at_m;
                    for (i64 x = 0; x < size; x++) {
                        i64 pos = p.SA[lb + x];
                        mem[2 * (i64)at_m + 0] = sample_of(p, pos);
                        mem[2 * (i64)at_m + 1] = pos;
                        at_m++;
                    }
                    at_r++;
                });
            } else {
                i64 l = 0, x = 0, y = 0;
                if (pair_test(p, (i64)i, l, x, y)) {
                    rows[3 * (i64)at_r + 0] = l;
                    rows[3 * (i64)at_r + 1] = x;
                    rows[3 * (i64)at_r + 2] = y;
                    at_r++;
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.out[4 * c + 0] = tr;
        a.out[4 * c + 1] = tm;
        a.out[4 * c + 2] = fits ? cur : -1;  // word offset of the rows, -1: did not fit (host re-sweeps)
        if (fits) *s_cursor = cur + need;
    }
    __syncthreads();
}

// One block per recursion step: block b works on args[b] (a whole frontier of small sub-indexes goes through ONE launch).
__global__ void __launch_bounds__(SM_THREADS) small_step_kernel(const SmallStepArgs *__restrict__ args) {
    __shared__ SmallStepArgs a;
    {
        const int words = (int)(sizeof(SmallStepArgs) / 4);
        const u32 *src = (const u32 *)(args + blockIdx.x);
        for (int i = (int)threadIdx.x; i < words; i += SM_THREADS) ((u32 *)&a)[i] = src[i];
    }
    __syncthreads();
    __shared__ unsigned char sD[SM_MAXN];
    __shared__ int s_cand[BB_CAP];
    __shared__ int s_cnt[12];
    __shared__ SplitState s_warp[32];
    __shared__ u32 s_scan_a[33], s_scan_b[33];
    __shared__ i64 s_cursor;
    const int tid = (int)threadIdx.x;
    const int n = a.n;
    // ---- labels (reveal.c:1005-1117) ----
    for (int i = tid; i < n; i += SM_THREADS) sD[i] = 0;
    if (tid == 0) s_cursor = 16;
    __syncthreads();
    for (i64 g = tid; g < a.total1; g += SM_THREADS) {
        int k = 0;
        while (k + 1 < a.m1 && a.pre[k + 1] <= g) k++;
        int r = a.SAi[a.ibeg[k] + (g - a.pre[k])];
        if ((unsigned)r < (unsigned)n) sD[r] = a.lab[k];
    }
    __syncthreads();
    const i64 mtotal = (i64)a.m2 * a.mum_l;
    for (i64 g = tid; g < mtotal; g += SM_THREADS) {
        int r = a.SAi[a.mbeg[g / a.mum_l] + g % a.mum_l];
        if ((unsigned)r < (unsigned)n) sD[r] = 3;
    }
    __syncthreads();
    // ---- split (reveal.c:582-664) ----
    {
        const int chunk = (n + SM_THREADS - 1) / SM_THREADS;
        const int lo = tid * chunk, hi = lo + chunk < n ? lo + chunk : n;
        SplitState st = split_identity();
        for (int i = lo; i < hi; i++) {
            int dummy;
            int v = (i == 0 || !(sD[i - 1] >= 1 && sD[i - 1] <= 4)) ? INF_LCP : a.pLCP[i];
            split_step(st, v, class_of(sD[i]), dummy);
        }
        SplitState run = block_excl_scan_state(st, s_warp, nullptr);
        for (int i = lo; i < hi; i++) {
            int cls = class_of(sD[i]);
            int out_m;
            u32 rank = cls >= 0 ? run.cnt[cls] : 0u;
            int v = (i == 0 || !(sD[i - 1] >= 1 && sD[i - 1] <= 4)) ? INF_LCP : a.pLCP[i];
            split_step(run, v, cls, out_m);
            if (cls >= 0 && (int)rank < a.cn[cls]) {
                int s = a.pSA[i];
                a.cSA[cls][rank] = s;
                a.cLCP[cls][rank] = rank == 0 ? 0 : out_m;
                a.SAi[s] = (int)rank;
            }
        }
        // class sizes for the host-side consistency check
        if (tid == SM_THREADS - 1) {
            SplitState fin = run;  // state after the last slot of the last thread = whole range
            a.out[12] = fin.cnt[0];
            a.out[13] = fin.cnt[1];
            a.out[14] = fin.cnt[2];
        }
    }
    __syncthreads();
    // ---- mark the matched bases (reveal.c:1230-1234) ----
    for (i64 g = tid; g < mtotal; g += SM_THREADS) {
        i64 j = a.mbeg[g / a.mum_l] + g % a.mum_l;
        unsigned char c = a.T[j];
        if (c >= 'A' && c <= 'Z') a.T[j] = (unsigned char)(c + 32);
    }
    __syncthreads();
    // ---- bubble_sort of the leading child (reveal.c:1250-1252) ----
    if (a.cn[0] > 0 && a.nb > 0) bubble_block(a.cSA[0], a.cLCP[0], a.SAi, (i64)a.cn[0], a.bbeg, a.nb, s_cand, s_cnt, a.bb_cap);
    __syncthreads();
    // ---- the children's MUM sweeps (reveal.c:802-829 of their own steps) ----
    for (int c = 0; c < 3; c++) {
        if (a.cn[c] > 0 && a.do_sweep[c]) {
            small_sweep(a, c, s_scan_a, s_scan_b, &s_cursor);
        } else if (tid == 0) {
            a.out[4 * c + 0] = 0;
            a.out[4 * c + 1] = 0;
            a.out[4 * c + 2] = -2;  // not swept
        }
    }
}

}  // namespace rv

using namespace rv;

// ---- device buffer pool: children come and go thousands of times per alignment ------------------
struct DevPool {
    // size-class free lists over blocks carved from a few big slabs: a recursion step never calls cudaMalloc
    // (which synchronises the device and costs ~1 ms) except when a new slab is needed
    std::multimap<size_t, void *> free_;
    std::map<void *, size_t> size_;
    std::vector<std::pair<void *, size_t>> slabs_;   // every slab of the pool
    std::vector<std::pair<void *, size_t>> spare_;   // slabs of a previous recursion, not carved yet (see recycle)
    unsigned char *cur_ = nullptr;
    size_t cur_left_ = 0;
    static size_t round_up(size_t b) {
        size_t c = 1024;
        while (c < b) c <<= 1;
        return c;
    }
    int take(size_t bytes, void **out) {
        size_t c = round_up(bytes ? bytes : 1);
        auto it = free_.find(c);
        if (it != free_.end()) {
            *out = it->second;
            free_.erase(it);
            return RV_OK;
        }
        if (c > cur_left_) {
            int best = -1;
            for (int i = 0; i < (int)spare_.size(); i++)
                if (spare_[i].second >= c && (best < 0 || spare_[i].second < spare_[best].second)) best = i;
            if (best >= 0) {
                cur_ = (unsigned char *)spare_[best].first;
                cur_left_ = spare_[best].second;
                spare_.erase(spare_.begin() + best);
            } else {
                size_t slab = (size_t)64 << 20;
                if (slab < 2 * c) slab = 2 * c;
                void *p = nullptr;
                cudaError_t e = cudaMalloc(&p, slab);
                if (e != cudaSuccess && slab > c) {
                    cudaGetLastError();
                    slab = c;
                    e = cudaMalloc(&p, slab);
                }
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    set_error("cudaMalloc(%zu) failed: %s", slab, cudaGetErrorString(e));
                    return RV_ERR_NOMEM;
                }
                // the unused tail of the previous slab is abandoned (at most one block of the largest class seen)
                slabs_.push_back({p, slab});
                cur_ = (unsigned char *)p;
                cur_left_ = slab;
            }
        }
        *out = cur_;
        size_[cur_] = c;
        cur_ += c;
        cur_left_ -= c;
        return RV_OK;
    }
    void give(void *p) {
        if (!p) return;
        free_.insert({size_[p], p});
    }
    size_t bytes() const {
        size_t b = 0;
        for (auto &sl : slabs_) b += sl.second;
        return b;
    }
    // every block is back (the recursion that used the pool is over, its stream synchronised): the slabs wait for the next one
    void recycle() {
        free_.clear();
        size_.clear();
        spare_ = slabs_;
        cur_ = nullptr;
        cur_left_ = 0;
    }
    void destroy() {
        for (auto &sl : slabs_) cudaFree(sl.first);
        slabs_.clear();
        spare_.clear();
        free_.clear();
        size_.clear();
        cur_ = nullptr;
        cur_left_ = 0;
    }
};

// per main index: buffer pool + the host-mapped result buffer of the single-block step
struct RecCtx {
    DevPool pool;
    i64 *h_out = nullptr, *d_out = nullptr;
    i64 out_words = 0;
    SmallStepArgs *h_args = nullptr, *d_args = nullptr;  // host-mapped argument blocks of one batch (SM_BATCH steps)
    struct rv_step_batch *open_ticket = nullptr;         // rv_sub_step_batch_begin without its _end yet
    int device = 0;
    // step statistics: [0] single-launch path, [1] general path
    long long steps[2] = {0, 0};
    long long launches[2] = {0, 0};
    double host_s[2] = {0, 0};
};
#include <chrono>
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct rv_sub {
    rv_index *main;
    int *SA = nullptr, *LCP = nullptr;
    i64 n = 0;
    bool owns = false;  // false: the root view over the main index's arrays
    // MUM sweep already done by the step that created this child (small_step_kernel)
    bool cached = false;
    int cminl = 0, cminn = 0;
    std::vector<i64> rows, members;
};

// accessors into rv_index implemented in rv_api.cu
namespace rv {
struct MainView {
    Stream *st;
    unsigned char *T;
    int *SA, *ISA, *LCP;
    unsigned short *SO;
    i64 n, nsep0;
    int nsamples, rc;
    const i64 *nsep_host;  // the separators (nsamples - 1 of them), host memory of the handle
    int nsep_count;
    int device;            // the GPU the handle lives on
    void **pool_slot;  // where the handle keeps its DevPool*
};
int main_view(rv_index *h, MainView *out);
int sub_sweep_pair(rv_index *h, const SweepArgs &a, int64_t *count);
int sub_sweep_multi(rv_index *h, const SweepArgs &a, int64_t *nrec, int64_t *nmem);
}  // namespace rv

// The recursion context of an index -- device slabs of the children's arrays, 32 MB of host-mapped result buffer, the argument
// blocks -- outlives the index in a small cache: one `rem` per index object would otherwise pay a dozen cudaMalloc / cudaFree and
// two cudaHostAlloc per alignment (tens of milliseconds, and serialised in the driver between the ranks of a box, which all
// reach the top of their recursion at the same moment).
#include <mutex>
static std::mutex g_ctx_mu;
static std::vector<RecCtx *> g_ctx_cache;
static const size_t CTX_SLOTS = 2;
static const size_t CTX_MAX_BYTES = (size_t)16 << 30;

static RecCtx *ctx_of(MainView &v) {
    if (!*v.pool_slot) {
        const int dev = v.device;
        RecCtx *c = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_ctx_mu);
            for (size_t i = 0; i < g_ctx_cache.size(); i++)
                if (g_ctx_cache[i]->device == dev) {
                    c = g_ctx_cache[i];
                    g_ctx_cache.erase(g_ctx_cache.begin() + (long)i);
                    break;
                }
        }
        if (!c) {
            c = new RecCtx();
            c->device = dev;
        }
        *v.pool_slot = c;
    }
    return (RecCtx *)*v.pool_slot;
}
static void ctx_destroy(RecCtx *c) {
    c->pool.destroy();
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_args) cudaFreeHost(c->h_args);
    delete c;
}
static DevPool *pool_of(MainView &v) { return &ctx_of(v)->pool; }

extern "C" {

void rv_pool_destroy(void *pool) {  // called by rv_index_free, after the index's stream was synchronised
    if (!pool) return;
    RecCtx *c = (RecCtx *)pool;
    if (!c->open_ticket && c->pool.bytes() <= CTX_MAX_BYTES) {
        c->pool.recycle();
        for (int k = 0; k < 2; k++) c->steps[k] = c->launches[k] = 0, c->host_s[k] = 0;
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (g_ctx_cache.size() < CTX_SLOTS) {
            g_ctx_cache.push_back(c);
            return;
        }
    }
    ctx_destroy(c);
}
void rv_pool_trim(void) {  // rv_trim: give the cached recursion contexts back
    std::vector<RecCtx *> drop;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        drop.swap(g_ctx_cache);
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (RecCtx *c : drop) {
        cudaSetDevice(c->device);
        ctx_destroy(c);
    }
    cudaSetDevice(cur);
}

int rv_sub_root(rv_index *h, rv_sub **out) {
    MainView v;
    RV_TRY(main_view(h, &v));
    if (!out) return RV_ERR_ARG;
    rv_sub *s = new rv_sub();
    s->main = h;
    s->SA = v.SA;
    s->LCP = v.LCP;
    s->n = v.n;
    s->owns = false;
    *out = s;
    return RV_OK;
}

int64_t rv_sub_n(const rv_sub *s) { return s ? s->n : 0; }

void rv_sub_free(rv_sub *s) {
    if (!s) return;
    if (s->owns) {
        MainView v;
        if (main_view(s->main, &v) == RV_OK) {
            DevPool *p = pool_of(v);
            p->give(s->SA);
            p->give(s->LCP);
        }
    }
    delete s;
}

int rv_sub_get(rv_sub *s, int32_t which, int32_t *out) {  // which: 0 SA, 1 LCP (int32 entries), for tests and getters
    if (!s || !out) return RV_ERR_ARG;
    MainView v;
    RV_TRY(main_view(s->main, &v));
    if (s->n > 0) RV_CUDA(cudaMemcpyAsync(out, which == 0 ? s->SA : s->LCP, (size_t)s->n * 4, cudaMemcpyDeviceToHost, v.st->s));
    RV_CUDA(cudaStreamSynchronize(v.st->s));
    return RV_OK;
}

static SweepArgs sub_args(const rv_sub *s, const MainView &v) {
    SweepArgs a;
    a.T = v.T;
    a.SA = s->SA;
    a.LCP = s->LCP;
    a.SO = v.SO;
    a.n = s->n;
    a.nT = v.n;
    a.nsep0 = v.nsep0;
    a.rc = v.rc;
    a.flavour = 1;  // getmums_rem: the aligner's pair sweep (reveal.c:818-822)
    a.minl = 0;
    a.minn = 2;
    a.main_nsamples = v.nsamples;
    if (v.nsamples > 2 && v.nsep_count <= SW_NSEP_INLINE) {
        a.nsep_n = v.nsep_count;
        for (int k = 0; k < a.nsep_n; k++) a.nsep_v[k] = v.nsep_host[k];
    }
    return a;
}

int rv_sub_mums_pair(rv_sub *s, int32_t minl, int64_t *count) {
    if (!s || !count) return RV_ERR_ARG;
    if (s->cached && s->cminl == minl) {
        *count = (int64_t)(s->rows.size() / 3);
        return RV_OK;
    }
    s->cached = false;
    MainView v;
    RV_TRY(main_view(s->main, &v));
    SweepArgs a = sub_args(s, v);
    a.minl = minl;
    return sub_sweep_pair(s->main, a, count);
}

int rv_sub_mums_multi(rv_sub *s, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem) {
    if (!s || !nrec || !nmem) return RV_ERR_ARG;
    if (s->cached && s->cminl == minl && s->cminn == minn) {
        *nrec = (int64_t)(s->rows.size() / 3);
        *nmem = (int64_t)(s->members.size() / 2);
        return RV_OK;
    }
    s->cached = false;
    MainView v;
    RV_TRY(main_view(s->main, &v));
    SweepArgs a = sub_args(s, v);
    a.minl = minl;
    a.minn = minn;
    return sub_sweep_multi(s->main, a, nrec, nmem);
}

// One recursion step on the device.  Intervals are (begin, end) pairs of text positions, end exclusive.
// mum_sp[0..mum_n) are the start positions of the chosen MUM, mum_l its length; matching[] the (begin,end)
// pairs graphalign returned, in its iteration order (bubble_sort replays them in that order).
// children[0..2] = leading, trailing, parallel (NULL when that class is empty).
static int split_general(rv_sub *parent, const int64_t *lead, int32_t nlead, const int64_t *trail, int32_t ntrail, const int64_t *par, int32_t npar,
                         const int64_t *mum_sp, int32_t mum_n, int64_t mum_l, const int64_t *matching, int32_t nmatch, rv_sub **children) {
    if (!parent || !children) return RV_ERR_ARG;
    children[0] = children[1] = children[2] = nullptr;
    MainView v;
    RV_TRY(main_view(parent->main, &v));
    DevPool *pool = pool_of(v);
    Stream &st = *v.st;
    const i64 n = parent->n;
    if (n <= 0) return RV_OK;
    // scratch blocks go back to the pool and unfinished children are released on EVERY way out of this function
    struct Lease {
        DevPool *pool;
        std::vector<void *> blocks;
        rv_sub *kids[3] = {nullptr, nullptr, nullptr};
        int take(size_t bytes, void **out) {
            int r = pool->take(bytes, out);
            if (r == RV_OK) blocks.push_back(*out);
            return r;
        }
        ~Lease() {
            for (void *p : blocks) pool->give(p);  // stream-ordered reuse: later users are enqueued behind the kernels above
            for (int q = 0; q < 3; q++) rv_sub_free(kids[q]);
        }
    } lease;
    lease.pool = pool;

    // ---- flatten the intervals on the host (host logic: counts like reveal.c:1017-1117) ----
    std::vector<i64> ibeg, pre;
    std::vector<unsigned char> lab;
    i64 total = 0, cls_n[3] = {0, 0, 0};
    const int64_t *src[3] = {lead, trail, par};
    const int32_t cnts[3] = {nlead, ntrail, npar};
    const unsigned char labels[3] = {1, 2, 4};
    for (int c = 0; c < 3; c++)
        for (int k = 0; k < cnts[c]; k++) {
            i64 b = src[c][2 * k], e = src[c][2 * k + 1];
            if (e <= b) continue;
            if (b < 0 || e > v.n) { set_error("rv_sub_split: interval out of range"); return RV_ERR_ARG; }
            ibeg.push_back(b);
            pre.push_back(total);
            lab.push_back(labels[c]);
            total += e - b;
            cls_n[c] += e - b;
        }
    const int m1 = (int)ibeg.size();
    const i64 total1 = total;
    i64 mtotal = 0;
    std::vector<i64> mbeg, mpre;
    for (int k = 0; k < mum_n; k++) {
        if (mum_l <= 0) break;
        if (mum_sp[k] < 0 || mum_sp[k] + mum_l > v.n) { set_error("rv_sub_split: mum out of range"); return RV_ERR_ARG; }
        mbeg.push_back(mum_sp[k]);
        mpre.push_back(mtotal);
        mtotal += mum_l;
    }
    const int m2 = (int)mbeg.size();
    std::vector<i64> bbeg;
    for (int k = 0; k < nmatch; k++) bbeg.push_back(matching[2 * k]);

    // ---- upload the small tables in one buffer ----
    size_t words = (size_t)2 * m1 + 2 * m2 + bbeg.size() + 8;
    size_t bytes = (words * 8 + (size_t)m1 + 64 + 15) / 16 * 16;  // d_counts lives in the (aligned) last 32 bytes
    void *dtab = nullptr;
    RV_TRY(lease.take(bytes, &dtab));
    std::vector<unsigned char> host(bytes, 0);
    i64 *hw = (i64 *)host.data();
    i64 *h_ibeg = hw, *h_pre = hw + m1, *h_mbeg = hw + 2 * m1, *h_mpre = hw + 2 * m1 + m2, *h_bbeg = hw + 2 * m1 + 2 * m2;
    for (int k = 0; k < m1; k++) { h_ibeg[k] = ibeg[k]; h_pre[k] = pre[k]; }
    for (int k = 0; k < m2; k++) { h_mbeg[k] = mbeg[k]; h_mpre[k] = mpre[k]; }
    for (size_t k = 0; k < bbeg.size(); k++) h_bbeg[k] = bbeg[k];
    unsigned char *h_lab = host.data() + words * 8;
    for (int k = 0; k < m1; k++) h_lab[k] = lab[k];
    RV_CUDA(cudaMemcpyAsync(dtab, host.data(), bytes, cudaMemcpyHostToDevice, st.s));
    i64 *d_w = (i64 *)dtab;
    const i64 *d_ibeg = d_w, *d_pre = d_w + m1, *d_mbeg = d_w + 2 * m1, *d_mpre = d_w + 2 * m1 + m2, *d_bbeg = d_w + 2 * m1 + 2 * m2;
    const unsigned char *d_lab = (const unsigned char *)dtab + words * 8;

    // ---- D labels ----
    void *dD = nullptr;
    RV_TRY(lease.take((size_t)n, &dD));
    unsigned char *D = (unsigned char *)dD;
    RV_CUDA(cudaMemsetAsync(D, 0, (size_t)n, st.s));
    if (total1 > 0) {
        RV_LAUNCH(label_kernel, (unsigned)((total1 + 255) / 256), 256, 0, st.s, d_ibeg, d_pre, d_lab, m1, total1, v.ISA, D, n);
        st.launches++;
    }
    if (mtotal > 0) {  // matched bases override (reveal.c:1112-1116)
        void *d3 = nullptr;
        RV_TRY(lease.take((size_t)m2 + 64, &d3));
        RV_CUDA(cudaMemsetAsync(d3, 3, (size_t)m2, st.s));
        RV_LAUNCH(label_kernel, (unsigned)((mtotal + 255) / 256), 256, 0, st.s, d_mbeg, d_mpre, (const unsigned char *)d3, m2, mtotal, v.ISA, D, n);
        st.launches++;
    }

    // ---- children ----
    rv_sub **kids = lease.kids;
    for (int c = 0; c < 3; c++)
        if (cls_n[c] > 0) {
            rv_sub *k = new rv_sub();
            k->main = parent->main;
            k->n = cls_n[c];
            k->owns = true;
            kids[c] = k;  // from here on the lease frees it (rv_sub_free gives its arrays back to the pool)
            void *p1 = nullptr, *p2 = nullptr;
            RV_TRY(pool->take((size_t)(cls_n[c] + 2) * 4, &p1));
            k->SA = (int *)p1;
            RV_TRY(pool->take((size_t)(cls_n[c] + 2) * 4, &p2));
            k->LCP = (int *)p2;
        }
    const i64 ntiles = (n + SP_TILE - 1) / SP_TILE;
    void *dtiles = nullptr;
    RV_TRY(lease.take((size_t)ntiles * sizeof(SplitState) + 64, &dtiles));
    SplitState *tiles = (SplitState *)dtiles;
    u32 *d_counts = (u32 *)((unsigned char *)dtab + bytes - 32);
    RV_LAUNCH(split_reduce_kernel, (unsigned)ntiles, SP_THREADS, 0, st.s, parent->LCP, D, n, tiles);
    RV_LAUNCH(split_tilescan_kernel, 1, TS_THREADS, 0, st.s, tiles, ntiles, d_counts);
    RV_LAUNCH(split_apply_kernel, (unsigned)ntiles, SP_THREADS, 0, st.s, parent->SA, parent->LCP, D, n, tiles, v.ISA,
              kids[0] ? kids[0]->SA : nullptr, kids[0] ? kids[0]->LCP : nullptr, kids[1] ? kids[1]->SA : nullptr,
              kids[1] ? kids[1]->LCP : nullptr, kids[2] ? kids[2]->SA : nullptr, kids[2] ? kids[2]->LCP : nullptr, (u32)cls_n[0], (u32)cls_n[1],
              (u32)cls_n[2]);
    st.launches += 3;
    // the labelled positions must be exactly the parent's suffixes of each class: check the device counts
    u32 *hc = st.pinned + 300;  // pinned: the copy is really asynchronous, the kernels below are enqueued behind it at once
    RV_CUDA(cudaMemcpyAsync(hc, d_counts, 12, cudaMemcpyDeviceToHost, st.s));

    // ---- mark the matched bases in T (reveal.c:1230-1234) ----
    if (mtotal > 0) {
        RV_LAUNCH(lower_kernel, (unsigned)((mtotal + 255) / 256), 256, 0, st.s, d_mbeg, d_mpre, m2, mtotal, v.T);
        st.launches++;
    }
    // ---- bubble_sort(i_leading, matching_intervals)  (reveal.c:1250-1252) ----
    void *dcand = nullptr;
    if (kids[0] && !bbeg.empty()) {
        i64 block_maxn = 8192;
        if (const char *e = getenv("RV_BUBBLE_BLOCK_MAXN")) block_maxn = atoll(e);  // test hook
        if (kids[0]->n <= block_maxn) {
            RV_LAUNCH(bubble_kernel, 1, BB_THREADS, 0, st.s, kids[0]->SA, kids[0]->LCP, v.ISA, kids[0]->n, d_bbeg, (int)bbeg.size(), bubble_cap());
            st.launches++;
        } else {
            RV_TRY(lease.take((size_t)(BB_CAP + 16) * 4, &dcand));
            int *cand = (int *)dcand, *cand_cnt = cand + BB_CAP;
            RV_CUDA(cudaMemsetAsync(cand_cnt, 0, 4, st.s));
            i64 blocks = (kids[0]->n + 255) / 256;
            if (blocks > 148 * 8) blocks = 148 * 8;
            for (int b = 0; b < (int)bbeg.size(); b++) {
                RV_LAUNCH(bubble_detect_kernel, (unsigned)blocks, 256, 0, st.s, kids[0]->SA, kids[0]->LCP, kids[0]->n, d_bbeg, b, cand, cand_cnt, bubble_cap());
                RV_LAUNCH(bubble_apply_kernel, 1, BB_THREADS, 0, st.s, kids[0]->SA, kids[0]->LCP, v.ISA, kids[0]->n, d_bbeg, b, cand, cand_cnt, bubble_cap());
                st.launches += 2;
            }
        }
    }
    RV_CUDA(cudaStreamSynchronize(st.s));
    RV_KCHECK();
    for (int c = 0; c < 3; c++)
        if ((i64)hc[c] != cls_n[c]) {
            set_error("rv_sub_split: class %d has %u suffixes in the parent but the intervals cover %lld positions", c, hc[c], (long long)cls_n[c]);
            return RV_ERR_ARG;
        }
    for (int c = 0; c < 3; c++) {
        children[c] = kids[c];
        kids[c] = nullptr;  // handed over
    }
    return RV_OK;
}

// ---- the single-launch path for small parents, one or many steps per launch -----------------------------------------------------
static const int SM_BATCH = 256;  // steps per launch at most

struct SmallPrep {   // host side of one step between prepare and collect
    rv_sub *kids[3] = {nullptr, nullptr, nullptr};
    i64 cls_n[3] = {0, 0, 0};
    i64 out_off = 0, out_words = 0;   // this step's share of the host-mapped result buffer
};

static int small_ctx(MainView &v, RecCtx **out) {
    RecCtx *ctx = ctx_of(v);
    if (!ctx->h_out) {
        const size_t bytes = (size_t)32 << 20;
        void *hp = nullptr, *dp = nullptr;
        RV_CUDA(cudaHostAlloc(&hp, bytes, cudaHostAllocMapped));
        RV_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
        ctx->h_out = (i64 *)hp;
        ctx->d_out = (i64 *)dp;
        ctx->out_words = (i64)(bytes / 8);
    }
    if (!ctx->h_args) {
        void *hp = nullptr, *dp = nullptr;
        RV_CUDA(cudaHostAlloc(&hp, sizeof(SmallStepArgs) * SM_BATCH, cudaHostAllocMapped));
        RV_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
        ctx->h_args = (SmallStepArgs *)hp;
        ctx->d_args = (SmallStepArgs *)dp;
    }
    *out = ctx;
    return RV_OK;
}

// fills `a` and allocates the children; *handled = 0 when the step does not qualify for the single-block path
static int small_prepare(const rv_step_desc &d, MainView &v, RecCtx *ctx, int32_t minl, int32_t minn, SmallStepArgs &a, SmallPrep &pp, int *handled) {
    *handled = 0;
    rv_sub *parent = d.parent;
    const i64 n = parent->n;
    i64 small_maxn = SM_MAXN;
    if (const char *e = getenv("RV_SMALL_MAXN")) {  // test hook: 0 forces the general path
        i64 x = atoll(e);
        if (x < small_maxn) small_maxn = x;
    }
    if (n <= 0 || n > small_maxn || d.mum_n > SM_MAXMUM || d.nmatch > SM_MAXMUM || d.mum_l <= 0) return RV_OK;
    memset(&a, 0, sizeof a);
    i64 total = 0;
    const int64_t *src[3] = {d.lead, d.trail, d.par};
    const int32_t cnts[3] = {d.nlead, d.ntrail, d.npar};
    const unsigned char labels[3] = {1, 2, 4};
    int m1 = 0;
    for (int c = 0; c < 3; c++)
        for (int k = 0; k < cnts[c]; k++) {
            i64 b = src[c][2 * k], e = src[c][2 * k + 1];
            if (e <= b) continue;
            if (b < 0 || e > v.n) { set_error("rv_sub_step: interval out of range"); return RV_ERR_ARG; }
            if (m1 >= SM_MAXIV) return RV_OK;  // too many intervals for the argument block: general path
            a.ibeg[m1] = b;
            a.pre[m1] = total;
            a.lab[m1] = labels[c];
            m1++;
            total += e - b;
            pp.cls_n[c] += e - b;
        }
    for (int k = 0; k < d.mum_n; k++) {
        if (d.mum_sp[k] < 0 || d.mum_sp[k] + d.mum_l > v.n) { set_error("rv_sub_step: mum out of range"); return RV_ERR_ARG; }
        a.mbeg[k] = d.mum_sp[k];
    }
    for (int k = 0; k < d.nmatch; k++) a.bbeg[k] = d.matching[2 * k];
    DevPool *pool = &ctx->pool;
    for (int c = 0; c < 3; c++)
        if (pp.cls_n[c] > 0) {
            rv_sub *k = new rv_sub();
            k->main = parent->main;
            k->n = pp.cls_n[c];
            k->owns = true;
            pp.kids[c] = k;
            void *p1 = nullptr, *p2 = nullptr;
            int r1 = pool->take((size_t)(pp.cls_n[c] + 2) * 4, &p1);
            k->SA = (int *)p1;
            int r2 = r1 == RV_OK ? pool->take((size_t)(pp.cls_n[c] + 2) * 4, &p2) : r1;
            k->LCP = (int *)p2;
            if (r1 != RV_OK || r2 != RV_OK) {
                for (int q = 0; q < 3; q++) { rv_sub_free(pp.kids[q]); pp.kids[q] = nullptr; }
                return RV_ERR_NOMEM;
            }
        }
    a.T = v.T;
    a.SAi = v.ISA;
    a.SO = v.SO;
    a.nT = v.n;
    a.nsep0 = v.nsep0;
    a.main_nsamples = v.nsamples;
    a.rc = v.rc;
    a.minl = minl;
    a.minn = minn;
    a.bb_cap = bubble_cap();
    a.pSA = parent->SA;
    a.pLCP = parent->LCP;
    a.n = (int)n;
    a.m1 = m1;
    a.total1 = total;
    a.m2 = d.mum_n;
    a.mum_l = d.mum_l;
    a.nb = d.nmatch;
    for (int c = 0; c < 3; c++) {
        a.cSA[c] = pp.kids[c] ? pp.kids[c]->SA : nullptr;
        a.cLCP[c] = pp.kids[c] ? pp.kids[c]->LCP : nullptr;
        a.cn[c] = (int)pp.cls_n[c];
        a.do_sweep[c] = d.sweep[c] ? 1 : 0;
    }
    *handled = 1;
    return RV_OK;
}

// after the launch + synchronisation: checks the class counts, hands the children (with their stored sweep results) over
static int small_collect(rv_step_desc &d, RecCtx *ctx, const SmallPrep &pp, int32_t minl, int32_t minn) {
    const i64 *o = ctx->h_out + pp.out_off;
    for (int c = 0; c < 3; c++)
        if (o[12 + c] != pp.cls_n[c]) {
            set_error("rv_sub_step: class %d has %lld suffixes in the parent but the intervals cover %lld positions", c, (long long)o[12 + c],
                      (long long)pp.cls_n[c]);
            for (int q = 0; q < 3; q++) rv_sub_free(pp.kids[q]);
            return RV_ERR_ARG;
        }
    for (int c = 0; c < 3; c++) {
        rv_sub *k = pp.kids[c];
        if (!k) continue;
        const i64 nr = o[4 * c + 0], nm = o[4 * c + 1], off = o[4 * c + 2];
        if (off >= 0) {  // swept and it fitted
            k->rows.assign(o + off, o + off + nr * 3);
            k->members.assign(o + off + nr * 3, o + off + nr * 3 + nm * 2);
            k->cached = true;
            k->cminl = minl;
            k->cminn = minn;
        }
        d.children[c] = k;
    }
    return RV_OK;
}

static int split_general(rv_sub *parent, const int64_t *lead, int32_t nlead, const int64_t *trail, int32_t ntrail, const int64_t *par, int32_t npar,
                         const int64_t *mum_sp, int32_t mum_n, int64_t mum_l, const int64_t *matching, int32_t nmatch, rv_sub **children);

// Recursion steps of one main index, as many as the caller has ready (a frontier of independent sub-indexes, reveal.c:1296-1324
// pushes up to three per step): the steps whose parent fits a thread block go through ONE launch and ONE synchronisation, one
// block per step; the others take the general path one after the other.  Every step reports its own status.
//
// Two halves so that a caller can overlap the device part of one batch with the host part (callbacks) of the next one:
//   rv_sub_step_batch_begin  checks and stages the steps, enqueues the launch and returns a ticket; when the batch needs several
//                            launches (more than SM_BATCH block-sized steps) only the last one is left in flight
//   rv_sub_step_batch_end    waits for that launch (an event, not the whole stream), hands children and statuses over, runs
//                            the general-path steps, frees the ticket
// `steps` must stay valid and untouched in between, and at most ONE ticket of a main index may be open at a time: the argument
// blocks and the result buffer of the launch are single host-mapped arrays of the index.
struct rv_step_batch {
    rv_step_desc *steps = nullptr;
    int32_t nsteps = 0, minl = 0, minn = 0;
    MainView v;
    RecCtx *ctx = nullptr;
    int worst = RV_OK;
    std::vector<int> general, which;
    std::vector<SmallPrep> prep;
    cudaEvent_t done = 0;
    bool launched = false;
    double t0 = 0;
};

static int batch_collect(rv_step_batch *b) {
    RecCtx *ctx = b->ctx;
    const int m = (int)b->which.size();
    for (int k = 0; k < m; k++) {
        int r = small_collect(b->steps[b->which[k]], ctx, b->prep[k], b->minl, b->minn);
        if (r != RV_OK) {
            b->steps[b->which[k]].status = r;
            b->worst = r;
        }
    }
    ctx->steps[0] += m;
    ctx->launches[0]++;
    ctx->host_s[0] += now_s() - b->t0;
    b->which.clear();
    b->prep.clear();
    b->launched = false;
    return RV_OK;
}

int rv_sub_step_batch_begin(rv_step_desc *steps, int32_t nsteps, int32_t minl, int32_t minn, rv_step_batch **ticket) {
    if (!ticket) return RV_ERR_ARG;
    *ticket = nullptr;
    if (!steps || nsteps < 0) return RV_ERR_ARG;
    for (int i = 0; i < nsteps; i++) {
        if (!steps[i].parent || steps[i].parent->main != steps[0].parent->main) { set_error("rv_sub_step_batch: steps of different indexes"); return RV_ERR_ARG; }
        steps[i].children[0] = steps[i].children[1] = steps[i].children[2] = nullptr;
        steps[i].status = RV_OK;
    }
    rv_step_batch *b = new rv_step_batch();
    b->steps = steps;
    b->nsteps = nsteps;
    b->minl = minl;
    b->minn = minn;
    *ticket = b;
    if (nsteps == 0) return RV_OK;
    int r0 = main_view(steps[0].parent->main, &b->v);
    if (r0 == RV_OK) r0 = small_ctx(b->v, &b->ctx);
    if (r0 == RV_OK && b->ctx->open_ticket) {
        set_error("rv_sub_step_batch_begin: another batch of this index is still open");
        r0 = RV_ERR_STATE;
    }
    if (r0 != RV_OK) {
        delete b;
        *ticket = nullptr;
        return r0;
    }
    RecCtx *ctx = b->ctx;
    Stream &st = *b->v.st;
    for (int base = 0; base < nsteps;) {
        if (b->launched) {  // a batch of more than SM_BATCH block-sized steps: the argument blocks are needed again
            cudaError_t e = cudaEventSynchronize(b->done);
            if (e != cudaSuccess) { set_error("rv_sub_step_batch: %s", cudaGetErrorString(e)); b->worst = RV_ERR_CUDA; b->launched = false; break; }
            batch_collect(b);
        }
        // ---- one launch: the next run of (at most SM_BATCH) steps ----
        b->t0 = now_s();
        int i = base;
        for (; i < nsteps && (int)b->which.size() < SM_BATCH; i++) {
            SmallPrep pp;
            int handled = 0;
            int r = small_prepare(steps[i], b->v, ctx, minl, minn, ctx->h_args[b->which.size()], pp, &handled);
            if (r != RV_OK) {
                steps[i].status = r;
                b->worst = r;
                continue;
            }
            if (!handled) {
                b->general.push_back(i);
                continue;
            }
            b->which.push_back(i);
            b->prep.push_back(pp);
        }
        base = i;
        const int m = (int)b->which.size();
        if (m > 0) {
            // every step gets an equal share of the host-mapped result buffer (a child whose MUMs do not fit is swept again later)
            const i64 share = (ctx->out_words / m) & ~(i64)7;
            for (int k = 0; k < m; k++) {
                b->prep[k].out_off = share * k;
                b->prep[k].out_words = share;
                ctx->h_args[k].out = ctx->d_out + share * k;
                ctx->h_args[k].out_words = share;
            }
            cudaError_t e = cudaSuccess;
            if (!b->done) e = cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming);
            if (e == cudaSuccess) {
                RV_LAUNCH(small_step_kernel, (unsigned)m, SM_THREADS, 0, st.s, (const SmallStepArgs *)ctx->d_args);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaEventRecord(b->done, st.s);
            if (e != cudaSuccess) {
                set_error("rv_sub_step_batch: launch failed: %s", cudaGetErrorString(e));
                for (int k = 0; k < m; k++) {
                    for (int q = 0; q < 3; q++) rv_sub_free(b->prep[k].kids[q]);
                    steps[b->which[k]].status = RV_ERR_CUDA;
                }
                b->which.clear();
                b->prep.clear();
                b->worst = RV_ERR_CUDA;
                break;
            }
            st.launches++;
            b->launched = true;
        }
    }
    ctx->open_ticket = b;
    return RV_OK;
}

int rv_sub_step_batch_end(rv_step_batch *b) {
    if (!b) return RV_ERR_ARG;
    if (b->nsteps == 0 || !b->ctx) {
        delete b;
        return RV_OK;
    }
    RecCtx *ctx = b->ctx;
    if (b->launched) {
        cudaError_t e = cudaEventSynchronize(b->done);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("rv_sub_step_batch: %s", cudaGetErrorString(e));
            for (size_t k = 0; k < b->which.size(); k++) {
                for (int q = 0; q < 3; q++) rv_sub_free(b->prep[k].kids[q]);
                b->steps[b->which[k]].status = RV_ERR_CUDA;
            }
            b->worst = RV_ERR_CUDA;
            b->launched = false;
        } else {
            batch_collect(b);
        }
    }
    for (int i : b->general) {
        const double t0 = now_s();
        rv_step_desc &d = b->steps[i];
        int r = split_general(d.parent, d.lead, d.nlead, d.trail, d.ntrail, d.par, d.npar, d.mum_sp, d.mum_n, d.mum_l, d.matching, d.nmatch, d.children);
        if (r != RV_OK) {
            d.status = r;
            b->worst = r;
        }
        ctx->steps[1]++;
        ctx->launches[1]++;
        ctx->host_s[1] += now_s() - t0;
    }
    const int worst = b->worst;
    if (b->done) cudaEventDestroy(b->done);
    ctx->open_ticket = nullptr;
    delete b;
    return worst;
}

int rv_sub_step_batch(rv_step_desc *steps, int32_t nsteps, int32_t minl, int32_t minn) {
    rv_step_batch *b = nullptr;
    RV_TRY(rv_sub_step_batch_begin(steps, nsteps, minl, minn, &b));
    return rv_sub_step_batch_end(b);
}

int rv_sub_step(rv_sub *parent, const int64_t *lead, int32_t nlead, const int64_t *trail, int32_t ntrail, const int64_t *par, int32_t npar,
                const int64_t *mum_sp, int32_t mum_n, int64_t mum_l, const int64_t *matching, int32_t nmatch, const int32_t *sweep, int32_t minl,
                int32_t minn, rv_sub **children) {
    if (!parent || !children) return RV_ERR_ARG;
    rv_step_desc d;
    memset(&d, 0, sizeof d);
    d.parent = parent;
    d.lead = lead; d.nlead = nlead;
    d.trail = trail; d.ntrail = ntrail;
    d.par = par; d.npar = npar;
    d.mum_sp = mum_sp; d.mum_n = mum_n; d.mum_l = mum_l;
    d.matching = matching; d.nmatch = nmatch;
    for (int c = 0; c < 3; c++) d.sweep[c] = sweep ? sweep[c] : 0;
    int r = rv_sub_step_batch(&d, 1, minl, minn);
    for (int c = 0; c < 3; c++) children[c] = d.children[c];
    return r;
}

int rv_sub_split(rv_sub *parent, const int64_t *lead, int32_t nlead, const int64_t *trail, int32_t ntrail, const int64_t *par, int32_t npar,
                 const int64_t *mum_sp, int32_t mum_n, int64_t mum_l, const int64_t *matching, int32_t nmatch, rv_sub **children) {
    return rv_sub_step(parent, lead, nlead, trail, ntrail, par, npar, mum_sp, mum_n, mum_l, matching, nmatch, nullptr, 0, 2, children);
}

// recursion statistics of a main index: steps and host seconds of the single-launch / general step paths
int rv_rec_stats(rv_index *h, int64_t *steps2, double *seconds2) {
    MainView v;
    RV_TRY(main_view(h, &v));
    RecCtx *ctx = ctx_of(v);
    for (int k = 0; k < 2; k++) {
        if (steps2) steps2[k] = ctx->steps[k];
        if (seconds2) seconds2[k] = ctx->host_s[k];
    }
    return RV_OK;
}

// extract (reveal.c:1386-1505): takes the text positions of `intervals` out of the sub-index IN PLACE -- their suffixes are
// dropped from SA / LCP (the LCP of a survivor is the minimum over the dropped run before it, :1452-1478), the inverse is
// rewritten to the new ranks, the bases are lower-cased (:1432) and bubble_sort runs on the result with the intervals (:1497).
// It is the split of rv_sub_step with ONE class: every slot is labelled "keep" first, the intervals then "matched".
// Slot 0: the reference starts its copy loop at slot 1 and leaves SA[0] of the result uninitialised (:1448); here slot 0
// holds the suffix that belongs there.
int rv_sub_extract(rv_sub *sub, const int64_t *intervals, int32_t nintervals) {
    if (!sub || (nintervals > 0 && !intervals) || nintervals < 0) return RV_ERR_ARG;
    MainView v;
    RV_TRY(main_view(sub->main, &v));
    DevPool *pool = pool_of(v);
    Stream &st = *v.st;
    const i64 n = sub->n;
    if (n <= 0 || nintervals == 0) return RV_OK;
    struct Lease {
        DevPool *pool;
        std::vector<void *> blocks;
        int take(size_t bytes, void **out) {
            int r = pool->take(bytes, out);
            if (r == RV_OK) blocks.push_back(*out);
            return r;
        }
        ~Lease() {
            for (void *p : blocks) pool->give(p);
        }
    } lease;
    lease.pool = pool;
    std::vector<i64> ibeg, pre, bbeg;
    i64 total = 0;
    for (int k = 0; k < nintervals; k++) {
        const i64 b = intervals[2 * k], e = intervals[2 * k + 1];
        if (b < 0 || e > v.n || e < b) { set_error("rv_sub_extract: interval out of range"); return RV_ERR_ARG; }
        bbeg.push_back(b);
        if (e == b) continue;
        ibeg.push_back(b);
        pre.push_back(total);
        total += e - b;
    }
    const int m = (int)ibeg.size();
    if (total > n) { set_error("rv_sub_extract: the intervals cover more positions than the index has suffixes"); return RV_ERR_ARG; }
    const i64 cn = n - total;
    // tables: [ibeg m][pre m][bbeg nb] then m label bytes, d_counts in the last 32 bytes
    const size_t words = (size_t)2 * m + bbeg.size() + 8;
    const size_t bytes = (words * 8 + (size_t)m + 64 + 15) / 16 * 16;
    void *dtab = nullptr;
    RV_TRY(lease.take(bytes, &dtab));
    std::vector<unsigned char> host(bytes, 0);
    i64 *hw = (i64 *)host.data();
    for (int k = 0; k < m; k++) { hw[k] = ibeg[k]; hw[m + k] = pre[k]; }
    for (size_t k = 0; k < bbeg.size(); k++) hw[2 * m + k] = bbeg[k];
    memset(host.data() + words * 8, 3, (size_t)m);
    RV_CUDA(cudaMemcpyAsync(dtab, host.data(), bytes, cudaMemcpyHostToDevice, st.s));
    const i64 *d_ibeg = (const i64 *)dtab, *d_pre = d_ibeg + m, *d_bbeg = d_ibeg + 2 * m;
    const unsigned char *d_lab = (const unsigned char *)dtab + words * 8;
    u32 *d_counts = (u32 *)((unsigned char *)dtab + bytes - 32);
    void *dD = nullptr;
    RV_TRY(lease.take((size_t)n, &dD));
    unsigned char *D = (unsigned char *)dD;
    RV_CUDA(cudaMemsetAsync(D, 1, (size_t)n, st.s));  // every slot: class 0 ("keep")
    if (total > 0) {
        RV_LAUNCH(label_kernel, (unsigned)((total + 255) / 256), 256, 0, st.s, d_ibeg, d_pre, d_lab, m, total, v.ISA, D, n);
        st.launches++;
    }
    void *p1 = nullptr, *p2 = nullptr;
    RV_TRY(pool->take((size_t)(cn + 2) * 4, &p1));
    if (pool->take((size_t)(cn + 2) * 4, &p2) != RV_OK) { pool->give(p1); return RV_ERR_NOMEM; }
    int *nSA = (int *)p1, *nLCP = (int *)p2;
    const i64 ntiles = (n + SP_TILE - 1) / SP_TILE;
    void *dtiles = nullptr;
    if (lease.take((size_t)ntiles * sizeof(SplitState) + 64, &dtiles) != RV_OK) { pool->give(p1); pool->give(p2); return RV_ERR_NOMEM; }
    SplitState *tiles = (SplitState *)dtiles;
    RV_LAUNCH(split_reduce_kernel, (unsigned)ntiles, SP_THREADS, 0, st.s, sub->LCP, D, n, tiles);
    RV_LAUNCH(split_tilescan_kernel, 1, TS_THREADS, 0, st.s, tiles, ntiles, d_counts);
    RV_LAUNCH(split_apply_kernel, (unsigned)ntiles, SP_THREADS, 0, st.s, sub->SA, sub->LCP, D, n, tiles, v.ISA, nSA, nLCP, (int *)nullptr, (int *)nullptr,
              (int *)nullptr, (int *)nullptr, (u32)cn, 0u, 0u);
    st.launches += 3;
    u32 *hc = st.pinned + 300;
    RV_CUDA(cudaMemcpyAsync(hc, d_counts, 12, cudaMemcpyDeviceToHost, st.s));
    if (total > 0) {
        RV_LAUNCH(lower_kernel, (unsigned)((total + 255) / 256), 256, 0, st.s, d_ibeg, d_pre, m, total, v.T);
        st.launches++;
    }
    void *dcand = nullptr;
    if (cn > 0 && !bbeg.empty()) {  // bubble_sort(idx, intervals) (reveal.c:1497)
        if (cn <= 8192) {
            RV_LAUNCH(bubble_kernel, 1, BB_THREADS, 0, st.s, nSA, nLCP, v.ISA, cn, d_bbeg, (int)bbeg.size(), bubble_cap());
            st.launches++;
        } else {
            if (lease.take((size_t)(BB_CAP + 16) * 4, &dcand) != RV_OK) { pool->give(p1); pool->give(p2); return RV_ERR_NOMEM; }
            int *cand = (int *)dcand, *cand_cnt = cand + BB_CAP;
            RV_CUDA(cudaMemsetAsync(cand_cnt, 0, 4, st.s));
            i64 blocks = (cn + 255) / 256;
            if (blocks > 148 * 8) blocks = 148 * 8;
            for (int b = 0; b < (int)bbeg.size(); b++) {
                RV_LAUNCH(bubble_detect_kernel, (unsigned)blocks, 256, 0, st.s, nSA, nLCP, cn, d_bbeg, b, cand, cand_cnt, bubble_cap());
                RV_LAUNCH(bubble_apply_kernel, 1, BB_THREADS, 0, st.s, nSA, nLCP, v.ISA, cn, d_bbeg, b, cand, cand_cnt, bubble_cap());
                st.launches += 2;
            }
        }
    }
    RV_CUDA(cudaStreamSynchronize(st.s));
    RV_KCHECK();
    if ((i64)hc[0] != cn) {
        pool->give(p1);
        pool->give(p2);
        set_error("rv_sub_extract: the intervals cover %lld positions but only %lld of them are suffixes of this index", (long long)total,
                  (long long)(n - (i64)hc[0]));
        return RV_ERR_ARG;
    }
    if (sub->owns) {
        pool->give(sub->SA);
        pool->give(sub->LCP);
    }
    sub->SA = nSA;
    sub->LCP = nLCP;
    sub->n = cn;
    sub->owns = true;
    sub->cached = false;
    return RV_OK;
}

// kernel launches behind those steps: [0] single-launch path (one launch serves a whole batch), [1] general path (calls)
int rv_rec_launches(rv_index *h, int64_t *launches2) {
    MainView v;
    RV_TRY(main_view(h, &v));
    RecCtx *ctx = ctx_of(v);
    if (!launches2) return RV_ERR_ARG;
    launches2[0] = ctx->launches[0];
    launches2[1] = ctx->launches[1];
    return RV_OK;
}

// rows / hdr (int64 triples) and members (int64 pairs) of the sub-index's last sweep
int rv_sub_fetch(rv_sub *s, int64_t *rows, int64_t cap_rows, int64_t *members, int64_t cap_members) {
    if (!s) return RV_ERR_ARG;
    if (s->cached) {
        i64 r = (i64)(s->rows.size() / 3), m = (i64)(s->members.size() / 2);
        if (r > cap_rows) r = cap_rows;
        if (m > cap_members) m = cap_members;
        if (r > 0 && rows) memcpy(rows, s->rows.data(), (size_t)r * 24);
        if (m > 0 && members) memcpy(members, s->members.data(), (size_t)m * 16);
        return RV_OK;
    }
    MainView v;
    RV_TRY(main_view(s->main, &v));
    if (v.nsamples > 2) return rv_mums_multi_fetch(s->main, rows, cap_rows, members, cap_members);
    return rv_mums_pair_fetch(s->main, rows, cap_rows);
}

}  // extern "C"
