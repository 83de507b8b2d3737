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
                                                   int m, i64 total, const int *__restrict__ SAi, unsigned char *__restrict__ D) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int lo = 0, hi = m - 1;  // last k with pre[k] <= g
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pre[mid] <= g) lo = mid; else hi = mid - 1;
    }
    i64 j = ibeg[lo] + (g - pre[lo]);
    D[SAi[j]] = lab[lo];
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

// single thread block: exclusive scan of the tile states (sequential over tiles; tiles are few)
__global__ void __launch_bounds__(32) split_tilescan_kernel(SplitState *__restrict__ tiles, i64 ntiles, u32 *__restrict__ counts) {
    if (threadIdx.x != 0) return;
    SplitState run = split_identity();
    for (i64 t = 0; t < ntiles; t++) {
        SplitState x = tiles[t];
        tiles[t] = run;
        run = split_combine(run, x);
    }
    counts[0] = run.cnt[0];
    counts[1] = run.cnt[1];
    counts[2] = run.cnt[2];
}

__global__ void __launch_bounds__(SP_THREADS)
split_apply_kernel(const int *__restrict__ SA, const int *__restrict__ LCP, const unsigned char *__restrict__ D, i64 n,
                   const SplitState *__restrict__ tiles, int *__restrict__ SAi, int *__restrict__ sa0, int *__restrict__ lcp0,
                   int *__restrict__ sa1, int *__restrict__ lcp1, int *__restrict__ sa2, int *__restrict__ lcp2) {
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
        if (cls >= 0) {
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

// the reference's loop body for slot i (reveal.c:686-722), on the child's arrays
__device__ __forceinline__ void bubble_body(int *SA, int *LCP, int *SAi, i64 n, i64 i, i64 begin) {
    if ((SA[i] < begin) && ((i64)SA[i] + LCP[i] > begin)) {  // the match overlaps the start position
        i64 x = i;
        int tmpSA = SA[i], tmpLCP = LCP[i];
        while (((i64)LCP[x] >= begin - tmpSA) && x > 0) {
            SAi[SA[x - 1]] = (int)x;
            SA[x] = SA[x - 1];
            LCP[x] = LCP[x - 1];
            x--;
        }
        SAi[tmpSA] = (int)x;
        SA[x] = tmpSA;
        LCP[x + 1] = (int)(begin - tmpSA);
        if (i < n - 1) {
            if (tmpLCP < LCP[i + 1]) LCP[i + 1] = tmpLCP;
        }
    } else if (i < n - 1) {
        if ((SA[i] < begin) && ((i64)SA[i] + LCP[i + 1] > begin)) {
            if (LCP[i + 1] > LCP[i]) LCP[i + 1] = (int)(begin - SA[i]);
        }
    }
}

__global__ void __launch_bounds__(BB_THREADS) bubble_kernel(int *SA, int *LCP, int *SAi, i64 n, const i64 *__restrict__ begins, int nbegins) {
    __shared__ int s_cand[BB_CAP];
    __shared__ int s_cnt;
    for (int b = 0; b < nbegins; b++) {
        const i64 begin = begins[b];
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        // candidates: LCP values only decrease during the pass, so a slot whose original values do not
        // reach across `begin` can never take either branch
        for (i64 i = threadIdx.x; i < n; i += BB_THREADS) {
            i64 s = SA[i];
            if (s < begin) {
                bool c = s + LCP[i] > begin || (i < n - 1 && s + LCP[i + 1] > begin);
                if (c) {
                    int at = atomicAdd(&s_cnt, 1);
                    if (at < BB_CAP) s_cand[at] = (int)i;
                }
            }
        }
        __syncthreads();
        const int cnt = s_cnt;
        if (cnt > BB_CAP) {  // rare: too many candidates for shared memory, replay the whole loop
            if (threadIdx.x == 0)
                for (i64 i = 0; i < n; i++) bubble_body(SA, LCP, SAi, n, i, begin);
        } else if (cnt > 0) {
            // bitonic sort of the candidate slots (ascending), padded with INT_MAX
            int m = 1;
            while (m < cnt) m <<= 1;
            for (int i = cnt + threadIdx.x; i < m; i += BB_THREADS) s_cand[i] = 0x7fffffff;
            __syncthreads();
            for (int k = 2; k <= m; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < m; i += BB_THREADS) {
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
            if (threadIdx.x == 0)
                for (int c = 0; c < cnt; c++) bubble_body(SA, LCP, SAi, n, (i64)s_cand[c], begin);
        }
        __syncthreads();
    }
}

}  // namespace rv

using namespace rv;

// ---- device buffer pool: children come and go thousands of times per alignment ------------------
struct DevPool {
    std::multimap<size_t, void *> free_;
    std::map<void *, size_t> size_;
    size_t held = 0;
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
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, c);
        if (e != cudaSuccess) {  // give cached blocks back and retry once
            trim();
            e = cudaMalloc(&p, c);
        }
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", c, cudaGetErrorString(e));
            return RV_ERR_NOMEM;
        }
        size_[p] = c;
        held += c;
        *out = p;
        return RV_OK;
    }
    void give(void *p) {
        if (!p) return;
        free_.insert({size_[p], p});
    }
    void trim() {
        for (auto &kv : free_) {
            cudaFree(kv.second);
            held -= size_[kv.second];
            size_.erase(kv.second);
        }
        free_.clear();
    }
};

struct rv_sub {
    rv_index *main;
    int *SA = nullptr, *LCP = nullptr;
    i64 n = 0;
    bool owns = false;  // false: the root view over the main index's arrays
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
    void **pool_slot;  // where the handle keeps its DevPool*
};
int main_view(rv_index *h, MainView *out);
int sub_sweep_pair(rv_index *h, const SweepArgs &a, int64_t *count);
int sub_sweep_multi(rv_index *h, const SweepArgs &a, int64_t *nrec, int64_t *nmem);
}  // namespace rv

static DevPool *pool_of(MainView &v) {
    if (!*v.pool_slot) *v.pool_slot = new DevPool();
    return (DevPool *)*v.pool_slot;
}

extern "C" {

void rv_pool_destroy(void *pool) {  // called by rv_index_free
    if (!pool) return;
    DevPool *p = (DevPool *)pool;
    p->trim();
    for (auto &kv : p->size_) cudaFree(kv.first);
    delete p;
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
    return a;
}

int rv_sub_mums_pair(rv_sub *s, int32_t minl, int64_t *count) {
    if (!s) return RV_ERR_ARG;
    MainView v;
    RV_TRY(main_view(s->main, &v));
    SweepArgs a = sub_args(s, v);
    a.minl = minl;
    return sub_sweep_pair(s->main, a, count);
}

int rv_sub_mums_multi(rv_sub *s, int32_t minl, int32_t minn, int64_t *nrec, int64_t *nmem) {
    if (!s) return RV_ERR_ARG;
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
int rv_sub_split(rv_sub *parent, const int64_t *lead, int32_t nlead, const int64_t *trail, int32_t ntrail, const int64_t *par, int32_t npar,
                 const int64_t *mum_sp, int32_t mum_n, int64_t mum_l, const int64_t *matching, int32_t nmatch, rv_sub **children) {
    if (!parent || !children) return RV_ERR_ARG;
    children[0] = children[1] = children[2] = nullptr;
    MainView v;
    RV_TRY(main_view(parent->main, &v));
    DevPool *pool = pool_of(v);
    Stream &st = *v.st;
    const i64 n = parent->n;
    if (n <= 0) return RV_OK;

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
    RV_TRY(pool->take(bytes, &dtab));
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
    RV_TRY(pool->take((size_t)n, &dD));
    unsigned char *D = (unsigned char *)dD;
    RV_CUDA(cudaMemsetAsync(D, 0, (size_t)n, st.s));
    if (total1 > 0) {
        RV_LAUNCH(label_kernel, (unsigned)((total1 + 255) / 256), 256, 0, st.s, d_ibeg, d_pre, d_lab, m1, total1, v.ISA, D);
        st.launches++;
    }
    if (mtotal > 0) {  // matched bases override (reveal.c:1112-1116)
        void *d3 = nullptr;
        RV_TRY(pool->take((size_t)m2 + 64, &d3));
        RV_CUDA(cudaMemsetAsync(d3, 3, (size_t)m2, st.s));
        RV_LAUNCH(label_kernel, (unsigned)((mtotal + 255) / 256), 256, 0, st.s, d_mbeg, d_mpre, (const unsigned char *)d3, m2, mtotal, v.ISA, D);
        st.launches++;
        pool->give(d3);  // stream-ordered reuse: later users are enqueued behind this kernel
    }

    // ---- children ----
    rv_sub *kids[3] = {nullptr, nullptr, nullptr};
    for (int c = 0; c < 3; c++)
        if (cls_n[c] > 0) {
            rv_sub *k = new rv_sub();
            k->main = parent->main;
            k->n = cls_n[c];
            k->owns = true;
            void *p1 = nullptr, *p2 = nullptr;
            int r1 = pool->take((size_t)(cls_n[c] + 2) * 4, &p1);
            int r2 = r1 == RV_OK ? pool->take((size_t)(cls_n[c] + 2) * 4, &p2) : r1;
            if (r1 != RV_OK || r2 != RV_OK) { delete k; return RV_ERR_NOMEM; }
            k->SA = (int *)p1;
            k->LCP = (int *)p2;
            kids[c] = k;
        }
    const i64 ntiles = (n + SP_TILE - 1) / SP_TILE;
    void *dtiles = nullptr;
    RV_TRY(pool->take((size_t)ntiles * sizeof(SplitState) + 64, &dtiles));
    SplitState *tiles = (SplitState *)dtiles;
    u32 *d_counts = (u32 *)((unsigned char *)dtab + bytes - 32);
    RV_LAUNCH(split_reduce_kernel, (unsigned)ntiles, SP_THREADS, 0, st.s, parent->LCP, D, n, tiles);
    RV_LAUNCH(split_tilescan_kernel, 1, 32, 0, st.s, tiles, ntiles, d_counts);
    RV_LAUNCH(split_apply_kernel, (unsigned)ntiles, SP_THREADS, 0, st.s, parent->SA, parent->LCP, D, n, tiles, v.ISA,
              kids[0] ? kids[0]->SA : nullptr, kids[0] ? kids[0]->LCP : nullptr, kids[1] ? kids[1]->SA : nullptr,
              kids[1] ? kids[1]->LCP : nullptr, kids[2] ? kids[2]->SA : nullptr, kids[2] ? kids[2]->LCP : nullptr);
    st.launches += 3;
    // the labelled positions must be exactly the parent's suffixes of each class: check the device counts
    u32 hc[3] = {0, 0, 0};
    RV_CUDA(cudaMemcpyAsync(hc, d_counts, 12, cudaMemcpyDeviceToHost, st.s));

    // ---- mark the matched bases in T (reveal.c:1230-1234) ----
    if (mtotal > 0) {
        RV_LAUNCH(lower_kernel, (unsigned)((mtotal + 255) / 256), 256, 0, st.s, d_mbeg, d_mpre, m2, mtotal, v.T);
        st.launches++;
    }
    // ---- bubble_sort(i_leading, matching_intervals)  (reveal.c:1250-1252) ----
    if (kids[0] && !bbeg.empty()) {
        RV_LAUNCH(bubble_kernel, 1, BB_THREADS, 0, st.s, kids[0]->SA, kids[0]->LCP, v.ISA, kids[0]->n, d_bbeg, (int)bbeg.size());
        st.launches++;
    }
    RV_CUDA(cudaStreamSynchronize(st.s));
    RV_KCHECK();
    pool->give(dtab);
    pool->give(dD);
    pool->give(dtiles);
    for (int c = 0; c < 3; c++)
        if ((i64)hc[c] != cls_n[c]) {
            set_error("rv_sub_split: class %d has %u suffixes in the parent but the intervals cover %lld positions", c, hc[c], (long long)cls_n[c]);
            for (int q = 0; q < 3; q++) rv_sub_free(kids[q]);
            return RV_ERR_ARG;
        }
    for (int c = 0; c < 3; c++) children[c] = kids[c];
    return RV_OK;
}

}  // extern "C"
