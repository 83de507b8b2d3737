// rv_tiny.cu -- getmums for MANY tiny two-sample indexes in one launch.
//
// `finish` / `transform` extend every anchor by indexing its two <= 200 bp flanks on their own: one index() / addsequence x 2 /
// construct() / getmums() per flank pair (reveal/transformold.py:1170-1240 `extend`, four such blocks per anchor).  A launch per
// 400-character index would be pure latency, so the flank pairs of a whole anchor list go through ONE launch, one thread block
// per pair, everything in shared memory:
//   suffix array   bitonic sort of the suffix starts by direct text comparison (plain byte order, a suffix that is a prefix of
//                  another sorts first -- the order divsufsort produces, interface.c:213-222)
//   LCP            direct comparison of neighbours with the reference's '$'/'N' barrier (compute_lcp, interface.c:97-114)
//   getmums        the per-slot test of reveal.c:55-116 (pair_test, flavour 0), rows (l, a, b) in SA-rank order
// The rows are exactly what index.getmums(minl) returns for that pair on its own.
#include "rv_internal.h"
#include "rv_sweep.h"
#include "rv_sweep_dev.cuh"

namespace rv {

static const int TY_THREADS = 256;
static const int TY_MAXN = 1024;   // characters per unit at most (a flank pair: 2 * (200 + 1))

// suffix a < suffix b of the unit's text t[0..n)
__device__ __forceinline__ bool tiny_less(const unsigned char *t, int n, int a, int b) {
    if (a == b) return false;
    const int lim = n - (a > b ? a : b);
    for (int h = 0; h < lim; h++) {
        const unsigned char x = t[a + h], y = t[b + h];
        if (x != y) return x < y;
    }
    return a > b;  // the shorter suffix (larger start) is a prefix of the other: it sorts first
}

__global__ void __launch_bounds__(TY_THREADS)
tiny_mums_kernel(const unsigned char *__restrict__ T, const i64 *__restrict__ off, const i64 *__restrict__ nsep0, int minl, int cap,
                 i64 *__restrict__ rows, int *__restrict__ counts) {
    __shared__ unsigned char s_t[TY_MAXN + 8];
    __shared__ int s_sa[TY_MAXN];
    __shared__ int s_lcp[TY_MAXN];
    __shared__ u32 s_scan[33];
    const int u = (int)blockIdx.x;
    const i64 o = off[u];
    const int n = (int)(off[u + 1] - o);
    const int tid = (int)threadIdx.x;
    if (n <= 0 || n > TY_MAXN) {
        if (tid == 0) counts[u] = n <= 0 ? 0 : -1;  // -1: too long for this path
        return;
    }
    for (int i = tid; i < n + 8; i += TY_THREADS) s_t[i] = i < n ? T[o + i] : 0;
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = tid; i < m; i += TY_THREADS) s_sa[i] = i < n ? i : 0x7fffffff;  // padding sorts last
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += TY_THREADS) {
                const int x = i ^ j;
                if (x > i) {
                    const int a = s_sa[i], b = s_sa[x];
                    const bool up = (i & k) == 0;
                    // a > b in suffix order (padding is the largest)
                    const bool gt = a == 0x7fffffff ? b != 0x7fffffff : (b == 0x7fffffff ? false : tiny_less(s_t, n, b, a));
                    if (gt == up) {
                        s_sa[i] = b;
                        s_sa[x] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (int r = tid; r < n; r += TY_THREADS) {
        int h = 0;
        if (r > 0) {
            const int a = s_sa[r], b = s_sa[r - 1];
            while (a + h < n && b + h < n) {
                const unsigned char c = s_t[a + h];
                if (c != s_t[b + h] || c == '$' || c == 'N') break;
                h++;
            }
        }
        s_lcp[r] = h;
    }
    __syncthreads();
    SweepArgs p;
    p.T = s_t;
    p.SA = s_sa;
    p.LCP = s_lcp;
    p.SO = nullptr;
    p.n = n;
    p.nT = n;
    p.nsep0 = nsep0[u];
    p.rc = 0;
    p.flavour = 0;
    p.minl = minl;
    p.minn = 2;
    p.main_nsamples = 2;
    // ordered emission: a chunk of consecutive slots per thread, block-wide exclusive scan of the hit counts
    const int chunk = (n + TY_THREADS - 1) / TY_THREADS;
    const int lo = tid * chunk, hi = lo + chunk < n ? lo + chunk : n;
    u32 mine = 0;
    for (int i = lo; i < hi; i++) {
        i64 l, a, b;
        mine += pair_test(p, (i64)i, l, a, b) ? 1u : 0u;
    }
    u32 total;
    const u32 inc = block_incl_sum<TY_THREADS, u32>(mine, s_scan, &total);
    u32 at = inc - mine;
    i64 *out = rows + (i64)u * cap * 3;
    for (int i = lo; i < hi && mine; i++) {
        i64 l = 0, a = 0, b = 0;
        if (pair_test(p, (i64)i, l, a, b)) {
            if ((int)at < cap) {
                out[3 * at + 0] = l;
                out[3 * at + 1] = a;
                out[3 * at + 2] = b;
            }
            at++;
        }
    }
    if (tid == 0) counts[u] = (int)total;
}

}  // namespace rv

using namespace rv;

extern "C++" {
namespace rv {
struct ChainView {
    Stream *st;
    Arena *ws;
};
int chain_view(rv_index *h, ChainView *out);
}  // namespace rv
}

extern "C" int rv_mums_tiny_batch(rv_index *h, const uint8_t *T, const int64_t *off, const int64_t *nsep0, int32_t nunits, int32_t minl, int32_t cap,
                                  int64_t *rows, int32_t *counts) {
    if (!h || nunits < 0 || cap < 0 || (nunits > 0 && (!T || !off || !nsep0 || !counts || (cap > 0 && !rows)))) return RV_ERR_ARG;
    if (nunits == 0) return RV_OK;
    for (int u = 0; u < nunits; u++)
        if (off[u + 1] < off[u] || off[u + 1] - off[u] > TY_MAXN) {
            set_error("rv_mums_tiny_batch: unit %d has %lld characters (limit %d): build it as an index of its own", u, (long long)(off[u + 1] - off[u]), TY_MAXN);
            return RV_ERR_UNSUPPORTED;
        }
    ChainView v;  // any handle: its stream and staging area
    RV_TRY(chain_view(h, &v));
    Stream &st = *v.st;
    const size_t tbytes = (size_t)(off[nunits] - off[0]), nrows = (size_t)nunits * (size_t)cap;
    RV_TRY(v.ws->reserve(tbytes + 256 + (size_t)(nunits + 1) * 16 + nrows * 24 + (size_t)nunits * 4 + 16 * 256));  // every take() is padded to 256 bytes
    v.ws->reset();
    unsigned char *dT = v.ws->take<unsigned char>(tbytes + 8);
    i64 *d_off = v.ws->take<i64>((size_t)nunits + 1), *d_sep = v.ws->take<i64>((size_t)nunits), *d_rows = v.ws->take<i64>(nrows * 3 + 3);
    int *d_counts = v.ws->take<int>((size_t)nunits);
    if (!dT || !d_off || !d_sep || !d_rows || !d_counts) { set_error("rv_mums_tiny_batch: workspace"); return RV_ERR_NOMEM; }
    std::vector<i64> rel((size_t)nunits + 1);
    for (int u = 0; u <= nunits; u++) rel[(size_t)u] = off[u] - off[0];
    RV_CUDA(cudaMemcpyAsync(dT, T + off[0], tbytes, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemcpyAsync(d_off, rel.data(), (size_t)(nunits + 1) * 8, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemcpyAsync(d_sep, nsep0, (size_t)nunits * 8, cudaMemcpyHostToDevice, st.s));
    RV_LAUNCH(tiny_mums_kernel, (unsigned)nunits, TY_THREADS, 0, st.s, (const unsigned char *)dT, (const i64 *)d_off, (const i64 *)d_sep, (int)minl, (int)cap,
              d_rows, d_counts);
    st.launches++;
    if (nrows) RV_CUDA(cudaMemcpyAsync(rows, d_rows, nrows * 24, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaMemcpyAsync(counts, d_counts, (size_t)nunits * 4, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));  // (rel[] is pageable: its copy was staged before the call returned)
    RV_KCHECK();
    return RV_OK;
}
