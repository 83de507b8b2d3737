// rv_chain.cu -- the chaining recurrence of the REM driver's mumpicker on the device, many anchor lists per launch.
//
// Replaces, for the batched recursion, the O(m^2) dynamic program of schemes.chain (reveal/schemes.py:20-104) with its gap
// cost (utils.gapcost, reveal/utils.py:162-180): the anchors of one sub-index, ordered by their coordinate in the first path,
// are chained between the sub-index's bounds; anchor r may follow every earlier anchor that ends at or before it in EVERY
// path, and scores  best_i( score[i] + gain[r] - wpen * gapcost(i, r) ).  The recurrence is sequential in r but every r looks
// at all earlier rows: one thread block per list walks r, its threads share the earlier rows, an arg-max reduction picks the
// predecessor with the reference's tie-breaks (higher score, then the row that became reachable first, then the earlier row).
// Frontier batching (rv_sub_step_batch) makes many lists available at once; they run side by side, one block each.
// Bit-identical to the host recurrence rv_chain_dp (csrc/ext/chain_dp.h), which is its oracle (tests/test_chain_device.py).
#include "rv_internal.h"

namespace rv {

static const int CH_THREADS = 256;
static const int CH_MAXK = 64;

struct Cand {      // a predecessor candidate of the current row
    i64 total, score, joined;
    int row;       // -1: none
};
__device__ __forceinline__ bool cand_better(const Cand &a, const Cand &b) {  // a beats b
    if (b.row < 0) return a.row >= 0;
    if (a.row < 0) return false;
    if (a.total != b.total) return a.total > b.total;
    if (a.score != b.score) return a.score > b.score;
    if (a.joined != b.joined) return a.joined < b.joined;
    return a.row < b.row;
}
__device__ __forceinline__ Cand cand_shfl_down(const Cand &c, int d) {
    Cand r;
    r.total = __shfl_down_sync(FULL, c.total, d);
    r.score = __shfl_down_sync(FULL, c.score, d);
    r.joined = __shfl_down_sync(FULL, c.joined, d);
    r.row = __shfl_down_sync(FULL, c.row, d);
    return r;
}

// smem_words: 64-bit words of dynamic shared memory per block.  A list whose rows x (k + 3) words fit works entirely out of
// shared memory (coordinates, lengths, scores, "reachable since" -- every row re-reads all earlier rows, and the recurrence is a
// chain of dependent steps, so the latency of those reads IS the run time); longer lists read global memory.
__global__ void __launch_bounds__(CH_THREADS)
chain_dp_kernel(const i64 *__restrict__ start, const i64 *__restrict__ length, const i64 *__restrict__ gain, const i64 *__restrict__ row_off,
                const i64 *__restrict__ start_off, const int *__restrict__ kk, i64 wpen, int model, i64 *__restrict__ link, i64 *__restrict__ score,
                i64 *__restrict__ joined, int smem_words) {
    __shared__ Cand s_best[CH_THREADS / 32];
    RV_DYN_SMEM(i64, dyn);
    const int job = (int)blockIdx.x;
    const i64 r0 = row_off[job];
    const int rows = (int)(row_off[job + 1] - r0);
    const int k = kk[job];
    const i64 *st = start + start_off[job];
    const i64 *len = length + r0, *gn = gain + r0;
    i64 *lk = link + r0, *sc = score + r0, *jn = joined + r0;
    const int tid = (int)threadIdx.x;
    const bool in_smem = (i64)rows * (k + 3) <= (i64)smem_words;
    if (in_smem) {
        i64 *s_st = dyn, *s_len = dyn + (i64)rows * k, *s_sc = s_len + rows, *s_jn = s_sc + rows;
        for (int i = tid; i < rows * k; i += CH_THREADS) s_st[i] = st[i];
        for (int i = tid; i < rows; i += CH_THREADS) {
            s_len[i] = len[i];
            s_jn[i] = i == 0 ? 0 : -1;
            s_sc[i] = 0;
        }
        st = s_st;
        len = s_len;
        sc = s_sc;
        jn = s_jn;
        if (tid == 0 && rows > 0) lk[0] = 0;
    } else {
        for (int i = tid; i < rows; i += CH_THREADS) jn[i] = i == 0 ? 0 : -1;
        if (tid == 0 && rows > 0) {
            lk[0] = 0;
            sc[0] = 0;
        }
    }
    __syncthreads();
    i64 dist[CH_MAXK];
    for (int r = 1; r < rows; r++) {
        const i64 *sr = st + (i64)r * k;
        Cand best;
        best.row = -1;
        best.total = best.score = best.joined = 0;
        for (int i = tid; i < r; i += CH_THREADS) {
            const i64 *si = st + (i64)i * k;
            const i64 li = len[i];
            if (k == 2) {  // two paths (pairwise alignment, the usual case): registers only
                const i64 d0 = sr[0] - (si[0] + li), d1 = sr[1] - (si[1] + li);
                if (d0 < 0 || d1 < 0) continue;
                i64 j = jn[i];
                if (j < 0) {
                    j = r;
                    jn[i] = r;
                }
                const i64 hi = d0 > d1 ? d0 : d1, lo = d0 > d1 ? d1 : d0;
                const i64 pen = model == 0 ? hi - lo : (model == 1 ? (d0 + d1) / 2 : hi);
                Cand c;
                c.total = sc[i] + gn[r] - wpen * pen;
                c.score = sc[i];
                c.joined = j;
                c.row = i;
                if (cand_better(c, best)) best = c;
                continue;
            }
            bool ok = true;
            for (int c = 0; c < k; c++) {
                const i64 d = sr[c] - (si[c] + li);
                if (d < 0) {
                    ok = false;
                    break;
                }
                dist[c] = d;
            }
            if (!ok) continue;
            i64 j = jn[i];
            if (j < 0) {
                j = r;
                jn[i] = r;  // row i only ever belongs to this thread: i = tid (mod CH_THREADS)
            }
            i64 pen = 0;
            if (model == 0) {
                for (int a = 0; a < k; a++)
                    for (int b = a + 1; b < k; b++) pen += dist[a] > dist[b] ? dist[a] - dist[b] : dist[b] - dist[a];
            } else if (model == 1) {
                i64 sum = 0;
                for (int c = 0; c < k; c++) sum += dist[c];
                pen = sum / k;
            } else {
                for (int a = 1; a < k; a++) {  // insertion sort: k is the number of paths
                    const i64 v = dist[a];
                    int b = a;
                    while (b > 0 && dist[b - 1] > v) {
                        dist[b] = dist[b - 1];
                        b--;
                    }
                    dist[b] = v;
                }
                pen = dist[k / 2];
            }
            Cand c;
            c.total = sc[i] + gn[r] - wpen * pen;
            c.score = sc[i];
            c.joined = j;
            c.row = i;
            if (cand_better(c, best)) best = c;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const Cand o = cand_shfl_down(best, d);
            if (cand_better(o, best)) best = o;
        }
        if ((tid & 31) == 0) s_best[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            Cand b = s_best[0];
            for (int wv = 1; wv < CH_THREADS / 32; wv++)
                if (cand_better(s_best[wv], b)) b = s_best[wv];
            if (b.row < 0) {  // cannot happen with a proper left bound
                b.row = 0;
                b.total = 0;
            }
            lk[r] = b.row;
            sc[r] = b.total;
        }
        __syncthreads();
    }
    if (in_smem) {  // the scores go back to global memory in one sweep
        i64 *gsc = score + r0;
        for (int i = tid; i < rows; i += CH_THREADS) gsc[i] = sc[i];
    }
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

extern "C" int rv_chain_batch(rv_index *h, int32_t nlists, const int64_t *row_off, const int64_t *start_off, const int32_t *kk, const int64_t *start,
                              const int64_t *length, const int64_t *gain, int64_t wpen, int32_t model, int64_t *link, int64_t *score) {
    if (!h || nlists < 0 || (nlists > 0 && (!row_off || !start_off || !kk || !start || !length || !gain || !link || !score))) return RV_ERR_ARG;
    if (nlists == 0) return RV_OK;
    if (model < 0 || model > 2) { set_error("rv_chain_batch: unknown gap model %d", model); return RV_ERR_ARG; }
    for (int i = 0; i < nlists; i++)
        if (kk[i] < 1 || kk[i] > CH_MAXK || row_off[i + 1] < row_off[i] || start_off[i + 1] - start_off[i] != (row_off[i + 1] - row_off[i]) * kk[i]) {
            set_error("rv_chain_batch: list %d is inconsistent (k = %d)", i, kk[i]);
            return RV_ERR_ARG;
        }
    ChainView v;
    RV_TRY(chain_view(h, &v));
    Stream &st = *v.st;
    const size_t rows = (size_t)row_off[nlists], cells = (size_t)start_off[nlists];
    const size_t words = cells + 5 * rows + 2 * (size_t)(nlists + 1) + 64;
    const size_t bytes = words * 8 + (size_t)nlists * 4 + 16 * 256;  // every take() is padded to 256 bytes
    RV_TRY(v.ws->reserve(bytes));
    v.ws->reset();
    i64 *d_start = v.ws->take<i64>(cells), *d_len = v.ws->take<i64>(rows), *d_gain = v.ws->take<i64>(rows), *d_link = v.ws->take<i64>(rows),
        *d_score = v.ws->take<i64>(rows), *d_joined = v.ws->take<i64>(rows), *d_roff = v.ws->take<i64>((size_t)nlists + 1),
        *d_soff = v.ws->take<i64>((size_t)nlists + 1);
    int *d_kk = v.ws->take<int>((size_t)nlists);
    if (!d_start || !d_len || !d_gain || !d_link || !d_score || !d_joined || !d_roff || !d_soff || !d_kk) { set_error("rv_chain_batch: workspace"); return RV_ERR_NOMEM; }
    RV_CUDA(cudaMemcpyAsync(d_start, start, cells * 8, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemcpyAsync(d_len, length, rows * 8, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemcpyAsync(d_gain, gain, rows * 8, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemcpyAsync(d_roff, row_off, (size_t)(nlists + 1) * 8, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemcpyAsync(d_soff, start_off, (size_t)(nlists + 1) * 8, cudaMemcpyHostToDevice, st.s));
    RV_CUDA(cudaMemcpyAsync(d_kk, kk, (size_t)nlists * 4, cudaMemcpyHostToDevice, st.s));
    // dynamic shared memory: enough for the longest list of the batch, at most 96 KB per block
    i64 need_words = 0;
    for (int i = 0; i < nlists; i++) {
        const i64 wds = (row_off[i + 1] - row_off[i]) * (kk[i] + 3);
        if (wds > need_words && wds * 8 <= 96 * 1024) need_words = wds;
    }
    const size_t smem = (size_t)need_words * 8;
#ifndef RV_EMU
    if (smem > 48 * 1024) {
        static bool attr_done[64] = {false};
        int dev = 0;
        RV_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !attr_done[dev]) {
            RV_CUDA(cudaFuncSetAttribute(chain_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_done[dev] = true;
        }
    }
#endif
    RV_LAUNCH(chain_dp_kernel, (unsigned)nlists, CH_THREADS, smem, st.s, (const i64 *)d_start, (const i64 *)d_len, (const i64 *)d_gain, (const i64 *)d_roff,
              (const i64 *)d_soff, (const int *)d_kk, (i64)wpen, (int)model, d_link, d_score, d_joined, (int)need_words);
    st.launches++;
    RV_CUDA(cudaMemcpyAsync(link, d_link, rows * 8, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaMemcpyAsync(score, d_score, rows * 8, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));
    RV_KCHECK();
    return RV_OK;
}
