// rv_sweep.cu -- SA/LCP sweeps that emit (multi-genome) Maximal Unique Matches.
//
// sweep_pair  replaces getmums (reveallib/reveal.c:55-116) and getmums_rem
//             (reveal.c:119-180): one independent test per SA slot.
// sweep_multi replaces getmultimums (reveal.c:436-580) + ismultimum
//             (reveal.c:227-259).  The reference enumerates lcp-intervals
//             bottom-up with a sequential stack; only intervals of at most
//             main->nsamples members can be reported (reveal.c:476), so the GPU
//             form gives every closing slot `ub` its own thread that walks the
//             left boundary back at most nsamples-1 steps with a running
//             minimum: [lb,ub] is an lcp-interval of value l iff
//             l = min LCP[lb+1..ub] > max(LCP[lb], LCP[ub+1]).
//             Pop order of the stack walk == (ub ascending, lb descending),
//             which is exactly thread order then walk order, so the emitted
//             list is identical including its order.
// Both run count -> exclusive scan -> write so the output order is deterministic.
#include "rv_internal.h"
#include "rv_sweep.h"
#include "rv_sweep_dev.cuh"

namespace rv {

static const int SW_THREADS = 256;
static const int SW_CHUNKS = 4;
static const int SW_TILE = SW_THREADS * SW_CHUNKS;

__global__ void __launch_bounds__(SW_THREADS) pair_count_kernel(SweepArgs p, u64 *__restrict__ tile_rec) {
    __shared__ u64 scratch[33];
    u64 mine = 0;
    for (int c = 0; c < SW_CHUNKS; c++) {
        i64 i = (i64)blockIdx.x * SW_TILE + c * SW_THREADS + threadIdx.x;
        i64 l, a, b;
        mine += pair_test(p, i, l, a, b) ? 1u : 0u;
    }
    u64 total;
    block_incl_sum<SW_THREADS, u64>(mine, scratch, &total);
    if (threadIdx.x == 0) tile_rec[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SW_THREADS) pair_write_kernel(SweepArgs p, const u64 *__restrict__ tile_rec, i64 *__restrict__ out, i64 cap) {
    __shared__ u64 scratch[33];
    u64 carry = tile_rec[blockIdx.x];
    for (int c = 0; c < SW_CHUNKS; c++) {
        i64 i = (i64)blockIdx.x * SW_TILE + c * SW_THREADS + threadIdx.x;
        i64 l = 0, a = 0, b = 0;
        bool hit = pair_test(p, i, l, a, b);
        u64 total;
        u64 inc = block_incl_sum<SW_THREADS, u64>(hit ? 1u : 0u, scratch, &total);
        if (hit) {
            u64 at = carry + inc - 1;
            if ((i64)at < cap) {
                out[3 * at + 0] = l;
                out[3 * at + 1] = a;
                out[3 * at + 2] = b;
            }
        }
        carry += total;
        __syncthreads();
    }
}

// ---- multi sweep --------------------------------------------------------------
__global__ void __launch_bounds__(SW_THREADS) multi_count_kernel(SweepArgs p, u64 *__restrict__ tile_rec, u64 *__restrict__ tile_mem) {
    __shared__ u64 s1[33], s2[33];
    u64 nr = 0, nm = 0;
    for (int c = 0; c < SW_CHUNKS; c++) {
        i64 ub = (i64)blockIdx.x * SW_TILE + c * SW_THREADS + threadIdx.x;
        multi_visit(p, ub, [&](i64, i64, i64 size) { nr++; nm += (u64)size; });
    }
    u64 tr, tm;
    block_incl_sum<SW_THREADS, u64>(nr, s1, &tr);
    block_incl_sum<SW_THREADS, u64>(nm, s2, &tm);
    if (threadIdx.x == 0) {
        tile_rec[blockIdx.x] = tr;
        tile_mem[blockIdx.x] = tm;
    }
}

__global__ void __launch_bounds__(SW_THREADS)
multi_write_kernel(SweepArgs p, const u64 *__restrict__ tile_rec, const u64 *__restrict__ tile_mem, i64 *__restrict__ hdr, i64 hdr_cap,
                   i64 *__restrict__ members, i64 mem_cap) {
    __shared__ u64 s1[33], s2[33];
    u64 carry_r = tile_rec[blockIdx.x], carry_m = tile_mem[blockIdx.x];
    for (int c = 0; c < SW_CHUNKS; c++) {
        i64 ub = (i64)blockIdx.x * SW_TILE + c * SW_THREADS + threadIdx.x;
        u64 nr = 0, nm = 0;
        multi_visit(p, ub, [&](i64, i64, i64 size) { nr++; nm += (u64)size; });
        u64 tr, tm;
        u64 ir = block_incl_sum<SW_THREADS, u64>(nr, s1, &tr);
        u64 im = block_incl_sum<SW_THREADS, u64>(nm, s2, &tm);
        u64 at_r = carry_r + ir - nr, at_m = carry_m + im - nm;
        if (nr) {
            multi_visit(p, ub, [&](i64 l, i64 lb, i64 size) {
                if ((i64)at_r < hdr_cap) {
                    hdr[3 * at_r + 0] = l;
                    hdr[3 * at_r + 1] = size;
                    hdr[3 * at_r + 2] = (i64)at_m;
                }
                for (i64 x = 0; x < size; x++) {
                    if ((i64)at_m < mem_cap) {
                        i64 pos = p.SA[lb + x];
                        members[2 * at_m + 0] = sample_of(p, pos);
                        members[2 * at_m + 1] = pos;
                    }
                    at_m++;
                }
                at_r++;
            });
        }
        carry_r += tr;
        carry_m += tm;
        __syncthreads();
    }
}

// single block: in-place exclusive scan of up to two u64 arrays; totals -> out[0], out[1]
__global__ void __launch_bounds__(1024) sweep_tilescan_kernel(u64 *__restrict__ a, u64 *__restrict__ b, i64 tiles, u64 *__restrict__ out) {
    __shared__ u64 s1[33], s2[33];
    u64 ca = 0, cb = 0;
    for (i64 b0 = 0; b0 < tiles; b0 += 1024) {
        i64 t = b0 + threadIdx.x;
        u64 va = t < tiles ? a[t] : 0ull;
        u64 vb = (b && t < tiles) ? b[t] : 0ull;
        u64 ta, tb;
        u64 ia = block_incl_sum<1024, u64>(va, s1, &ta);
        u64 ib = block_incl_sum<1024, u64>(vb, s2, &tb);
        if (t < tiles) {
            a[t] = ca + ia - va;
            if (b) b[t] = cb + ib - vb;
        }
        ca += ta;
        cb += tb;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = ca;
        out[1] = cb;
    }
}

size_t sweep_scratch_bytes(i64 n) {
    i64 tiles = (n + SW_TILE - 1) / SW_TILE;
    return (size_t)(2 * tiles + 8) * 8 + 512;
}

int sweep_pair_count(Stream &st, const SweepArgs &p, void *scratch, i64 *count) {
    *count = 0;
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    u64 *tile_rec = (u64 *)scratch;
    u64 *totals = tile_rec + 2 * tiles;
    RV_TRY(prof_begin(st));
    RV_LAUNCH(pair_count_kernel, (unsigned)tiles, SW_THREADS, 0, st.s, p, tile_rec);
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)p.n * 9));
    RV_LAUNCH(sweep_tilescan_kernel, 1, 1024, 0, st.s, tile_rec, (u64 *)nullptr, tiles, totals);
    st.launches += 2;
    u64 h[2];
    RV_CUDA(cudaMemcpyAsync(h, totals, 16, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));
    *count = (i64)h[0];
    return RV_OK;
}

int sweep_pair_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_out, i64 cap) {
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    RV_TRY(prof_begin(st));
    RV_LAUNCH(pair_write_kernel, (unsigned)tiles, SW_THREADS, 0, st.s, p, (const u64 *)scratch, d_out, cap);
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)p.n * 9));
    st.launches++;
    RV_KCHECK();
    return RV_OK;
}

int sweep_multi_count(Stream &st, const SweepArgs &p, void *scratch, i64 *nrec, i64 *nmem) {
    *nrec = *nmem = 0;
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    u64 *tile_rec = (u64 *)scratch, *tile_mem = tile_rec + tiles;
    u64 *totals = tile_rec + 2 * tiles;
    RV_TRY(prof_begin(st));
    RV_LAUNCH(multi_count_kernel, (unsigned)tiles, SW_THREADS, 0, st.s, p, tile_rec, tile_mem);
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)p.n * 11));
    RV_LAUNCH(sweep_tilescan_kernel, 1, 1024, 0, st.s, tile_rec, tile_mem, tiles, totals);
    st.launches += 2;
    u64 h[2];
    RV_CUDA(cudaMemcpyAsync(h, totals, 16, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));
    *nrec = (i64)h[0];
    *nmem = (i64)h[1];
    return RV_OK;
}

int sweep_multi_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_hdr, i64 hdr_cap, i64 *d_mem, i64 mem_cap) {
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    const u64 *tile_rec = (const u64 *)scratch, *tile_mem = tile_rec + tiles;
    RV_TRY(prof_begin(st));
    RV_LAUNCH(multi_write_kernel, (unsigned)tiles, SW_THREADS, 0, st.s, p, tile_rec, tile_mem, d_hdr, hdr_cap, d_mem, mem_cap);
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)p.n * 11));
    st.launches++;
    RV_KCHECK();
    return RV_OK;
}

}  // namespace rv
