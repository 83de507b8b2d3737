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

// count pass: per-tile number of hits + one hit bit per slot (a word per 32 slots), so that the write pass only touches the
// (sparse) hits.  A thread takes 4 CONSECUTIVE slots: their LCP and SA entries arrive as two 128-bit loads issued before anything
// is tested (32 bytes in flight per thread keep HBM busy; one 4-byte load per thread and test did not), then the text gathers of
// the slots that pass the LCP / sample tests go out together.
__global__ void __launch_bounds__(SW_THREADS) pair_count_kernel(SweepArgs p, u64 *__restrict__ tile_rec, u32 *__restrict__ hitbits) {
    grid_dep_wait();
    __shared__ u64 scratch[33];
    const i64 i0 = (i64)blockIdx.x * SW_TILE + (i64)threadIdx.x * SW_CHUNKS;  // SW_CHUNKS == 4
    u32 h = 0;  // bit j: slot i0 + j is a hit
    const bool vec = ((((size_t)p.SA) | ((size_t)p.LCP)) & 15u) == 0;
    if (vec && i0 + 4 <= p.n) {
        const int4 lv = *(const int4 *)(p.LCP + i0);
        const int4 sv = *(const int4 *)(p.SA + i0);
        const int lprev = i0 > 0 ? p.LCP[i0 - 1] : 0;
        const int sprev = i0 > 0 ? p.SA[i0 - 1] : 0;
        const int lnext = i0 + 4 < p.n ? p.LCP[i0 + 4] : (int)0x80000000;  // no slot after the last one (pair_candidate)
        const int l[6] = {lprev, lv.x, lv.y, lv.z, lv.w, lnext};
        const int s[5] = {sprev, sv.x, sv.y, sv.z, sv.w};
        bool cand[4];
        i64 a[4], b[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int li = l[j + 1];
            cand[j] = i0 + j >= 1 && li >= p.minl && li > 0 && l[j] < li && l[j + 2] < li && (((i64)s[j + 1] > p.nsep0) != ((i64)s[j] > p.nsep0));
            a[j] = s[j + 1] < s[j] ? s[j + 1] : s[j];
            b[j] = s[j + 1] < s[j] ? s[j] : s[j + 1];
        }
        unsigned char ca[4], cb[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool g = cand[j] && a[j] > 0 && b[j] > 0;
            ca[j] = g ? p.T[a[j] - 1] : (unsigned char)0;
            cb[j] = g ? p.T[b[j] - 1] : (unsigned char)1;  // (no gather: left-maximal, as in left_maximal())
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (cand[j] && (ca[j] != cb[j] || ca[j] == 'N' || ca[j] == '$' || is_lower(ca[j]))) h |= 1u << j;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            i64 l, a, b;
            if (pair_test(p, i0 + j, l, a, b)) h |= 1u << j;
        }
    }
    // hit word of 32 slots = the nibbles of 8 neighbouring lanes
    u32 word = h << (4u * (threadIdx.x & 7u));
    word |= __shfl_xor_sync(FULL, word, 1);
    word |= __shfl_xor_sync(FULL, word, 2);
    word |= __shfl_xor_sync(FULL, word, 4);
    if ((threadIdx.x & 7u) == 0) hitbits[i0 >> 5] = word;  // (the bitmap is padded to whole tiles)
    u64 total;
    block_incl_sum<SW_THREADS, u64>((u64)__popc(h), scratch, &total);
    if (threadIdx.x == 0) tile_rec[blockIdx.x] = total;
}

// write pass: one warp per tile of SW_TILE slots (= 32 ballot words, one per lane)
__global__ void __launch_bounds__(SW_THREADS) pair_write_kernel(SweepArgs p, const u64 *__restrict__ tile_rec, const u32 *__restrict__ hitbits,
                                                                i64 tiles, i64 *__restrict__ out, i64 cap) {
    grid_dep_wait();
    const i64 tile = (i64)blockIdx.x * (SW_THREADS / 32) + (threadIdx.x >> 5);
    if (tile >= tiles) return;  // warp-uniform
    const unsigned lane = threadIdx.x & 31u;
    const i64 word = tile * (SW_TILE / 32) + lane;
    u32 bits = word * 32 < p.n ? hitbits[word] : 0u;
    u32 c = (u32)__popc(bits);
    u32 inc = warp_incl_sum(c);
    u64 at = tile_rec[tile] + (u64)(inc - c);
    while (bits) {
        int bpos = __ffs((int)bits) - 1;
        bits &= bits - 1u;
        i64 i = word * 32 + bpos;
        i64 l = 0, a = 0, b = 0;
        pair_test(p, i, l, a, b);
        if ((i64)at < cap) {
            out[3 * at + 0] = l;
            out[3 * at + 1] = a;
            out[3 * at + 2] = b;
        }
        at++;
    }
}

// ---- multi sweep --------------------------------------------------------------
// Two phases per tile of SW_TILE slots.  (1) every thread looks at 4 consecutive slots by one 128-bit LCP load and keeps the ones
// where an interval of at least minl closes (LCP[ub] > LCP[ub+1]) -- a third of the slots of similar genomes, fewer elsewhere;
// their tile offsets are compacted into a shared list.  (2) the threads take the listed slots one after the other, so the walks
// with their gathers run with every lane busy instead of the 10 of 32 a slot-per-thread mapping leaves active.
__global__ void __launch_bounds__(SW_THREADS) multi_count_kernel(SweepArgs p, u64 *__restrict__ tile_rec, u64 *__restrict__ tile_mem,
                                                                 u32 *__restrict__ hitbits) {
    grid_dep_wait();
    __shared__ u64 s1[33], s2[33];
    __shared__ u32 s3[33];
    __shared__ unsigned short s_list[SW_TILE];
    __shared__ u32 s_hit[SW_TILE / 32];
    const i64 tile0 = (i64)blockIdx.x * SW_TILE;
    const i64 i0 = tile0 + (i64)threadIdx.x * SW_CHUNKS;  // SW_CHUNKS == 4
    if (threadIdx.x < SW_TILE / 32) s_hit[threadIdx.x] = 0;
    u32 closing = 0;  // bit j: an interval may close at slot i0 + j
    if ((((size_t)p.LCP) & 15u) == 0 && i0 + 4 < p.n) {
        const int4 lv = *(const int4 *)(p.LCP + i0);
        const int l[5] = {lv.x, lv.y, lv.z, lv.w, p.LCP[i0 + 4]};
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (i0 + j >= 1 && l[j] > l[j + 1] && l[j] > 0 && l[j] >= p.minl) closing |= 1u << j;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const i64 ub = i0 + j;
            if (ub < 1 || ub >= p.n) continue;
            const i64 next = ub + 1 < p.n ? (i64)p.LCP[ub + 1] : -1;
            const i64 m = p.LCP[ub];
            if (m > next && m > 0 && m >= p.minl) closing |= 1u << j;
        }
    }
    u32 total;
    const u32 cnt = (u32)__popc(closing);
    u32 at = block_incl_sum<SW_THREADS, u32>(cnt, s3, &total) - cnt;
    for (u32 bits = closing; bits; bits &= bits - 1u) s_list[at++] = (unsigned short)(threadIdx.x * SW_CHUNKS + (__ffs((int)bits) - 1));
    __syncthreads();
    u64 nr = 0, nm = 0;
    for (u32 k = threadIdx.x; k < total; k += SW_THREADS) {
        const u32 off = s_list[k];
        const u64 r0 = nr;
        multi_visit(p, tile0 + off, [&](i64, i64, i64 size) { nr++; nm += (u64)size; });
        if (nr != r0) atomicOr(&s_hit[off >> 5], 1u << (off & 31u));
    }
    __syncthreads();
    if (threadIdx.x < SW_TILE / 32) hitbits[(tile0 >> 5) + threadIdx.x] = s_hit[threadIdx.x];  // (the bitmap is padded to whole tiles)
    u64 tr, tm;
    block_incl_sum<SW_THREADS, u64>(nr, s1, &tr);
    block_incl_sum<SW_THREADS, u64>(nm, s2, &tm);
    if (threadIdx.x == 0) {
        tile_rec[blockIdx.x] = tr;
        tile_mem[blockIdx.x] = tm;
    }
}

// write pass: one warp per tile; a lane re-walks only the closing slots flagged in its ballot word
__global__ void __launch_bounds__(SW_THREADS)
multi_write_kernel(SweepArgs p, const u64 *__restrict__ tile_rec, const u64 *__restrict__ tile_mem, const u32 *__restrict__ hitbits, i64 tiles,
                   i64 *__restrict__ hdr, i64 hdr_cap, i64 *__restrict__ members, i64 mem_cap) {
    grid_dep_wait();
    const i64 tile = (i64)blockIdx.x * (SW_THREADS / 32) + (threadIdx.x >> 5);
    if (tile >= tiles) return;  // warp-uniform
    const unsigned lane = threadIdx.x & 31u;
    const i64 word = tile * (SW_TILE / 32) + lane;
    const u32 bits0 = word * 32 < p.n ? hitbits[word] : 0u;
    u64 nr = 0, nm = 0;
    for (u32 bits = bits0; bits; bits &= bits - 1u) {
        i64 ub = word * 32 + (__ffs((int)bits) - 1);
        multi_visit(p, ub, [&](i64, i64, i64 size) { nr++; nm += (u64)size; });
    }
    u64 ir = warp_incl_sum(nr), im = warp_incl_sum(nm);
    u64 at_r = tile_rec[tile] + ir - nr, at_m = tile_mem[tile] + im - nm;
    for (u32 bits = bits0; bits; bits &= bits - 1u) {
        i64 ub = word * 32 + (__ffs((int)bits) - 1);
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
}

// ---- multi-MEM sweep: one thread per segment start (see mem_walk) ----------------------------------------------
__global__ void __launch_bounds__(SW_THREADS) mems_count_kernel(SweepArgs p, int *__restrict__ st_l, int *__restrict__ st_lb,
                                                                u64 *__restrict__ tile_rec, u64 *__restrict__ tile_mem, u32 *__restrict__ hitbits) {
    grid_dep_wait();
    __shared__ u64 s1[33], s2[33];
    u64 nr = 0, nm = 0;
    for (int c = 0; c < SW_CHUNKS; c++) {
        i64 i = (i64)blockIdx.x * SW_TILE + c * SW_THREADS + threadIdx.x;
        u64 r0 = nr;
        if (mem_segment_start(p, i)) mem_walk(p, i, st_l, st_lb, [&](i64, i64, i64, i64 size) { nr++; nm += (u64)size; });
        unsigned m = __ballot_sync(FULL, nr != r0);
        if ((threadIdx.x & 31u) == 0) hitbits[i >> 5] = m;
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
mems_write_kernel(SweepArgs p, int *__restrict__ st_l, int *__restrict__ st_lb, const u64 *__restrict__ tile_rec, const u64 *__restrict__ tile_mem,
                  const u32 *__restrict__ hitbits, i64 tiles, i64 *__restrict__ hdr, i64 hdr_cap, i64 *__restrict__ members, i64 mem_cap) {
    grid_dep_wait();
    const i64 tile = (i64)blockIdx.x * (SW_THREADS / 32) + (threadIdx.x >> 5);
    if (tile >= tiles) return;  // warp-uniform
    const unsigned lane = threadIdx.x & 31u;
    const i64 word = tile * (SW_TILE / 32) + lane;
    const u32 bits0 = word * 32 < p.n ? hitbits[word] : 0u;
    u64 nr = 0, nm = 0;
    for (u32 bits = bits0; bits; bits &= bits - 1u) {
        i64 i = word * 32 + (__ffs((int)bits) - 1);
        mem_walk(p, i, st_l, st_lb, [&](i64, i64, i64, i64 size) { nr++; nm += (u64)size; });
    }
    u64 ir = warp_incl_sum(nr), im = warp_incl_sum(nm);
    u64 at_r = tile_rec[tile] + ir - nr, at_m = tile_mem[tile] + im - nm;
    for (u32 bits = bits0; bits; bits &= bits - 1u) {
        i64 i = word * 32 + (__ffs((int)bits) - 1);
        mem_walk(p, i, st_l, st_lb, [&](i64 l, i64 c, i64 lb, i64 size) {
            if ((i64)at_r < hdr_cap) {
                hdr[3 * at_r + 0] = l;
                hdr[3 * at_r + 1] = c;          // number of distinct samples (reveal.c:353), NOT the member count
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
}

// single block: in-place exclusive scan of up to two u64 arrays; totals -> out[0], out[1].  A thread takes TS_K consecutive tiles
// (all its loads in flight at once, a serial sum in registers), so a text of 8 million slots is one block-wide scan instead of eight.
static const int TS_K = 8;
__global__ void __launch_bounds__(1024) sweep_tilescan_kernel(u64 *__restrict__ a, u64 *__restrict__ b, i64 tiles, u64 *__restrict__ out) {
    grid_dep_wait();
    __shared__ u64 s1[33], s2[33];
    u64 ca = 0, cb = 0;
    for (i64 b0 = 0; b0 < tiles; b0 += 1024 * TS_K) {
        const i64 t0 = b0 + (i64)threadIdx.x * TS_K;
        u64 va[TS_K], vb[TS_K];
#pragma unroll
        for (int j = 0; j < TS_K; j++) {
            va[j] = t0 + j < tiles ? a[t0 + j] : 0ull;
            vb[j] = (b && t0 + j < tiles) ? b[t0 + j] : 0ull;
        }
        u64 sa = 0, sb = 0;
#pragma unroll
        for (int j = 0; j < TS_K; j++) {
            sa += va[j];
            sb += vb[j];
        }
        u64 ta, tb = 0, ib = 0;
        const u64 ia = block_incl_sum<1024, u64>(sa, s1, &ta);
        if (b) ib = block_incl_sum<1024, u64>(sb, s2, &tb);  // (b is the same for every thread)
        u64 ra = ca + ia - sa, rb = cb + ib - sb;
#pragma unroll
        for (int j = 0; j < TS_K; j++) {
            if (t0 + j < tiles) {
                a[t0 + j] = ra;
                if (b) b[t0 + j] = rb;
            }
            ra += va[j];
            rb += vb[j];
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
    // [tile_rec][tile_mem][totals 8 words][hit bits: one u32 per 32 slots, padded to whole tiles]
    return (size_t)(2 * tiles + 8) * 8 + (size_t)tiles * (SW_TILE / 32) * 4 + 512;
}

// count pass + tile scan; when a result buffer is already at hand (d_spec, cap_spec rows) the write pass is
// enqueued right behind them, before the host reads the count back -- one synchronisation instead of two.
// The caller re-runs sweep_pair_write if the count turns out to exceed cap_spec.
int sweep_pair_count(Stream &st, const SweepArgs &p, void *scratch, i64 *count, i64 *d_spec, i64 cap_spec) {
    *count = 0;
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    u64 *tile_rec = (u64 *)scratch;
    u64 *totals = tile_rec + 2 * tiles;
    RV_TRY(prof_begin(st, RV_PROF_SWEEP));
    RV_LAUNCH(pair_count_kernel, (unsigned)tiles, SW_THREADS, 0, st.s, p, tile_rec, (u32 *)(tile_rec + 2 * tiles + 8));
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)p.n * 9));
    RV_LAUNCH_PDL(sweep_tilescan_kernel, 1, 1024, 0, st.s, tile_rec, (u64 *)nullptr, tiles, totals);
    st.launches += 2;
    u64 *h = (u64 *)(st.pinned + 304);  // pinned: the read-back is asynchronous and the speculative write pass is enqueued before the host waits
    RV_CUDA(cudaMemcpyAsync(h, totals, 16, cudaMemcpyDeviceToHost, st.s));
    if (d_spec && cap_spec > 0) RV_TRY(sweep_pair_write(st, p, scratch, d_spec, cap_spec));
    RV_CUDA(cudaStreamSynchronize(st.s));
    *count = (i64)h[0];
    if (st.prof && d_spec && cap_spec > 0) st.prof_bytes[RV_PROF_SWEEP] += (long long)(*count < cap_spec ? *count : cap_spec) * (24 + 13);  // rows written + the slots re-read for them
    return RV_OK;
}

int sweep_pair_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_out, i64 cap) {
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    RV_TRY(prof_begin(st, RV_PROF_SWEEP));
    const u64 *tile_rec = (const u64 *)scratch;
    const unsigned wblocks = (unsigned)((tiles + SW_THREADS / 32 - 1) / (SW_THREADS / 32));
    RV_LAUNCH_PDL(pair_write_kernel, wblocks, SW_THREADS, 0, st.s, p, tile_rec, (const u32 *)(tile_rec + 2 * tiles + 8), tiles, d_out, cap);
    // the sparse pass reads one hit bit per slot and a tile offset per 1024 slots; the rows it writes are booked by the caller
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)(p.n / 8 + tiles * 8)));
    st.launches++;
    RV_KCHECK();
    return RV_OK;
}

int sweep_multi_count(Stream &st, const SweepArgs &p, void *scratch, i64 *nrec, i64 *nmem, i64 *d_hdr_spec, i64 hdr_cap_spec, i64 *d_mem_spec,
                      i64 mem_cap_spec) {
    *nrec = *nmem = 0;
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    u64 *tile_rec = (u64 *)scratch, *tile_mem = tile_rec + tiles;
    u64 *totals = tile_rec + 2 * tiles;
    RV_TRY(prof_begin(st, RV_PROF_SWEEP));
    RV_LAUNCH(multi_count_kernel, (unsigned)tiles, SW_THREADS, 0, st.s, p, tile_rec, tile_mem, (u32 *)(tile_rec + 2 * tiles + 8));
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)p.n * 11));
    RV_LAUNCH_PDL(sweep_tilescan_kernel, 1, 1024, 0, st.s, tile_rec, tile_mem, tiles, totals);
    st.launches += 2;
    u64 *h = (u64 *)(st.pinned + 304);
    RV_CUDA(cudaMemcpyAsync(h, totals, 16, cudaMemcpyDeviceToHost, st.s));
    if (d_hdr_spec && hdr_cap_spec > 0 && mem_cap_spec > 0)
        RV_TRY(sweep_multi_write(st, p, scratch, d_hdr_spec, hdr_cap_spec, d_mem_spec, mem_cap_spec));
    RV_CUDA(cudaStreamSynchronize(st.s));
    *nrec = (i64)h[0];
    *nmem = (i64)h[1];
    if (st.prof && d_hdr_spec && hdr_cap_spec > 0 && mem_cap_spec > 0) st.prof_bytes[RV_PROF_SWEEP] += (long long)*nrec * 24 + (long long)*nmem * (16 + 15);
    return RV_OK;
}

int sweep_multi_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_hdr, i64 hdr_cap, i64 *d_mem, i64 mem_cap) {
    if (p.n < 2) return RV_OK;
    i64 tiles = (p.n + SW_TILE - 1) / SW_TILE;
    const u64 *tile_rec = (const u64 *)scratch, *tile_mem = tile_rec + tiles;
    RV_TRY(prof_begin(st, RV_PROF_SWEEP));
    const unsigned wblocks = (unsigned)((tiles + SW_THREADS / 32 - 1) / (SW_THREADS / 32));
    RV_LAUNCH_PDL(multi_write_kernel, wblocks, SW_THREADS, 0, st.s, p, tile_rec, tile_mem, (const u32 *)(tile_rec + 2 * tiles + 8), tiles, d_hdr, hdr_cap,
              d_mem, mem_cap);
    RV_TRY(prof_end(st, RV_PROF_SWEEP, 1, (long long)(p.n / 8 + tiles * 16)));
    st.launches++;
    RV_KCHECK();
    return RV_OK;
}

}  // namespace rv

namespace rv {

// getmultimems: scratch = sweep scratch followed by the two stack arrays (n ints each)
size_t mems_scratch_bytes(i64 n) { return sweep_scratch_bytes(n) + (size_t)(n + 64) * 8 + 512; }

static void mems_layout(void *scratch, i64 n, u64 **tile_rec, u64 **tile_mem, u64 **totals, u32 **hitbits, int **st_l, int **st_lb, i64 *tiles) {
    *tiles = (n + SW_TILE - 1) / SW_TILE;
    *tile_rec = (u64 *)scratch;
    *tile_mem = *tile_rec + *tiles;
    *totals = *tile_rec + 2 * *tiles;
    *hitbits = (u32 *)(*tile_rec + 2 * *tiles + 8);
    unsigned char *after = (unsigned char *)scratch + (sweep_scratch_bytes(n) + 255) / 256 * 256;
    *st_l = (int *)after;
    *st_lb = *st_l + (n + 32);
}

int sweep_mems_count(Stream &st, const SweepArgs &p, void *scratch, i64 *nrec, i64 *nmem) {
    *nrec = *nmem = 0;
    if (p.n < 2) return RV_OK;
    u64 *tile_rec, *tile_mem, *totals;
    u32 *hitbits;
    int *st_l, *st_lb;
    i64 tiles;
    mems_layout(scratch, p.n, &tile_rec, &tile_mem, &totals, &hitbits, &st_l, &st_lb, &tiles);
    RV_LAUNCH(mems_count_kernel, (unsigned)tiles, SW_THREADS, 0, st.s, p, st_l, st_lb, tile_rec, tile_mem, hitbits);
    RV_LAUNCH_PDL(sweep_tilescan_kernel, 1, 1024, 0, st.s, tile_rec, tile_mem, tiles, totals);
    st.launches += 2;
    u64 *h = (u64 *)(st.pinned + 304);
    RV_CUDA(cudaMemcpyAsync(h, totals, 16, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));
    *nrec = (i64)h[0];
    *nmem = (i64)h[1];
    return RV_OK;
}

int sweep_mems_write(Stream &st, const SweepArgs &p, void *scratch, i64 *d_hdr, i64 hdr_cap, i64 *d_mem, i64 mem_cap) {
    if (p.n < 2) return RV_OK;
    u64 *tile_rec, *tile_mem, *totals;
    u32 *hitbits;
    int *st_l, *st_lb;
    i64 tiles;
    mems_layout(scratch, p.n, &tile_rec, &tile_mem, &totals, &hitbits, &st_l, &st_lb, &tiles);
    const unsigned wblocks = (unsigned)((tiles + SW_THREADS / 32 - 1) / (SW_THREADS / 32));
    RV_LAUNCH_PDL(mems_write_kernel, wblocks, SW_THREADS, 0, st.s, p, st_l, st_lb, tile_rec, tile_mem, hitbits, tiles, d_hdr, hdr_cap, d_mem, mem_cap);
    st.launches++;
    RV_KCHECK();
    return RV_OK;
}

}  // namespace rv
