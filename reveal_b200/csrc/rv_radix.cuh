// rv_radix.cuh -- hand-written LSD radix sort of (key, 32-bit value) pairs for
// sm_100a: one histogram kernel for all digit positions, then one single-pass
// ("onesweep"-style) kernel per digit: tiles are claimed in order through an
// atomic ticket, ranked with warp match/ballot, staged in shared memory in
// bucket order and scattered to HBM in contiguous per-bucket runs; the
// cross-tile bucket offsets come from a decoupled look-back over per-tile
// (flag|count) words, so every pass reads and writes each pair exactly once.
//
// Replaces, together with rv_sa.cu, the reference's divsufsort() call
// (reveallib/interface.c:213-222 -> divsufsort/divsufsort.c:333).
#pragma once
#include "rv_internal.h"

namespace rv {

#ifndef RV_TMA_ON
#define RV_TMA_ON 1
#endif
#ifndef RV_RS_MATCH
#define RV_RS_MATCH 0
#endif
static const int RS_THREADS = 256;
static const int RS_WARPS = RS_THREADS / 32;
#ifndef RV_RS_IPT
#define RV_RS_IPT 16
#endif
static const int RS_IPT = RV_RS_IPT;
static const int RS_TILE = RS_THREADS * RS_IPT;  // 4096 pairs per tile
static const int RS_BINS = 256;
static const int RS_MAXPASS = 8;

static const u32 RS_FLAG_AGG = 1u << 30;   // tile aggregate available
static const u32 RS_FLAG_INCL = 2u << 30;  // inclusive prefix available
static const u32 RS_FLAG_MASK = 3u << 30;
static const u32 RS_VAL_MASK = ~RS_FLAG_MASK;  // => item counts must stay below 2^30

struct RadixPlan {
    int npass;
    int shift[RS_MAXPASS];
    u32 mask[RS_MAXPASS];
};

// Digits over the bit ranges [lo0,hi0) and [lo1,hi1) of the key (second range
// optional), each range split into the fewest <=8-bit digits of even width.
inline RadixPlan make_plan(int lo0, int hi0, int lo1 = 0, int hi1 = 0) {
    RadixPlan p;
    p.npass = 0;
    int lo[2] = {lo0, lo1}, hi[2] = {hi0, hi1};
    for (int r = 0; r < 2; r++) {
        int bits = hi[r] - lo[r];
        if (bits <= 0) continue;
        int np = (bits + 7) / 8;
        int at = lo[r];
        for (int k = 0; k < np; k++) {
            int w = (bits - (at - lo[r]) + (np - k) - 1) / (np - k);
            p.shift[p.npass] = at;
            p.mask[p.npass] = (1u << w) - 1u;
            p.npass++;
            at += w;
        }
    }
    return p;
}

inline int mask_bits(u32 m) {
    int b = 0;
    while (m) { b++; m >>= 1; }
    return b;
}

inline size_t radix_scratch_bytes(i64 n) {
    i64 tiles = (n + RS_TILE - 1) / RS_TILE;
    // [hist npass*256][base npass*256][ticket 8][status npass*tiles*256]
    return (size_t)(2 * RS_MAXPASS * RS_BINS + 8 + 32) * 4 + (size_t)RS_MAXPASS * tiles * RS_BINS * 4 + 512;
}

// Keys that are never materialised: key[i] = the first k symbols of suffix i of the text as a base-`base` number over the
// symbol CLASSES (rv_sa.cu stage 2), value[i] = i.  The histogram kernel and the first digit pass compute them on the fly, so
// the 8-12 bytes per suffix of a key/value array are neither written nor read back.
//
// Classes: every frequent symbol (the ones that carry the entropy: ACGT) has a digit of its own, in byte order.  A rare symbol
// ('$', N, a stray IUPAC letter) and "past the end of the text" END the key: a rare symbol takes the digit of the next frequent
// symbol above it and the remaining digits are 0; with no frequent symbol above it takes the largest digit and the remaining
// digits are the largest, too; past the end is digit 0 followed by zeros.  This keeps the key order-preserving
// (suffix a < suffix b  =>  key a <= key b) with `base` = number of frequent symbols -- 2 bits per base for DNA whatever else
// occurs in the text -- and ties mean nothing: equal keys only say "not ordered yet".
struct TextKeySrc {
    const unsigned char *T;
    i64 n;
    u32 base;
    int k;
    u64 top;                  // base^(k-1)
    u64 pw[64];               // base^j (j < k)
    unsigned short code[256]; // bits 0-7 digit, bit 8 rare (ends the key), bit 9 fill the rest with the largest digit
};
static const int TK_PER = 16;  // consecutive positions whose keys one thread rolls
static const u32 TK_RARE = 0x100u, TK_MAXFILL = 0x200u;

template <typename KeyT> struct KeyRoller {
    KeyT raw;     // all k digits, no truncation
    u64 rm, xm;   // bit t: symbol t of the window is rare / rare with max fill
    __device__ __forceinline__ u32 sym(const TextKeySrc &s, const unsigned short *s_code, i64 p) const { return p < s.n ? (u32)s_code[s.T[p]] : TK_RARE; }
    __device__ __forceinline__ void first(const TextKeySrc &s, const unsigned short *s_code, i64 i0) {
        raw = 0;
        rm = xm = 0;
        for (int t = 0; t < s.k; t++) {
            u32 c = sym(s, s_code, i0 + t);
            raw = raw * (KeyT)s.base + (KeyT)(c & 0xffu);
            rm |= (u64)((c >> 8) & 1u) << t;
            xm |= (u64)((c >> 9) & 1u) << t;
        }
    }
    // window i -> i+1
    __device__ __forceinline__ void next(const TextKeySrc &s, const unsigned short *s_code, i64 p) {
        u32 out = sym(s, s_code, p), in = sym(s, s_code, p + s.k);
        raw = (raw - (KeyT)(out & 0xffu) * (KeyT)s.top) * (KeyT)s.base + (KeyT)(in & 0xffu);
        rm = (rm >> 1) | ((u64)((in >> 8) & 1u) << (s.k - 1));
        xm = (xm >> 1) | ((u64)((in >> 9) & 1u) << (s.k - 1));
    }
    __device__ __forceinline__ KeyT key(const TextKeySrc &s) const {
        if (rm == 0) return raw;
        const int d = __ffsll((long long)rm) - 1;  // first rare symbol of the window: the key ends there
        const KeyT P = (KeyT)s.pw[s.k - 1 - d];
        KeyT kk = raw - raw % P;
        if ((xm >> d) & 1ull) kk += P - 1;
        return kk;
    }
};

// DNA fast path of the roller: four frequent symbols (2-bit digits) and 32-bit keys (k <= 16).  The TK_PER = 16 consecutive
// suffixes of a thread need the 16 + k - 1 <= 31 symbols from its first position on: two 128-bit loads of the text, the
// classes of the 32 symbols packed two bits each into one 64-bit word (first symbol in the top bits) and the rare / max-fill
// flags one bit per symbol; every key is then a shift and a mask of that word -- the same value KeyRoller::key() yields.
struct KeyBlock4 {
    u64 P;       // class of symbol t in bits [62 - 2t, 64 - 2t)
    u32 rm, xm;  // bit t: symbol t is rare / rare with max fill
    // usable when the whole 32-byte window lies inside the text and is 16-byte aligned
    static __device__ __forceinline__ bool fits(const TextKeySrc &s, i64 i0) {
        return TK_PER == 16 && s.base == 4u && s.k <= 16 && i0 + 32 <= s.n && ((((size_t)s.T) + (size_t)i0) & 15u) == 0u;
    }
    __device__ __forceinline__ void load(const unsigned char *p, const unsigned short *s_code) {
        const uint4 a = *(const uint4 *)p, b = *(const uint4 *)(p + 16);
        const u32 w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        P = 0;
        rm = xm = 0;
#pragma unroll
        for (int t = 0; t < 32; t++) {
            const u32 c = s_code[(w[t >> 2] >> (8 * (t & 3))) & 0xffu];
            P = (P << 2) | (u64)(c & 3u);
            rm |= ((c >> 8) & 1u) << t;
            xm |= ((c >> 9) & 1u) << t;
        }
    }
    __device__ __forceinline__ u32 key(int j, int k) const {  // suffix at symbol j (0..15) of the window
        const u32 raw = (u32)(P >> (2 * (32 - j - k))) & (k >= 16 ? 0xffffffffu : ((1u << (2 * k)) - 1u));
        const u32 r = (rm >> j) & ((1u << k) - 1u);
        if (r == 0u) return raw;
        const int d = __ffs((int)r) - 1;                    // first rare symbol of the window: the key ends there
        const u32 low = (1u << (2 * (k - 1 - d))) - 1u;     // the digits behind it
        return ((xm >> (j + d)) & 1u) ? (raw | low) : (raw & ~low);
    }
};

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_text_kernel(TextKeySrc src, RadixPlan plan, u32 *__restrict__ ghist) {
    grid_dep_wait();
    __shared__ u32 sh[RS_MAXPASS * RS_BINS];
    __shared__ unsigned short s_code[256];
    s_code[threadIdx.x] = src.code[threadIdx.x];  // RS_THREADS == 256
    for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const i64 stride = (i64)gridDim.x * RS_THREADS * TK_PER;
    int byte_digits = plan.npass == 3 || plan.npass == 4 ? plan.npass : 0;  // npass if digit p is byte p of the key
    for (int p = 0; p < plan.npass; p++)
        if (plan.shift[p] != 8 * p || plan.mask[p] != 0xffu) byte_digits = 0;
    for (i64 i0 = ((i64)blockIdx.x * RS_THREADS + threadIdx.x) * TK_PER; i0 < src.n; i0 += stride) {
        if (sizeof(KeyT) == 4 && KeyBlock4::fits(src, i0)) {
            KeyBlock4 kb;
            kb.load(src.T + i0, s_code);
            if (byte_digits == 3) {  // the digits are the bytes of the key (12- and 16-mers of DNA): no plan look-ups, no loop
#pragma unroll
                for (int j = 0; j < TK_PER; j++) {
                    const u32 key = kb.key(j, src.k);
                    atomicAdd(&sh[key & 0xffu], 1u);
                    atomicAdd(&sh[RS_BINS + ((key >> 8) & 0xffu)], 1u);
                    atomicAdd(&sh[2 * RS_BINS + ((key >> 16) & 0xffu)], 1u);
                }
            } else if (byte_digits == 4) {
#pragma unroll
                for (int j = 0; j < TK_PER; j++) {
                    const u32 key = kb.key(j, src.k);
                    atomicAdd(&sh[key & 0xffu], 1u);
                    atomicAdd(&sh[RS_BINS + ((key >> 8) & 0xffu)], 1u);
                    atomicAdd(&sh[2 * RS_BINS + ((key >> 16) & 0xffu)], 1u);
                    atomicAdd(&sh[3 * RS_BINS + (key >> 24)], 1u);
                }
            } else {
#pragma unroll
                for (int j = 0; j < TK_PER; j++) {
                    const u32 key = kb.key(j, src.k);
                    for (int p = 0; p < plan.npass; p++) atomicAdd(&sh[p * RS_BINS + ((key >> plan.shift[p]) & plan.mask[p])], 1u);
                }
            }
            continue;
        }
        KeyRoller<KeyT> kr;
        kr.first(src, s_code, i0);
        for (int j = 0; j < TK_PER && i0 + j < src.n; j++) {
            const KeyT key = kr.key(src);
            for (int p = 0; p < plan.npass; p++) atomicAdd(&sh[p * RS_BINS + ((u32)(key >> plan.shift[p]) & plan.mask[p])], 1u);
            kr.next(src, s_code, i0 + j);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += RS_THREADS) {
        u32 c = sh[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const KeyT *__restrict__ keys, i64 n, RadixPlan plan, u32 *__restrict__ ghist) {
    grid_dep_wait();
    __shared__ u32 sh[RS_MAXPASS * RS_BINS];
    for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    i64 stride = (i64)gridDim.x * RS_THREADS;
    for (i64 i = (i64)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        KeyT k = keys[i];
        for (int p = 0; p < plan.npass; p++) {
            u32 d = (u32)(k >> plan.shift[p]) & plan.mask[p];
            atomicAdd(&sh[p * RS_BINS + d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npass * RS_BINS; i += RS_THREADS) {
        u32 c = sh[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// one block per pass: exclusive scan of the 256 bucket counts
__global__ void __launch_bounds__(RS_BINS) rs_scan_kernel(const u32 *__restrict__ ghist, u32 *__restrict__ gbase) {
    grid_dep_wait();
    __shared__ u32 scratch[33];
    u32 c = ghist[blockIdx.x * RS_BINS + threadIdx.x];
    u32 total;
    u32 inc = block_incl_sum<RS_BINS>(c, scratch, &total);
    gbase[blockIdx.x * RS_BINS + threadIdx.x] = inc - c;
}

#ifndef RV_RS_MINBLOCKS
#define RV_RS_MINBLOCKS 5
#endif
template <typename KeyT, bool HAS_VAL, bool FROM_TEXT>
__global__ void __launch_bounds__(RS_THREADS, RV_RS_MINBLOCKS)
rs_pass_kernel(const KeyT *__restrict__ kin, KeyT *__restrict__ kout, const u32 *__restrict__ vin, u32 *__restrict__ vout,
               i64 n, int shift, u32 mask, int dbits, const u32 *__restrict__ gbase, u32 *status, u32 *ticket, TextKeySrc src) {
    grid_dep_wait();
    __shared__ u32 s_whist[RS_WARPS * RS_BINS];  // per-warp bucket counts -> running per-warp offsets inside the bucket
    __shared__ u32 s_start[RS_BINS];             // first tile-local slot of each bucket
    __shared__ u32 s_off[RS_BINS];               // global slot of a bucket's first item minus s_start (mod 2^32)
    __shared__ u32 s_scan[33];
    __shared__ u32 s_tile;
    __shared__ __align__(8) u64 s_bar;  // mbarrier of the tile's TMA loads
    RV_DYN_SMEM(unsigned char, smem);
    KeyT *s_keys = (KeyT *)smem;
    u32 *s_vals = (u32 *)(smem + (size_t)RS_TILE * sizeof(KeyT));

    const unsigned tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
    if (tid == 0) {
        s_tile = atomicAdd(ticket, 1u);
        mbar_init(&s_bar, 1);
    }
    for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) s_whist[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const i64 base = (i64)tile * RS_TILE;
    const int cnt = (int)((n - base) < (i64)RS_TILE ? (n - base) : (i64)RS_TILE);
    // full tiles of real pairs arrive by TMA: two bulk copies (keys, values) land in the staging area in input
    // order, the threads pick their striped items out of shared memory instead of issuing 32 global loads each
    const bool by_tma = !FROM_TEXT && cnt == RS_TILE && RV_TMA_ON;
    if (by_tma && tid == 0) {
        mbar_expect_tx(&s_bar, (u32)(RS_TILE * (sizeof(KeyT) + (HAS_VAL ? 4 : 0))));
        tma_load_1d(s_keys, kin + base, (u32)(RS_TILE * sizeof(KeyT)), &s_bar);
        if (HAS_VAL) tma_load_1d(s_vals, vin + base, (u32)(RS_TILE * 4), &s_bar);
    }

    // ---- load (warp-striped: a warp owns 32*IPT consecutive pairs) + early per-warp bucket counts ----
    KeyT key[RS_IPT];
    const int wbase = (int)w * 32 * RS_IPT;
    u32 *wh = s_whist + w * RS_BINS;
    if (FROM_TEXT) {
        // keys of RS_IPT consecutive suffixes rolled per thread, exchanged through shared memory into the striped order
        unsigned short *s_code = (unsigned short *)s_off;  // 512 B: s_off is written much later
        s_code[tid] = src.code[tid];
        __syncthreads();
        // (one pad element per 32 keeps the blocked writes -- a thread's RS_IPT consecutive keys -- and the striped reads free of
        //  bank conflicts; the tail of the padded array lies in the value area, which is not in use yet)
        const i64 i0 = base + (i64)tid * RS_IPT;
        if (i0 < n) {
            if (sizeof(KeyT) == 4 && RS_IPT == TK_PER && KeyBlock4::fits(src, i0)) {
                KeyBlock4 kb;
                kb.load(src.T + i0, s_code);
#pragma unroll
                for (int j = 0; j < RS_IPT; j++) {
                    const int e = (int)tid * RS_IPT + j;
                    s_keys[e + (e >> 5)] = (KeyT)kb.key(j, src.k);
                }
            } else {
                KeyRoller<KeyT> kr;
                kr.first(src, s_code, i0);
#pragma unroll
                for (int j = 0; j < RS_IPT; j++) {
                    if (i0 + j < n) {
                        const int e = (int)tid * RS_IPT + j;
                        s_keys[e + (e >> 5)] = kr.key(src);
                        kr.next(src, s_code, i0 + j);
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RS_IPT; k++) {
            int idx = wbase + k * 32 + (int)l;
            key[k] = idx < cnt ? s_keys[idx + (idx >> 5)] : (KeyT)0;
        }
        __syncthreads();  // s_keys is reused as the bucket-ordered staging area below
    } else if (by_tma) {
        mbar_wait(&s_bar, 0);
#pragma unroll
        for (int k = 0; k < RS_IPT; k++) key[k] = s_keys[wbase + k * 32 + (int)l];
        // (the __syncthreads after the early counts separates these reads from the bucket-ordered staging writes)
    } else {
#pragma unroll
        for (int k = 0; k < RS_IPT; k++) {
            int idx = wbase + k * 32 + (int)l;
            key[k] = idx < cnt ? kin[base + idx] : (KeyT)0;
        }
    }
#pragma unroll
    for (int k = 0; k < RS_IPT; k++) {
        int idx = wbase + k * 32 + (int)l;
        if (idx < cnt) atomicAdd(&wh[(u32)(key[k] >> shift) & mask], 1u);
    }
    __syncthreads();

    // ---- per bucket: exclusive warp offsets, tile count (published at once), tile-local start ----
    u32 run = 0;
    u32 *my = status + (size_t)tile * RS_BINS + tid;  // RS_THREADS == RS_BINS: thread d owns bucket d
    {
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            u32 t = s_whist[ww * RS_BINS + tid];
            s_whist[ww * RS_BINS + tid] = run;
            run += t;
        }
        // successors can add this tile's count while it is still ranking
        st_volatile(my, (tile == 0 ? RS_FLAG_INCL : RS_FLAG_AGG) | run);
        u32 total;
        u32 inc = block_incl_sum<RS_THREADS>(run, s_scan, &total);
        s_start[tid] = inc - run;
    }
    __syncthreads();

    // ---- rank with ballots (peers = lanes of the warp holding the same digit) and stage in bucket order ----
    unsigned short slot[RS_IPT];  // tile-local destination of every pair (the values follow in one batch of loads)
#pragma unroll
    for (int k = 0; k < RS_IPT; k++) {
        int idx = wbase + k * 32 + (int)l;
        bool valid = idx < cnt;
        u32 d = (u32)(key[k] >> shift) & mask;
#if RV_RS_MATCH
        // one match instruction instead of a ballot per digit bit; lanes past the end of the tile share a value no digit has
        unsigned peers = __match_any_sync(FULL, valid ? d : 0xffffffffu);
#else
        unsigned peers = __ballot_sync(FULL, valid);
        for (int b = 0; b < dbits; b++) {
            unsigned bal = __ballot_sync(FULL, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? bal : ~bal;
        }
#endif
        u32 old = 0;
        const int leader = __ffs((int)peers) - 1;  // invalid lanes: garbage, unused
        if (valid && (int)l == leader) {
            old = wh[d];
            wh[d] = old + (u32)__popc(peers);
        }
        old = __shfl_sync(FULL, old, leader & 31);
        if (valid) {
            u32 p = s_start[d] + old + (u32)__popc(peers & lanemask_lt());
            s_keys[p] = key[k];
            slot[k] = (unsigned short)p;
        }
        __syncwarp();
    }
    if (HAS_VAL) {
        u32 val[RS_IPT];
#pragma unroll
        for (int k = 0; k < RS_IPT; k++) {
            int idx = wbase + k * 32 + (int)l;
            val[k] = FROM_TEXT ? (u32)(base + idx) : (by_tma ? s_vals[idx] : (idx < cnt ? vin[base + idx] : 0u));
        }
        if (by_tma) __syncthreads();  // every thread has its values out of the input-ordered area
#pragma unroll
        for (int k = 0; k < RS_IPT; k++) {
            int idx = wbase + k * 32 + (int)l;
            if (idx < cnt) s_vals[slot[k]] = val[k];
        }
    }

    // ---- decoupled look-back for the bucket's global offset ----
    {
        u32 excl = 0;
        if (tile > 0) {
            for (i64 t = (i64)tile - 1;; t--) {
                const u32 *q = status + (size_t)t * RS_BINS + tid;
                u32 v;
                while (((v = ld_volatile(q)) & RS_FLAG_MASK) == 0u) { RV_SPIN(); }
                excl += v & RS_VAL_MASK;
                if ((v & RS_FLAG_MASK) == RS_FLAG_INCL) break;
            }
            st_volatile(my, RS_FLAG_INCL | (excl + run));
        }
        s_off[tid] = gbase[tid] + excl - s_start[tid];
    }
    __syncthreads();

    // ---- scatter contiguous per-bucket runs ----
    for (int p = tid; p < cnt; p += RS_THREADS) {
        KeyT kk = s_keys[p];
        u32 d = (u32)(kk >> shift) & mask;
        u32 dst = s_off[d] + (u32)p;
        kout[dst] = kk;
        if (HAS_VAL) vout[dst] = s_vals[p];
    }
}

// Sorts n pairs by the digits of `plan` (least significant first), ping-ponging
// between (k0,v0) and (k1,v1).  On return *result_in_0 tells where the sorted
// pairs are.  `scratch` needs radix_scratch_bytes(n).  Stable.
// With `text` the input pairs are virtual (TextKeySrc): nothing is read from (k0,v0), the first pass writes (k1,v1).
template <typename KeyT>
int radix_sort_pairs(Stream &st, KeyT *k0, KeyT *k1, u32 *v0, u32 *v1, i64 n, const RadixPlan &plan, void *scratch, bool *result_in_0,
                     const TextKeySrc *text = nullptr) {
    *result_in_0 = true;
    if (n <= 0 || plan.npass == 0 || (n <= 1 && !text)) return RV_OK;
    if (n >= (i64)1 << 30) {
        set_error("radix_sort_pairs: n=%lld exceeds the 2^30 look-back word limit", (long long)n);
        return RV_ERR_UNSUPPORTED;
    }
    const i64 tiles = (n + RS_TILE - 1) / RS_TILE;
    u32 *ghist = (u32 *)scratch;
    u32 *gbase = ghist + RS_MAXPASS * RS_BINS;
    u32 *ticket = gbase + RS_MAXPASS * RS_BINS;
    u32 *status = ticket + 32;
    size_t zero_bytes = (size_t)(2 * RS_MAXPASS * RS_BINS + 32) * 4 + (size_t)plan.npass * tiles * RS_BINS * 4;
    RV_CUDA(cudaMemsetAsync(scratch, 0, zero_bytes, st.s));
#ifndef RV_HIST_BPS
#define RV_HIST_BPS 8   // blocks per SM of the (grid-stride) histogram kernels (measured at C2: 5 = one resident wave and 10 are 0.5 % slower)
#endif
    int hist_blocks = (int)(tiles < 148 * RV_HIST_BPS ? tiles : 148 * RV_HIST_BPS);
    TextKeySrc none;
    memset(&none, 0, sizeof none);
    if (text) {
        RV_LAUNCH((rs_hist_text_kernel<KeyT>), hist_blocks, RS_THREADS, 0, st.s, *text, plan, ghist);
    } else {
        RV_LAUNCH((rs_hist_kernel<KeyT>), hist_blocks, RS_THREADS, 0, st.s, k0, n, plan, ghist);
    }
    RV_LAUNCH_PDL(rs_scan_kernel, plan.npass, RS_BINS, 0, st.s, ghist, gbase);
    st.launches += 2;
    const size_t smem = (size_t)RS_TILE * (sizeof(KeyT) + 4);
    static bool attr_done[64] = {false};  // per device: a process may drive several GPUs (rv_set_device)
    int dev = 0;
    RV_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr_done[dev]) {
        auto kfn = rs_pass_kernel<KeyT, true, false>;
        auto kft = rs_pass_kernel<KeyT, true, true>;
        (void)kfn;
        (void)kft;
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done[dev] = true;
    }
    bool in0 = true;
    RV_TRY(prof_begin(st, RV_PROF_RADIX_PASS));
    for (int p = 0; p < plan.npass; p++) {
        if (p == 0 && text) {
            RV_LAUNCH_PDL((rs_pass_kernel<KeyT, true, true>), (unsigned)tiles, RS_THREADS, smem, st.s, (const KeyT *)nullptr, k1, (const u32 *)nullptr, v1, n,
                      plan.shift[p], plan.mask[p], mask_bits(plan.mask[p]), gbase + p * RS_BINS, status + (size_t)p * tiles * RS_BINS, ticket + p,
                      *text);
        } else {
            RV_LAUNCH_PDL((rs_pass_kernel<KeyT, true, false>), (unsigned)tiles, RS_THREADS, smem, st.s, in0 ? k0 : k1, in0 ? k1 : k0, in0 ? v0 : v1,
                      in0 ? v1 : v0, n, plan.shift[p], plan.mask[p], mask_bits(plan.mask[p]), gbase + p * RS_BINS,
                      status + (size_t)p * tiles * RS_BINS, ticket + p, none);
        }
        st.launches++;
        in0 = !in0;
    }
    RV_KCHECK();
    // algorithmic bytes: a virtual first pass reads one text byte per pair instead of key + value
    long long pair_bytes = (long long)(sizeof(KeyT) + 4);
    long long bytes = (long long)plan.npass * n * pair_bytes * 2 - (text ? n * (pair_bytes - 1) : 0);
    RV_TRY(prof_end(st, RV_PROF_RADIX_PASS, plan.npass, bytes));
    *result_in_0 = in0;
    return RV_OK;
}

}  // namespace rv
