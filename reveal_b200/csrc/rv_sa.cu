// rv_sa.cu -- GPU suffix-array builder: k-mer radix sort, direct comparison inside small
// groups, prefix doubling for the rest.
//
// Replaces the reference's  divsufsort(T, SA, n)  (reveallib/interface.c:213-222,
// divsufsort/divsufsort.c:333), the inverse fill  SAi[SA[i]] = i  (interface.c:235-238)
// and -- for every suffix the comparison stage places -- compute_lcp (interface.c:97-114).
// The suffix array of a byte string is unique (plain unsigned-byte lexicographic order, a
// suffix that is a proper prefix of another sorts first), so any correct builder is
// bit-identical to divsufsort.
//
// Stages
//   1. 256-bin histogram of T -> dense symbol codes 1..sigma (0 = past the end of the text).
//   2. key[i] = first k symbols of suffix i as a base-(sigma+1) number (order preserving); k is
//      the shortest prefix that separates ~4n random k-mers, in a 32-bit key when that fits
//      (DNA + '$': 12 symbols) else a 64-bit key; hand-written onesweep radix sort of (key, i).
//      The keys are virtual: the histogram kernel and the first digit pass roll them from the
//      text (TextKeySrc in rv_radix.cuh), no key array is written before the first scatter.
//   3. Equal-key runs are "groups".  Similar genomes give groups of a few homologous positions
//      whose suffixes agree for ~1/divergence characters.  Groups of <= SA_SMALL_G suffixes are
//      ordered by ALL-PAIRS direct comparison: one thread per member walks its earlier group
//      mates 8 text bytes per step (funnel-shifted aligned words), the warp keeps iterating
//      until every lane's queue is empty, so lanes only idle at the tail.  Every pair yields
//      who is smaller (-> a per-member count = its place in the group) and their common
//      prefix; the largest common prefix with a smaller mate is the member's LCP entry, with
//      the reference's '$'/'N' barrier applied.  A comparison longer than SA_CMP_CAP bytes
//      defers its group to stage 4.
//   4. Groups of more than SA_SMALL_G suffixes and deferred groups: prefix doubling
//      (Manber-Myers / Larsson-Sadakane) over the still-tied suffixes only, sort key
//      (rank[i] : rank[i+h]+1), h = k, 2k, 4k, ...; bounded work for long repeats.
//   When stage 4 is not needed (no repeats beyond SA_SMALL_G copies) the LCP array is complete
//   after stage 3; otherwise the caller runs the Kasai kernel of rv_lcp.cu.
#include "rv_radix.cuh"
#include <stdlib.h>

namespace rv {

static const int AP_THREADS = 256;
static const int AP_IPT = 8;
static const int AP_TILE = AP_THREADS * AP_IPT;  // 2048 active entries per tile

#ifndef RV_SA_SMALL_G
#define RV_SA_SMALL_G 16
#endif
static const int SA_SMALL_G = RV_SA_SMALL_G;     // largest group finished by direct comparison (<= 16: group geometry lives in 32-key windows)
#ifndef RV_SA_CMP_CAP
#define RV_SA_CMP_CAP 65536
#endif
static const int SA_CMP_CAP = RV_SA_CMP_CAP;   // symbols after which a comparison gives up and its group is left to stage 4

struct CodeTable {
    unsigned short code[256];  // 0 is reserved for "past the end of the text"
};

__global__ void __launch_bounds__(256) sa_bytehist_kernel(const unsigned char *__restrict__ T, i64 n, u32 *__restrict__ hist) {
    __shared__ u32 sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&sh[T[i]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// ---- group geometry ---------------------------------------------------------------------
// L / R: equal keys to the left / right of entry e, each capped at G.
template <typename KeyT>
__device__ __forceinline__ void run_lengths(const KeyT *__restrict__ keys, i64 A, i64 e, int G, int &L, int &R) {
    KeyT k = keys[e];
    L = 0;
    R = 0;
    while (L < G && e - L - 1 >= 0 && keys[e - L - 1] == k) L++;
    while (R < G && e + R + 1 < A && keys[e + R + 1] == k) R++;
}

// ---- stage 3: all-pairs comparison inside small groups -------------------------------------
// Three-level barrier bitmap: bit i of b0[] is set iff T[i] is '$' or 'N' (the reference's LCP barrier, interface.c:107);
// bit w of b1[] iff word w of b0 is non-zero; bit v of b2[] iff word v of b1 is non-zero (one b2 word covers 32768 positions,
// so the '$'/'N' cut of an LCP value of a million costs ~30 loads, not 30000).
struct Barriers {
    const u32 *b0, *b1, *b2;
};
// One pass over T: the barrier bitmaps and (PACK) the 4-bit packed text (8 symbols per word, symbol i in bits 4i..4i+3).
// A thread takes 16 consecutive characters (one 128-bit load; T is 16-byte aligned and readable, zero padded, up to n + 64),
// a block 4096 characters = 4 level-1 words.  b2 must be zeroed before the launch; blocks run one past the end of the text so
// that the packed text is zero padded.
static const int TP_THREADS = 256;
static const int TP_BLOCK = TP_THREADS * 16;
template <bool PACK>
__global__ void __launch_bounds__(TP_THREADS) sa_textprep_kernel(const unsigned char *__restrict__ T, i64 n, CodeTable tab, u32 *__restrict__ bar0,
                                                                u32 *__restrict__ bar1, u32 *__restrict__ bar2, u32 *__restrict__ packed) {
    __shared__ u32 s_nz[TP_THREADS / 32];
    __shared__ unsigned char s_code[256];
    if (PACK) {
        s_code[threadIdx.x] = (unsigned char)(tab.code[threadIdx.x] & 15u);
        __syncthreads();
    }
    const i64 i0 = (i64)blockIdx.x * TP_BLOCK + (i64)threadIdx.x * 16;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (i0 < n) q = *(const uint4 *)(T + i0);   // bytes past n (inside the padding) are zero
    const u32 wv[4] = {q.x, q.y, q.z, q.w};
    u32 bits = 0, p0 = 0, p1 = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const u32 ch = (wv[j >> 2] >> (8 * (j & 3))) & 0xffu;
        bits |= (ch == (u32)'$' || ch == (u32)'N') ? (1u << j) : 0u;
        if (PACK) {
            const u32 cd = s_code[ch];
            if (j < 8) p0 |= cd << (4 * j);
            else p1 |= cd << (4 * (j - 8));
        }
    }
    if (i0 + 16 > n) {  // the last characters of the text: nothing beyond n counts
        const int keep = i0 < n ? (int)(n - i0) : 0;
        bits &= keep >= 16 ? 0xffffu : ((1u << keep) - 1u);
        if (PACK && keep < 16) {
            if (keep <= 8) { p1 = 0; p0 &= keep == 8 ? 0xffffffffu : ((1u << (4 * keep)) - 1u); }
            else p1 &= (1u << (4 * (keep - 8))) - 1u;
        }
    }
    if (PACK) *(uint2 *)(packed + (i0 >> 3)) = make_uint2(p0, p1);
    // two neighbouring threads share one level-0 word
    const u32 hi = __shfl_down_sync(FULL, bits, 1);
    const u32 word = bits | (hi << 16);
    if ((threadIdx.x & 1u) == 0) bar0[i0 >> 5] = word;
    const unsigned nz = __ballot_sync(FULL, (threadIdx.x & 1u) == 0 && word != 0u);  // bit 2k: word k of this warp's 16
    if ((threadIdx.x & 31u) == 0) {
        u32 m = 0;
        for (int k = 0; k < 16; k++) m |= ((nz >> (2 * k)) & 1u) << k;
        s_nz[threadIdx.x >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x < TP_THREADS / 64) {  // one level-1 word per 1024 characters = two warps
        const u32 w1 = s_nz[2 * threadIdx.x] | (s_nz[2 * threadIdx.x + 1] << 16);
        const i64 v = (i64)blockIdx.x * (TP_BLOCK / 1024) + threadIdx.x;
        bar1[v] = w1;
        if (w1) atomicOr(&bar2[v >> 5], 1u << (v & 31));
    }
}

// offset of the first barrier character in T[x .. x+len), or len
__device__ __forceinline__ u32 first_barrier(const Barriers &B, u32 x, u32 len) {
    if (len == 0u) return 0u;
    u32 w = x >> 5;
    const u32 wl = (x + len - 1u) >> 5;  // last level-0 word of the range
    // The usual answer is "none".  Level 2 (one bit per 1024 characters, a few hundred words: always cached) says so alone for
    // every text that is not littered with '$' / 'N'; else level 1 (n/1024 words) when the range lies within two of its words:
    // level 0 is only touched when a barrier is near.
    if ((wl >> 5) - (w >> 5) <= 1u) {
        const u32 va = w >> 5, vb = wl >> 5;   // level-1 words = level-2 bits of the range
        const u32 qa = __ldg(B.b2 + (va >> 5)) >> (va & 31u), qb = __ldg(B.b2 + (vb >> 5)) >> (vb & 31u);
        if (((qa | qb) & 1u) == 0u) return len;
    }
    if ((wl >> 5) - (w >> 5) <= 1u) {
        const u32 v = w >> 5, vl = wl >> 5;
        u32 m = __ldg(B.b1 + v) >> (w & 31u);           // words w .. end of level-1 word v
        if (v == vl) {
            const u32 span = wl - w;
            if (span < 31u) m &= (2u << span) - 1u;
        } else {
            const u32 last = wl & 31u;                   // words 0 .. last of level-1 word vl
            u32 m2 = __ldg(B.b1 + vl);
            if (last < 31u) m2 &= (2u << last) - 1u;
            m |= m2;
        }
        if (m == 0u) return len;
    }
    {
        const u32 bits = __ldg(B.b0 + w) >> (x & 31u);
        if (bits) {
            const u32 at = (u32)(__ffs((int)bits) - 1);
            return at < len ? at : len;
        }
    }
    w++;
    while (w <= wl) {
        u32 v = w >> 5;
        const u32 m = B.b1[v] >> (w & 31u);  // level-0 words w .. end of level-1 word v
        if (m) {
            w += (u32)(__ffs((int)m) - 1);
            if (w > wl) return len;
            const u32 at = (w << 5) + (u32)(__ffs((int)B.b0[w]) - 1) - x;
            return at < len ? at : len;
        }
        v++;  // next non-empty level-1 word, through level 2
        for (;;) {
            if ((v << 5) > wl) return len;
            const u32 m2 = B.b2[v >> 5] >> (v & 31u);
            if (m2) {
                v += (u32)(__ffs((int)m2) - 1);
                break;
            }
            v = ((v >> 5) + 1u) << 5;
        }
        w = v << 5;
    }
    return len;
}

// Text as the comparison loops see it: SB bits per symbol in little-endian 32-bit words (symbol i of a word
// in bits [SB*i, SB*i+SB)).  SB = 8: the raw bytes of T.  SB = 4: order-preserving dense codes packed two per
// byte (sa_textprep_kernel) when the alphabet has at most 15 symbols -- DNA with '$', 'N' and a few IUPAC codes --
// which halves the words a comparison has to fetch.  Zero padded past the end.
template <int SB> struct Sym {
    static const u32 LOG_SPW = SB == 8 ? 2u : 3u;   // log2(symbols per word)
    static const u32 SPW = 1u << LOG_SPW;
    static const u32 LOG_SB = SB == 8 ? 3u : 2u;
    static const u32 STEP = 4u * SPW;                // symbols per comparison step (4 words per suffix)
    static const u32 MASK = (1u << SB) - 1u;
};

// Suffixes p and q are known to agree on their first h0 symbols: extends the comparison, one step = 4 words per suffix,
// funnel-shifted to the suffix start, XOR, first set bit.  Returns false when `cap` more symbols brought no decision;
// else lcp = common prefix and p_less = "suffix p sorts before suffix q" (a suffix that is a prefix of the other sorts first).
template <int SB>
__device__ __forceinline__ bool fwd_compare(const u32 *__restrict__ W, u32 n32, u32 p, u32 q, u32 h0, u32 cap, u32 &lcp, bool &p_less) {
    typedef Sym<SB> S;
    const u32 lenmin = n32 - (p > q ? p : q);
    if (h0 >= lenmin) {
        lcp = lenmin;
        p_less = p > q;
        return true;
    }
    const u32 pp = p + h0, qq = q + h0;
    const u32 *pa = W + (pp >> S::LOG_SPW), *pb = W + (qq >> S::LOG_SPW);
    const unsigned sha = (pp & (S::SPW - 1u)) * SB, shb = (qq & (S::SPW - 1u)) * SB;
    u32 lo_a = *pa, lo_b = *pb;
    u32 h = h0;
    const u32 stop = h0 + cap;
    for (;;) {
        const u32 a1 = pa[1], a2 = pa[2], a3 = pa[3], a4 = pa[4];
        const u32 b1 = pb[1], b2 = pb[2], b3 = pb[3], b4 = pb[4];
        const u32 wa0 = __funnelshift_r(lo_a, a1, sha), wb0 = __funnelshift_r(lo_b, b1, shb);
        const u32 wa1 = __funnelshift_r(a1, a2, sha), wb1 = __funnelshift_r(b1, b2, shb);
        const u32 wa2 = __funnelshift_r(a2, a3, sha), wb2 = __funnelshift_r(b2, b3, shb);
        const u32 wa3 = __funnelshift_r(a3, a4, sha), wb3 = __funnelshift_r(b3, b4, shb);
        const u32 d0 = wa0 ^ wb0, d1 = wa1 ^ wb1, d2 = wa2 ^ wb2, d3 = wa3 ^ wb3;
        if (d0 | d1 | d2 | d3) {
            const u32 wsel = d0 ? 0u : (d1 ? 1u : (d2 ? 2u : 3u));
            const u32 dd = d0 ? d0 : (d1 ? d1 : (d2 ? d2 : d3));
            const u32 va = d0 ? wa0 : (d1 ? wa1 : (d2 ? wa2 : wa3));
            const u32 vb = d0 ? wb0 : (d1 ? wb1 : (d2 ? wb2 : wb3));
            const u32 bsh = (u32)(__ffs((int)dd) - 1) & ~(u32)(SB - 1);  // bit offset of the first differing symbol
            const u32 at = h + wsel * S::SPW + (bsh >> S::LOG_SB);
            if (at >= lenmin) {  // the difference lies beyond the end of the shorter suffix
                lcp = lenmin;
                p_less = p > q;
            } else {
                lcp = at;
                p_less = ((va >> bsh) & S::MASK) < ((vb >> bsh) & S::MASK);
            }
            return true;
        }
        h += S::STEP;
        if (h >= lenmin) {
            lcp = lenmin;
            p_less = p > q;
            return true;
        }
        if (h >= stop) return false;
        pa += 4;
        pb += 4;
        lo_a = a4;
        lo_b = b4;
    }
}

// barrier-aware LCP of suffixes p and q (p = the one whose entry it is) by direct comparison
template <int SB>
__device__ __forceinline__ int direct_lcp(const u32 *__restrict__ W, u32 n32, u32 p, u32 q, const Barriers &B) {
    u32 lcp = 0;
    bool less;
    fwd_compare<SB>(W, n32, p, q, 0u, 0xC0000000u, lcp, less);
    return (int)first_barrier(B, p, lcp);
}

#ifndef RV_PR_THREADS
#define RV_PR_THREADS 128
#endif
#ifndef RV_PR_CHUNK
#define RV_PR_CHUNK 256
#endif
#ifndef RV_PR_MINBLOCKS
#define RV_PR_MINBLOCKS 9
#endif
static const int PR_THREADS = RV_PR_THREADS;
static const int PR_WARPS = PR_THREADS / 32;
static const int PR_CHUNK = RV_PR_CHUNK;               // nominal SA slots per warp
static const int PR_MAXT = PR_CHUNK + 32;              // a chunk is stretched to whole groups (<= SA_SMALL_G more), padded to rounds
static const int PR_ROUNDS = PR_MAXT / 32;
static_assert(PR_ROUNDS <= 10, "pair_chunk_setup packs one 6-bit field per round into 64 bits");
static const int PL_CAP = 512;                         // pair items per warp (a power of two >= 32 * 15 + 31)
#ifndef RV_ET_WAYS
#define RV_ET_WAYS 4
#endif
static const int ET_WAYS = RV_ET_WAYS;                          // entries per bucket of the sampled-pair table

// The comparison stage works on runs of whole groups: one warp owns about PR_CHUNK consecutive slots of the sorted (key, suffix)
// list, stretched at both ends to group boundaries.  Both kernels of the stage start with the same staging: suffixes into shared
// memory, one "starts a group" bit per slot (ballots over neighbouring keys, no loads beyond the keys themselves).
struct PairChunk {
    u32 *ssa;      // [PR_MAXT] staged suffixes
    u32 *head;     // [-1 .. PR_ROUNDS] bit t: slot t starts a group (head[-1] and head[PR_ROUNDS] are zero: 64-bit windows)
    i64 s;         // first slot of the run
    int nt;        // slots in the run
    int rounds;
    int end_closed;  // the last group of the run really ends with the run
};
// equal keys to the left (L, capped at 16) and right (R, capped at 15) of slot e, by one 32-key window around it: every lane
// loads one key, a ballot does the rest (the serial walk of run_lengths would be a chain of dependent loads at the very start
// of the warp's life).  Warp-uniform result.
template <typename KeyT>
__device__ __forceinline__ void window_run(const KeyT *__restrict__ keys, i64 n, i64 e, bool wanted, int &L, int &R) {
    const unsigned lane = threadIdx.x & 31u;
    const i64 at = e - 16 + (i64)lane;
    const bool in = wanted && at >= 0 && at < n;
    const KeyT k = in ? keys[at] : (KeyT)0;
    const KeyT k0 = __shfl_sync(FULL, k, 16);
    const unsigned eq = __ballot_sync(FULL, in && k == k0);
    L = __clz((int)~(eq << 16));             // consecutive equal keys below bit 16, downwards
    if (L > 16) L = 16;
    const unsigned up = eq >> 17;            // bits for e+1 .. e+15
    R = __ffs((int)~up) - 1;
    if (R > 15 || R < 0) R = 15;
}

// keyclz: for slot t = 32 r + lane, bits [6r, 6r+6) = leading zero bits of (key of slot t) ^ (key of slot t-1), capped at 32 --
// what sa_place_kernel needs of the keys once the staging is over (PR_ROUNDS <= 10)
template <typename KeyT>
__device__ __forceinline__ void pair_chunk_setup(PairChunk &c, const KeyT *__restrict__ keys, const u32 *__restrict__ sa, i64 n, u32 *s_head_raw /*[PR_ROUNDS+2]*/,
                                                 u64 &keyclz) {
    keyclz = 0ull;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    c.head = s_head_raw + 1;
    const i64 c0 = ((i64)blockIdx.x * PR_WARPS + w) * PR_CHUNK;
    i64 s = c0, end = c0 + PR_CHUNK < n ? c0 + PR_CHUNK : n;
    int end_closed = 1;
    {
        int L0, R0, L1, R1;
        window_run(keys, n, c0, c0 < n && c0 > 0, L0, R0);
        window_run(keys, n, end, c0 < n && end < n, L1, R1);
        if (c0 < n && c0 > 0 && L0 > 0 && L0 + R0 + 1 <= SA_SMALL_G) s = c0 + R0 + 1;  // that group belongs to the previous warp
        if (c0 < n && end < n && L1 > 0) {
            if (L1 + R1 + 1 <= SA_SMALL_G) end = end + R1 + 1;     // finish the group that straddles the nominal end
            else end_closed = 0;                                     // a large group runs through the end
        }
    }
    c.end_closed = end_closed;
    c.s = s;
    c.nt = c0 < n && end > s ? (int)(end - s) : 0;
    c.rounds = (c.nt + 31) / 32;
    if (c.nt == 0) return;  // warp-uniform
    if (lane < 2) s_head_raw[lane ? PR_ROUNDS + 1 : 0] = 0u;
    // every load of the run in flight at once (the round count is small and fixed), then the ballots
    KeyT kreg[PR_ROUNDS];
    u32 sreg[PR_ROUNDS];
    const KeyT kfirst = (lane == 0 && s > 0) ? keys[s - 1] : (KeyT)0;
#pragma unroll
    for (int r = 0; r < PR_ROUNDS; r++) {
        const int t = r * 32 + (int)lane;
        const bool valid = t < c.nt;
        kreg[r] = valid ? keys[s + t] : (KeyT)0;
        sreg[r] = valid ? sa[s + t] : 0u;
    }
    KeyT carry = kfirst;  // key of the slot before this round's first (lane 0's left neighbour)
#pragma unroll
    for (int r = 0; r < PR_ROUNDS; r++) {
        const int t = r * 32 + (int)lane;
        const i64 e = s + t;
        const bool valid = t < c.nt;
        KeyT kp = __shfl_up_sync(FULL, kreg[r], 1);
        if (lane == 0) kp = carry;
        const bool is_head = valid && (e == 0 || kreg[r] != kp);
        if (sizeof(KeyT) == 4) keyclz |= (u64)(u32)__clz((int)((u32)kreg[r] ^ (u32)kp)) << (6 * r);
        const unsigned hm = __ballot_sync(FULL, is_head);
        if (lane == 0) c.head[r] = hm;
        if (t < PR_MAXT) c.ssa[t] = sreg[r];
        carry = __shfl_sync(FULL, kreg[r], 31);
    }
    __syncwarp();
}
// members to the left / right of slot t inside its group; false when the group has more than SA_SMALL_G members (or is cut by the run's end)
__device__ __forceinline__ bool pair_chunk_group(const PairChunk &c, int t, int &L, int &R) {
    const int round = t >> 5;
    const unsigned b = (unsigned)t & 31u;
    const u64 v = ((u64)c.head[round] << 32) | (u64)c.head[round - 1];
    const u64 below = v & ((2ull << (32u + b)) - 1ull);  // heads at or before t
    L = below ? (int)(32u + b) - (63 - __clzll((long long)below)) : 0xFF;
    const u64 v2 = (((u64)c.head[round + 1] << 32) | (u64)c.head[round]) >> (b + 1u);  // heads after t
    R = v2 ? __ffsll((long long)v2) - 1 : 0xFF;
    if (R != 0xFF && t + R + 1 > c.nt) R = 0xFF;                                        // (cannot happen: no head bits beyond nt)
    if (R == 0xFF && t + 33 >= c.nt) R = c.end_closed ? c.nt - 1 - t : 0xFF;            // the run ends the group
    return L != 0xFF && R != 0xFF && L + R + 1 <= SA_SMALL_G;
}

// Similar genomes put homologous positions x (one genome) and y (the other) into one group for as long as the genomes agree:
// the pairs (x, y), (x+1, y+1), ... lie on one DIAGONAL and lcp(x+1, y+1) = lcp(x, y) - 1.  Comparing every pair down to its
// mismatch costs  (match length)^2 / 2  symbols per diagonal -- and far more inside repeats.  Instead:
//   sa_lead_kernel   only pairs whose smaller position x is a multiple of S (= one comparison step, 32 packed symbols) are
//                    compared to the end (cap SA_CMP_CAP); the result goes into a table keyed by (x, y).
//   sa_place_kernel  every pair compares ONE step; if that step finds no mismatch, the pair's diagonal has passed the sampled
//                    position x' = next multiple of S, and  lcp(x, y) = (x' - x) + lcp(x', y + x' - x)  is one table look-up.
//                    A look-up that finds nothing (the sampled pair sits in a large or postponed group, its bucket overflowed,
//                    or it never was a pair because the k-mers at x' differ) just continues the comparison.
// Table: (n/S + 1) buckets of ET_WAYS 64-bit entries  [y : 32 | lcp : 31 | x sorts first : 1], 0 = empty, bucket = x / S.
template <typename KeyT, int SB>
__global__ void __launch_bounds__(PR_THREADS, RV_PR_MINBLOCKS)
sa_lead_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ sa, i64 n, const u32 *__restrict__ W, u64 *__restrict__ etab) {
    grid_dep_wait();
    __shared__ u32 s_sa[PR_WARPS][PR_MAXT];
    __shared__ u32 s_head[PR_WARPS][PR_ROUNDS + 2];
    __shared__ unsigned short s_list[PR_WARPS][PR_MAXT];
    typedef Sym<SB> S;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    PairChunk c;
    c.ssa = s_sa[w];
    u64 keyclz_unused;
    pair_chunk_setup(c, keys, sa, n, s_head[w], keyclz_unused);
    if (c.nt == 0) return;
    unsigned short *list = s_list[w];
    // the sampled members of the run (their groups are looked at when a lane takes them)
    u32 nlist = 0;
    for (int r = 0; r < c.rounds; r++) {
        const int t = r * 32 + (int)lane;
        const bool samp = t < c.nt && (c.ssa[t] & (S::STEP - 1u)) == 0u;
        const unsigned m = __ballot_sync(FULL, samp);
        if (samp) list[nlist + (u32)__popc(m & lanemask_lt())] = (unsigned short)t;
        nlist += (u32)__popc(m);
    }
    __syncwarp();
    const u32 n32 = (u32)n;
    for (u32 i = lane; i < nlist; i += 32) {
        const int t = list[i];
        int L, R;
        if (!pair_chunk_group(c, t, L, R) || L + R == 0) continue;
        const u32 x = c.ssa[t];
        // bucket x / S belongs to this x alone (x is a multiple of S): its entries are written by this lane only, in order
        u64 *bucket = etab + (size_t)(x >> (S::LOG_SPW + 2u)) * ET_WAYS;
        int used = 0;
        for (int m = t - L; m <= t + R && used < ET_WAYS; m++) {
            const u32 y = c.ssa[m];
            if (y <= x) continue;
            u32 lcp;
            bool x_less;
            if (!fwd_compare<SB>(W, n32, x, y, 0u, (u32)SA_CMP_CAP, lcp, x_less)) continue;  // too long: not recorded
            bucket[used++] = ((u64)y << 32) | ((u64)lcp << 1) | (x_less ? 1ull : 0ull);
        }
    }
}

__device__ __forceinline__ bool etab_find(const u64 *__restrict__ etab, u32 bucket, u64 first, u32 y, u32 &lcp, bool &x_less) {
    const u64 *b = etab + (size_t)bucket * ET_WAYS;
#pragma unroll 1
    for (int e = 0; e < ET_WAYS; e++) {
        const u64 v = e ? b[e] : first;  // entry 0 was loaded ahead by the caller
        if (v == 0ull) return false;
        if ((u32)(v >> 32) == y) {
            lcp = (u32)(v >> 1) & 0x7fffffffu;
            x_less = (v & 1ull) != 0ull;
            return true;
        }
    }
    return false;
}

// lcp and order of the suffixes at positions x < y of one group: the pair's own record if x is a sampled position; else one
// comparison step and, if that step found no difference, the record of the sampled pair further down the diagonal; else (no
// record) the comparison goes on.  false: still undecided SA_CMP_CAP symbols later.
template <int SB>
__device__ __forceinline__ bool resolve_pair(const u32 *__restrict__ W, u32 n32, const u64 *__restrict__ etab, u32 x, u32 y, u32 &lcp, bool &x_less) {
    typedef Sym<SB> S;
    const u32 xo = x & (S::STEP - 1u);
    const u32 cc = xo ? S::STEP - xo : 0u;  // the diagonal reaches its sampled position after cc matching symbols (a sampled x: at once)
    const u32 lenmin = n32 - y;
    const u32 bucket = (x + cc) >> (S::LOG_SPW + 2u);
    const u64 first = cc < lenmin ? etab[(size_t)bucket * ET_WAYS] : 0ull;  // in flight while the step below runs
    bool decided = fwd_compare<SB>(W, n32, x, y, 0u, S::STEP, lcp, x_less);  // one step
    if (!decided && cc < lenmin) {
        u32 l2;
        if (etab_find(etab, bucket, first, y + cc, l2, x_less)) {
            lcp = cc + l2;
            decided = true;
        }
    }
    if (!decided) decided = fwd_compare<SB>(W, n32, x, y, S::STEP, (u32)SA_CMP_CAP, lcp, x_less);
    return decided;
}

// Every member of a small group meets each earlier member once; the larger suffix of a pair gains one smaller mate (-> its place
// inside the group) and the pair's common prefix (-> its LCP entry = the largest over its smaller mates, cut at the first '$'/'N'
// when it is written); both live in shared memory because a group never leaves its warp.  The warp then places its groups: SA,
// inverse SA and LCP.  A pair that is still undecided SA_CMP_CAP symbols after its look-up postpones its group to stage 4.
template <typename KeyT, int SB>
__global__ void __launch_bounds__(PR_THREADS, RV_PR_MINBLOCKS)
sa_place_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ sa, i64 n, const u32 *__restrict__ W, Barriers bars,
                const u64 *__restrict__ etab, int *__restrict__ SA, int *__restrict__ rank, int *__restrict__ LCP, unsigned char *__restrict__ deferred,
                u32 *__restrict__ flag_large, int *__restrict__ chunk_start, u32 *__restrict__ needbits, int key_digits2) {
    grid_dep_wait();
    __shared__ u32 s_sa[PR_WARPS][PR_MAXT];
    __shared__ u32 s_lcp[PR_WARPS][PR_MAXT];
    __shared__ u32 s_cnt[PR_WARPS][PR_MAXT / 4];        // one byte per slot: smaller mates seen so far
    __shared__ unsigned char s_L[PR_WARPS][PR_MAXT];    // members to the left inside the group; 0xFF: not a small group
    __shared__ u32 s_head[PR_WARPS][PR_ROUNDS + 2];
    __shared__ u32 s_def[PR_WARPS][PR_ROUNDS];          // bit t0: the group starting at t0 is postponed to stage 4
    __shared__ unsigned short s_pairs[PR_WARPS][PL_CAP];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    PairChunk c;
    c.ssa = s_sa[w];
    u32 *ssa = c.ssa, *lcpv = s_lcp[w], *cnt = s_cnt[w], *sdef = s_def[w];
    unsigned char *sL = s_L[w];
    u64 keyclz;
    pair_chunk_setup(c, keys, sa, n, s_head[w], keyclz);
    const int nt = c.nt;
    const i64 s = c.s;
    if (lane == 0) chunk_start[(i64)blockIdx.x * PR_WARPS + w] = nt ? (int)s : -1;  // its LCP entry: sa_chunkhead_kernel
    if (nt == 0) return;  // warp-uniform
    const int rounds = c.rounds;
    for (int r = 0; r < rounds; r++) {
        const int t = r * 32 + (int)lane;
        int L = 0xFF, R;
        bool small = false;
        if (t < nt) small = pair_chunk_group(c, t, L, R);
        sL[t] = small ? (unsigned char)L : (unsigned char)0xFF;
        if (t < nt && !small && L == 0) *flag_large = 1u;
        lcpv[t] = 0u;
        if ((lane & 3u) == 0) cnt[t >> 2] = 0u;
        if (lane == 0) sdef[r] = 0u;
    }
    __syncwarp();

    // ---- pairs: every (member, earlier mate) of a small group becomes an item of a shared-memory list; the lanes take the
    //      items 32 at a time, so they stay busy whatever the group sizes are ----
    const u32 n32 = (u32)n;
    unsigned short *plist = s_pairs[w];
    u32 qn = 0, next = 0;  // items [next, qn) of the ring
    for (int r = 0; r <= rounds; r++) {
        if (r < rounds) {
            const int t = r * 32 + (int)lane;
            const int L = t < nt ? (int)sL[t] : 0xFF;
            const u32 cpairs = L == 0xFF ? 0u : (u32)L;
            const u32 inc = warp_incl_sum(cpairs);
            const u32 total = __shfl_sync(FULL, inc, 31);
            const u32 at = qn + inc - cpairs;
            // (groups of two to four members are the rule: their items are written without a loop)
            if (cpairs >= 1u) plist[at & (PL_CAP - 1)] = (unsigned short)(((u32)t << 4) | 1u);
            if (cpairs >= 2u) plist[(at + 1u) & (PL_CAP - 1)] = (unsigned short)(((u32)t << 4) | 2u);
            if (cpairs >= 3u) plist[(at + 2u) & (PL_CAP - 1)] = (unsigned short)(((u32)t << 4) | 3u);
            for (u32 d = 4; d <= cpairs; d++) plist[(at + d - 1u) & (PL_CAP - 1)] = (unsigned short)(((u32)t << 4) | d);
            qn += total;
            __syncwarp();
        }
        // drain whole warps' worth (everything after the last round); a round adds at most 32 * 15 items, PL_CAP holds that plus a rest
        while (qn - next >= 32u || (r == rounds && qn != next)) {
            const u32 idx = next + lane;
            bool have = (int)(qn - idx) > 0;
            int tx = 0, ty = 0, t0 = 0;
            if (have) {
                const u32 it = plist[idx & (PL_CAP - 1)];
                tx = (int)(it >> 4);
                ty = tx - (int)(it & 15u);
                t0 = tx - (int)sL[tx];
                have = !((sdef[t0 >> 5] >> (t0 & 31)) & 1u);  // else given up: stage 4 orders this group
            }
            u32 lcp = 0;
            bool p_less = false, ok = false;
            if (have) {
                const u32 p = ssa[tx], q = ssa[ty];
                const u32 x = p < q ? p : q, y = p < q ? q : p;  // positions: x < y
                bool x_less = false;
                ok = resolve_pair<SB>(W, n32, etab, x, y, lcp, x_less);
                p_less = (p == x) == x_less;
            }
            __syncwarp();  // the comparisons end at different times: the updates below are issued once for the whole warp
            if (have) {
                if (ok) {
                    const int big = p_less ? ty : tx;  // the larger suffix gains a smaller mate
                    atomicAdd(&cnt[big >> 2], 1u << (8 * (big & 3)));
                    atomicMax(&lcpv[big], lcp);
                } else {  // too long: let the doubling rounds order this group
                    atomicOr(&sdef[t0 >> 5], 1u << (t0 & 31));
                }
            }
            next += (qn - next) < 32u ? (qn - next) : 32u;
            __syncwarp();
        }
    }
    __syncwarp();

    // ---- place the warp's groups: slot = group start + number of smaller mates ----
    // The final order of the chunk's suffixes replaces ssa[] (shared memory is also L1 capacity here: no second
    // array); every lane keeps its <= PR_ROUNDS entries in registers across the overwrite.
    u32 my_suf[PR_ROUNDS];
    int my_f[PR_ROUNDS];
#pragma unroll
    for (int k = 0; k < PR_ROUNDS; k++) {
        const int t = k * 32 + (int)lane;
        my_f[k] = -1;
        my_suf[k] = 0;
        if (t >= nt) continue;
        unsigned L = sL[t];
        if (L == 0xFFu) {  // member of a group with more than SA_SMALL_G suffixes: stage 4
            SA[s + t] = -1;  // "not placed": what the neighbours' LCP passes see until stage 4 fills the slot
            continue;
        }
        const int t0 = t - (int)L;
        if ((sdef[t0 >> 5] >> (t0 & 31)) & 1u) {
            if (L == 0) {
                deferred[s + t] = 1;
                *flag_large = 1u;
            }
            SA[s + t] = -1;
            continue;
        }
        u32 r = (cnt[t >> 2] >> (8 * (t & 3))) & 0xffu;
        i64 slot = s + t0 + (i64)r;
        u32 suf = ssa[t];
        SA[slot] = (int)suf;
        rank[suf] = (int)slot;
        if (r > 0) LCP[slot] = (int)first_barrier(bars, suf, lcpv[t]);  // the reference's '$'/'N' cut (interface.c:107)
        my_f[k] = t0 + (int)r;
        my_suf[k] = suf;
    }
    __syncwarp();
    u32 *fin = ssa;
    for (int t = (int)lane; t < nt; t += 32) fin[t] = 0xFFFFFFFFu;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < PR_ROUNDS; k++)
        if (my_f[k] >= 0) fin[my_f[k]] = my_suf[k];
    __syncwarp();
    // ---- the first slot of every group: its left neighbour belongs to another group, one short direct comparison.
    //      (The first slot of the chunk has its neighbour in another warp: sa_chunkhead_kernel.) ----
    for (int f = (int)lane; f < nt; f += 32) {
        if (f == 0 || !((c.head[f >> 5] >> (f & 31)) & 1u)) continue;
        u32 a = fin[f], b = fin[f - 1];
        if (a == 0xFFFFFFFFu) continue;  // stage 4 places the slot and marks the suffix for the LCP pass that follows it
        if (b == 0xFFFFFFFFu) {          // the left neighbour is not known yet: this suffix's LCP entry comes from that pass too
            atomicOr(&needbits[a >> 5], 1u << (a & 31u));
            continue;
        }
        if (sizeof(KeyT) == 4 && key_digits2 > 0) {
            // Neighbouring groups differ inside their keys.  With 2-bit digits and every rare symbol a barrier symbol, the equal
            // leading digits of the two keys ARE the common prefix, provided no '$'/'N' (and not the end of the text) lies within
            // those symbols of either suffix -- the leading zeros of key XOR previous key, kept by this lane since the staging, and
            // two looks at the cache-resident bitmap level instead of a comparison on the text.
            const u32 lk = (((u32)(keyclz >> (6 * (f >> 5))) & 63u) - (u32)(32 - 2 * key_digits2)) >> 1;  // f = 32 r + lane: this lane staged the slot
            if (a + lk + 1u <= n32 && b + lk + 1u <= n32 && first_barrier(bars, a, lk + 1u) == lk + 1u && first_barrier(bars, b, lk + 1u) == lk + 1u) {
                LCP[s + f] = (int)lk;
                continue;
            }
        }
        LCP[s + f] = direct_lcp<SB>(W, n32, a, b, bars);
    }
}

// LCP entry of the first slot of every warp chunk of sa_pairs_kernel
__global__ void __launch_bounds__(256)
sa_chunkhead_kernel(const int *__restrict__ chunk_start, i64 nchunks, i64 n, const unsigned char *__restrict__ T, Barriers bars,
                    const int *__restrict__ SA, int *__restrict__ LCP, u32 *__restrict__ needbits) {
    grid_dep_wait();
    i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    int j = chunk_start[c];
    if (j < 0) return;
    if (j == 0) {
        LCP[0] = 0;
        return;
    }
    // this pass is enqueued before the host knows whether stage 4 is needed: slots it will fill hold -1 until then
    const int p = SA[j], q = SA[j - 1];
    if (p < 0) return;
    if (q < 0) {
        atomicOr(&needbits[(u32)p >> 5], 1u << ((u32)p & 31u));
        return;
    }
    LCP[j] = direct_lcp<8>((const u32 *)T, (u32)n, (u32)p, (u32)q, bars);
}

// ---- LCP entries of the suffixes stage 4 placed (and of their right neighbours in the suffix array) ----------------------------
// needbits: one bit per TEXT position whose LCP entry is still missing.  Kasai's amortisation needs text order: consecutive marked
// positions i, i+1 have lcp(i+1, Phi(i+1)) >= lcp(i, Phi(i)) - 1, so inside a long repeat every suffix costs one comparison step
// instead of a comparison of the whole remaining repeat.  One thread walks LS_WORDS words of the bitmap; unmarked stretches
// (entries the comparison stage already delivered) cost one load per 32 positions.  With every bit set this IS Kasai et al. in
// chunks of 32*LS_WORDS positions on the packed text; it replaces compute_lcp (interface.c:97-114) for repeat-heavy inputs.
static const int LS_WORDS = 2;
template <int SB>
__global__ void __launch_bounds__(128)
lcp_sparse_kernel(const u32 *__restrict__ needbits, i64 n, const u32 *__restrict__ W, Barriers bars, const int *__restrict__ SA, const int *__restrict__ ISA, int *__restrict__ LCP) {
    grid_dep_wait();
    typedef Sym<SB> S;
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 w0 = t * LS_WORDS, nwords = (n + 31) / 32;
    if (w0 >= nwords) return;
    const u32 n32 = (u32)n;
    u32 h = 0;
    i64 prev = -2;
    for (i64 w = w0; w < w0 + LS_WORDS && w < nwords; w++) {
        for (u32 bits = needbits[w]; bits; bits &= bits - 1u) {
            const i64 i = w * 32 + (__ffs((int)bits) - 1);
            if (i >= n) break;
            const int r = ISA[i];
            if (r == 0) {
                LCP[0] = 0;
                prev = -2;
                continue;
            }
            const u32 j = (u32)SA[r - 1];
            h = (i == prev + 1 && h > 0u) ? h - 1u : 0u;
            prev = i;
            // extend the match from offset h, 4 words per suffix and step
            const u32 p = (u32)i + h, q = j + h;
            const u32 lenmin = n32 - (p > q ? p : q);
            const u32 *pa = W + (p >> S::LOG_SPW), *pb = W + (q >> S::LOG_SPW);
            const unsigned sha = (p & (S::SPW - 1u)) * SB, shb = (q & (S::SPW - 1u)) * SB;
            u32 lo_a = *pa, lo_b = *pb, g = 0, match = lenmin;
            while (g < lenmin) {
                u32 a1 = pa[1], a2 = pa[2], a3 = pa[3], a4 = pa[4];
                u32 b1 = pb[1], b2 = pb[2], b3 = pb[3], b4 = pb[4];
                u32 d0 = __funnelshift_r(lo_a, a1, sha) ^ __funnelshift_r(lo_b, b1, shb);
                u32 d1 = __funnelshift_r(a1, a2, sha) ^ __funnelshift_r(b1, b2, shb);
                u32 d2 = __funnelshift_r(a2, a3, sha) ^ __funnelshift_r(b2, b3, shb);
                u32 d3 = __funnelshift_r(a3, a4, sha) ^ __funnelshift_r(b3, b4, shb);
                if (d0 | d1 | d2 | d3) {
                    u32 wsel = d0 ? 0u : (d1 ? 1u : (d2 ? 2u : 3u));
                    u32 dd = d0 ? d0 : (d1 ? d1 : (d2 ? d2 : d3));
                    u32 at = g + wsel * S::SPW + ((u32)(__ffs((int)dd) - 1) >> S::LOG_SB);
                    match = at < lenmin ? at : lenmin;
                    break;
                }
                g += S::STEP;
                pa += 4;
                pb += 4;
                lo_a = a4;
                lo_b = b4;
            }
            h += match;
            LCP[r] = (int)first_barrier(bars, (u32)i, h);  // the reference's '$'/'N' cut (interface.c:107) on the way out only
        }
    }
}

// ---- stage 4: prefix doubling ---------------------------------------------------------------
// key[e] = rank[sa[e]] : (rank[sa[e]+h]+1, or 0 when the suffix ends first)
__global__ void __launch_bounds__(256) sa_gather_kernel(const u32 *__restrict__ sa, const u32 *__restrict__ grp, const int *__restrict__ rank,
                                                       i64 A, i64 n, i64 h, u64 *__restrict__ keys) {
    grid_dep_wait();
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A) return;
    i64 p = (i64)sa[e] + h;
    u32 k2 = p < n ? (u32)rank[p] + 1u : 0u;
    keys[e] = ((u64)grp[e] << 32) | (u64)k2;
}

// Exact refinement round of stage 4: second key = the kx symbols at sa[e]+off as a base-(sigma+1) number over the EXACT codes
// (0 = past the end).  The k-mer keys of stage 2 merge rare symbols into a neighbouring digit, so a group of equal keys shares
// only the CLASS string of its first k symbols; before the doubling rounds may treat ranks as "order by the first h symbols"
// the groups that reach stage 4 are refined by the true symbols of those positions.
__global__ void __launch_bounds__(256) sa_exact_gather_kernel(const u32 *__restrict__ sa, const u32 *__restrict__ grp, const unsigned char *__restrict__ T,
                                                             CodeTable tab, i64 A, i64 n, i64 off, int kx, u32 xbase, u64 *__restrict__ keys) {
    grid_dep_wait();
    __shared__ unsigned short s_code[256];
    s_code[threadIdx.x] = tab.code[threadIdx.x];
    __syncthreads();
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A) return;
    i64 p = (i64)sa[e] + off;
    u32 k2 = 0;
    for (int t = 0; t < kx; t++, p++) k2 = k2 * xbase + (p < n ? (u32)s_code[T[p]] : 0u);
    keys[e] = ((u64)grp[e] << 32) | (u64)k2;
}

// head: first entry of its equal-key run.  active: the run still needs doubling rounds, i.e. it has more
// than G members (G = 1 in the doubling rounds: any tie) or the comparison stage deferred it.
template <typename KeyT>
__device__ __forceinline__ void ap_flags(const KeyT *__restrict__ keys, i64 A, i64 e, int G, const unsigned char *__restrict__ deferred,
                                         bool &head, bool &active) {
    int L, R;
    run_lengths(keys, A, e, G, L, R);
    head = L == 0;
    active = (L + R + 1 > G) || (deferred && L + R > 0 && deferred[e - L]);
}

// per tile: (largest slot+1 of a group head, number of entries that stay active)
template <typename KeyT>
__global__ void __launch_bounds__(AP_THREADS) sa_reduce_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ pos, i64 A, int G,
                                                               const unsigned char *__restrict__ deferred,
                                                               u32 *__restrict__ tile_max, u32 *__restrict__ tile_cnt) {
    grid_dep_wait();
    __shared__ u32 s1[33], s2[33];
    i64 base = (i64)blockIdx.x * AP_TILE + (i64)threadIdx.x * AP_IPT;
    u32 mx = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < AP_IPT; k++) {
        i64 e = base + k;
        if (e < A) {
            bool head, active;
            ap_flags(keys, A, e, G, deferred, head, active);
            if (head && (active || G == 1)) mx = (pos ? pos[e] : (u32)e) + 1u;
            cnt += active ? 1u : 0u;
        }
    }
    u32 tmx, tcnt;
    block_incl_max<AP_THREADS>(mx, s1, &tmx);
    block_incl_sum<AP_THREADS>(cnt, s2, &tcnt);
    if (threadIdx.x == 0) {
        tile_max[blockIdx.x] = tmx;
        tile_cnt[blockIdx.x] = tcnt;
    }
}

// single block: exclusive max-scan / sum-scan over the tile aggregates; total -> *out_total
__global__ void __launch_bounds__(1024) sa_tilescan_kernel(u32 *__restrict__ tile_max, u32 *__restrict__ tile_cnt, i64 tiles, u32 *__restrict__ out_total) {
    grid_dep_wait();
    __shared__ u32 s1[33], s2[33];
    __shared__ u32 s_im[1024];
    u32 carry_max = 0, carry_sum = 0;
    for (i64 b0 = 0; b0 < tiles; b0 += 1024) {
        i64 t = b0 + threadIdx.x;
        u32 m = t < tiles ? tile_max[t] : 0u;
        u32 c = t < tiles ? tile_cnt[t] : 0u;
        u32 tm, tc;
        u32 im = block_incl_max<1024>(m, s1, &tm);
        u32 ic = block_incl_sum<1024>(c, s2, &tc);
        if (t < tiles) tile_cnt[t] = carry_sum + ic - c;
        // exclusive max needs the inclusive max of the previous thread: stage through shared memory
        s_im[threadIdx.x] = im;
        __syncthreads();
        if (t < tiles) {
            u32 ex = threadIdx.x > 0 ? s_im[threadIdx.x - 1] : 0u;
            tile_max[t] = ex > carry_max ? ex : carry_max;
        }
        carry_max = tm > carry_max ? tm : carry_max;
        carry_sum += tc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_total = carry_sum;
}

// per tile: finish the two scans with the tile carries and apply the round:
//   SA[slot] = suffix, rank[suffix] = slot of its group head, and append the
//   entries of groups that still have >= 2 members to the next active list.
template <typename KeyT>
__global__ void __launch_bounds__(AP_THREADS)
sa_apply_kernel(const KeyT *__restrict__ keys, const u32 *__restrict__ sa, const u32 *__restrict__ pos, i64 A, int G,
                const unsigned char *__restrict__ deferred, const u32 *__restrict__ tile_max, const u32 *__restrict__ tile_cnt,
                int *__restrict__ SA, int *__restrict__ rank, u32 *__restrict__ sa2, u32 *__restrict__ pos2, u32 *__restrict__ grp2,
                u32 *__restrict__ needbits) {
    grid_dep_wait();
    __shared__ u32 s1[33], s2[33];
    __shared__ u32 s_im[AP_THREADS];
    i64 base = (i64)blockIdx.x * AP_TILE + (i64)threadIdx.x * AP_IPT;
    u32 hp[AP_IPT];
    bool act[AP_IPT];
    u32 mx = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < AP_IPT; k++) {
        i64 e = base + k;
        hp[k] = 0;
        act[k] = false;
        if (e < A) {
            bool head, active;
            ap_flags(keys, A, e, G, deferred, head, active);
            if (head && (active || G == 1)) hp[k] = (pos ? pos[e] : (u32)e) + 1u;
            act[k] = active;
            cnt += active ? 1u : 0u;
        }
        mx = hp[k] > mx ? hp[k] : mx;
        hp[k] = mx;  // thread-local inclusive max
    }
    u32 tm, tc;
    u32 imax = block_incl_max<AP_THREADS>(mx, s1, &tm);
    u32 isum = block_incl_sum<AP_THREADS>(cnt, s2, &tc);
    // exclusive max over the threads before this one
    s_im[threadIdx.x] = imax;
    __syncthreads();
    u32 pre_max = threadIdx.x > 0 ? s_im[threadIdx.x - 1] : 0u;
    u32 cm = tile_max[blockIdx.x];
    pre_max = pre_max > cm ? pre_max : cm;
    u32 dst = tile_cnt[blockIdx.x] + isum - cnt;
#pragma unroll
    for (int k = 0; k < AP_IPT; k++) {
        i64 e = base + k;
        if (e < A && (act[k] || G == 1)) {  // G > 1: the other entries were placed by the comparison stage
            u32 g = (hp[k] > pre_max ? hp[k] : pre_max) - 1u;  // slot of the group head
            u32 slot = pos ? pos[e] : (u32)e;
            u32 s = sa[e];
            SA[slot] = (int)s;
            rank[s] = (int)g;
            if (act[k]) {
                sa2[dst] = s;
                pos2[dst] = slot;
                grp2[dst] = g;
                dst++;
                if (needbits) atomicOr(&needbits[s >> 5], 1u << (s & 31u));  // round 0: every suffix stage 4 orders needs its LCP entry
            }
        }
    }
}

static inline int bits_for(u64 v) {  // number of bits needed to hold v
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b;
}

size_t sa_workspace_bytes(i64 n) {
    size_t a = (size_t)((n + 63) / 64 * 64);
    i64 tiles = (n + AP_TILE - 1) / AP_TILE;
    // keys x2 (u64), vals x2, pos x2, grp x2 (u32), deferred (u8), tile aggregates, radix scratch, small stuff
    return a * (8 + 8 + 4 + 4 + 4 + 4 + 4 + 4 + 2) + a / 8 + a / 256 + a / 64 + a / 2 + a * 4 + 65536 + (size_t)tiles * 8 + radix_scratch_bytes(n) + 16 * 256 * 16 + (1 << 16);
}

struct SaBuffers {
    u64 *k0, *k1;
    u32 *v0, *v1, *posA, *posB, *grpA, *grpB, *tile_max, *tile_cnt, *small;
    unsigned char *deferred;
    u32 *needbits;     // one bit per text position: LCP entry still missing after the comparison stage (lcp_sparse_kernel)
    int *chunk_start;  // first slot of every warp chunk of sa_pairs_kernel
    u32 *packed;       // 4-bit packed text, n/8 + 16 words
    u32 *bar, *bar1, *bar2;  // three-level barrier bitmap: n/32 + 300, n/1024 + 16, n/32768 + 8 words
    u64 *etab;         // sampled-pair table of the comparison stage: (n/16 + 2) buckets of ET_WAYS entries
    void *rscratch;
};

// stages 2-3 for one key width; on return *keys_out / *sa_out hold the sorted keys and suffixes
template <typename KeyT>
static int sort_and_compare(Stream &st, const SaBuffers &B, const unsigned char *dT, i64 n, const CodeTable &tab, int sigma, u32 base, int k,
                            int key_bits, int *dSA, int *dISA, int *dLCP, KeyT **keys_out, u32 **sa_out, u32 **sa_free, bool *packed_out,
                            PhaseTimes *pt, cudaStream_t aux) {
    KeyT *k0 = (KeyT *)B.k0, *k1 = (KeyT *)B.k1;
    // the (key, suffix) pairs are virtual: the histogram kernel and the first digit pass roll the k-mer keys
    // straight from the text (TextKeySrc), so no key array is written before the first scatter
    TextKeySrc src;
    src.T = dT;
    src.n = n;
    src.base = base;
    src.k = k;
    src.top = 1;
    for (int t = 1; t < k; t++) src.top *= (u64)base;
    src.pw[0] = 1;
    for (int t = 1; t < 64; t++) src.pw[t] = t < k ? src.pw[t - 1] * (u64)base : 0;
    memcpy(src.code, st.alpha.kcls, sizeof src.code);  // key digits = classes (rare symbols merged), not the exact codes
    // ---- beside the digit passes (stream `aux`, joined below): the tables of the comparison stage are cleared and the text pass
    //      (barrier bitmaps, 4-bit packed text) runs; neither depends on the sort ----
    RV_CUDA(cudaMemsetAsync(B.deferred, 0, (size_t)n, aux));
    RV_CUDA(cudaMemsetAsync(B.needbits, 0, (size_t)(n / 32 + 2) * 4, aux));
    const i64 pr_per_block = (i64)PR_WARPS * PR_CHUNK;
    // text for the comparisons: 4-bit packed codes when the alphabet allows (sigma <= 15), else the raw bytes
    const bool packed = sigma <= 15 && !getenv("RV_SA_NO_PACK");  // env: test hook for the byte path
    // (Equal keys do not promise equal first k symbols -- a rare symbol ends the key -- so the comparisons start at the suffix itself.)
    const unsigned pblocks = (unsigned)((n + pr_per_block - 1) / pr_per_block);
    const unsigned prep_blocks = (unsigned)((n + TP_BLOCK - 1) / TP_BLOCK + 1);  // one block past the end: zero padding of the packed text
    const Barriers bars = {B.bar, B.bar1, B.bar2};
    // keys whose equal leading digits can stand for the common prefix of neighbouring groups (sa_place_kernel): 2-bit digits,
    // a 32-bit key, and no rare symbol that is not a '$'/'N' barrier (the bitmap then rules rare symbols out)
    int key_digits2 = 0;
    if (sizeof(KeyT) == 4 && base == 4 && !getenv("RV_SA_NO_KEYLCP")) {
        bool ok = true;
        for (int c = 0; c < 256; c++)
            if ((st.alpha.kcls[c] & 0x100) && st.alpha.code[c] && c != '$' && c != 'N') ok = false;
        if (ok) key_digits2 = k;
    }
    RV_CUDA(cudaMemsetAsync(B.bar2, 0, (size_t)(n / 32768 + 8) * 4, aux));
    const int step_syms = packed ? 32 : 16;
    RV_CUDA(cudaMemsetAsync(B.etab, 0, (size_t)(n / step_syms + 2) * ET_WAYS * 8, aux));
    RV_TRY(prof_begin(st, RV_PROF_TEXT, aux));
    if (packed) {
        RV_LAUNCH((sa_textprep_kernel<true>), prep_blocks, TP_THREADS, 0, aux, dT, n, tab, B.bar, B.bar1, B.bar2, B.packed);
    } else {
        RV_LAUNCH((sa_textprep_kernel<false>), prep_blocks, TP_THREADS, 0, aux, dT, n, tab, B.bar, B.bar1, B.bar2, B.packed);
    }
    RV_TRY(prof_end(st, RV_PROF_TEXT, 1, (long long)n, aux));
    bool in0;
    RV_TRY(radix_sort_pairs<KeyT>(st, k0, k1, B.v0, B.v1, n, make_plan(0, key_bits), B.rscratch, &in0, &src));
    if (pt) pt->sa_sorted_items += n;
    KeyT *keys = in0 ? k0 : k1;
    u32 *sa = in0 ? B.v0 : B.v1;
    if (aux != st.s) RV_TRY(side_join(st));
    const u32 *W = packed ? (const u32 *)B.packed : (const u32 *)dT;
    RV_TRY(prof_begin(st, RV_PROF_LEAD));
    if (packed) {
        RV_LAUNCH_PDL((sa_lead_kernel<KeyT, 4>), pblocks, PR_THREADS, 0, st.s, keys, sa, n, W, B.etab);
    } else {
        RV_LAUNCH_PDL((sa_lead_kernel<KeyT, 8>), pblocks, PR_THREADS, 0, st.s, keys, sa, n, W, B.etab);
    }
    RV_TRY(prof_end(st, RV_PROF_LEAD, 1, (long long)n * (long long)(sizeof(KeyT) + 4)));
    RV_TRY(prof_begin(st, RV_PROF_PAIRS));
    if (packed) {
        RV_LAUNCH_PDL((sa_place_kernel<KeyT, 4>), pblocks, PR_THREADS, 0, st.s, keys, sa, n, W, bars, (const u64 *)B.etab, dSA, dISA, dLCP, B.deferred,
                  B.small + 257, B.chunk_start, B.needbits, key_digits2);
    } else {
        RV_LAUNCH_PDL((sa_place_kernel<KeyT, 8>), pblocks, PR_THREADS, 0, st.s, keys, sa, n, W, bars, (const u64 *)B.etab, dSA, dISA, dLCP, B.deferred,
                  B.small + 257, B.chunk_start, B.needbits, key_digits2);
    }
    RV_TRY(prof_end(st, RV_PROF_PAIRS, 1, (long long)n * (long long)(sizeof(KeyT) + 4 + 12)));
    st.launches += 3;
    {   // the chunk-head LCP pass is enqueued before anyone knows whether stage 4 is needed: slots that stage 4 still has to
        // fill read -1 and the entry is left to lcp_sparse_kernel
        const i64 nchunks = ((n + pr_per_block - 1) / pr_per_block) * PR_WARPS;
        RV_LAUNCH_PDL(sa_chunkhead_kernel, (unsigned)((nchunks + 255) / 256), 256, 0, st.s, B.chunk_start, nchunks, n, dT, bars, dSA, dLCP, B.needbits);
        st.launches++;
    }
    RV_KCHECK();
    *keys_out = keys;
    *packed_out = packed;
    *sa_out = sa;
    *sa_free = in0 ? B.v1 : B.v0;
    return RV_OK;
}

// Stage 4.  Round 0 works on the stage-2 keys (KeyT): it places nothing new but collects the entries of the groups the comparison
// stage left alone into the active list.  Then exact refinement rounds over the first k symbols (sa_exact_gather_kernel), then
// the doubling rounds on (rank : rank) u64 keys.
template <typename KeyT>
static int doubling(Stream &st, const SaBuffers &B, const unsigned char *dT, const CodeTable &tab, int sigma, i64 n, int k, const KeyT *keys0, u32 *sa,
                    u32 *sa_alt, int *dSA, int *dISA, i64 *first_active, PhaseTimes *pt) {
    u32 *pos = nullptr, *grp = nullptr;  // current active list is (sa, pos, grp)
    u32 *pos_next = B.posA, *grp_next = B.grpA;
    u64 *keys = B.k0, *keys_alt = B.k1;  // free once round 0 has consumed keys0 (which may alias one of them)
    i64 A = n;
    const int nbits = bits_for((u64)n);  // ranks < n, second key <= n
    const u32 xbase = (u32)sigma + 1;
    int kx = 0, xbits = 0;               // symbols (and bits) of one exact second key
    {
        u64 v = 1;
        while (v * xbase <= 4294967296ull && kx < 32) { v *= xbase; kx++; }
        xbits = bits_for(v - 1);
    }
    i64 covered = 0;  // leading symbols of every active group known to be equal (exactly)
    i64 h = 0;
    for (int round = 0;; round++) {
        const i64 tiles = (A + AP_TILE - 1) / AP_TILE;
        if (round == 0) {
            RV_LAUNCH_PDL((sa_reduce_kernel<KeyT>), (unsigned)tiles, AP_THREADS, 0, st.s, keys0, pos, A, SA_SMALL_G, B.deferred, B.tile_max, B.tile_cnt);
            RV_LAUNCH_PDL(sa_tilescan_kernel, 1, 1024, 0, st.s, B.tile_max, B.tile_cnt, tiles, B.small + 256);
            RV_LAUNCH_PDL((sa_apply_kernel<KeyT>), (unsigned)tiles, AP_THREADS, 0, st.s, keys0, sa, pos, A, SA_SMALL_G, B.deferred, B.tile_max,
                      B.tile_cnt, dSA, dISA, sa_alt, pos_next, grp_next, B.needbits);
        } else {
            RV_LAUNCH_PDL((sa_reduce_kernel<u64>), (unsigned)tiles, AP_THREADS, 0, st.s, keys, pos, A, 1, (const unsigned char *)nullptr, B.tile_max,
                      B.tile_cnt);
            RV_LAUNCH_PDL(sa_tilescan_kernel, 1, 1024, 0, st.s, B.tile_max, B.tile_cnt, tiles, B.small + 256);
            RV_LAUNCH_PDL((sa_apply_kernel<u64>), (unsigned)tiles, AP_THREADS, 0, st.s, keys, sa, pos, A, 1, (const unsigned char *)nullptr, B.tile_max,
                      B.tile_cnt, dSA, dISA, sa_alt, pos_next, grp_next, (u32 *)nullptr);
        }
        st.launches += 3;
        RV_CUDA(cudaMemcpyAsync(st.pinned + 310, B.small + 256, 4, cudaMemcpyDeviceToHost, st.s));
        RV_CUDA(cudaStreamSynchronize(st.s));
        const u32 nactive = st.pinned[310];
        if (pt) pt->sa_rounds = round + 1;
        if (round == 0) *first_active = nactive;
        if (nactive == 0) break;
        if (covered >= k && h >= n) {
            set_error("sa_build: internal error, %u suffixes still tied at h=%lld >= n", nactive, (long long)h);
            return RV_ERR_STATE;
        }
        // the next active list lives in (sa_alt, pos_next, grp_next)
        A = nactive;
        { u32 *t = sa; sa = sa_alt; sa_alt = t; }
        pos = pos_next;
        grp = grp_next;
        pos_next = (pos == B.posA) ? B.posB : B.posA;
        grp_next = (grp == B.grpA) ? B.grpB : B.grpA;
        RadixPlan plan;
        if (covered < k) {  // exact refinement of the symbols [covered, covered + kx)
            RV_LAUNCH_PDL(sa_exact_gather_kernel, (unsigned)((A + 255) / 256), 256, 0, st.s, sa, grp, dT, tab, A, n, covered, kx, xbase, keys);
            plan = make_plan(0, xbits, 32, 32 + nbits);
            covered += kx;
            h = covered;
        } else {
            RV_LAUNCH_PDL(sa_gather_kernel, (unsigned)((A + 255) / 256), 256, 0, st.s, sa, grp, dISA, A, n, h, keys);
            plan = make_plan(0, nbits, 32, 32 + nbits);
            h *= 2;
        }
        st.launches++;
        bool r0;
        RV_TRY(radix_sort_pairs<u64>(st, keys, keys_alt, sa, sa_alt, A, plan, B.rscratch, &r0));
        if (pt) pt->sa_sorted_items += A;
        if (!r0) {
            { u64 *t = keys; keys = keys_alt; keys_alt = t; }
            { u32 *t = sa; sa = sa_alt; sa_alt = t; }
        }
    }
    RV_KCHECK();
    return RV_OK;
}

int sa_build(Stream &st, Arena &ws, const unsigned char *dT, i64 n, int *dSA, int *dISA, int *dLCP, bool *lcp_done, PhaseTimes *pt,
             bool fresh_alphabet) {
    *lcp_done = false;
    if (n <= 0) return RV_OK;
    if (n >= ((i64)1 << 30)) {
        set_error("sa_build: n=%lld not supported yet (limit 2^30-1)", (long long)n);
        return RV_ERR_UNSUPPORTED;
    }
    const size_t ws_mark = ws.off;
    SaBuffers B;
    B.k0 = ws.take<u64>(n);
    B.k1 = ws.take<u64>(n);
    B.v0 = ws.take<u32>(n);
    B.v1 = ws.take<u32>(n);
    B.posA = ws.take<u32>(n);
    B.posB = ws.take<u32>(n);
    B.grpA = ws.take<u32>(n);
    B.grpB = ws.take<u32>(n);
    const i64 tiles_n = (n + AP_TILE - 1) / AP_TILE;
    B.tile_max = ws.take<u32>(tiles_n);
    B.tile_cnt = ws.take<u32>(tiles_n);
    B.rscratch = ws.take<unsigned char>(radix_scratch_bytes(n));
    B.small = ws.take<u32>(512);  // [0..255] byte histogram, [256] active count, [257] "stage 4 needed"
    B.deferred = ws.take<unsigned char>(n);
    B.needbits = ws.take<u32>(n / 32 + 2);
    B.chunk_start = ws.take<int>(n / PR_CHUNK + 2 * PR_WARPS + 8);
    B.packed = ws.take<u32>(n / 8 + 1100);  // sa_textprep_kernel writes whole 4096-symbol blocks, one block past the end
    B.bar = ws.take<u32>(n / 32 + 300);
    B.bar1 = ws.take<u32>(n / 1024 + 16);
    B.bar2 = ws.take<u32>(n / 32768 + 8);
    B.etab = ws.take<u64>((size_t)(n / 16 + 2) * ET_WAYS);
    if (!B.deferred || !B.needbits || !B.chunk_start || !B.packed || !B.bar || !B.bar1 || !B.bar2 || !B.etab || !B.k0 || !B.k1 || !B.v0 || !B.v1 || !B.posA || !B.posB || !B.grpA || !B.grpB || !B.tile_max || !B.tile_cnt ||
        !B.rscratch || !B.small) {
        set_error("sa_build: workspace too small");
        return RV_ERR_NOMEM;
    }

    // 1. alphabet.  The histogram always runs, but when the handle has built a text before, the build starts at once
    //    with that text's code table and the histogram is only checked at the single synchronisation point below
    //    (a symbol the table lacks -> rebuild with the fresh table).  The first build on a handle waits for it.
    //    Work that does not depend on the radix sort -- this histogram, the text pass and the clearing of the comparison
    //    stage's tables -- runs on the handle's side stream, beside the digit passes (RV_SA_NO_SIDE=1: everything on one stream).
    u32 *hist = st.pinned;  // [0..255] counts, [256] stage-4 flag
    const bool speculative = st.alpha.valid && !fresh_alphabet;
    const bool use_side = !getenv("RV_SA_NO_SIDE");
    cudaStream_t aux = st.s;
    if (use_side && speculative) {
        RV_TRY(side_fork(st));
        aux = st.side ? st.side : st.s;
    }
    RV_CUDA(cudaMemsetAsync(B.small, 0, 512 * 4, aux));
    {
        i64 blocks = (n + 256 * 64 - 1) / (256 * 64);
        if (blocks > 148 * 8) blocks = 148 * 8;
        RV_LAUNCH(sa_bytehist_kernel, (unsigned)blocks, 256, 0, aux, dT, n, B.small);
        st.launches++;
    }
    RV_CUDA(cudaMemcpyAsync(hist, B.small, 256 * 4, cudaMemcpyDeviceToHost, aux));
    CodeTable tab;
    if (!speculative) {
        RV_CUDA(cudaStreamSynchronize(st.s));
        AlphaCache &al = st.alpha;
        memset(al.code, 0, sizeof al.code);
        memset(al.kcls, 0, sizeof al.kcls);
        al.sigma = al.sigma_eff = 0;
        // Exact codes: every symbol present, in byte order (the packed text of the comparisons).
        // Key classes: only the symbols that carry the entropy (>= 1/32 of the text: ACGT, not '$' / N / a stray IUPAC
        // letter) get a digit of their own; every other symbol shares the digit of the nearest frequent symbol below
        // it (digit 0 if there is none), and so does "past the end of the text".  The map is monotone in the byte
        // value, so equal-or-smaller suffixes get equal-or-smaller keys and the comparison stage settles the ties --
        // for DNA the key is the 2-bit packed k-mer: 16 symbols per 32-bit key whatever else occurs in the text.
        bool freq[256];
        for (int c = 0; c < 256; c++) {
            freq[c] = hist[c] && (u64)hist[c] * 32 >= (u64)n;
            if (hist[c]) al.code[c] = (unsigned short)(++al.sigma);
            if (freq[c]) al.sigma_eff++;
        }
        al.kbase = al.sigma_eff < 2 ? 2 : al.sigma_eff;
        // key digit of every byte value (TextKeySrc in rv_radix.cuh): frequent symbols 0..f-1 in byte order; any other byte ends the
        // key with the digit of the next frequent symbol above it + zero fill, or the largest digit + max fill if there is none
        int above = -1;  // digit of the nearest frequent symbol above c
        {
            int cls = al.sigma_eff;
            for (int c = 255; c >= 0; c--) {
                if (freq[c]) {
                    above = --cls;
                    al.kcls[c] = (unsigned short)cls;
                } else {
                    al.kcls[c] = above >= 0 ? (unsigned short)(above | 0x100) : (unsigned short)((al.kbase - 1) | 0x300);
                }
            }
        }
        if (al.sigma_eff < 2) al.sigma_eff = 2;
        al.valid = true;
        if (use_side) {
            RV_TRY(side_fork(st));
            aux = st.side ? st.side : st.s;
        }
    }
    memcpy(tab.code, st.alpha.code, sizeof tab.code);
    const int sigma = st.alpha.sigma, sigma_eff = st.alpha.sigma_eff;
    const u32 base = (u32)st.alpha.kbase;  // digits 0..base-1
    // shortest k with sigma_eff^k >= 4n: random k-mers are then mostly unique
    int k_need = 1;
    {
        double want = 4.0 * (double)n, have = sigma_eff;
        while (have < want && k_need < 64) { have *= sigma_eff; k_need++; }
    }
    auto capacity = [&](int bits) {  // largest k with base^k <= 2^bits
        int kk = 0;
        double v = 1.0, lim = bits == 32 ? 4294967296.0 : 18446744073709551616.0;
        while (v * base <= lim && kk < 64) { v *= base; kk++; }
        return kk;
    };
    auto bits_of_k = [&](int kk) {  // bits of the largest k-symbol key
        u64 m = 1;
        for (int t = 0; t < kk; t++) m *= base;  // base^64 may wrap to 0 for base 2: 64 bits
        return m == 0 ? 64 : bits_for(m - 1);
    };
    auto passes_of_k = [&](int kk) { return (bits_of_k(kk) + 7) / 8; };
    const int k32 = capacity(32), k64 = capacity(64);
    // 32-bit keys whenever k_need - 1 symbols fit (a key one symbol short of k_need has 'base' times fewer distinct
    // values: the comparison stage absorbs the slightly larger groups).  Within the 32-bit range k is chosen by digit
    // passes: one symbol less if that saves a whole 8-bit pass, else as many symbols as the passes hold anyway.
    bool use32 = k32 >= 1 && k32 + 1 >= k_need;
    if (const char *force = getenv("RV_SA_KEY_BITS")) {  // test hook: exercise both key widths on small inputs
        if (force[0] == '6') use32 = false;
        if (force[0] == '3') use32 = true;
    }
    int k;
    if (use32) {
        int kn = k_need < k32 ? k_need : k32;
        int target = passes_of_k(kn);
        if (kn > 1 && passes_of_k(kn - 1) < target) target = passes_of_k(kn - 1);
        k = kn > 1 ? kn - 1 : 1;
        while (k + 1 <= k32 && passes_of_k(k + 1) <= target) k++;
    } else {
        k = k_need < k64 ? k_need : k64;
    }
    if (k < 1) k = 1;
    const int key_bits = bits_of_k(k);

    u32 *sa = nullptr, *sa_free = nullptr;
    i64 first_active = 0;
    bool packed = false;
    u32 *keys32 = nullptr;
    u64 *keys64 = nullptr;
    if (use32) {
        RV_TRY(sort_and_compare<u32>(st, B, dT, n, tab, sigma, base, k, key_bits, dSA, dISA, dLCP, &keys32, &sa, &sa_free, &packed, pt, aux));
    } else {
        RV_TRY(sort_and_compare<u64>(st, B, dT, n, tab, sigma, base, k, key_bits, dSA, dISA, dLCP, &keys64, &sa, &sa_free, &packed, pt, aux));
    }
    // ---- the one synchronisation point of a build: is stage 4 needed, and (speculative start) was the alphabet right ----
    RV_CUDA(cudaMemcpyAsync(hist + 256, B.small + 257, 4, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));
    if (speculative) {
        bool ok = true;
        for (int c = 0; c < 256; c++)
            if (hist[c] && !tab.code[c]) ok = false;  // a symbol the cached table has no code for
        if (!ok) {
            ws.off = ws_mark;  // give the workspace back and start over with this text's own alphabet
            return sa_build(st, ws, dT, n, dSA, dISA, dLCP, lcp_done, pt, true);
        }
    }
    const bool large = hist[256] != 0;
    if (large) {
        if (use32) {
            // round 0 reads the u32 keys that live in the first half of k0 or k1; later rounds reuse both
            // buffers as u64 keys, so move the u32 keys out of the way (posB is free until round 2)
            RV_CUDA(cudaMemcpyAsync(B.posB, keys32, (size_t)n * 4, cudaMemcpyDeviceToDevice, st.s));
            RV_TRY(doubling<u32>(st, B, dT, tab, sigma, n, k, B.posB, sa, sa_free, dSA, dISA, &first_active, pt));
        } else {
            RV_TRY(doubling<u64>(st, B, dT, tab, sigma, n, k, keys64, sa, sa_free, dSA, dISA, &first_active, pt));
        }
        // LCP entries of everything stage 4 placed, and of the slots right after such a suffix: Kasai's walk over the marked
        // text positions (a handful for a stray long match, all of them for a text that is one big repeat)
        const i64 lthreads = ((n + 31) / 32 + LS_WORDS - 1) / LS_WORDS;
        const Barriers bars = {B.bar, B.bar1, B.bar2};
        RV_TRY(prof_begin(st, RV_PROF_LCP));
        if (packed) {
            RV_LAUNCH_PDL((lcp_sparse_kernel<4>), (unsigned)((lthreads + 127) / 128), 128, 0, st.s, B.needbits, n, (const u32 *)B.packed, bars, dSA, dISA, dLCP);
        } else {
            RV_LAUNCH_PDL((lcp_sparse_kernel<8>), (unsigned)((lthreads + 127) / 128), 128, 0, st.s, B.needbits, n, (const u32 *)dT, bars, dSA, dISA, dLCP);
        }
        RV_TRY(prof_end(st, RV_PROF_LCP, 1, (long long)first_active * 13 + n / 8));
        st.launches++;
        RV_KCHECK();
    }
    *lcp_done = true;
    return RV_OK;
}

// compute_lcp (interface.c:97-114) for a suffix array that did not come from sa_build (cache files): every position marked,
// byte text.  The two-level barrier bitmap is built on the way.
int lcp_build(Stream &st, Arena &ws, const unsigned char *dT, i64 n, const int *dSA, const int *dISA, int *dLCP) {
    if (n <= 0) return RV_OK;
    const size_t ws_mark = ws.off;
    u32 *bar = ws.take<u32>(n / 32 + 300), *bar1 = ws.take<u32>(n / 1024 + 16), *bar2 = ws.take<u32>(n / 32768 + 8), *needbits = ws.take<u32>(n / 32 + 2);
    if (!bar || !bar1 || !bar2 || !needbits) {
        set_error("lcp_build: workspace too small");
        return RV_ERR_NOMEM;
    }
    CodeTable tab;
    memset(&tab, 0, sizeof tab);
    const unsigned prep_blocks = (unsigned)((n + TP_BLOCK - 1) / TP_BLOCK + 1);
    RV_CUDA(cudaMemsetAsync(bar2, 0, (size_t)(n / 32768 + 8) * 4, st.s));
    RV_LAUNCH((sa_textprep_kernel<false>), prep_blocks, TP_THREADS, 0, st.s, dT, n, tab, bar, bar1, bar2, (u32 *)nullptr);
    const Barriers bars = {bar, bar1, bar2};
    RV_CUDA(cudaMemsetAsync(needbits, 0xff, (size_t)(n / 32 + 2) * 4, st.s));
    const i64 lthreads = ((n + 31) / 32 + LS_WORDS - 1) / LS_WORDS;
    RV_TRY(prof_begin(st, RV_PROF_LCP));
    RV_LAUNCH_PDL((lcp_sparse_kernel<8>), (unsigned)((lthreads + 127) / 128), 128, 0, st.s, needbits, n, (const u32 *)dT, bars, dSA, dISA, dLCP);
    RV_TRY(prof_end(st, RV_PROF_LCP, 1, (long long)n * 13));
    st.launches += 2;
    RV_KCHECK();
    ws.off = ws_mark;  // stream-ordered: later users of the workspace are enqueued behind these kernels
    return RV_OK;
}

}  // namespace rv
