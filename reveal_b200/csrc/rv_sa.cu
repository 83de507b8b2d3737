// rv_sa.cu -- GPU suffix-array builder (prefix doubling over radix-sorted keys).
//
// Replaces the reference's  divsufsort(T, SA, n)  (reveallib/interface.c:213-222,
// divsufsort/divsufsort.c:333)  and the inverse fill  SAi[SA[i]] = i
// (reveallib/interface.c:235-238).  The suffix array of a byte string is unique
// (plain unsigned-byte lexicographic order, a suffix that is a proper prefix of
// another sorts first), so any correct builder is bit-identical to divsufsort.
//
// Algorithm (Manber-Myers / Larsson-Sadakane doubling, GPU form):
//   1. 256-bin histogram of T -> dense symbol codes 1..sigma (0 = past the end),
//      b = bits per code, k = 64/b symbols fit one 64-bit key;
//   2. key[i] = first k symbols of suffix i; radix sort (key, i);
//   3. equal-key runs are "groups"; rank[i] = SA slot of the first member of
//      i's group; groups of one suffix are final and leave the active list;
//   4. round h = k, 2k, 4k...: for every active suffix i the sort key is
//      (rank[i] : rank[i+h]+1 or 0 past the end); radix sort the active list,
//      which permutes suffixes only inside their groups; split groups where the
//      second half differs; write SA slots + ranks; compact the active list.
//   When the active list is empty every rank is the suffix's final SA slot,
//   i.e. the rank array IS the inverse suffix array.
#include "rv_radix.cuh"

namespace rv {

static const int AP_THREADS = 256;
static const int AP_IPT = 8;
static const int AP_TILE = AP_THREADS * AP_IPT;  // 2048 active entries per tile

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

// key[i] = codes of T[i..i+k) packed most-significant-first, b bits each; val[i] = i
__global__ void __launch_bounds__(256) sa_keygen_kernel(const unsigned char *__restrict__ T, i64 n, CodeTable tab, int b, int k,
                                                       u64 *__restrict__ keys, u32 *__restrict__ vals) {
    __shared__ unsigned short s_code[256];
    s_code[threadIdx.x] = tab.code[threadIdx.x];
    __syncthreads();
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 key = 0;
    for (int t = 0; t < k; t++) {
        i64 p = i + t;
        u64 c = p < n ? (u64)s_code[T[p]] : 0ull;
        key = (key << b) | c;
    }
    keys[i] = key;
    vals[i] = (u32)i;
}

// key[e] = rank[sa[e]] : (rank[sa[e]+h]+1, or 0 when the suffix ends first)
__global__ void __launch_bounds__(256) sa_gather_kernel(const u32 *__restrict__ sa, const u32 *__restrict__ grp, const int *__restrict__ rank,
                                                       i64 A, i64 n, i64 h, u64 *__restrict__ keys) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A) return;
    i64 p = (i64)sa[e] + h;
    u32 k2 = p < n ? (u32)rank[p] + 1u : 0u;
    keys[e] = ((u64)grp[e] << 32) | (u64)k2;
}

// head: first entry of its equal-key run.  active: the run still needs doubling rounds, i.e. it has more
// than G members (G = 1 in the doubling rounds: any tie) or the comparison kernel deferred it.
__device__ __forceinline__ void ap_flags(const u64 *__restrict__ keys, i64 A, i64 e, int G, const unsigned char *__restrict__ deferred,
                                         bool &head, bool &active) {
    u64 k = keys[e];
    int L = 0, R = 0;
    while (L < G && e - L - 1 >= 0 && keys[e - L - 1] == k) L++;
    while (R < G && e + R + 1 < A && keys[e + R + 1] == k) R++;
    head = L == 0;
    active = (L + R + 1 > G) || (deferred && L + R > 0 && deferred[e - L]);
}

// per tile: (largest slot+1 of a group head, number of entries that stay active)
__global__ void __launch_bounds__(AP_THREADS) sa_reduce_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ pos, i64 A, int G,
                                                               const unsigned char *__restrict__ deferred,
                                                               u32 *__restrict__ tile_max, u32 *__restrict__ tile_cnt) {
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
    __shared__ u32 s1[33], s2[33];
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
        __shared__ u32 s_im[1024];
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
__global__ void __launch_bounds__(AP_THREADS)
sa_apply_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ sa, const u32 *__restrict__ pos, i64 A, int G,
                const unsigned char *__restrict__ deferred, const u32 *__restrict__ tile_max, const u32 *__restrict__ tile_cnt, int *__restrict__ SA, int *__restrict__ rank,
                u32 *__restrict__ sa2, u32 *__restrict__ pos2, u32 *__restrict__ grp2) {
    __shared__ u32 s1[33], s2[33];
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
    __shared__ u32 s_im[AP_THREADS];
    s_im[threadIdx.x] = imax;
    __syncthreads();
    u32 pre_max = threadIdx.x > 0 ? s_im[threadIdx.x - 1] : 0u;
    u32 cm = tile_max[blockIdx.x];
    pre_max = pre_max > cm ? pre_max : cm;
    u32 dst = tile_cnt[blockIdx.x] + isum - cnt;
#pragma unroll
    for (int k = 0; k < AP_IPT; k++) {
        i64 e = base + k;
        if (e < A && (act[k] || G == 1)) {  // G > 1: the other entries were finished by sa_finish_small_kernel
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
            }
        }
    }
}

// ---- small groups: finish by direct suffix comparison ---------------------------------
// After the k-mer sort nearly every group of similar genomes is a handful of homologous
// positions whose suffixes agree for ~1/divergence characters.  One thread orders such a
// group (<= SA_SMALL_G members) by comparing the suffixes themselves, four text bytes per
// step through funnel-shifted aligned words, instead of log(LCP) doubling rounds over the
// whole array.  Comparisons longer than SA_CMP_CAP bytes defer the group to the doubling
// rounds, which bound the work for long repeats / identical sequences.
static const int SA_SMALL_G = 16;
static const int SA_CMP_CAP = 4096;

// order of suffixes a and b (a != b) that agree on their first `skip` characters.
// returns -1 (a < b), +1 (a > b), 0 (undecided within SA_CMP_CAP bytes).  T is 4-byte
// aligned and readable (zero padded) up to n + 8.
__device__ __forceinline__ int suffix_cmp(const unsigned char *__restrict__ T, i64 n, u32 a, u32 b, int skip) {
    const u32 *__restrict__ W = (const u32 *)T;
    i64 p = (i64)a + skip, q = (i64)b + skip;
    i64 lenmin = (n - p) < (n - q) ? (n - p) : (n - q);  // >= 0
    i64 ia = p >> 2, ib = q >> 2;
    const unsigned sha = (unsigned)(p & 3) * 8u, shb = (unsigned)(q & 3) * 8u;
    u32 lo_a = W[ia], lo_b = W[ib];
    for (i64 h = 0; h < lenmin; h += 4) {
        u32 hi_a = W[++ia], hi_b = W[++ib];
        u32 wa = __funnelshift_r(lo_a, hi_a, sha), wb = __funnelshift_r(lo_b, hi_b, shb);
        lo_a = hi_a;
        lo_b = hi_b;
        u32 x = wa ^ wb;
        if (x) {
            int byte = (__ffs((int)x) - 1) >> 3;  // first differing byte in text order (little-endian words)
            if (h + byte >= lenmin) break;        // the difference lies past the end of the shorter suffix
            u32 ca = (wa >> (8 * byte)) & 0xffu, cb = (wb >> (8 * byte)) & 0xffu;
            return ca < cb ? -1 : 1;
        }
        if (h >= SA_CMP_CAP) return 0;
    }
    return p > q ? -1 : 1;  // one suffix is a proper prefix of the other: the shorter one sorts first
}

__global__ void __launch_bounds__(256)
sa_finish_small_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ sa, i64 n, const unsigned char *__restrict__ T, int skip,
                       int *__restrict__ SA, int *__restrict__ rank, unsigned char *__restrict__ deferred, u32 *__restrict__ flag_large) {
    i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    u64 key = keys[j];
    if (j > 0 && keys[j - 1] == key) return;  // not the head of its group
    int g = 1;
    while (g <= SA_SMALL_G && j + g < n && keys[j + g] == key) g++;
    if (g == 1) {
        u32 s = sa[j];
        SA[j] = (int)s;
        rank[s] = (int)j;
        return;
    }
    if (g > SA_SMALL_G) {
        *flag_large = 1u;
        return;
    }
    u32 ids[SA_SMALL_G];
    ids[0] = sa[j];
    bool undecided = false;
    for (int t = 1; t < g; t++) {  // insertion sort; every comparison starts after the shared k-mer
        u32 x = sa[j + t];
        int at = t;
        while (at > 0) {
            int c = suffix_cmp(T, n, x, ids[at - 1], skip);
            if (c == 0) undecided = true;
            if (c >= 0) break;
            ids[at] = ids[at - 1];
            at--;
        }
        ids[at] = x;
        if (undecided) break;
    }
    if (undecided) {
        deferred[j] = 1;
        *flag_large = 1u;
        return;
    }
    for (int t = 0; t < g; t++) {
        SA[j + t] = (int)ids[t];
        rank[ids[t]] = (int)(j + t);
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
    // keys x2 (u64), vals x2, pos x2, grp x2 (u32), tile aggregates, radix scratch, small stuff
    return a * (8 + 8 + 4 + 4 + 4 + 4 + 4 + 4 + 1) + (size_t)tiles * 8 + radix_scratch_bytes(n) + 16 * 256 * 16 + (1 << 16);
}

int sa_build(Stream &st, Arena &ws, const unsigned char *dT, i64 n, int *dSA, int *dISA, PhaseTimes *pt) {
    if (n <= 0) return RV_OK;
    if (n >= ((i64)1 << 30)) {
        set_error("sa_build: n=%lld not supported yet (limit 2^30-1)", (long long)n);
        return RV_ERR_UNSUPPORTED;
    }
    u64 *k0 = ws.take<u64>(n), *k1 = ws.take<u64>(n);
    u32 *v0 = ws.take<u32>(n), *v1 = ws.take<u32>(n);
    u32 *posA = ws.take<u32>(n), *posB = ws.take<u32>(n);
    u32 *grpA = ws.take<u32>(n), *grpB = ws.take<u32>(n);
    const i64 tiles_n = (n + AP_TILE - 1) / AP_TILE;
    u32 *tile_max = ws.take<u32>(tiles_n), *tile_cnt = ws.take<u32>(tiles_n);
    void *rscratch = ws.take<unsigned char>(radix_scratch_bytes(n));
    u32 *small = ws.take<u32>(512);  // [0..255] byte histogram, [256] active count, [257] "large groups exist"
    unsigned char *deferred = ws.take<unsigned char>(n);
    if (!deferred || !k0 || !k1 || !v0 || !v1 || !posA || !posB || !grpA || !grpB || !tile_max || !tile_cnt || !rscratch || !small) {
        set_error("sa_build: workspace too small");
        return RV_ERR_NOMEM;
    }

    // 1. alphabet
    u32 hist[256];
    RV_CUDA(cudaMemsetAsync(small, 0, 512 * 4, st.s));
    {
        i64 blocks = (n + 256 * 64 - 1) / (256 * 64);
        if (blocks > 148 * 8) blocks = 148 * 8;
        RV_LAUNCH(sa_bytehist_kernel, (unsigned)blocks, 256, 0, st.s, dT, n, small);
        st.launches++;
    }
    RV_CUDA(cudaMemcpyAsync(hist, small, 256 * 4, cudaMemcpyDeviceToHost, st.s));
    RV_CUDA(cudaStreamSynchronize(st.s));
    CodeTable tab;
    memset(&tab, 0, sizeof tab);
    int sigma = 0;
    for (int c = 0; c < 256; c++)
        if (hist[c]) tab.code[c] = (unsigned short)(++sigma);
    const int b = bits_for((u64)sigma);  // codes 0..sigma
    const int k = 64 / b;

    // 2. initial keys + sort
    RV_LAUNCH(sa_keygen_kernel, (unsigned)((n + 255) / 256), 256, 0, st.s, dT, n, tab, b, k, k0, v0);
    st.launches++;
    bool in0;
    RV_TRY(radix_sort_pairs<u64>(st, k0, k1, v0, v1, n, make_plan(0, b * k), rscratch, &in0));
    if (pt) pt->sa_sorted_items += n;
    u64 *keys = in0 ? k0 : k1;
    u32 *sa = in0 ? v0 : v1;
    u64 *keys_alt = in0 ? k1 : k0;
    u32 *sa_alt = in0 ? v1 : v0;

    // 3. finish singletons and small groups by direct comparison
    RV_CUDA(cudaMemsetAsync(deferred, 0, (size_t)n, st.s));
    RV_LAUNCH(sa_finish_small_kernel, (unsigned)((n + 255) / 256), 256, 0, st.s, keys, sa, n, dT, k, dSA, dISA, deferred, small + 257);
    st.launches++;
    {
        u32 large = 0;
        RV_CUDA(cudaMemcpyAsync(&large, small + 257, 4, cudaMemcpyDeviceToHost, st.s));
        RV_CUDA(cudaStreamSynchronize(st.s));
        if (!large) {
            RV_KCHECK();
            return RV_OK;
        }
    }

    // 4. prefix doubling for what is left (groups of more than SA_SMALL_G suffixes, deferred groups)
    u32 *pos = nullptr, *grp = nullptr;   // current active list is (sa, pos, grp)
    u32 *pos_next = posA, *grp_next = grpA;
    i64 A = n;
    const int nbits = bits_for((u64)n);  // ranks < n, second key <= n
    i64 h = k;
    for (int round = 0;; round++) {
        const i64 tiles = (A + AP_TILE - 1) / AP_TILE;
        const int G = round == 0 ? SA_SMALL_G : 1;
        const unsigned char *dfr = round == 0 ? deferred : nullptr;
        RV_LAUNCH(sa_reduce_kernel, (unsigned)tiles, AP_THREADS, 0, st.s, keys, pos, A, G, dfr, tile_max, tile_cnt);
        RV_LAUNCH(sa_tilescan_kernel, 1, 1024, 0, st.s, tile_max, tile_cnt, tiles, small + 256);
        // the compacted entries go to the buffers not holding the current list
        RV_LAUNCH(sa_apply_kernel, (unsigned)tiles, AP_THREADS, 0, st.s, keys, sa, pos, A, G, dfr, tile_max, tile_cnt, dSA, dISA, sa_alt, pos_next,
                  grp_next);
        st.launches += 3;
        u32 nactive = 0;
        RV_CUDA(cudaMemcpyAsync(&nactive, small + 256, 4, cudaMemcpyDeviceToHost, st.s));
        RV_CUDA(cudaStreamSynchronize(st.s));
        if (pt) pt->sa_rounds = round + 1;
        if (nactive == 0) break;
        if (h >= n) {
            set_error("sa_build: internal error, %u suffixes still tied at h=%lld >= n", nactive, (long long)h);
            return RV_ERR_STATE;
        }
        // the next active list lives in (sa_alt, pos_next, grp_next)
        A = nactive;
        { u32 *t = sa; sa = sa_alt; sa_alt = t; }
        pos = pos_next;
        grp = grp_next;
        pos_next = (pos == posA) ? posB : posA;
        grp_next = (grp == grpA) ? grpB : grpA;
        // keys for this round go to `keys` (free now), sorted ping-pong with keys_alt / (sa, sa_alt)
        RV_LAUNCH(sa_gather_kernel, (unsigned)((A + 255) / 256), 256, 0, st.s, sa, grp, dISA, A, n, h, keys);
        st.launches++;
        bool r0;
        RV_TRY(radix_sort_pairs<u64>(st, keys, keys_alt, sa, sa_alt, A, make_plan(0, nbits, 32, 32 + nbits), rscratch, &r0));
        if (pt) pt->sa_sorted_items += A;
        if (!r0) {
            { u64 *t = keys; keys = keys_alt; keys_alt = t; }
            { u32 *t = sa; sa = sa_alt; sa_alt = t; }
        }
        h *= 2;
    }
    RV_KCHECK();
    return RV_OK;
}

}  // namespace rv
