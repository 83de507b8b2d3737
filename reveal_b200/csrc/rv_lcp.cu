// rv_lcp.cu -- inverse suffix array from a cached suffix array, sample-id array and reverse complement.
//
// (compute_lcp, reveallib/interface.c:97-114, lives with the suffix-array builder: rv_sa.cu, lcp_sparse_kernel / lcp_build.)
// so_build replaces build_SO (interface.c:116-134); revcomp_suffix replaces
// revcomp + comp_tab (interface.c:136-158) as used by construct (interface.c:168-172).
#include "rv_internal.h"

namespace rv {

// SAi[SA[i]] = i  (interface.c:235-238) -- only used when the suffix array comes from a cache file
__global__ void __launch_bounds__(256) isa_scatter_kernel(const int *__restrict__ SA, i64 n, int *__restrict__ ISA, u32 *__restrict__ bad) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = SA[i];
    if (s < 0 || s >= n) {
        *bad = 1u;  // not a permutation of 0..n-1: a stale or foreign cache file
        return;
    }
    ISA[s] = (int)i;
}

// every slot must have won its scatter: a value that occurs twice leaves one of the two slots unanswered
__global__ void __launch_bounds__(256) isa_verify_kernel(const int *__restrict__ SA, i64 n, const int *__restrict__ ISA, u32 *__restrict__ bad) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = SA[i];
    if (s < 0 || s >= n || ISA[s] != (int)i) *bad = 1u;
}

int isa_build(Stream &st, i64 n, const int *dSA, int *dISA, u32 *d_bad) {
    if (n <= 0) return RV_OK;
    RV_LAUNCH(isa_scatter_kernel, (unsigned)((n + 255) / 256), 256, 0, st.s, dSA, n, dISA, d_bad);
    RV_LAUNCH(isa_verify_kernel, (unsigned)((n + 255) / 256), 256, 0, st.s, dSA, n, dISA, d_bad);
    st.launches += 2;
    RV_KCHECK();
    return RV_OK;
}

// SO[p] = number of sample separators strictly before p  (nsep[k] = position of the last '$' of sample k)
__global__ void __launch_bounds__(256) so_fill_kernel(i64 n, const i64 *__restrict__ nsep, int nsamples, unsigned short *__restrict__ SO) {
    i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int lo = 0, hi = nsamples - 1;  // count of k in [0,nsamples-1) with nsep[k] < p
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (nsep[mid] < p) lo = mid + 1; else hi = mid;
    }
    SO[p] = (unsigned short)lo;
}

int so_build(Stream &st, i64 n, const i64 *dNsep, int nsamples, unsigned short *dSO) {
    if (n <= 0) return RV_OK;
    RV_LAUNCH(so_fill_kernel, (unsigned)((n + 255) / 256), 256, 0, st.s, n, dNsep, nsamples, dSO);
    st.launches++;
    RV_KCHECK();
    return RV_OK;
}

// IUPAC-aware complement of interface.c:136-145 as a rule (identity below '@'):
// A<->T C<->G B<->V D<->H K<->M R<->Y in both cases, U->A, u->a, '`'->'@'.
__device__ __forceinline__ unsigned char comp_char(unsigned char c) {
    unsigned char up = c & 0xDFu, low = c & 0x20u;
    bool alpha = (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
    if (c == 96) return 64;
    if (!alpha) return c;
    unsigned char r = up;
    switch (up) {
        case 'A': r = 'T'; break;
        case 'T': r = 'A'; break;
        case 'U': r = 'A'; break;
        case 'C': r = 'G'; break;
        case 'G': r = 'C'; break;
        case 'B': r = 'V'; break;
        case 'V': r = 'B'; break;
        case 'D': r = 'H'; break;
        case 'H': r = 'D'; break;
        case 'K': r = 'M'; break;
        case 'M': r = 'K'; break;
        case 'R': r = 'Y'; break;
        case 'Y': r = 'R'; break;
        default: break;
    }
    return (unsigned char)(r | low);
}

__global__ void __launch_bounds__(256) revcomp_kernel(unsigned char *T, i64 start, i64 len) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 half = (len + 1) / 2;
    if (i >= half) return;
    i64 a = start + i, b = start + len - 1 - i;
    unsigned char ca = comp_char(T[a]), cb = comp_char(T[b]);
    T[a] = cb;
    T[b] = ca;  // a == b in the middle of an odd length: both hold comp(T[a])
}

int revcomp_suffix(Stream &st, unsigned char *dT, i64 start, i64 n) {
    i64 len = n - start;
    if (len <= 0) return RV_OK;
    i64 half = (len + 1) / 2;
    RV_LAUNCH(revcomp_kernel, (unsigned)((half + 255) / 256), 256, 0, st.s, dT, start, len);
    st.launches++;
    RV_KCHECK();
    return RV_OK;
}

}  // namespace rv
