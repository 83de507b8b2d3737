// rv_internal.h -- host-side plumbing shared by the translation units of
// libreveal_b200.so (workspace arena, error propagation, phase entry points).
#pragma once
#include "rv_platform.cuh"
#include "../../include/reveal_b200.h"
#include <stdio.h>
#include <string.h>
#include <vector>

namespace rv {

// ---- error propagation: every entry point returns 0 or a negative code -------
// (codes: enum rv_status in include/reveal_b200.h)

void set_error(const char *fmt, ...);

#define RV_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t rv_e_ = (call);                                                              \
        if (rv_e_ != cudaSuccess) {                                                              \
            rv::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(rv_e_)); \
            return RV_ERR_CUDA;                                                              \
        }                                                                                        \
    } while (0)
#define RV_TRY(call)                 \
    do {                             \
        int rv_r_ = (call);          \
        if (rv_r_ != 0) return rv_r_; \
    } while (0)
#define RV_KCHECK() RV_CUDA(cudaGetLastError())

// ---- grow-only device workspace (bump allocator, 256-byte aligned) ------------
// One slab per index handle: repeated builds of the same size never touch
// cudaMalloc again, and everything a build needs is contiguous in HBM.
struct Arena {
    unsigned char *base = nullptr;
    size_t cap = 0, off = 0;
    int reserve(size_t bytes);
    void reset() { off = 0; }
    void release();
    template <class T> T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
        if (off + bytes > cap) return nullptr;
        T *p = (T *)(base + off);
        off += bytes;
        return p;
    }
};

struct PhaseTimes {  // device milliseconds (CUDA events on the build stream)
    float pack = 0, sa = 0, lcp = 0, so = 0, total = 0;
    int sa_rounds = 0;
    long long sa_sorted_items = 0;  // sum over radix-sort calls of the item count
    int launches = 0;
};

struct ProfRec {
    cudaEvent_t e0, e1;
    int slot;
    long long launches, bytes;
};

// alphabet of the last text built on a handle: lets the next build start without waiting for its histogram
struct AlphaCache {
    bool valid = false;
    unsigned short code[256];  // exact dense codes 1..sigma of the symbols present (0: absent), byte order
    unsigned short kcls[256];  // k-mer key class of every symbol: monotone in the byte value, rare symbols merged into a neighbour
    int sigma = 0, sigma_eff = 0, kbase = 2;
};

struct Stream {
    cudaStream_t s = 0;
    // Side stream of a build: work that does not depend on the radix sort (byte histogram, text pass, clearing of the comparison
    // stage's tables) runs beside the digit passes; forked from / joined into `s` with the two events.  Created on first use.
    cudaStream_t side = 0;
    cudaEvent_t ev_fork = 0, ev_join = 0;
    u32 *pinned = nullptr;      // 2 KB of pinned host memory for small asynchronous read-backs
    AlphaCache alpha;
    int launches = 0;           // kernels launched by this library on the stream since the last reset
    long long launches_total = 0;
    // optional per-kernel profile: event pairs recorded around runs of one kernel, resolved lazily
    // (prof_collect) so that profiling adds no synchronisation to the measured step
    bool prof = false;
    unsigned prof_mask = 0xffffffffu;   // bit k: slot k is bracketed (rv_profile)
    std::vector<ProfRec> pending;
    std::vector<cudaEvent_t> free_events;
    cudaEvent_t cur0 = 0;
    double prof_ms[RV_PROF_SLOTS] = {0};          // per slot (enum rv_prof_slot): summed kernel time
    long long prof_launches[RV_PROF_SLOTS] = {0};
    long long prof_bytes[RV_PROF_SLOTS] = {0};    // algorithmic bytes moved by those launches
};

inline int prof_event(Stream &st, cudaEvent_t *e) {
    if (!st.free_events.empty()) {
        *e = st.free_events.back();
        st.free_events.pop_back();
        return RV_OK;
    }
    RV_CUDA(cudaEventCreate(e));
    return RV_OK;
}
// (`on`: the stream the bracketed kernels are launched on, st.s by default; begin / end pairs do not nest)
inline int prof_begin(Stream &st, int slot, cudaStream_t on = 0) {
    if (!st.prof || !((st.prof_mask >> slot) & 1u)) return RV_OK;
    RV_TRY(prof_event(st, &st.cur0));
    RV_CUDA(cudaEventRecord(st.cur0, on ? on : st.s));
    return RV_OK;
}
inline int prof_end(Stream &st, int slot, long long launches, long long bytes, cudaStream_t on = 0) {
    if (!st.prof || !((st.prof_mask >> slot) & 1u)) return RV_OK;
    ProfRec r;
    r.e0 = st.cur0;
    RV_TRY(prof_event(st, &r.e1));
    RV_CUDA(cudaEventRecord(r.e1, on ? on : st.s));
    r.slot = slot;
    r.launches = launches;
    r.bytes = bytes;
    st.pending.push_back(r);
    return RV_OK;
}
// fork: the side stream waits for everything enqueued on st.s so far; join: st.s waits for the side stream
inline int side_fork(Stream &st) {
    if (!st.ev_fork) {
        RV_CUDA(cudaStreamCreateWithFlags(&st.side, cudaStreamNonBlocking));
        RV_CUDA(cudaEventCreateWithFlags(&st.ev_fork, cudaEventDisableTiming));
        RV_CUDA(cudaEventCreateWithFlags(&st.ev_join, cudaEventDisableTiming));
    }
    RV_CUDA(cudaEventRecord(st.ev_fork, st.s));
    RV_CUDA(cudaStreamWaitEvent(st.side, st.ev_fork, 0));
    return RV_OK;
}
inline int side_join(Stream &st) {
    RV_CUDA(cudaEventRecord(st.ev_join, st.side));
    RV_CUDA(cudaStreamWaitEvent(st.s, st.ev_join, 0));
    return RV_OK;
}
// resolve the recorded pairs (synchronises the stream)
inline int prof_collect(Stream &st) {
    if (st.pending.empty()) return RV_OK;
    RV_CUDA(cudaStreamSynchronize(st.s));
    if (st.side) RV_CUDA(cudaStreamSynchronize(st.side));
    for (ProfRec &r : st.pending) {
        float ms = 0;
        RV_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        st.prof_ms[r.slot] += ms;
        st.prof_launches[r.slot] += r.launches;
        st.prof_bytes[r.slot] += r.bytes;
        st.free_events.push_back(r.e0);
        st.free_events.push_back(r.e1);
    }
    st.pending.clear();
    return RV_OK;
}

// ---- phase entry points (each in its own .cu) ---------------------------------
size_t sa_workspace_bytes(i64 n);
// Builds SA and ISA (=final ranks) of the byte string dT[0..n) into dSA/dISA (int32).  dT must be 8-byte
// aligned and readable (zero padded) up to n+16.  When the comparison stage could place every suffix the
// barrier-aware LCP array is complete as well (*lcp_done); otherwise the caller runs lcp_build.
int sa_build(Stream &st, Arena &ws, const unsigned char *dT, i64 n, int *dSA, int *dISA, int *dLCP, bool *lcp_done, PhaseTimes *pt,
             bool fresh_alphabet = false);

int lcp_build(Stream &st, Arena &ws, const unsigned char *dT, i64 n, const int *dSA, const int *dISA, int *dLCP);
int isa_build(Stream &st, i64 n, const int *dSA, int *dISA, u32 *d_bad);
int so_build(Stream &st, i64 n, const i64 *dNsep, int nsamples, unsigned short *dSO);
int revcomp_suffix(Stream &st, unsigned char *dT, i64 start, i64 n);

}  // namespace rv
