// cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never part of the product library).
//
// A tiny single-OS-thread emulation of the CUDA execution model, just enough
// to run reveal_b200/csrc/*.cu compiled with g++ (-DRV_EMU) inside this
// GPU-less build container so that kernel LOGIC (indexing, ranking, scans,
// look-back protocols, warp collectives) can be checked against the oracle
// before GPU minutes are spent.  It does NOT model the memory system, races
// between blocks or performance.
//
// Model: blocks of a launch run one after another in blockIdx order; the
// threads of a block are fibers (ucontext) scheduled round-robin; a fiber
// yields at __syncthreads / __syncwarp / warp collectives / spin-waits.
// `__shared__` maps to `static` (valid because one block runs at a time).
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{a, b}; }

namespace emu {

struct WarpState {
    unsigned arrived = 0;
    unsigned gen = 0;
    unsigned long long slot[32];
};

// Context of a fiber.  x86-64: the stack pointer of a suspended fiber (callee-saved registers on its stack, emu_switch in
// cuda_emu.cpp) -- swapcontext makes a signal-mask system call per switch, a third of the emulator's run time; elsewhere ucontext.
#if defined(__x86_64__)
#define RV_EMU_ASM_SWITCH 1
struct FiberCtx { void *sp = nullptr; };
#else
#define RV_EMU_ASM_SWITCH 0
typedef ucontext_t FiberCtx;
#endif

struct Fiber {
    FiberCtx ctx;
    char *stack = nullptr;
    bool done = true;
    uint3 tid;
};

struct State {
    uint3 t_idx{0, 0, 0}, b_idx{0, 0, 0};
    dim3 b_dim, g_dim;
    FiberCtx sched;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    int cur = -1;
    int nthreads = 0;
    int alive = 0;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    unsigned long long progress = 0;  // bumped whenever any fiber makes observable progress
    unsigned char *dyn_smem = nullptr;
    size_t dyn_smem_cap = 0;
    const std::function<void()> *body = nullptr;
    const char *kname = "?";
};

State &S();
void yield();
void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
[[noreturn]] void die(const char *msg);

inline int lane() { return (int)(S().t_idx.x & 31u); }
inline int linear_tid() { return (int)S().t_idx.x; }
inline WarpState &warp() { return S().warps[S().t_idx.x >> 5]; }

// barrier among the lanes of `mask` in the current warp
inline void warp_barrier(unsigned mask) {
    WarpState &w = warp();
    unsigned me = 1u << lane();
    if (!(mask & me)) die("warp collective: calling lane not in mask");
    {   // lanes beyond the end of a partial last warp do not exist
        int nl = S().nthreads - (int)(S().t_idx.x & ~31u);
        if (nl < 32) mask &= (1u << nl) - 1u;
    }
    unsigned g = w.gen;
    w.arrived |= me;
    S().progress++;
    if ((w.arrived & mask) == mask) {
        w.arrived &= ~mask;
        w.gen++;
        return;
    }
    while (w.gen == g) yield();
}

template <class T> inline unsigned long long to_bits(T v) {
    static_assert(sizeof(T) <= 8, "emu: shuffle payload too wide");
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> inline T from_bits(unsigned long long b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

// post value, run f(slots) once everyone in mask has posted, then re-sync
template <class T, class F> inline auto collective(unsigned mask, T v, F f) -> decltype(f((const unsigned long long *)nullptr)) {
    WarpState &w = warp();
    w.slot[lane()] = to_bits(v);
    warp_barrier(mask);
    auto r = f((const unsigned long long *)w.slot);
    warp_barrier(mask);
    return r;
}

}  // namespace emu

#define threadIdx (emu::S().t_idx)
#define blockIdx (emu::S().b_idx)
#define blockDim (emu::S().b_dim)
#define gridDim (emu::S().g_dim)
static const int warpSize = 32;

// ---- synchronisation --------------------------------------------------------
static inline void __syncthreads() {
    emu::State &s = emu::S();
    unsigned g = s.bar_gen;
    s.bar_arrived++;
    s.progress++;
    if (s.bar_arrived >= s.alive) {
        s.bar_arrived = 0;
        s.bar_gen++;
        return;
    }
    while (s.bar_gen == g) emu::yield();
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    emu::warp_barrier(mask);
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { emu::yield(); }
// spin-wait hint used by look-back loops: lets other fibers run
static inline void rv_emu_spin() { emu::yield(); }

// ---- warp collectives -------------------------------------------------------
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int l = emu::lane();
    int base = l & ~(width - 1);
    int s = base + (src & (width - 1));
    return emu::collective(mask, v, [&](const unsigned long long *sl) { return emu::from_bits<T>(sl[s]); });
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int l = emu::lane();
    int base = l & ~(width - 1);
    int s = l - (int)d;
    if (s < base) s = l;
    return emu::collective(mask, v, [&](const unsigned long long *sl) { return emu::from_bits<T>(sl[s]); });
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int l = emu::lane();
    int base = l & ~(width - 1);
    int s = l + (int)d;
    if (s >= base + width) s = l;
    return emu::collective(mask, v, [&](const unsigned long long *sl) { return emu::from_bits<T>(sl[s]); });
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    int l = emu::lane();
    int s = l ^ x;
    if ((s & ~(width - 1)) != (l & ~(width - 1))) s = l;
    return emu::collective(mask, v, [&](const unsigned long long *sl) { return emu::from_bits<T>(sl[s]); });
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    return emu::collective(mask, (unsigned)(pred != 0), [&](const unsigned long long *sl) {
        unsigned r = 0;
        for (int i = 0; i < 32; i++)
            if (((mask >> i) & 1u) && sl[i]) r |= 1u << i;
        return r;
    });
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __activemask() { return 0xffffffffu; }
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    unsigned long long mine = emu::to_bits(v);
    return emu::collective(mask, v, [&](const unsigned long long *sl) {
        unsigned r = 0;
        for (int i = 0; i < 32; i++)
            if (((mask >> i) & 1u) && sl[i] == mine) r |= 1u << i;
        return r;
    });
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    return emu::collective(mask, v, [&](const unsigned long long *sl) {
        unsigned r = 0;
        for (int i = 0; i < 32; i++)
            if ((mask >> i) & 1u) r += (unsigned)sl[i];
        return r;
    });
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    return emu::collective(mask, v, [&](const unsigned long long *sl) {
        unsigned r = 0;
        for (int i = 0; i < 32; i++)
            if (((mask >> i) & 1u) && (unsigned)sl[i] > r) r = (unsigned)sl[i];
        return r;
    });
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    return emu::collective(mask, v, [&](const unsigned long long *sl) {
        unsigned r = 0xffffffffu;
        for (int i = 0; i < 32; i++)
            if (((mask >> i) & 1u) && (unsigned)sl[i] < r) r = (unsigned)sl[i];
        return r;
    });
}

// ---- bit intrinsics ----------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned long long __brevll(unsigned long long x) {
    unsigned long long r = 0;
    for (int i = 0; i < 64; i++) r |= ((x >> i) & 1ull) << (63 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (s & 31u));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (s & 31u)) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned s = (sel >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)((v >> (8 * (s & 7u))) & 0xffu);
        if (s & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldcv(const T *p) { return *(const volatile T *)p; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }

// ---- atomics (single OS thread => plain read-modify-write) --------------------
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; emu::S().progress++; return o; }
template <class T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; emu::S().progress++; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; emu::S().progress++; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; emu::S().progress++; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; emu::S().progress++; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; emu::S().progress++; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; emu::S().progress++; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; emu::S().progress++; return o; }

// ---- host runtime subset -----------------------------------------------------
typedef int cudaError_t;
typedef struct emu_stream_st *cudaStream_t;
struct emu_event_st { double t; };
typedef emu_event_st *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };

cudaError_t cudaMalloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
cudaError_t emu_check_guards();
static inline cudaError_t cudaMallocHost(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
enum { cudaHostAllocMapped = 2 };
static inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return emu_check_guards(); }
static inline cudaError_t cudaDeviceSynchronize() { return emu_check_guards(); }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int *c) { *c = 1; return 0; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)8 << 30; return 0; }
static inline double emu_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event_st{0}; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event_st{0}; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = emu_now(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return 0; }
#define cudaFuncSetAttribute(...) (0)
#define cudaFuncAttributeMaxDynamicSharedMemorySize 0
