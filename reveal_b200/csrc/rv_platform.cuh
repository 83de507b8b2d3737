// rv_platform.cuh -- one place where the kernels meet the toolchain.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a (real CUDA).
// Test-only build: g++ -DRV_EMU with tests/emu/cuda_emu.h, a fiber-based
// emulation of blocks/warps used in the GPU-less container to check kernel
// logic against the oracle (never shipped, never loaded by the product path).
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef RV_EMU
#include "cuda_emu.h"
#define RV_LAUNCH(kern, grid, block, smem, stream, ...)                                    \
    do {                                                                                   \
        auto rv_k_ = kern;                                                                 \
        emu::launch(#kern, dim3(grid), dim3(block), (size_t)(smem), [&]() { rv_k_(__VA_ARGS__); }); \
    } while (0)
#define RV_LAUNCH_PDL RV_LAUNCH
#define RV_DYN_SMEM(T, name) T *name = (T *)emu::S().dyn_smem
#define RV_SPIN() rv_emu_spin()
#else
#include <cuda_runtime.h>
#define RV_LAUNCH(kern, grid, block, smem, stream, ...)                \
    do {                                                               \
        auto rv_k_ = kern;                                             \
        rv_k_<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);     \
    } while (0)
// Programmatic dependent launch: the kernel may be made resident while its predecessor on the stream is still draining (its
// blocks wait in grid_dep_wait(), the first statement of every kernel launched this way), which takes the launch latency and
// the ramp of the first wave out of the gap between two kernels of a chain.  RV_PDL=0 builds plain launches.
#ifndef RV_PDL
#define RV_PDL 1
#endif
#if RV_PDL
#define RV_LAUNCH_PDL(kern, grid_, block_, smem_, stream_, ...)                                   \
    do {                                                                                        \
        auto rv_k_ = kern;                                                                      \
        cudaLaunchConfig_t rv_cfg_ = {};                                                        \
        rv_cfg_.gridDim = dim3(grid_);                                                         \
        rv_cfg_.blockDim = dim3(block_);                                                       \
        rv_cfg_.dynamicSmemBytes = (size_t)(smem_);                                            \
        rv_cfg_.stream = (stream_);                                                            \
        cudaLaunchAttribute rv_at_[1];                                                          \
        rv_at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
        rv_at_[0].val.programmaticStreamSerializationAllowed = 1;                               \
        rv_cfg_.attrs = rv_at_;                                                                 \
        rv_cfg_.numAttrs = 1;                                                                   \
        cudaLaunchKernelEx(&rv_cfg_, rv_k_, __VA_ARGS__);                                       \
    } while (0)
#else
#define RV_LAUNCH_PDL RV_LAUNCH
#endif
#define RV_DYN_SMEM(T, name)                                    \
    extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    T *name = (T *)name##_raw_
#define RV_SPIN() ((void)0)
#endif

namespace rv {

typedef unsigned int u32;
typedef unsigned long long u64;
typedef int64_t i64;

static const unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }

// first statement of a kernel that may be launched with RV_LAUNCH_PDL: everything the predecessor wrote is visible after it
// (a no-op for a plain launch); then the successor, if it was launched the same way, may be made resident
__device__ __forceinline__ void grid_dep_wait() {
#if !defined(RV_EMU) && RV_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// volatile single-word global accesses for inter-block protocols
__device__ __forceinline__ u32 ld_volatile(const u32 *p) { return *(const volatile u32 *)p; }
__device__ __forceinline__ void st_volatile(u32 *p, u32 v) { *(volatile u32 *)p = v; }

// ---- TMA: 1-D bulk copy global -> shared, completion on an mbarrier (sm_90+: cp.async.bulk / UBLKCP) ----------
// dst, src 16-byte aligned, bytes a multiple of 16.  One thread arms the barrier with the byte count and issues
// the copies; every thread of the block then waits on the barrier's phase.
#ifdef RV_EMU
// emulation: word = [phase bit 63 | pending bytes]; the copy happens at issue, waiting threads yield until the phase flips
__device__ __forceinline__ void mbar_init(u64 *bar, u32) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) { *bar += bytes; }
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, u64 *bar) {
    memcpy(dst, src, bytes);
    *bar -= bytes;
    if ((*bar & 0xffffffffull) == 0) *bar ^= 1ull << 63;
    emu::S().progress++;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    while (((*(volatile u64 *)bar) >> 63) == (u64)parity) rv_emu_spin();
}
#else
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
#endif

// ---- warp / block scan helpers (warp-shuffle based) ---------------------------
template <class T> __device__ __forceinline__ T warp_incl_sum(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= (unsigned)d) v += t;
    }
    return v;
}
__device__ __forceinline__ u32 warp_incl_max(u32 v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= (unsigned)d) v = v > t ? v : t;
    }
    return v;
}

// Block-wide scans for blocks of NT threads (NT multiple of 32, <= 1024).
// `scratch` must hold 33 words and must not be reused by another scan without
// a __syncthreads in between.  Returns the inclusive scan value of the calling
// thread; *total receives the block aggregate.  Contains two __syncthreads.
template <int NT, class T> __device__ __forceinline__ T block_incl_sum(T v, T *scratch, T *total) {
    const int W = NT / 32;
    T inc = warp_incl_sum(v);
    unsigned w = threadIdx.x >> 5;
    if (lane_id() == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        T t = lane_id() < (unsigned)W ? scratch[lane_id()] : (T)0;
        T ti = warp_incl_sum(t);
        if (lane_id() < (unsigned)W) scratch[lane_id()] = ti - t;  // exclusive warp offsets
        if (lane_id() == 31) scratch[32] = ti;
    }
    __syncthreads();
    T r = inc + scratch[w];
    *total = scratch[32];
    return r;
}
template <int NT> __device__ __forceinline__ u32 block_incl_max(u32 v, u32 *scratch, u32 *total) {
    const int W = NT / 32;
    u32 inc = warp_incl_max(v);
    unsigned w = threadIdx.x >> 5;
    if (lane_id() == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 t = lane_id() < (unsigned)W ? scratch[lane_id()] : 0u;
        u32 ti = warp_incl_max(t);
        u32 ex = __shfl_up_sync(FULL, ti, 1);
        if (lane_id() == 0) ex = 0u;
        if (lane_id() < (unsigned)W) scratch[lane_id()] = ex;  // max over the preceding warps
        if (lane_id() == 31) scratch[32] = ti;
    }
    __syncthreads();
    u32 pre = scratch[w];
    u32 r = inc > pre ? inc : pre;
    *total = scratch[32];
    return r;
}

}  // namespace rv
