// radix_yardstick.cu -- how far is the hand-written onesweep sort of rv_radix.cuh from a tuned library sort on this GPU?
// Sorts n (24-bit key, 32-bit value) pairs with radix_sort_pairs (the product's sort, stored keys) and with
// cub::DeviceRadixSort::SortPairs (library; a yardstick only, never on the product path) and prints both times.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I reveal_b200/csrc scripts/radix_yardstick.cu -o build/radix_yardstick
#include "rv_radix.cuh"
#include <cub/cub.cuh>
#include <stdarg.h>
#include <stdlib.h>
#include <vector>

namespace rv {
void set_error(const char *fmt, ...) {
    va_list va;
    va_start(va, fmt);
    vfprintf(stderr, fmt, va);
    va_end(va);
    fputc('\n', stderr);
}
}  // namespace rv
using namespace rv;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv) {
    const i64 n = argc > 1 ? atoll(argv[1]) : 10002906;
    const int bits = argc > 2 ? atoi(argv[2]) : 24;
    std::vector<u32> hk((size_t)n), hv((size_t)n);
    u64 x = 88172645463325252ull;
    for (i64 i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        hk[(size_t)i] = (u32)(x >> 20) & ((1u << bits) - 1u);
        hv[(size_t)i] = (u32)i;
    }
    u32 *k0, *k1, *v0, *v1, *kin, *flush;
    CK(cudaMalloc(&k0, n * 4)); CK(cudaMalloc(&k1, n * 4)); CK(cudaMalloc(&v0, n * 4)); CK(cudaMalloc(&v1, n * 4)); CK(cudaMalloc(&kin, n * 4));
    CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemcpy(kin, hk.data(), n * 4, cudaMemcpyHostToDevice));
    void *scratch;
    CK(cudaMalloc(&scratch, radix_scratch_bytes(n)));
    Stream st;
    CK(cudaStreamCreate(&st.s));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, k1, v0, v1, (int)n, 0, bits, st.s);
    void *tmp;
    CK(cudaMalloc(&tmp, tmp_bytes));
    float ours = 0, lib = 0;
    const int reps = 7;
    for (int r = 0; r < reps + 2; r++) {
        CK(cudaMemcpyAsync(k0, kin, n * 4, cudaMemcpyDeviceToDevice, st.s));
        CK(cudaMemcpyAsync(v0, hv.data(), n * 4, cudaMemcpyHostToDevice, st.s));
        CK(cudaMemsetAsync(flush, 1, 256 << 20, st.s));
        CK(cudaEventRecord(e0, st.s));
        bool in0;
        if (radix_sort_pairs<u32>(st, k0, k1, v0, v1, n, make_plan(0, bits), scratch, &in0) != 0) return 1;
        CK(cudaEventRecord(e1, st.s));
        CK(cudaStreamSynchronize(st.s));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) ours += ms;
        CK(cudaMemsetAsync(flush, 1, 256 << 20, st.s));
        CK(cudaEventRecord(e0, st.s));
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, k1, v0, v1, (int)n, 0, bits, st.s);
        CK(cudaEventRecord(e1, st.s));
        CK(cudaStreamSynchronize(st.s));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) lib += ms;
    }
    printf("{\"n\": %lld, \"key_bits\": %d, \"rv_radix_ms\": %.4f, \"cub_ms\": %.4f}\n", (long long)n, bits, ours / reps, lib / reps);
    return 0;
}
