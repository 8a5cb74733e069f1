// Shared helpers for the sm_100a kernels (guitar-audio-transcriber hot path).
#pragma once

#ifndef GAT_CPU_EMU
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#define GAT_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define GAT_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

namespace gat {

constexpr int kWarp = 32;
constexpr double kPi = 3.14159265358979323846264338327950288;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u < v ? u : v;
    }
    return v;
}

// numpy's reflect padding (no edge repeat): index i of the padded signal maps to a source index in [0, n).
__device__ __forceinline__ long long reflect_index(long long i, long long n) {
    if (n == 1) return 0;
    long long period = 2 * (n - 1);
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - i;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Asynchronous 4-byte global -> shared copies (LDGSTS): no register staging, completion by group.
#ifndef GAT_CPU_EMU
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
inline void cp_async4(float* dst_smem, const float* src) { *dst_smem = *src; }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
#endif

}  // namespace gat
