// TEST INFRASTRUCTURE ONLY - minimal host emulation of the CUDA execution model.
//
// The build container has nvcc but no GPU.  To debug kernel LOGIC (indexing, FFT decomposition, scans)
// before spending GPU minutes, tests/emu compiles the very same csrc/*.cuh sources with g++ against this
// header: one std::thread per CUDA thread, blocks run one after another, __syncthreads/__syncwarp are
// real barriers, warp shuffles go through a per-warp exchange buffer.  It is slow and exists only for
// tests/test_emu_*.py; libgat.so (the product) is built by nvcc and never contains or loads any of this.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define GAT_CPU_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };

namespace emu {
struct BlockState {
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<std::array<uint64_t, 32>> xchg;  // per warp shuffle buffer
    std::vector<unsigned char> dyn_smem;
    unsigned nthreads = 0;
};
inline BlockState*& state() { static BlockState* s = nullptr; return s; }
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline thread_local unsigned t_linear = 0;

template <typename T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

inline void warp_sync() {
    BlockState* s = state();
    s->warp_bar[t_linear / 32]->arrive_and_wait();
}
inline unsigned warp_width() {  // threads present in this (possibly partial) warp
    BlockState* s = state();
    unsigned w = t_linear / 32;
    return std::min(32u, s->nthreads - w * 32);
}
template <typename T, typename F> inline T exchange(T v, F pick) {
    BlockState* s = state();
    auto& buf = s->xchg[t_linear / 32];
    buf[t_linear % 32] = to_bits(v);
    warp_sync();
    T r = from_bits<T>(buf[pick() % warp_width()]);
    warp_sync();
    return r;
}

template <typename Kernel, typename... Args>
void launch(Kernel kernel, dim3 grid, dim3 block, size_t smem, Args... args) {
    BlockState st;
    st.nthreads = block.x * block.y * block.z;
    st.block_bar = std::make_unique<std::barrier<>>(st.nthreads);
    unsigned nwarps = (st.nthreads + 31) / 32;
    for (unsigned w = 0; w < nwarps; ++w)
        st.warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32u, st.nthreads - w * 32)));
    st.xchg.resize(nwarps);
    st.dyn_smem.assign(smem + 64, 0);
    state() = &st;
    g_blockDim = block;
    g_gridDim = grid;
    std::vector<std::thread> threads;
    threads.reserve(st.nthreads);
    for (unsigned t = 0; t < st.nthreads; ++t) {
        threads.emplace_back([&, t]() {
            t_linear = t;
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        t_blockIdx = dim3(bx, by, bz);
                        kernel(args...);
                        state()->block_bar->arrive_and_wait();  // blocks are serialised: statics are per block
                    }
        });
    }
    for (auto& th : threads) th.join();
    state() = nullptr;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { emu::state()->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    return emu::exchange(v, [&] { return (unsigned)src; });
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    unsigned lane = emu::t_linear % 32;
    return emu::exchange(v, [&] { return lane ^ (unsigned)m; });
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    unsigned lane = emu::t_linear % 32;
    return emu::exchange(v, [&] { return lane + d < 32 ? lane + d : lane; });
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    unsigned lane = emu::t_linear % 32;
    return emu::exchange(v, [&] { return lane >= d ? lane - d : lane; });
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::BlockState* s = emu::state();
    auto& buf = s->xchg[emu::t_linear / 32];
    buf[emu::t_linear % 32] = pred ? 1 : 0;
    emu::warp_sync();
    unsigned r = 0;
    for (unsigned i = 0; i < emu::warp_width(); ++i) r |= (buf[i] ? 1u : 0u) << i;
    emu::warp_sync();
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == ((emu::warp_width() == 32) ? 0xffffffffu : ((1u << emu::warp_width()) - 1)); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }

// --- arithmetic intrinsics (IEEE, no contraction: build with -ffp-contract=off)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { volatile float x = a.x + b.x, y = a.y + b.y; return float2{x, y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { volatile float x = a.x * b.x, y = a.y * b.y; return float2{x, y}; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __ldg(const float* p) { return *p; }
static inline double __ldg(const double* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline float2 __ldg(const float2* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline double2 __ldg(const double2* p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
static inline float fminf_(float a, float b) { return fminf(a, b); }
static inline void sincospif(float x, float* s, float* c) { *s = (float)sin(M_PI * (double)x); *c = (float)cos(M_PI * (double)x); }

static std::mutex& emu_atomic_mutex() { static std::mutex m; return m; }
template <typename T> static inline T atomicAdd(T* p, T v) { std::lock_guard<std::mutex> g(emu_atomic_mutex()); T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { std::lock_guard<std::mutex> g(emu_atomic_mutex()); T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { std::lock_guard<std::mutex> g(emu_atomic_mutex()); T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { std::lock_guard<std::mutex> g(emu_atomic_mutex()); T o = *p; if (o == c) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { std::lock_guard<std::mutex> g(emu_atomic_mutex()); T o = *p; *p = v; return o; }

// --- the sliver of the runtime API that csrc/gat.cu uses
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
struct cudaEvent_st; typedef cudaEvent_st* cudaEvent_t;
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* t, cudaEvent_t, cudaEvent_t) { *t = 0; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };

#define GAT_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define GAT_DYN_SMEM(name) unsigned char* name = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(emu::state()->dyn_smem.data()) + 63) & ~uintptr_t(63))
