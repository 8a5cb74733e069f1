// Microbenchmark: throughput of scalar FFMA, packed FFMA2 (fma.rn.f32x2) and mixes of the two on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma2_probe fma2_probe.cu && ./fma2_probe
// Prints TFLOP/s (2 flops per FMA lane-op).  Register-only chains, 1024 threads x 2 CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int kPacked, int kScalar>     // per inner round: kPacked FFMA2 chains and kScalar FFMA chains
__global__ void __launch_bounds__(1024, 2) probe(float* out, int iters, float a, float b) {
    float2 p[kPacked > 0 ? kPacked : 1];
    float s[kScalar > 0 ? kScalar : 1];
    for (int i = 0; i < kPacked; ++i) p[i] = make_float2((float)(threadIdx.x + i), (float)(threadIdx.x + 2 * i));
    for (int i = 0; i < kScalar; ++i) s[i] = (float)(threadIdx.x + 3 * i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < kPacked; ++i) p[i] = __ffma2_rn(p[i], a2, b2);
#pragma unroll
            for (int i = 0; i < kScalar; ++i) s[i] = fmaf(s[i], a, b);
        }
    }
    float acc = 0.0f;
    for (int i = 0; i < kPacked; ++i) acc += p[i].x + p[i].y;
    for (int i = 0; i < kScalar; ++i) acc += s[i];
    if (acc == 12345.678f) out[0] = acc;
}

template <int kPacked, int kScalar>
void run(const char* name, float* d, int sms) {
    const int iters = 4096, grid = 2 * sms;
    probe<kPacked, kScalar><<<grid, 1024>>>(d, iters, 0.999f, 0.001f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<kPacked, kScalar><<<grid, 1024>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)(2 * kPacked + kScalar) * 8.0 * iters * 1024.0 * grid;
    printf("%-28s %7.2f TFLOP/s  (%d FFMA2 + %d FFMA chains per thread, %.3f ms)\n", name, 2.0 * fmas / (ms * 1e-3) * 1e-12, kPacked, kScalar, ms);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 64);
    run<0, 16>("scalar FFMA", d, sms);
    run<8, 0>("packed FFMA2", d, sms);
    run<16, 0>("packed FFMA2 (16 chains)", d, sms);
    run<4, 8>("mix 1 FFMA2 : 2 FFMA", d, sms);
    run<6, 6>("mix 1 FFMA2 : 1 FFMA", d, sms);
    run<8, 4>("mix 2 FFMA2 : 1 FFMA", d, sms);
    run<8, 2>("mix 4 FFMA2 : 1 FFMA", d, sms);
    return 0;
}
