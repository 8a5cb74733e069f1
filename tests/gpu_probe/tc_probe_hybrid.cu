// GPU probe (test infrastructure): can the two 3xTF32 correction passes run as BF16 MMAs?
//   D = tf32(a_hi) x tf32(b_hi)   [kind::tf32, K = 8 per MMA]
//     + bf16(a_lo) x bf16(b_hi) + bf16(a_hi) x bf16(b_lo)   [kind::f16 with BF16 inputs, K = 16 per MMA]
// all accumulating into the same FP32 TMEM tile.  Checks the BF16 instruction descriptor, the no-swizzle K-major
// layout for 16-bit operands with a row shift, mixing MMA kinds on one accumulator, and the resulting accuracy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I guitar_audio_transcriber_ai_b200/csrc tests/gpu_probe/tc_probe_hybrid.cu -o tc_probe_hybrid
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc05.cuh"
using namespace gat::tc;

constexpr int M = 128, K = 32, ROWS = 176;

__host__ __device__ constexpr uint32_t idesc_16(int m, int n, int afmt, int bfmt) {      // 0 = F16, 1 = BF16
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int N>
__global__ void probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int shift, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* a_hi = reinterpret_cast<float*>(smem);                                   // [K/4][ROWS][4] tf32
    float* b_hi = a_hi + (K / 4) * ROWS * 4;                                        // [K/4][N][4]
    __nv_bfloat16* a_hb = reinterpret_cast<__nv_bfloat16*>(b_hi + (K / 4) * N * 4); // [K/8][ROWS][8] bf16(a_hi)
    __nv_bfloat16* a_lb = a_hb + (K / 8) * ROWS * 8;                                // bf16(a_lo)
    __nv_bfloat16* b_hb = a_lb + (K / 8) * ROWS * 8;                                // [K/8][N][8]
    __nv_bfloat16* b_lb = b_hb + (K / 8) * N * 8;
    __half* a_hf = reinterpret_cast<__half*>(b_lb + (K / 8) * N * 8);               // [K/8][ROWS][8] fp16(a) (RNE)
    __nv_bfloat16* a_lf = reinterpret_cast<__nv_bfloat16*>(a_hf + (K / 8) * ROWS * 8); // bf16(a - fp16(a))
    __half* b_hf = reinterpret_cast<__half*>(a_lf + (K / 8) * ROWS * 8);
    __nv_bfloat16* b_lf = reinterpret_cast<__nv_bfloat16*>(b_hf + (K / 8) * N * 8);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < ROWS * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        const float v = A[i], h = tf32_hi(v);
        a_hi[((k / 4) * ROWS + r) * 4 + (k % 4)] = h;
        const __half hf = __float2half_rn(v);
        a_hb[((k / 8) * ROWS + r) * 8 + (k % 8)] = __float2bfloat16_rn(mode == 5 ? __half2float(hf) : h);
        a_lb[((k / 8) * ROWS + r) * 8 + (k % 8)] = __float2bfloat16_rn(v - h);
        a_hf[((k / 8) * ROWS + r) * 8 + (k % 8)] = hf;
        a_lf[((k / 8) * ROWS + r) * 8 + (k % 8)] = __float2bfloat16_rn(v - __half2float(hf));
    }
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        const float v = B[i], h = tf32_hi(v);
        b_hi[((k / 4) * N + n) * 4 + (k % 4)] = h;
        const __half hf = __float2half_rn(v);
        b_hb[((k / 8) * N + n) * 8 + (k % 8)] = __float2bfloat16_rn(mode == 5 ? __half2float(hf) : h);
        b_lb[((k / 8) * N + n) * 8 + (k % 8)] = __float2bfloat16_rn(v - h);
        b_hf[((k / 8) * N + n) * 8 + (k % 8)] = hf;
        b_lf[((k / 8) * N + n) * 8 + (k % 8)] = __float2bfloat16_rn(v - __half2float(hf));
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&tmem_slot, N < 32 ? 32 : N);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        uint32_t acc = 0;
        if (mode == 3) {                                   // all 16-bit: fp16 x fp16 main, mixed bf16 x fp16 / fp16 x bf16 corrections
            for (int s = 0; s < K / 16; ++s) {
                const uint32_t a_off = (uint32_t)(shift * 16 + 2 * s * ROWS * 16), b_off = (uint32_t)(2 * s * N * 16);
                mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_hf) + a_off, ROWS * 16, 128),
                        smem_desc_kmajor_noswizzle(smem_u32(b_hf) + b_off, N * 16, 128), idesc_16(M, N, 0, 0), acc);
                acc = 1;
                mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_lf) + a_off, ROWS * 16, 128),
                        smem_desc_kmajor_noswizzle(smem_u32(b_hf) + b_off, N * 16, 128), idesc_16(M, N, 1, 0), 1);
                mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_hf) + a_off, ROWS * 16, 128),
                        smem_desc_kmajor_noswizzle(smem_u32(b_lf) + b_off, N * 16, 128), idesc_16(M, N, 0, 1), 1);
            }
        }
        if (mode == 5) {                                   // fp16 x fp16 main, bf16 x bf16 corrections (bf16 copies of the fp16 hi parts)
            for (int s = 0; s < K / 16; ++s) {
                const uint32_t a_off = (uint32_t)(shift * 16 + 2 * s * ROWS * 16), b_off = (uint32_t)(2 * s * N * 16);
                mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_hf) + a_off, ROWS * 16, 128),
                        smem_desc_kmajor_noswizzle(smem_u32(b_hf) + b_off, N * 16, 128), idesc_16(M, N, 0, 0), acc);
                acc = 1;
                mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_lf) + a_off, ROWS * 16, 128),
                        smem_desc_kmajor_noswizzle(smem_u32(b_hb) + b_off, N * 16, 128), idesc_16(M, N, 1, 1), 1);
                mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_hb) + a_off, ROWS * 16, 128),
                        smem_desc_kmajor_noswizzle(smem_u32(b_lf) + b_off, N * 16, 128), idesc_16(M, N, 1, 1), 1);
            }
        }
        if (mode != 2 && mode != 3 && mode != 5) {         // main pass, TF32
            for (int s = 0; s < K / 8; ++s) {
                const uint64_t ad = smem_desc_kmajor_noswizzle(smem_u32(a_hi) + (uint32_t)(shift * 16 + 2 * s * ROWS * 16), ROWS * 16, 128);
                const uint64_t bd = smem_desc_kmajor_noswizzle(smem_u32(b_hi) + (uint32_t)(2 * s * N * 16), N * 16, 128);
                mma_tf32(tmem, ad, bd, idesc_tf32(M, N), acc);
                acc = 1;
            }
        }
        if (mode != 0 && mode != 3 && mode != 5) {         // corrections, BF16 (mode 2: ONLY a_hb x b_hb, to test the bf16 path alone)
            for (int s = 0; s < K / 16; ++s) {
                const uint32_t a_off = (uint32_t)(shift * 16 + 2 * s * ROWS * 16), b_off = (uint32_t)(2 * s * N * 16);
                if (mode == 2) {
                    mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_hb) + a_off, ROWS * 16, 128),
                            smem_desc_kmajor_noswizzle(smem_u32(b_hb) + b_off, N * 16, 128), idesc_bf16(M, N), acc);
                    acc = 1;
                } else {
                    mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_lb) + a_off, ROWS * 16, 128),
                            smem_desc_kmajor_noswizzle(smem_u32(b_hb) + b_off, N * 16, 128), idesc_bf16(M, N), 1);
                    mma_f16(tmem, smem_desc_kmajor_noswizzle(smem_u32(a_hb) + a_off, ROWS * 16, 128),
                            smem_desc_kmajor_noswizzle(smem_u32(b_lb) + b_off, N * 16, 128), idesc_bf16(M, N), 1);
                }
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_thread_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32; ++j) D[(warp * 32 + (tid & 31)) * N + c0 + j] = v[j];
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, N < 32 ? 32 : N);
}

static float bf16r(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7fffu + ((u >> 16) & 1u); u &= 0xffff0000u; float r; memcpy(&r, &u, 4); return r; }

template <int N>
int run(int shift, int mode, float scale) {
    std::vector<float> A(ROWS * K), B(N * K), D(M * N);
    srand(99 + shift + mode);
    for (auto& v : A) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * scale;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    const size_t smem = (size_t)((K / 4) * ROWS * 4 + (K / 4) * N * 4) * 4 + (size_t)(4 * (K / 8) * ROWS * 8 + 4 * (K / 8) * N * 8) * 2;
    cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, shift, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d shift=%d mode=%d: CUDA error %s\n", N, shift, mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double e_exact = 0, e_bf = 0, ref_max = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ex = 0, bf = 0;
            for (int k = 0; k < K; ++k) {
                const float a = A[(m + shift) * K + k], b = B[n * K + k];
                ex += (double)a * b;
                bf += (double)bf16r(tf32_hi(a)) * bf16r(tf32_hi(b));
            }
            const double d = D[m * N + n];
            e_exact = fmax(e_exact, fabs(d - ex)); e_bf = fmax(e_bf, fabs(d - bf)); ref_max = fmax(ref_max, fabs(ex));
        }
    printf("N=%3d shift=%2d mode=%d scale=%g: max|D-exact|=%.3e  max|D-bf16ref|=%.3e  (max|ref|=%.2f)\n", N, shift, mode, scale, e_exact, e_bf, ref_max);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    if (mode == 2) return e_bf < 1e-4 * scale ? 0 : 1;
    if (mode == 1 || mode == 3 || mode == 5) return e_exact < 2e-5 * scale ? 0 : 1;
    return 0;
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 1;     // a failing MMA poisons the context: one family per process
    int bad = 0;
    if (which == 1) {
        bad += run<64>(0, 2, 1.f);        // bf16 path alone: descriptor + layout check against a bf16 reference
        bad += run<64>(7, 2, 1.f);
        bad += run<128>(21, 2, 1.f);
        bad += run<64>(0, 0, 1.f);        // tf32 main pass alone (error ~1e-3, for scale)
        bad += run<64>(5, 1, 1.f);        // tf32 main + bf16 corrections (what conv_tc.cuh does)
        bad += run<128>(47, 1, 1.f);
        bad += run<128>(3, 1, 8.f);
    } else if (which == 5) {              // fp16 main + bf16 corrections, same-format MMAs only
        bad += run<64>(5, 5, 1.f);
        bad += run<128>(47, 5, 1.f);
        bad += run<128>(3, 5, 8.f);
        bad += run<128>(9, 5, 0.01f);
    } else {                              // mixed a/b formats inside one MMA
        bad += run<64>(5, 3, 1.f);
        bad += run<128>(47, 3, 1.f);
    }
    printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
    return bad;
}
