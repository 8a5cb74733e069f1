// GPU probe (test infrastructure): validates the tcgen05 primitives in csrc/tc05.cuh on a B200 before the
// conv kernel relies on them: no-swizzle K-major descriptors with an arbitrary 16-byte row shift, the tf32
// instruction descriptor, the TMEM lane/column mapping of tcgen05.ld, input truncation, 3xTF32 accuracy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I guitar_audio_transcriber_ai_b200/csrc tests/gpu_probe/tc_probe.cu -o tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "tc05.cuh"
using namespace gat::tc;

constexpr int M = 128, K = 32, ROWS = 176;   // A buffer holds ROWS rows; the MMA reads rows [shift, shift+128)

template <int N>
__global__ void probe_kernel(const float* __restrict__ A /*[ROWS][K]*/, const float* __restrict__ B /*[N][K]*/,
                             float* __restrict__ D /*[M][N]*/, int shift, int passes) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* a_hi = reinterpret_cast<float*>(smem);                 // [K/4 chunks][ROWS][4]
    float* a_lo = a_hi + (K / 4) * ROWS * 4;
    float* b_hi = a_lo + (K / 4) * ROWS * 4;                      // [K/4 chunks][N][4]
    float* b_lo = b_hi + (K / 4) * N * 4;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < ROWS * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        const float v = A[i], h = passes > 1 ? tf32_hi(v) : v;
        a_hi[((k / 4) * ROWS + r) * 4 + (k % 4)] = h;
        a_lo[((k / 4) * ROWS + r) * 4 + (k % 4)] = v - tf32_hi(v);
    }
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        const float v = B[i], h = passes > 1 ? tf32_hi(v) : v;
        b_hi[((k / 4) * N + n) * 4 + (k % 4)] = h;
        b_lo[((k / 4) * N + n) * 4 + (k % 4)] = v - tf32_hi(v);
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
    if (warp == 0) tmem_alloc(&tmem_slot, N < 32 ? 32 : N);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(M, N);
        uint32_t acc = 0;
        for (int s = 0; s < K / 8; ++s) {
            for (int p = 0; p < passes; ++p) {
                const float* ab = (p == 1) ? a_lo : a_hi;
                const float* bb = (p == 2) ? b_lo : b_hi;
                const uint64_t ad = smem_desc_kmajor_noswizzle(smem_u32(ab) + (uint32_t)(shift * 16 + 2 * s * ROWS * 16), ROWS * 16, 128);
                const uint64_t bd = smem_desc_kmajor_noswizzle(smem_u32(bb) + (uint32_t)(2 * s * N * 16), N * 16, 128);
                mma_tf32(tmem, ad, bd, idesc, acc);
                acc = 1;
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

static float trunc_tf32(float x) { return tf32_hi(x); }
static float rne_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0xfffu + ((u >> 13) & 1u); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

template <int N>
int run(int shift, int passes) {
    std::vector<float> A(ROWS * K), B(N * K), D(M * N);
    srand(1234 + shift + passes);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    const size_t smem = (size_t)(2 * (K / 4) * ROWS * 4 + 2 * (K / 4) * N * 4) * 4;
    cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, shift, passes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d shift=%d passes=%d: CUDA error %s\n", N, shift, passes, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double e_exact = 0, e_trunc = 0, e_rne = 0, ref_max = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ex = 0, tr = 0, rn = 0;
            for (int k = 0; k < K; ++k) {
                const float a = A[(m + shift) * K + k], b = B[n * K + k];
                ex += (double)a * b; tr += (double)trunc_tf32(a) * trunc_tf32(b); rn += (double)rne_tf32(a) * rne_tf32(b);
            }
            const double d = D[m * N + n];
            e_exact = fmax(e_exact, fabs(d - ex)); e_trunc = fmax(e_trunc, fabs(d - tr)); e_rne = fmax(e_rne, fabs(d - rn));
            ref_max = fmax(ref_max, fabs(ex));
        }
    printf("N=%3d shift=%2d passes=%d: max|D-exact|=%.3e  max|D-trunc_tf32|=%.3e  max|D-rne_tf32|=%.3e  (max|ref|=%.2f)\n",
           N, shift, passes, e_exact, e_trunc, e_rne, ref_max);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    const double tol = passes == 3 ? 2e-5 : 2e-2;
    return e_exact < tol ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run<64>(0, 1);
    bad += run<64>(5, 1);
    bad += run<128>(13, 1);
    bad += run<64>(5, 3);
    bad += run<128>(47, 3);
    printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
    return bad;
}
