// Classifier head of the CNN (training/cnn_trainer.py:105-131): AdaptiveAvgPool2d((4,4)) -> Flatten ->
// Linear(2048, 256) -> LeakyReLU -> Linear(256, classes) -> softmax (note_predictor.py:106-107).
//
// FC1 is a true dense contraction ([clips x 2048] . [2048 x 256]) and runs on the tensor cores with the same
// operand conventions as conv_tc.cuh: TF32 hi/lo "chunk planes" [k/4][row][4 floats], no-swizzle K-major
// descriptors, 3xTF32 accumulation in TMEM, bulk-TMA loads, warp-specialised producer / issuer / epilogue.
// The pooling that precedes it writes those planes directly; FC2 + softmax (12 kFLOP per clip) stay on CUDA cores.
#pragma once
#ifndef GAT_CPU_EMU
#include "common.cuh"
#include "tc05.cuh"

namespace gat {

// ---- AdaptiveAvgPool2d((4,4)) over act3 [N][H][W][C] -> FC1 operand planes.
// K index of FC1 = c*16 + i*4 + j (torch Flatten of [C][4][4]); chunk = c*4 + i holds j = 0..3.
struct AvgPoolPlanesParams {
    const float* act; int N, H, W, C;
    float* out_hi; float* out_lo;     // [C*4 chunks][rows_pad][4]
    long long rows_pad;
};

__global__ void __launch_bounds__(256) avgpool_planes_kernel(AvgPoolPlanesParams p) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;    // (clip, i, c) with c fastest
    const long long total = (long long)p.N * 4 * p.C;
    if (idx >= total) return;
    const int c = (int)(idx % p.C);
    const int i = (int)((idx / p.C) & 3);
    const long long clip = idx / (4 * p.C);
    const int y0 = (i * p.H) / 4, y1 = ((i + 1) * p.H + 3) / 4;
    const float* a = p.act + clip * p.H * p.W * p.C + c;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x0 = (j * p.W) / 4, x1 = ((j + 1) * p.W + 3) / 4;
        float s = 0.0f;
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) s += a[((long long)y * p.W + x) * p.C];
        o[j] = s / (float)((y1 - y0) * (x1 - x0));
    }
    const long long off = (((long long)(c * 4 + i)) * p.rows_pad + clip) * 4;
    const float4 hi = make_float4(tc::tf32_hi(o[0]), tc::tf32_hi(o[1]), tc::tf32_hi(o[2]), tc::tf32_hi(o[3]));
    *reinterpret_cast<float4*>(p.out_hi + off) = hi;
    *reinterpret_cast<float4*>(p.out_lo + off) = make_float4(o[0] - hi.x, o[1] - hi.y, o[2] - hi.z, o[3] - hi.w);
}

// ---- FC1 on tcgen05.  CTA (x, y) = 128 rows (clips) x one of `k_splits` slices of K, streamed in blocks of 32; the
// slices' partial sums are added, biased and activated by fc2_softmax_kernel (split-K keeps every SM busy when there
// are few clips: a single note used to wait 72 us for one CTA walking all of K).
struct FcTcParams {
    const float* a_hi; const float* a_lo;   // [K/4][rows_pad][4]
    long long rows_pad;                     // multiple of 128
    const float* w;                         // [K/32][hi|lo][8 chunks][NOUT][4]
    int n_rows, K;
    int k_splits;                           // gridDim.y; K/32 must be divisible by it
    float* out;                             // [k_splits][n_rows][NOUT] partial sums (no bias, no activation)
};

constexpr int kFcStages = 2;

template <int NOUT>
__host__ __device__ inline size_t fc_tc_smem_bytes() {
    return (size_t)kFcStages * (2 * 8 * 128 * 16 + 2 * 8 * NOUT * 16) + 256;
}

template <int NOUT>
__global__ void __launch_bounds__(192, 1) fc_tc_kernel(FcTcParams p) {
    using namespace tc;
    constexpr uint32_t A_STAGE = 2 * 8 * 128 * 16;         // hi|lo x 8 chunks x 128 rows x 16 B
    constexpr uint32_t W_STAGE = 2 * 8 * NOUT * 16;
    constexpr uint32_t STAGE = A_STAGE + W_STAGE;
    constexpr uint32_t TMEM_COLS = NOUT <= 32 ? 32 : NOUT <= 64 ? 64 : NOUT <= 128 ? 128 : NOUT <= 256 ? 256 : 512;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kFcStages * STAGE);
    uint64_t* full = bars; uint64_t* empty = bars + kFcStages; uint64_t* acc_full = bars + 2 * kFcStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kFcStages + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kFcStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_slot;
    const int nkb = p.K / 32 / p.k_splits;                  // K blocks of this slice
    const int kb0 = blockIdx.y * nkb;
    const long long row0 = (long long)blockIdx.x * 128;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t st = kb % kFcStages;
                mbar_wait(empty + st, ((kb / kFcStages) & 1) ^ 1);
                mbar_expect_tx(full + st, STAGE);
                unsigned char* dst = smem + (size_t)st * STAGE;
                for (int part = 0; part < 2; ++part) {
                    const float* src = part ? p.a_lo : p.a_hi;
                    for (int c = 0; c < 8; ++c)
                        bulk_g2s(dst + (size_t)(part * 8 + c) * 2048, src + (((long long)(kb0 + kb) * 8 + c) * p.rows_pad + row0) * 4, 2048, full + st);
                }
                bulk_g2s(dst + A_STAGE, p.w + (size_t)(kb0 + kb) * (W_STAGE / 4), W_STAGE, full + st);
            }
        }
    } else if (warp == 1) {
        // the warp runs the loop, one elected lane issues (see tc::elect_one)
        const bool leader = elect_one();
        const uint32_t idesc = idesc_tf32(128, NOUT);
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t st = kb % kFcStages;
            mbar_wait(full + st, (kb / kFcStages) & 1);
            fence_after_thread_sync();
            const uint32_t a_hi = smem_u32(smem) + st * STAGE, a_lo = a_hi + 8 * 2048;
            const uint32_t w_hi = a_hi + A_STAGE, w_lo = w_hi + 8 * NOUT * 16;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const uint64_t dah = smem_desc_kmajor_noswizzle(a_hi + 2 * s * 2048, 2048, 128);
                const uint64_t dal = smem_desc_kmajor_noswizzle(a_lo + 2 * s * 2048, 2048, 128);
                const uint64_t dbh = smem_desc_kmajor_noswizzle(w_hi + 2 * s * NOUT * 16, NOUT * 16, 128);
                const uint64_t dbl = smem_desc_kmajor_noswizzle(w_lo + 2 * s * NOUT * 16, NOUT * 16, 128);
                if (leader) {
                    mma_tf32(tmem, dah, dbh, idesc, (kb | s) != 0 ? 1u : 0u);
                    mma_tf32(tmem, dal, dbh, idesc, 1u);
                    mma_tf32(tmem, dah, dbl, idesc, 1u);
                }
            }
            if (leader) mma_commit(empty + st);
            __syncwarp();
        }
        if (leader) mma_commit(acc_full);
    } else {
        const int quarter = warp & 3;
        const long long row = row0 + quarter * 32 + lane;
        mbar_wait(acc_full, 0);
        fence_after_thread_sync();
        for (int cb = 0; cb < NOUT / 32; ++cb) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32), v);
            if (row < p.n_rows) {
                float* o = p.out + ((long long)blockIdx.y * p.n_rows + row) * NOUT + cb * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ---- FC1 finish (sum of the K slices + bias + LeakyReLU) -> FC2 + softmax: one warp per clip.
struct Fc2Params {
    const float* hid; int N, hidden;     // [k_splits][N][hidden] partial sums of FC1
    int k_splits;
    const float* b1; float slope;        // FC1 bias [hidden], LeakyReLU slope
    const float* w2; const float* b2;    // [hidden][classes], [classes]
    int classes;                         // <= 64
    float* logits; float* probs;         // [N][classes]
};

__global__ void __launch_bounds__(256) fc2_softmax_kernel(Fc2Params p) {
    GAT_DYN_SMEM(smem_raw);
    float* w = reinterpret_cast<float*>(smem_raw);                 // [hidden][classes]
    float* hbuf = w + ((p.hidden * p.classes + 3) & ~3);           // [warps][hidden] activated hidden vectors (16-byte aligned)
    // FC2's weights: 128-bit loads, four in flight per thread (one 4-byte load at a time, 47 dependent round trips to L2
    // per CTA, was most of this kernel's 50 us)
    {
        const int n4 = (p.hidden * p.classes) >> 2;
        const float4* src = reinterpret_cast<const float4*>(p.w2);
        float4* dst = reinterpret_cast<float4*>(w);
        for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < n4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; if (i < n4) dst[i] = v[u]; }
        }
        for (int i = 4 * n4 + threadIdx.x; i < p.hidden * p.classes; i += blockDim.x) w[i] = p.w2[i];
    }
    __syncthreads();
    const int lane = lane_id(), nwarps = blockDim.x >> 5;
    float* h = hbuf + warp_id() * p.hidden;
    for (int clip = blockIdx.x * nwarps + warp_id(); clip < p.N; clip += gridDim.x * nwarps) {
        // hidden unit k = lane + 32 j: the K slices of FC1 are added in slice order (a clip's value does not depend on the
        // batch); the loads of four units (up to 32) are issued together
        for (int k0 = lane; k0 < p.hidden; k0 += 128) {
            float part[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int k = k0 + 32 * j;
                    part[j][s] = (s < p.k_splits && k < p.hidden) ? __ldg(p.hid + ((long long)s * p.N + clip) * p.hidden + k) : 0.0f;
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + 32 * j;
                if (k < p.hidden) {
                    float acc = 0.0f;
#pragma unroll
                    for (int s = 0; s < 8; ++s) acc += part[j][s];
                    for (int s = 8; s < p.k_splits; ++s) acc += __ldg(p.hid + ((long long)s * p.N + clip) * p.hidden + k);
                    const float z = acc + __ldg(p.b1 + k);
                    h[k] = z > 0.0f ? z : z * p.slope;
                }
            }
        }
        __syncwarp();
        // logits: four interleaved partial sums per class (k mod 4), folded as (s0 + s1) + (s2 + s3): a fixed order, and a
        // dependent chain of hidden / 4 FMAs instead of hidden
        float a0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, a1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const bool has0 = lane < p.classes, has1 = lane + 32 < p.classes;
        const int c0 = has0 ? lane : 0, c1 = has1 ? lane + 32 : 0;          // clamped: inactive lanes read a valid weight
        int k = 0;
        for (; k + 4 <= p.hidden; k += 4) {
            const float4 x = *reinterpret_cast<const float4*>(h + k);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float xu = u == 0 ? x.x : u == 1 ? x.y : u == 2 ? x.z : x.w;
                a0[u] = fmaf(xu, w[(k + u) * p.classes + c0], a0[u]);
                a1[u] = fmaf(xu, w[(k + u) * p.classes + c1], a1[u]);
            }
        }
        for (; k < p.hidden; ++k) {
            a0[k & 3] = fmaf(h[k], w[k * p.classes + c0], a0[k & 3]);
            a1[k & 3] = fmaf(h[k], w[k * p.classes + c1], a1[k & 3]);
        }
        __syncwarp();
        const float v0 = has0 ? ((a0[0] + a0[1]) + (a0[2] + a0[3])) + p.b2[lane] : -3.0e38f;
        const float v1 = has1 ? ((a1[0] + a1[1]) + (a1[2] + a1[3])) + p.b2[lane + 32] : -3.0e38f;
        const float mx = warp_max(fmaxf(v0, v1));
        const float e0 = has0 ? expf(v0 - mx) : 0.0f;
        const float e1 = has1 ? expf(v1 - mx) : 0.0f;
        const float sum = warp_sum(e0 + e1);
        const long long o = (long long)clip * p.classes;
        if (has0) { p.logits[o + lane] = v0; p.probs[o + lane] = e0 / sum; }
        if (has1) { p.logits[o + lane + 32] = v1; p.probs[o + lane + 32] = e1 / sum; }
    }
}

}  // namespace gat
#endif  // GAT_CPU_EMU
