// TEST INFRASTRUCTURE ONLY (tests/emu): CUDA-core formulations of the CNN layers for the host-emulation build.
// The product runs conv2 / conv3 / FC1 on tcgen05 (csrc/conv_tc.cuh, csrc/fc_tc.cuh), which the emulation cannot
// execute; these plain kernels stand in for them so that everything AROUND the tensor-core kernels can be
// exercised in a GPU-less container.  Included by csrc/infer.cuh only when GAT_CPU_EMU is defined.
#pragma once

namespace gat {

// ---------------------------------------------------------------------------------------------------
// conv1: C_in = 1.  in [N][H][W] (mel-dB image), out padded NHWC [N][H/2+2][W/2+2][C] with zero border.
struct Conv1Params {
    const float* in; int N, H, W;
    const float* w;      // [9][C]
    const float* bias;   // [C]
    float* out; int C;   // C <= 64
    float slope;
};

__global__ void __launch_bounds__(256) conv1_pool_kernel(Conv1Params p) {
    __shared__ float ws[9 * 64];
    __shared__ float bs[64];
    for (int i = threadIdx.x; i < 9 * p.C; i += blockDim.x) ws[i] = p.w[i];
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) bs[i] = p.bias[i];
    __syncthreads();
    const int Hp = p.H / 2, Wp = p.W / 2;
    const int tiles = ceil_div(Hp * Wp, (int)blockDim.x);
    const int clip = blockIdx.x / tiles;
    const int q = (blockIdx.x - clip * tiles) * blockDim.x + threadIdx.x;     // pooled pixel
    if (q >= Hp * Wp) return;
    const int py = q / Wp, px = q - py * Wp;
    const float* img = p.in + (long long)clip * p.H * p.W;
    float patch[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int y = 2 * py - 1 + a, x = 2 * px - 1 + b;
            patch[a][b] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? img[y * p.W + x] : 0.0f;
        }
    float* o = p.out + (((long long)clip * (Hp + 2) + py + 1) * (Wp + 2) + px + 1) * p.C;
    for (int c = 0; c < p.C; ++c) {
        float best = -3.0e38f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float acc = 0.0f;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc = fmaf(patch[a + ky][b + kx], ws[(ky * 3 + kx) * p.C + c], acc);
                best = fmaxf(best, acc);
            }
        o[c] = leaky(best + bs[c], p.slope);
    }
}

// ---------------------------------------------------------------------------------------------------
// conv2 / conv3 on CUDA cores: in padded NHWC [N][H+2][W+2][CIN], weights [9][CIN][COUT], fused
// bias + LeakyReLU + MaxPool2.  A CTA of 256 threads owns kPix pooled pixels x all COUT channels; a thread
// owns one pooled pixel (its four conv positions) and eight output channels.
struct ConvParams {
    const float* in; int N, H, W;      // H, W: un-padded input size
    const float* w; const float* bias;
    float* out; int out_pad;           // 1: write padded NHWC [H/2+2][W/2+2][COUT]; 0: dense [H/2][W/2][COUT]
    float slope;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) conv3x3_pool_kernel(ConvParams p) {
    constexpr int kGroups = COUT / 8;          // channel groups per pixel
    constexpr int kPix = 256 / kGroups;        // pooled pixels per CTA
    __shared__ __align__(16) float ws[CIN * COUT];   // one tap of weights
    const int Hp = p.H / 2, Wp = p.W / 2;
    const int tiles = ceil_div(Hp * Wp, kPix);
    const int clip = blockIdx.x / tiles;
    const int g = threadIdx.x % kGroups;
    const int q = (blockIdx.x - clip * tiles) * kPix + threadIdx.x / kGroups;
    const bool active = q < Hp * Wp;
    const int py = active ? q / Wp : 0, px = active ? q - py * Wp : 0;
    const int Wpad = p.W + 2;
    const float* img = p.in + (long long)clip * (p.H + 2) * Wpad * CIN;
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[a][c] = 0.0f;

    for (int tap = 0; tap < 9; ++tap) {
        __syncthreads();
        for (int i = threadIdx.x; i < CIN * COUT / 4; i += blockDim.x)
            reinterpret_cast<float4*>(ws)[i] = reinterpret_cast<const float4*>(p.w + (long long)tap * CIN * COUT)[i];
        __syncthreads();
        const int ky = tap / 3, kx = tap - ky * 3;
        // conv position (2py+a, 2px+b) reads padded pixel (2py+a+ky, 2px+b+kx)
        const float* base = img + ((long long)(2 * py + ky) * Wpad + 2 * px + kx) * CIN;
        for (int ci = 0; ci < CIN; ci += 4) {
            float4 x[4];
            x[0] = *reinterpret_cast<const float4*>(base + ci);
            x[1] = *reinterpret_cast<const float4*>(base + CIN + ci);
            x[2] = *reinterpret_cast<const float4*>(base + (long long)Wpad * CIN + ci);
            x[3] = *reinterpret_cast<const float4*>(base + (long long)Wpad * CIN + CIN + ci);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 w0 = *reinterpret_cast<const float4*>(ws + (ci + u) * COUT + g * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(ws + (ci + u) * COUT + g * 8 + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float xv = u == 0 ? x[a].x : u == 1 ? x[a].y : u == 2 ? x[a].z : x[a].w;
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(xv, wv[c], acc[a][c]);
                }
            }
        }
    }
    if (!active) return;
    float* o = p.out_pad ? p.out + (((long long)clip * (Hp + 2) + py + 1) * (Wp + 2) + px + 1) * COUT
                         : p.out + (((long long)clip * Hp + py) * Wp + px) * COUT;
    float r[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float m = fmaxf(fmaxf(acc[0][c], acc[1][c]), fmaxf(acc[2][c], acc[3][c]));
        r[c] = leaky(m + p.bias[g * 8 + c], p.slope);
    }
    *reinterpret_cast<float4*>(o + g * 8) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(o + g * 8 + 4) = make_float4(r[4], r[5], r[6], r[7]);
}


// ---------------------------------------------------------------------------------------------------
// AdaptiveAvgPool2d((4,4)) + Flatten (C-major) + Linear + LeakyReLU + Linear + softmax.
// A CTA of 256 threads handles kHeadClips clips so the 2 MB of FC1 weights are streamed once per group.
constexpr int kHeadClips = 16;

struct HeadParams {
    const float* act;      // [N][H][W][C] dense NHWC (after the last pool)
    int N, H, W, C;        // C = 128
    const float* w1;       // [C*16][hidden] (input index c*16 + i*4 + j as torch Flatten gives it)
    const float* b1;       // [hidden]
    int hidden;            // 256
    const float* w2;       // [hidden][classes]
    const float* b2;
    int classes;           // <= 64
    float slope;
    float* logits;         // [N][classes]
    float* probs;          // [N][classes]
};

__global__ void __launch_bounds__(256) cnn_head_kernel(HeadParams p) {
    GAT_DYN_SMEM(smem_raw);
    const int F = p.C * 16;
    float* feat = reinterpret_cast<float*>(smem_raw);              // [F][kHeadClips]
    float* hid = feat + (size_t)F * kHeadClips;                    // [hidden][kHeadClips]
    float* lg = hid + (size_t)p.hidden * kHeadClips;               // [kHeadClips][64]
    const int clip0 = blockIdx.x * kHeadClips;
    const int nc = min(kHeadClips, p.N - clip0);

    // adaptive average pool: window [floor(i*L/4), ceil((i+1)*L/4))
    for (int idx = threadIdx.x; idx < F * kHeadClips; idx += blockDim.x) {
        const int cl = idx / F, f = idx - cl * F;
        const int i = f / (4 * p.C), j = (f / p.C) & 3, c = f % p.C;       // iterate c fastest: coalesced NHWC reads
        float v = 0.0f;
        if (cl < nc) {
            const int y0 = (i * p.H) / 4, y1 = ((i + 1) * p.H + 3) / 4;
            const int x0 = (j * p.W) / 4, x1 = ((j + 1) * p.W + 3) / 4;
            const float* a = p.act + (long long)(clip0 + cl) * p.H * p.W * p.C;
            float s = 0.0f;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) s += a[((long long)y * p.W + x) * p.C + c];
            v = s / (float)((y1 - y0) * (x1 - x0));
        }
        feat[(size_t)(c * 16 + i * 4 + j) * kHeadClips + cl] = v;
    }
    __syncthreads();
    // FC1: thread o owns hidden unit o for all clips of the group
    for (int o = threadIdx.x; o < p.hidden; o += blockDim.x) {
        float acc[kHeadClips];
#pragma unroll
        for (int c = 0; c < kHeadClips; ++c) acc[c] = 0.0f;
        for (int k = 0; k < F; ++k) {
            const float w = p.w1[(long long)k * p.hidden + o];
            const float4* fv = reinterpret_cast<const float4*>(feat + (size_t)k * kHeadClips);
#pragma unroll
            for (int c4 = 0; c4 < kHeadClips / 4; ++c4) {
                const float4 x = fv[c4];
                acc[4 * c4 + 0] = fmaf(x.x, w, acc[4 * c4 + 0]);
                acc[4 * c4 + 1] = fmaf(x.y, w, acc[4 * c4 + 1]);
                acc[4 * c4 + 2] = fmaf(x.z, w, acc[4 * c4 + 2]);
                acc[4 * c4 + 3] = fmaf(x.w, w, acc[4 * c4 + 3]);
            }
        }
        const float b = p.b1[o];
#pragma unroll
        for (int c = 0; c < kHeadClips; ++c) hid[(size_t)o * kHeadClips + c] = leaky(acc[c] + b, p.slope);
    }
    __syncthreads();
    // FC2: one (clip, class) per thread-iteration
    for (int idx = threadIdx.x; idx < kHeadClips * p.classes; idx += blockDim.x) {
        const int cl = idx / p.classes, o = idx - cl * p.classes;
        float acc = 0.0f;
        for (int k = 0; k < p.hidden; ++k) acc = fmaf(hid[(size_t)k * kHeadClips + cl], p.w2[(long long)k * p.classes + o], acc);
        lg[cl * 64 + o] = acc + p.b2[o];
    }
    __syncthreads();
    // softmax per clip: one warp per clip
    for (int cl = warp_id(); cl < nc; cl += (int)(blockDim.x >> 5)) {
        const int lane = lane_id();
        const float v0 = lane < p.classes ? lg[cl * 64 + lane] : -3.0e38f;
        const float v1 = lane + 32 < p.classes ? lg[cl * 64 + lane + 32] : -3.0e38f;
        const float mx = warp_max(fmaxf(v0, v1));
        const float e0 = lane < p.classes ? expf(v0 - mx) : 0.0f;
        const float e1 = lane + 32 < p.classes ? expf(v1 - mx) : 0.0f;
        const float sum = warp_sum(e0 + e1);
        const long long o = (long long)(clip0 + cl) * p.classes;
        if (lane < p.classes) { p.logits[o + lane] = v0; p.probs[o + lane] = e0 / sum; }
        if (lane + 32 < p.classes) { p.logits[o + lane + 32] = v1; p.probs[o + lane + 32] = e1 / sum; }
    }
}

}  // namespace gat
