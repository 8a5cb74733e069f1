// CNN + MLP forward passes and the ensemble (reference: note_predictor.py:84-135,
// training/cnn_trainer.py:30-139, training/mlp_trainer.py:32-105).
//
// The CNN layers live in conv_tc.cuh / fc_tc.cuh (tcgen05); this file holds the MLP + ensemble kernel.
// Eval-mode BatchNorm is folded into the conv weights on the host (checkpoint.pack_cnn).
#pragma once
#include "common.cuh"

namespace gat {

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.0f ? v : v * slope; }

}  // namespace gat
#ifdef GAT_CPU_EMU
#include "emu_cnn.cuh"   // tests/emu: CUDA-core stand-ins for the tensor-core layers (host-emulation build only)
#endif
namespace gat {


// ---------------------------------------------------------------------------------------------------
// MLP forward (Linear -> LayerNorm -> LeakyReLU(0.1))* -> Linear -> softmax, then the ensemble
// probs = w_mlp * mlp + w_cnn * cnn (two float32 products, one float32 add, no FMA: note_predictor.py:112),
// first-max argmax and confidence.  One warp per clip; weights stay in shared memory.
constexpr int kMlpMaxLayers = 6;
constexpr int kMlpMaxWidth = 256;

struct MlpParams {
    const float* x; int N, ld;         // [N][ld] features (first dims[0] columns used)
    const float* params;               // packed: per layer W^T[in][out], b[out], (gamma[out], beta[out])
    int n_params;
    int n_linear;                      // number of Linear layers
    int dims[kMlpMaxLayers + 1];
    float slope;                       // 0.1
    float ln_eps;                      // 1e-5
    const float* cnn_probs;            // [N][classes] or nullptr
    float w_mlp, w_cnn;
    float* mlp_logits;                 // [N][classes]
    float* mlp_probs;                  // [N][classes]
    float* probs;                      // [N][classes]
    long long* index;                  // [N]
    float* conf;                       // [N]
};

__global__ void __launch_bounds__(256) mlp_ensemble_kernel(MlpParams p) {
    GAT_DYN_SMEM(smem_raw);
    float* wsm = reinterpret_cast<float*>(smem_raw);                         // all parameters
    float* scratch = wsm + p.n_params;                                       // per warp 2 * kMlpMaxWidth
    {   // parameters: 128-bit loads, four in flight per thread (80 dependent 4-byte round trips to L2 per CTA before)
        const int n4 = p.n_params >> 2;
        const float4* src = reinterpret_cast<const float4*>(p.params);
        float4* dst = reinterpret_cast<float4*>(wsm);
        for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; v[u] = i < n4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; if (i < n4) dst[i] = v[u]; }
        }
        for (int i = 4 * n4 + threadIdx.x; i < p.n_params; i += blockDim.x) wsm[i] = p.params[i];
    }
    __syncthreads();
    const int lane = lane_id(), warp = warp_id(), nwarps = blockDim.x >> 5;
    float* bufa = scratch + (size_t)warp * 2 * kMlpMaxWidth;
    float* bufb = bufa + kMlpMaxWidth;
    const int classes = p.dims[p.n_linear];
    for (int clip = blockIdx.x * nwarps + warp; clip < p.N; clip += gridDim.x * nwarps) {
        for (int i = lane; i < p.dims[0]; i += 32) bufa[i] = p.x[(long long)clip * p.ld + i];
        __syncwarp();
        float* in = bufa; float* out = bufb;
        int off = 0;
        for (int l = 0; l < p.n_linear; ++l) {
            const int din = p.dims[l], dout = p.dims[l + 1];
            const float* W = wsm + off;            off += din * dout;
            const float* B = wsm + off;            off += dout;
            for (int o = lane; o < dout; o += 32) {
                float acc = 0.0f;
                for (int k = 0; k < din; ++k) acc = fmaf(in[k], W[k * dout + o], acc);
                out[o] = acc + B[o];
            }
            __syncwarp();
            if (l + 1 < p.n_linear) {
                const float* G = wsm + off;        off += dout;
                const float* Be = wsm + off;       off += dout;
                float s = 0.0f;
                for (int o = lane; o < dout; o += 32) s += out[o];
                const float mean = warp_sum(s) / (float)dout;
                float v = 0.0f;
                for (int o = lane; o < dout; o += 32) { const float d = out[o] - mean; v += d * d; }
                const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)dout + p.ln_eps);
                for (int o = lane; o < dout; o += 32) out[o] = leaky((out[o] - mean) * rstd * G[o] + Be[o], p.slope);
                __syncwarp();
            }
            float* t = in; in = out; out = t;
        }
        // `in` now holds the logits
        const float v0 = lane < classes ? in[lane] : -3.0e38f;
        const float v1 = lane + 32 < classes ? in[lane + 32] : -3.0e38f;
        const float mx = warp_max(fmaxf(v0, v1));
        const float e0 = lane < classes ? expf(v0 - mx) : 0.0f;
        const float e1 = lane + 32 < classes ? expf(v1 - mx) : 0.0f;
        const float sum = warp_sum(e0 + e1);
        const float m0 = e0 / sum, m1 = e1 / sum;
        const long long o = (long long)clip * classes;
        float q0 = m0, q1 = m1;
        if (p.cnn_probs) {
            const float c0 = lane < classes ? p.cnn_probs[o + lane] : 0.0f;
            const float c1 = lane + 32 < classes ? p.cnn_probs[o + lane + 32] : 0.0f;
            q0 = __fadd_rn(__fmul_rn(p.w_mlp, m0), __fmul_rn(p.w_cnn, c0));
            q1 = __fadd_rn(__fmul_rn(p.w_mlp, m1), __fmul_rn(p.w_cnn, c1));
        }
        if (lane < classes) { p.mlp_logits[o + lane] = v0; p.mlp_probs[o + lane] = m0; p.probs[o + lane] = q0; }
        if (lane + 32 < classes) { p.mlp_logits[o + lane + 32] = v1; p.mlp_probs[o + lane + 32] = m1; p.probs[o + lane + 32] = q1; }
        // np.argmax: first maximum
        float bv = lane < classes ? q0 : -1.0f; int bi = lane;
        if (lane + 32 < classes && q1 > bv) { bv = q1; bi = lane + 32; }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, s);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { p.index[clip] = bi; p.conf[clip] = bv; }
        __syncwarp();
    }
}

}  // namespace gat
