// YIN pitch baseline (reference: dsp/yin.py:39-75 -> librosa.yin(signal, fmin=50, fmax=1000, sr)).
//
// librosa builds the difference function from an FFT autocorrelation,
//     d[tau] = E[0] + E[tau] - 2 acf[tau],  acf[tau] = sum_{j=1..W} x[j] x[j+tau],  E[tau] = sum_{j=tau+1..tau+W} x[j]^2
// with W = 1024 inside frames of 2048 (hop 512, zero centre padding).  Consecutive frames overlap by half a window
// (hop = W/2), so every kernel here forms the sums per 512-sample BLOCK - a frame is the sum of two consecutive block
// partials - and a warp walks a segment of frames of one clip with one block of work per frame instead of two.
//   yin_fft_kernel   (bottom of the file; the one run_yin launches while max_period <= 512, i.e. up to 25.6 kHz): block
//                    partials through warp FFTs, one forward transform per block and one inverse per frame pair.
//   yin_pair_kernel  / yin_kernel: the block partials as a register-tiled sliding dot product (lane l owns kLPT lags and
//                    keeps the window x[j+lag] in a register ring, so one broadcast load + one window load feed kLPT FMAs);
//                    ten times the FLOPs, kept for higher sample rates and as the A/B reference (GAT_YIN_DIRECT).
// The cumulative-mean normalisation, trough search and parabolic refinement follow librosa's dtypes: d in float32, CMND
// and shifts in float64 (float32 / int64 promotes), f0 = sr / period in float64.
#pragma once
#include "common.cuh"
#include "fft.cuh"

namespace gat {

struct YinParams {
    const float* audio;       // [N][n]
    long long n;
    int N;
    const float* clip_scale;  // divide samples by this per-clip value first (memory path), or nullptr
    int T;                    // frames = 1 + n / 512
    int hop;                  // 512 (= kYinWin / 2: the block decomposition relies on it)
    int sr;
    int min_period, max_period;
    double trough_threshold;  // 0.1
    int seg_frames;           // frames per work item
    double* f0;               // [N][T]
    const Cpx<float>* tw;     // yin_fft_kernel: W_1024^(n1*k2) table of the warp FFT (fft.cuh), else unused
};

constexpr int kYinFrame = 2048;
constexpr int kYinWin = 1024;
constexpr int kYinBlock = 512;
constexpr int kYinBuf = 3072;         // floats of padded signal a warp keeps in shared memory

template <int kLPT>
__host__ __device__ inline size_t yin_smem_per_warp() {
    return kYinBuf * sizeof(float) + (size_t)32 * kLPT * sizeof(double);
}

// Tail of one frame, shared by both difference-function kernels: lane `lane` holds d[tau] for the kLPT CONSECUTIVE lags
// b .. b + kLPT - 1.  Cumulative-mean normalisation (float32 cumulative sum, float64 division as numpy promotes), first
// trough under the threshold else the global minimum, parabolic refinement, f0 = sr / period.
template <int kLPT>
__device__ __forceinline__ void yin_finish_frame(const float (&dl)[kLPT], int b, const YinParams& p, double* yv, int nl, int clip, int t) {
    const int lane = lane_id();
    const double tiny = 1.1754943508222875e-38;                    // np.finfo(float32).tiny
    // cumulative sum over tau = 1..max_period (blocked: in-lane sequential + warp scan of lane totals)
    float run = 0.0f, cl[kLPT];
#pragma unroll
    for (int i = 0; i < kLPT; ++i) {
        const int tau = b + i;
        if (tau >= 1 && tau <= p.max_period) run += dl[i];
        cl[i] = run;
    }
    float offs = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, offs, o);
        if (lane >= o) offs += u;
    }
    offs -= run;                                        // exclusive prefix of lane totals
#pragma unroll
    for (int i = 0; i < kLPT; ++i) {
        const int tau = b + i;
        if (tau >= p.min_period && tau <= p.max_period) {
            // d / (cumsum / tau + tiny) with ONE float64 division: d * tau / (cumsum + tiny * tau).  d * tau is exact in
            // float64 (24 + 10 bits); the quotient differs from numpy's two-division form by at most an ulp of float64,
            // nine orders below the float32 noise of d itself.  (The two divisions were ~17 % of the kernel.)
            const double dt = (double)tau;
            yv[tau - p.min_period] = ((double)dl[i] * dt) / ((double)(offs + cl[i]) + tiny * dt);
        }
    }
    __syncwarp();
    // first trough under the threshold, else the global minimum (first occurrence)
    int first = 0x7fffffff;
    double best = 1e300; int best_i = 0x7fffffff;
    for (int i = lane; i < nl; i += 32) {
        const double y0 = yv[i];
        bool trough;
        if (i == 0) trough = nl > 1 && y0 < yv[1];
        else if (i == nl - 1) trough = y0 < yv[i - 1];
        else trough = (y0 < yv[i - 1]) && (y0 <= yv[i + 1]);
        if (trough && y0 < p.trough_threshold && i < first) first = i;
        if (y0 < best) { best = y0; best_i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int f2 = __shfl_xor_sync(0xffffffffu, first, o);
        first = f2 < first ? f2 : first;
        const double b2 = __shfl_xor_sync(0xffffffffu, best, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (b2 < best || (b2 == best && i2 < best_i)) { best = b2; best_i = i2; }
    }
    if (lane == 0) {
        const int idx = first != 0x7fffffff ? first : (best_i != 0x7fffffff ? best_i : 0);
        double shift = 0.0;
        if (idx > 0 && idx < nl - 1) {
            const double ym = yv[idx - 1], y0 = yv[idx], yp = yv[idx + 1];
            const double aa = yp + ym - 2.0 * y0;
            const double bb = (yp - ym) / 2.0;
            if (!(fabs(bb) >= fabs(aa))) shift = -bb / aa;
        }
        p.f0[(long long)clip * p.T + t] = (double)p.sr / ((double)(p.min_period + idx) + shift);
    }
    __syncwarp();
}

template <int kLPT>
__global__ void __launch_bounds__(384, 1) yin_kernel(YinParams p) {
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int lane = lane_id(), warp = warp_id();
    unsigned char* base = smem_raw + (size_t)warp * yin_smem_per_warp<kLPT>();
    float* buf = reinterpret_cast<float*>(base);
    double* yv = reinterpret_cast<double*>(buf + kYinBuf);          // CMND, index tau - min_period
    const int nl = p.max_period - p.min_period + 1;
    // samples a block needs, counted from the start of the PREVIOUS block (whose frame is finalised with it)
    constexpr int kSpan = 2 * kYinBlock + 33 * kLPT + 1;
    constexpr int kBlocksPerFill = (kYinBuf - kSpan) / kYinBlock + 1;
    static_assert(kSpan <= kYinBuf, "kYinBuf too small for this lag tile");

    const int n_seg = (p.T + p.seg_frames - 1) / p.seg_frames;
    const long long n_work = (long long)p.N * n_seg;
    for (long long work = (long long)blockIdx.x * nwarps + warp; work < n_work; work += (long long)gridDim.x * nwarps) {
        const int clip = (int)(work / n_seg);
        const int f0 = (int)(work % n_seg) * p.seg_frames;
        const int nf = min(p.seg_frames, p.T - f0);
        const float* src = p.audio + (long long)clip * p.n;
        const float c = p.clip_scale ? p.clip_scale[clip] : 1.0f;
        auto padded = [&](long long i) {                            // centre-padded, normalised signal
            const long long s = i - kYinFrame / 2;
            float v = (s >= 0 && s < p.n) ? src[s] : 0.0f;
            if (p.clip_scale) v = __fdiv_rn(v, c);
            return v;
        };
        const int b = kLPT * lane;                                  // this lane's first lag
        float prev_acc[kLPT], prev_e = 0.0f;
#pragma unroll
        for (int q = 0; q < kLPT; ++q) prev_acc[q] = 0.0f;
        long long buf_base = 0;                                     // padded index of buf[0]
        int fill_left = 0;

        for (int blk = f0; blk <= f0 + nf; ++blk) {
            // ---- keep padded[512*(blk-1) .. +kSpan) resident
            if (fill_left == 0) {
                __syncwarp();
                const long long nb = (long long)kYinBlock * (blk - 1);
                // global loads eight at a time per lane: the refill is latency-bound, not bandwidth-bound
                auto fill = [&](int first) {
                    for (int i0 = first + lane; i0 < kYinBuf; i0 += 32 * 8) {
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) { const int i = i0 + 32 * u; v[u] = i < kYinBuf ? padded(nb + i) : 0.0f; }
#pragma unroll
                        for (int u = 0; u < 8; ++u) { const int i = i0 + 32 * u; if (i < kYinBuf) buf[i] = v[u]; }
                    }
                };
                if (blk == f0) {
                    fill(0);
                } else {                                            // slide: keep the tail, load the rest
                    // the shift is a multiple of 32, so every address is read and later overwritten by the SAME
                    // lane: program order makes the in-place forward copy safe without a staging array
                    const int keep = (int)(buf_base + kYinBuf - nb), off = kYinBuf - keep;
                    for (int i = lane; i < keep; i += 32) buf[i] = buf[i + off];
                    fill(keep);
                }
                buf_base = nb;
                fill_left = kBlocksPerFill;
                __syncwarp();
            }
            --fill_left;
            const float* xs = buf + (int)((long long)kYinBlock * blk - buf_base);    // xs[j] = padded[512*blk + j]

            // ---- block partials: acc[i] = sum_{j=1..512} xs[j] xs[j + b + i],  e_blk = sum_j xs[j + b]^2
            float ring[kLPT], acc[kLPT];
#pragma unroll
            for (int q = 0; q < kLPT; ++q) { ring[q] = xs[b + 1 + q]; acc[q] = 0.0f; }
            float e_blk = 0.0f;
            constexpr int kFull = kYinBlock / kLPT;                 // iterations that need no bounds check
            for (int it = 0; it < kFull; ++it) {
                const int j0 = 1 + it * kLPT;
#pragma unroll
                for (int s = 0; s < kLPT; ++s) {
                    const int j = j0 + s;
                    const float xj = xs[j];
#pragma unroll
                    for (int i = 0; i < kLPT; ++i) acc[i] = fmaf(xj, ring[(s + i) % kLPT], acc[i]);
                    e_blk = fmaf(ring[s], ring[s], e_blk);
                    ring[s] = xs[b + j + kLPT];
                }
            }
#pragma unroll
            for (int s = 0; s < kYinBlock - kFull * kLPT; ++s) {    // the last, partial iteration
                const int j = 1 + kFull * kLPT + s;
                const float xj = xs[j];
#pragma unroll
                for (int i = 0; i < kLPT; ++i) acc[i] = fmaf(xj, ring[(s + i) % kLPT], acc[i]);
                e_blk = fmaf(ring[s], ring[s], e_blk);
                ring[s] = xs[b + j + kLPT];
            }
            if (blk > f0) {
                // ---- frame t = blk - 1: acf = previous block + this block
                const int t = blk - 1;
                const float* fx = xs - kYinBlock;                   // fx[j] = padded[512*t + j]
                float e_b = prev_e + e_blk;                         // E[b] = sum_{j=1..1024} x[j+b]^2
                // d[tau] = (E0 + E[tau]) - 2 acf[tau] in float32, with librosa's 1e-6 dead zones
                float e0 = __shfl_sync(0xffffffffu, e_b, 0);
                if (fabsf(e0) < 1e-6f) e0 = 0.0f;
                float e_tau = e_b;
                float dl[kLPT];
#pragma unroll
                for (int i = 0; i < kLPT; ++i) {
                    const int tau = b + i;
                    float a = prev_acc[i] + acc[i];
                    if (fabsf(a) < 1e-6f) a = 0.0f;
                    float e = e_tau;
                    if (fabsf(e) < 1e-6f) e = 0.0f;
                    dl[i] = __fsub_rn(__fadd_rn(e0, e), __fmul_rn(2.0f, a));
                    // slide the energy window to the next lag: E[tau+1] = E[tau] + x[tau+1025]^2 - x[tau+1]^2
                    const float xin = fx[tau + kYinWin + 1], xout = fx[tau + 1];
                    e_tau = e_tau + xin * xin - xout * xout;
                }
                yin_finish_frame<kLPT>(dl, b, p, yv, nl, clip, t);
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < kLPT; ++q) prev_acc[q] = acc[q];
            prev_e = e_blk;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// The same difference function with the packed FP32 FMA of sm_100 (FFMA2, fma.rn.f32x2): round 2.
// The scalar kernel above issues one FFMA per (sample j, lag); ncu had it at 66 % issue utilisation with the FMA pipe half
// busy.  Here a lane's accumulators are PAIRS over two consecutive samples, (sum over odd j, sum over even j), updated by
//     (accA, accB)[lag] += (x[j], x[j+1]) * (x[j+lag], x[j+1+lag])
// - one instruction for two products.  The right-hand pair starts at index j + lag; for it to stay an aligned register pair
// while j advances by two, a lane owns lags of ONE parity, two apart: lane = 16*parity + m owns 2*(kLPT*m + i) + parity.
// The window of 2*kLPT samples lives in kLPT register pairs (a ring, as before).  acf = accA + accB at the end of a block:
// the summation order differs from the scalar kernel's (and from librosa's FFT autocorrelation) at float32 rounding level.
// After a frame's two blocks the d values go through shared memory once so that the cumulative sum runs with the
// consecutive-lag lane mapping of yin_finish_frame.
constexpr int kYinPairBuf = 2048;      // two blocks per refill; 13.6 KB of shared memory per warp -> 16 warps per SM

template <int kLPT>
__host__ __device__ inline size_t yin_pair_smem_per_warp() {
    return (kYinPairBuf + 2) * sizeof(float) + (size_t)32 * kLPT * sizeof(double) + (size_t)(32 * kLPT + 2) * sizeof(float);
}

template <int kLPT>
__global__ void __launch_bounds__(512, 1) yin_pair_kernel(YinParams p) {
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int lane = lane_id(), warp = warp_id();
    unsigned char* base = smem_raw + (size_t)warp * ((yin_pair_smem_per_warp<kLPT>() + 15) / 16 * 16);
    float* buf = reinterpret_cast<float*>(base);                            // buf[i] = padded[buf_base - 1 + i]: sample 512*blk + 1 sits at an even index
    double* yv = reinterpret_cast<double*>(buf + kYinPairBuf + 2);          // CMND, index tau - min_period
    float* dsm = reinterpret_cast<float*>(yv + 32 * kLPT);                  // d[tau] of the frame being finished
    const int nl = p.max_period - p.min_period + 1;
    constexpr int kLags = 32 * kLPT;
    // samples a block needs, counted from the start of the PREVIOUS block (whose frame is finalised with it): the energy
    // slide of the finish reads fx[tau + 1026] with fx = xs - 512, the inner loop xs[512 + kLags + 1]
    constexpr int kSpan = 2 * kYinBlock + kLags + 2 * kLPT + 4;
    constexpr int kBlocksPerFill = (kYinPairBuf - kSpan) / kYinBlock + 1;
    static_assert(kSpan <= kYinPairBuf, "kYinPairBuf too small for this lag tile");

    const int par = lane >> 4, m = lane & 15;
    const int b = 2 * kLPT * m + par;                                       // this lane's first lag; its lags are b, b + 2, ...

    const int n_seg = (p.T + p.seg_frames - 1) / p.seg_frames;
    const long long n_work = (long long)p.N * n_seg;
    for (long long work = (long long)blockIdx.x * nwarps + warp; work < n_work; work += (long long)gridDim.x * nwarps) {
        const int clip = (int)(work / n_seg);
        const int f0 = (int)(work % n_seg) * p.seg_frames;
        const int nf = min(p.seg_frames, p.T - f0);
        const float* src = p.audio + (long long)clip * p.n;
        const float c = p.clip_scale ? p.clip_scale[clip] : 1.0f;
        auto padded = [&](long long i) {                            // centre-padded, normalised signal
            const long long s = i - kYinFrame / 2;
            float v = (s >= 0 && s < p.n) ? src[s] : 0.0f;
            if (p.clip_scale) v = __fdiv_rn(v, c);
            return v;
        };
        float prev_acc[kLPT], prev_e = 0.0f;
#pragma unroll
        for (int q = 0; q < kLPT; ++q) prev_acc[q] = 0.0f;
        long long buf_base = 0;                                     // padded index of buf[1]
        int fill_left = 0;

        for (int blk = f0; blk <= f0 + nf; ++blk) {
            // ---- keep padded[512*(blk-1) .. +kSpan) resident (buf[0] is the sample before it)
            if (fill_left == 0) {
                __syncwarp();
                const long long nb = (long long)kYinBlock * (blk - 1);
                auto fill = [&](int first) {
                    for (int i0 = first + lane; i0 < kYinPairBuf; i0 += 32 * 8) {
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) { const int i = i0 + 32 * u; v[u] = i < kYinPairBuf ? padded(nb - 1 + i) : 0.0f; }
#pragma unroll
                        for (int u = 0; u < 8; ++u) { const int i = i0 + 32 * u; if (i < kYinPairBuf) buf[i] = v[u]; }
                    }
                };
                if (blk == f0) {
                    fill(0);
                } else {                                            // slide: keep the tail, load the rest (shift is a multiple of 32)
                    const int keep = (int)(buf_base + kYinPairBuf - nb), off = kYinPairBuf - keep;
                    for (int i = lane; i < keep; i += 32) buf[i] = buf[i + off];
                    fill(keep);
                }
                buf_base = nb;
                fill_left = kBlocksPerFill;
                __syncwarp();
            }
            --fill_left;
            const float* xs = buf + 1 + (int)((long long)kYinBlock * blk - buf_base);    // xs[j] = padded[512*blk + j]; &xs[1] is 8-byte aligned

            // ---- block partials over sample pairs (j, j+1), j = 1, 3, ..., 511
            float2 ring[kLPT], acc[kLPT];
#pragma unroll
            for (int q = 0; q < kLPT; ++q) { ring[q] = make_float2(xs[1 + b + 2 * q], xs[2 + b + 2 * q]); acc[q] = make_float2(0.0f, 0.0f); }
            // four independent energy accumulators: one would be a serial FFMA2 chain (the compiler groups the fifteen
            // updates of a round back to back: ncu's top stall was the fixed-latency wait on it)
            float2 e2[4] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
            constexpr int kSteps = kYinBlock / 2;                   // 256 sample pairs
            constexpr int kFull = kSteps / kLPT;
            for (int it = 0; it < kFull; ++it) {
                const int j0 = 1 + 2 * it * kLPT;
#pragma unroll
                for (int s = 0; s < kLPT; ++s) {
                    const int j = j0 + 2 * s;
                    const float2 xp = *reinterpret_cast<const float2*>(xs + j);          // broadcast load
#pragma unroll
                    for (int i = 0; i < kLPT; ++i) acc[i] = __ffma2_rn(xp, ring[(s + i) % kLPT], acc[i]);
                    e2[s & 3] = __ffma2_rn(ring[s], ring[s], e2[s & 3]);
                    ring[s] = make_float2(xs[j + b + 2 * kLPT], xs[j + b + 2 * kLPT + 1]);
                }
            }
#pragma unroll
            for (int s = 0; s < kSteps - kFull * kLPT; ++s) {       // the last, partial round
                const int j = 1 + 2 * (kFull * kLPT + s);
                const float2 xp = *reinterpret_cast<const float2*>(xs + j);
#pragma unroll
                for (int i = 0; i < kLPT; ++i) acc[i] = __ffma2_rn(xp, ring[(s + i) % kLPT], acc[i]);
                e2[s & 3] = __ffma2_rn(ring[s], ring[s], e2[s & 3]);
                ring[s] = make_float2(xs[j + b + 2 * kLPT], xs[j + b + 2 * kLPT + 1]);
            }
            float accs[kLPT];
#pragma unroll
            for (int q = 0; q < kLPT; ++q) accs[q] = acc[q].x + acc[q].y;
            const float e_blk = ((e2[0].x + e2[0].y) + (e2[1].x + e2[1].y)) + ((e2[2].x + e2[2].y) + (e2[3].x + e2[3].y));   // sum_{j=1..512} xs[j + b]^2
            if (blk > f0) {
                // ---- frame t = blk - 1: acf = previous block + this block
                const int t = blk - 1;
                const float* fx = xs - kYinBlock;                   // fx[j] = padded[512*t + j]
                float e_b = prev_e + e_blk;                         // E[b] = sum_{j=1..1024} x[j+b]^2
                float e0 = __shfl_sync(0xffffffffu, e_b, 0);        // lane 0 owns lag 0
                if (fabsf(e0) < 1e-6f) e0 = 0.0f;
                float e_tau = e_b;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kLPT; ++i) {
                    const int tau = b + 2 * i;
                    float a = prev_acc[i] + accs[i];
                    if (fabsf(a) < 1e-6f) a = 0.0f;
                    float e = e_tau;
                    if (fabsf(e) < 1e-6f) e = 0.0f;
                    dsm[tau] = __fsub_rn(__fadd_rn(e0, e), __fmul_rn(2.0f, a));
                    // slide the energy window two lags on: E[tau+1] = E[tau] + x[tau+1025]^2 - x[tau+1]^2, twice
                    const float in1 = fx[tau + kYinWin + 1], out1 = fx[tau + 1];
                    e_tau = e_tau + in1 * in1 - out1 * out1;
                    const float in2 = fx[tau + kYinWin + 2], out2 = fx[tau + 2];
                    e_tau = e_tau + in2 * in2 - out2 * out2;
                }
                __syncwarp();
                float dl[kLPT];
#pragma unroll
                for (int i = 0; i < kLPT; ++i) dl[i] = dsm[kLPT * lane + i];
                yin_finish_frame<kLPT>(dl, kLPT * lane, p, yv, nl, clip, t);
            }
#pragma unroll
            for (int q = 0; q < kLPT; ++q) prev_acc[q] = accs[q];
            prev_e = e_blk;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// The difference function through FFTs, as librosa forms it - but per 512-sample BLOCK instead of per frame.
//
// With hop = W/2 a frame's autocorrelation is the sum of two block partials, acf_t = P_t + P_{t+1},
//     P_b[tau] = sum_{j=1..512} xs[j] xs[j+tau],  xs[j] = padded[512 b + j],
// and P_b is a linear correlation of u = xs[1..512] with v = xs[1..1024]: for tau <= 512 nothing wraps in a cyclic
// correlation of length 1024.  Both are real, so ONE complex 1024-point FFT of z = v + i u_padded carries the two
// spectra (Z = V + iU), and with A = Z[k], Z' = Z[1024-k]
//     C_b[k] = conj(U[k]) V[k] = ( (A.x Z'.y + A.y Z'.x) / 2,  (|A|^2 - |Z'|^2) / 4 ).
// C_b is Hermitian (its inverse transform is real), so TWO frames share one inverse transform:
//     G = (C_t + C_{t+1}) + i (C_{t+1} + C_{t+2}),   IFFT(G) = acf_t + i acf_{t+1},   IFFT(G) = conj(FFT(conj G)) / 1024.
// A warp walks a segment of frames of one clip: one forward transform per block, one inverse per frame pair: ~1.6 FFTs
// of 51 kFLOP per frame where the direct form (yin_pair_kernel) spends 500 kFLOP.  The transform is the warp FFT of
// the STFT kernels (32 complex registers per lane, one transpose through shared memory); its output layout - bin
// 32 k1 + lane in register bitrev5(k1) - is the input layout of the next transform up to a register renaming, the
// partner bin Z[1024-k] sits in lane (32 - lane) % 32 (one shuffle), and the running G lives in a lane-private 8 KB
// strip of shared memory.  The energy terms slide along the lags as before (block sums of squares from the transform's
// own inputs, a warp scan over the per-lane increments).  float32 throughout, like numpy's rfft on float32 frames;
// the summation order differs from both numpy's and the direct kernels' at float32 rounding level.
// Valid while max_period <= 512 (sample rates up to 25.6 kHz at fmin = 50 Hz); run_yin falls back to the direct form.
// 12 warps: three per scheduler, 168 registers each.  (Shared memory would hold a 13th, but four warps on one scheduler
// cap every thread at 128 registers and the block prefetch below then spills.)
template <int kLPT>
__host__ __device__ constexpr int yin_fft_warps() { return 12; }
template <int kLPT>
__host__ __device__ constexpr size_t yin_fft_scratch_bytes() {        // per warp: transpose buffer, reused by the frame finish
    return (size_t)kXbufStride * 32 * sizeof(Cpx<float>) > (size_t)512 * kLPT ? (size_t)kXbufStride * 32 * sizeof(Cpx<float>) : (size_t)512 * kLPT;
}
template <int kLPT>
__host__ __device__ constexpr size_t yin_fft_smem_bytes() {
    return 1024 * sizeof(Cpx<float>) + (size_t)yin_fft_warps<kLPT>() * (yin_fft_scratch_bytes<kLPT>() + 1024 * sizeof(Cpx<float>)) + 64;
}

// Forward complex FFT of 1024 points held as v[r] = z[lane + 32 r]; on return v[bitrev5(k1)] = Z[32 k1 + lane].
// (The first half of frame_fft_power in stft2.cuh: two in-lane 32-point transforms around a twiddled transpose.)
__device__ __forceinline__ void warp_cfft1024(Cpx<float> (&v)[32], Cpx<float>* xbuf, const Cpx<float>* tw) {
    const int lane = lane_id();
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        fft_dif<float, 32, 0, 32>(v);
        if (pass == 0) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int k2 = bitrev5(r);
                const Cpx<float> w = tw[k2 * 32 + lane];
                xbuf[k2 * kXbufStride + lane] = (k2 == 0) ? v[r] : cmul(v[r], w);
            }
            __syncwarp();
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[n1] = xbuf[lane * kXbufStride + n1];
            __syncwarp();
        }
    }
}

// Tail of one frame for the FFT kernel: the arithmetic of yin_finish_frame (same operations in the same order, so the two
// give identical results on identical d), but the float64 divisions run in a ROLLED loop over shared memory - fourteen
// inlined divisions were a tenth of the kernel's code, and its warps, each somewhere else in an 80 KB loop, stalled on
// instruction fetch (ncu: "no instruction" 2.0 per issue).
template <int kLPT>
__device__ __forceinline__ void yin_finish_frame_rolled(const float (&dl)[kLPT], int b, const YinParams& p, double* yv, int nl, int clip, int t) {
    const int lane = lane_id();
    const double tiny = 1.1754943508222875e-38;                    // np.finfo(float32).tiny
    float run = 0.0f, cl[kLPT];
#pragma unroll
    for (int i = 0; i < kLPT; ++i) {
        const int tau = b + i;
        if (tau >= 1 && tau <= p.max_period) run += dl[i];
        cl[i] = run;
    }
    float offs = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, offs, o);
        if (lane >= o) offs += u;
    }
    offs -= run;                                        // exclusive prefix of lane totals
    float2* dc = reinterpret_cast<float2*>(yv);         // (d, cumulative sum) in the slot its quotient will take
#pragma unroll
    for (int i = 0; i < kLPT; ++i) {
        const int tau = b + i;
        if (tau >= p.min_period && tau <= p.max_period) dc[tau - p.min_period] = make_float2(dl[i], offs + cl[i]);
    }
    __syncwarp();
    // Quotients and trough tests 32 lags at a time, in increasing lag order, and NO FURTHER than the first trough under the
    // threshold: the result is that trough (librosa takes the first), so the lags behind it are never needed - for a pitched
    // frame that is most of them.  Chunk j is tested once chunk j + 1 has its quotients (a lag's test reads its right
    // neighbour).  The global minimum is only used when no chunk had such a trough, i.e. when all were visited.
    int first = 0x7fffffff;
    double best = 1e300; int best_i = 0x7fffffff;
    const int nch = (nl + 31) >> 5;
#pragma unroll 1
    for (int j = 0; j <= nch; ++j) {
        const int i = lane + 32 * j;
        if (i < nl) {                                   // every lane rewrites only the slot it reads
            const float2 q = dc[i];
            const double dt = (double)(i + p.min_period);
            yv[i] = ((double)q.x * dt) / ((double)q.y + tiny * dt);     // d / (cumsum / tau + tiny) with one division, as above
        }
        __syncwarp();
        const int it = i - 32;
        if (j >= 1 && it < nl) {
            const double y0 = yv[it];
            bool trough;
            if (it == 0) trough = nl > 1 && y0 < yv[1];
            else if (it == nl - 1) trough = y0 < yv[it - 1];
            else trough = (y0 < yv[it - 1]) && (y0 <= yv[it + 1]);
            if (trough && y0 < p.trough_threshold && it < first) first = it;
            if (y0 < best) { best = y0; best_i = it; }
        }
        if (__any_sync(0xffffffffu, first != 0x7fffffff)) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int f2 = __shfl_xor_sync(0xffffffffu, first, o);
        first = f2 < first ? f2 : first;
        const double b2 = __shfl_xor_sync(0xffffffffu, best, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (b2 < best || (b2 == best && i2 < best_i)) { best = b2; best_i = i2; }
    }
    if (lane == 0) {
        const int idx = first != 0x7fffffff ? first : (best_i != 0x7fffffff ? best_i : 0);
        double shift = 0.0;
        if (idx > 0 && idx < nl - 1) {
            const double ym = yv[idx - 1], y0 = yv[idx], yp = yv[idx + 1];
            const double aa = yp + ym - 2.0 * y0;
            const double bb = (yp - ym) / 2.0;
            if (!(fabs(bb) >= fabs(aa))) shift = -bb / aa;
        }
        p.f0[(long long)clip * p.T + t] = (double)p.sr / ((double)(p.min_period + idx) + shift);
    }
    __syncwarp();
}

template <int kLPT>
__global__ void __launch_bounds__(32 * yin_fft_warps<kLPT>(), 1) yin_fft_kernel(YinParams p) {
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int lane = lane_id(), warp = warp_id();
    Cpx<float>* tw = reinterpret_cast<Cpx<float>*>(smem_raw);
    unsigned char* mine = smem_raw + 1024 * sizeof(Cpx<float>) + (size_t)warp * (yin_fft_scratch_bytes<kLPT>() + 1024 * sizeof(Cpx<float>));
    Cpx<float>* xbuf = reinterpret_cast<Cpx<float>*>(mine);                 // transpose buffer / stash of the next inverse's input
    Cpx<float>* gs = reinterpret_cast<Cpx<float>*>(mine + yin_fft_scratch_bytes<kLPT>());   // running G, slot k1*32 + lane = bin 32 k1 + lane
    // the frame finish reuses the transpose buffer between transforms
    double* yv = reinterpret_cast<double*>(mine);                           // CMND, index tau - min_period
    float* dsm = reinterpret_cast<float*>(mine);                            // energy increments: consumed before yv is written
    float* acf_a = reinterpret_cast<float*>(yv + 32 * kLPT);                // autocorrelations of the pair's two frames
    float* acf_b = acf_a + 32 * kLPT;
    float* stage = reinterpret_cast<float*>(mine);                          // 1024 samples of a block that touches the padding
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tw[i] = p.tw[i];
    __syncthreads();           // the only block barrier

    const int nl = p.max_period - p.min_period + 1;
    const int partner = (32 - lane) & 31;
    const int n32 = (int)p.n;                                               // run_yin: n < 2^31 - 2^16
    const int n_seg = (p.T + p.seg_frames - 1) / p.seg_frames;
    const long long n_work = (long long)p.N * n_seg;
    for (long long work = (long long)blockIdx.x * nwarps + warp; work < n_work; work += (long long)gridDim.x * nwarps) {
        const int clip = (int)(work / n_seg);
        const int f0 = (int)(work % n_seg) * p.seg_frames;
        const int nf = min(p.seg_frames, p.T - f0);
        const float* src = p.audio + (long long)clip * p.n;
        // normalised path: multiply by 1/c (the direct kernels divide every sample: an ulp of difference per sample)
        const float ic = p.clip_scale ? 1.0f / p.clip_scale[clip] : 1.0f;
        auto padded = [&](int i) {                                  // centre-padded (zeros), UNSCALED signal
            const int s = i - kYinFrame / 2;
            return ((unsigned)s < (unsigned)n32) ? src[s] : 0.0f;
        };
        // Samples of a block are fetched one step ahead (the loads of block idx + 1 fly during block idx's spectrum
        // product and, for every second block, the inverse transform): ncu's top stall was the wait for these loads.
        float nx[32];
        bool pre = false;                                           // nx holds the next block (it touches no padding)
        auto prefetch = [&](int blk) {
            const int s0 = kYinBlock * blk + 1 - kYinFrame / 2;     // source index of xs[1]
            pre = s0 >= 0 && s0 + 1024 <= n32;
            if (pre) {
                const float* g = src + s0 + lane;
#pragma unroll
                for (int r = 0; r < 32; ++r) nx[r] = g[32 * r];
            }
        };
        prefetch(f0);
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) gs[k1 * 32 + lane] = Cpx<float>{0.0f, 0.0f};
        float tb0 = 0.0f, tb1 = 0.0f, tb2 = 0.0f;                   // sum_{j=1..512} xs[j]^2 of this block and the two before
        float e0_a = 0.0f, e0_b = 0.0f;                             // E[0] of the frames of the pending inverse
        int idx = 0, inv_frames = 0;
        bool inv = false;
        Cpx<float> v[32];
#pragma unroll 1
        while (true) {
            if (!inv) {
                // ---- block f0 + idx: z[j] = xs[1 + j] * (1 + i [j < 512])
                if (!pre) {                                         // rolled: stage the padded block through shared memory
                    const int base = kYinBlock * (f0 + idx) + 1;
#pragma unroll 1
                    for (int e0 = lane; e0 < 1024; e0 += 32 * 8) {
                        float val[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) val[u] = padded(base + e0 + 32 * u);
#pragma unroll
                        for (int u = 0; u < 8; ++u) stage[e0 + 32 * u] = val[u];
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < 32; ++r) nx[r] = stage[lane + 32 * r];
                    __syncwarp();
                }
                float es = 0.0f;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const float x = nx[r] * ic;
                    v[r] = Cpx<float>{x, r < 16 ? x : 0.0f};
                    if (r < 16) es = fmaf(x, x, es);
                }
                tb2 = tb1; tb1 = tb0;
                tb0 = warp_sum(es);
            } else {
#pragma unroll
                for (int r = 0; r < 32; ++r) v[r] = xbuf[r * 32 + lane];
                __syncwarp();                                       // the stash is consumed before the transpose overwrites it
            }
            warp_cfft1024(v, xbuf, tw);
            if (!inv) {
                // ---- C_b from Z, folded into the running G.  Position of block idx in its frame pair:
                //   first block of the segment (gs = 0)   gs += C
                //   odd idx, more blocks follow           gs += (1 + i) C
                //   odd idx, last block (nf odd)          stash = gs + C            -> inverse, one frame  (idx - 1)
                //   even idx >= 2                         stash = gs + i C, gs = C  -> inverse, two frames (idx - 2, idx - 1)
                const bool first = idx == 0, odd = (idx & 1) != 0, last = idx == nf;
                const bool to_gs = first || (odd && !last);
                const bool end = !odd && !first;
                // out = t + m C with m = 1, 1 + i, 1, i, conjugated for the stash (it holds conj(G): input n = lane + 32 k1
                // of the inverse).  With C = (kH P, kQ Q) the multiplier, the conjugation and the inverse transform's 1/1024
                // (kH, kQ: exact powers of two) fold into four warp-uniform scalars: out = (tx + a1 P - a2 Q, sg ty + b1 Q + b2 P).
                constexpr float kH = 0.5f / 1024.0f, kQ = 0.25f / 1024.0f;
                const float mx = end ? 0.0f : 1.0f, my = (odd && last) || first ? 0.0f : 1.0f;
                const float sg = to_gs ? 1.0f : -1.0f;
                const float a1 = mx * kH, a2 = my * kQ, b1 = sg * mx * kQ, b2 = sg * my * kH;
                Cpx<float>* const dst = (to_gs ? gs : xbuf) + lane;
                Cpx<float>* const g = gs + lane;
                Cpx<float>* const dst2 = end ? g : dst;             // second store: C for the next pair, else the first one again
                                                                    // (a conditional store made the compiler clone the loop)
                if (!last) prefetch(f0 + idx + 1);
                // straight-line code (selects and one predicated store, no branches), eight bins at a time: loads and
                // shuffles first, stores last
#pragma unroll
                for (int c8 = 0; c8 < 32; c8 += 8) {
                    float zx[8], zy[8];
                    Cpx<float> t[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k1 = c8 + q;
                        const Cpx<float> other = v[bitrev5(31 - k1)];
                        const Cpx<float> self = v[bitrev5((32 - k1) & 31)];
                        const float sx = __shfl_sync(0xffffffffu, other.x, partner);
                        const float sy = __shfl_sync(0xffffffffu, other.y, partner);
                        zx[q] = lane == 0 ? self.x : sx;
                        zy[q] = lane == 0 ? self.y : sy;
                        t[q] = g[k1 * 32];
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k1 = c8 + q;
                        const Cpx<float> a = v[bitrev5(k1)];
                        const float P = a.x * zy[q] + a.y * zx[q];
                        const float Q = (a.x * a.x + a.y * a.y) - (zx[q] * zx[q] + zy[q] * zy[q]);
                        const float ox = fmaf(a1, P, fmaf(-a2, Q, t[q].x)), oy = fmaf(b1, Q, fmaf(b2, P, sg * t[q].y));
                        dst[k1 * 32] = Cpx<float>{ox, oy};
                        dst2[k1 * 32] = Cpx<float>{end ? P * kH : ox, end ? Q * kQ : oy};
                    }
                }
                __syncwarp();
                if (to_gs) {
                    ++idx;
                } else {
                    inv = true;
                    inv_frames = odd ? 1 : 2;
                    if (odd) { e0_a = tb1 + tb0; } else { e0_a = tb2 + tb1; e0_b = tb1 + tb0; }
                }
            } else {
                // ---- FFT(conj G) = conj(acf_a + i acf_b); lag 32 k1 + lane in register bitrev5(k1)
#pragma unroll
                for (int k1 = 0; k1 < kLPT; ++k1) {
                    acf_a[32 * k1 + lane] = v[bitrev5(k1)].x;
                    acf_b[32 * k1 + lane] = -v[bitrev5(k1)].y;
                }
                __syncwarp();
#pragma unroll 1
                for (int fi = 0; fi < inv_frames; ++fi) {
                    const int t = f0 + idx - inv_frames + fi;
                    const float* ac = fi ? acf_b : acf_a;
                    const float e_first = fi ? e0_b : e0_a;                    // E[0] = sum_{j=1..1024} fx[j]^2, fx[j] = padded[512 t + j]
                    // E[tau+1] - E[tau] = fx[tau+1025]^2 - fx[tau+1]^2: coalesced loads, then through shared memory to
                    // the consecutive-lag mapping of the finish
                    const int fb = kYinBlock * t + 1 + lane;                   // padded index of fx[1 + lane]
                    const int s_lo = fb - lane - kYinFrame / 2;
                    if (s_lo >= 0 && s_lo + kYinWin + 32 * kLPT <= n32) {
                        const float* g = src + s_lo + lane;
#pragma unroll
                        for (int i = 0; i < kLPT; ++i) {
                            const float xin = g[32 * i + kYinWin] * ic, xout = g[32 * i] * ic;
                            dsm[32 * i + lane] = xin * xin - xout * xout;
                        }
                    } else {
#pragma unroll 2
                        for (int i = 0; i < kLPT; ++i) {
                            const float xin = padded(fb + 32 * i + kYinWin) * ic, xout = padded(fb + 32 * i) * ic;
                            dsm[32 * i + lane] = xin * xin - xout * xout;
                        }
                    }
                    __syncwarp();
                    const int b = kLPT * lane;
                    float inc[kLPT], tot = 0.0f;
#pragma unroll
                    for (int i = 0; i < kLPT; ++i) { inc[i] = dsm[b + i]; tot += inc[i]; }
                    float offs = tot;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float u = __shfl_up_sync(0xffffffffu, offs, o);
                        if (lane >= o) offs += u;
                    }
                    float e_tau = e_first + (offs - tot);                      // E[b]
                    const float e0 = fabsf(e_first) < 1e-6f ? 0.0f : e_first;  // librosa's 1e-6 dead zones
                    float dl[kLPT];
#pragma unroll
                    for (int i = 0; i < kLPT; ++i) {
                        float a = ac[b + i];
                        if (fabsf(a) < 1e-6f) a = 0.0f;
                        float e = e_tau;
                        if (fabsf(e) < 1e-6f) e = 0.0f;
                        dl[i] = __fsub_rn(__fadd_rn(e0, e), __fmul_rn(2.0f, a));
                        e_tau += inc[i];
                    }
                    __syncwarp();                                              // dsm (aliasing yv) is consumed
                    yin_finish_frame_rolled<kLPT>(dl, b, p, yv, nl, clip, t);
                }
                inv = false;
                if (idx == nf) break;
                ++idx;
            }
        }
        __syncwarp();
    }
}

// Median of the frame-wise f0 per clip (dsp/yin.py:57-67: NaNs dropped, np.median, float64) and the MLP's
// pitch feature log10(hz) (audio/features.py:204-206).  One warp per clip, rank selection.
struct YinMedianParams {
    const double* f0;   // [N][T]
    int N, T;
    double* hz_out;     // [N] (NaN when no valid frame: the reference returns None)
    float* feat_out;    // optional: feat_out[clip*ld + col] = (float)log10(hz)
    int ld, col;
};

__global__ void yin_median_kernel(YinMedianParams p) {
    const int clip = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (clip >= p.N) return;
    const int lane = lane_id();
    const double* f = p.f0 + (long long)clip * p.T;
    int nvalid = 0;
    for (int i = lane; i < p.T; i += 32) nvalid += (f[i] == f[i]) ? 1 : 0;
    nvalid = warp_sum(nvalid);
    double lo = 0.0, hi = 0.0;
    const int k_lo = (nvalid - 1) / 2, k_hi = nvalid / 2;
    for (int i = lane; i < p.T; i += 32) {
        const double v = f[i];
        if (!(v == v)) continue;
        int rank = 0;
        for (int j = 0; j < p.T; ++j) {
            const double u = f[j];
            if (u == u && (u < v || (u == v && j < i))) ++rank;
        }
        if (rank == k_lo) lo = v;
        if (rank == k_hi) hi = v;
    }
    // exactly one lane holds each of lo / hi; the others contribute 0
    lo = warp_sum(lo);
    hi = warp_sum(hi);
    if (lane == 0) {
        double hz = nvalid == 0 ? __longlong_as_double(0x7ff8000000000000LL) : (k_lo == k_hi ? lo : (lo + hi) / 2.0);
        p.hz_out[clip] = hz;
        if (p.feat_out) p.feat_out[(long long)clip * p.ld + p.col] = (float)log10(hz);
    }
}

}  // namespace gat
