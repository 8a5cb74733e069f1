// YIN pitch baseline (reference: dsp/yin.py:39-75 -> librosa.yin(signal, fmin=50, fmax=1000, sr)).
//
// librosa builds the difference function from an FFT autocorrelation,
//     d[tau] = E[0] + E[tau] - 2 acf[tau],  acf[tau] = sum_{j=1..W} x[j] x[j+tau],  E[tau] = sum_{j=tau+1..tau+W} x[j]^2
// with W = 1024 inside frames of 2048 (hop 512, zero centre padding).  Here each warp owns one frame held
// in shared memory and evaluates acf directly as a register-tiled sliding dot product: lane l owns the
// kLPT consecutive lags [kLPT*l, kLPT*l + kLPT) and keeps the kLPT-sample window x[j+lag] in a register
// ring, so one broadcast load + one window load feed kLPT FMAs.  The cumulative-mean normalisation,
// trough search and parabolic refinement follow librosa's dtypes: d in float32, CMND and shifts in
// float64 (float32 / int64 promotes), f0 = sr / period in float64.
#pragma once
#include "common.cuh"

namespace gat {

struct YinParams {
    const float* audio;       // [N][n]
    long long n;
    int N;
    const float* clip_scale;  // divide samples by this per-clip value first (memory path), or nullptr
    int T;                    // frames = 1 + n / 512
    int hop;                  // 512
    int sr;
    int min_period, max_period;
    double trough_threshold;  // 0.1
    double* f0;               // [N][T]
};

constexpr int kYinFrame = 2048;
constexpr int kYinWin = 1024;
constexpr int kYinXs = 2240;          // frame + zero tail so the register ring can read ahead

template <int kLPT>
__host__ __device__ inline size_t yin_smem_per_warp() {
    return kYinXs * sizeof(float) + (size_t)32 * kLPT * sizeof(double);
}

template <int kLPT>
__global__ void __launch_bounds__(384, 1) yin_kernel(YinParams p) {
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int lane = lane_id(), warp = warp_id();
    unsigned char* base = smem_raw + (size_t)warp * yin_smem_per_warp<kLPT>();
    float* xs = reinterpret_cast<float*>(base);
    double* yv = reinterpret_cast<double*>(xs + kYinXs);            // CMND, index tau - min_period
    const int nl = p.max_period - p.min_period + 1;
    const double tiny = 1.1754943508222875e-38;                    // np.finfo(float32).tiny

    const long long n_work = (long long)p.N * p.T;
    for (long long work = (long long)blockIdx.x * nwarps + warp; work < n_work; work += (long long)gridDim.x * nwarps) {
        const int clip = (int)(work / p.T), t = (int)(work % p.T);
        const float* src = p.audio + (long long)clip * p.n;
        const float c = p.clip_scale ? p.clip_scale[clip] : 1.0f;
        for (int i = lane; i < kYinXs; i += 32) {
            const long long s = (long long)t * p.hop + i - kYinFrame / 2;
            float v = (i < kYinFrame && s >= 0 && s < p.n) ? src[s] : 0.0f;
            if (p.clip_scale) v = __fdiv_rn(v, c);
            xs[i] = v;
        }
        __syncwarp();

        // ---- sliding dot products for this lane's lags b .. b+kLPT-1
        const int b = kLPT * lane;
        float ring[kLPT], acc[kLPT];
#pragma unroll
        for (int q = 0; q < kLPT; ++q) { ring[q] = xs[b + 1 + q]; acc[q] = 0.0f; }
        float e_b = 0.0f;                                           // E[b] = sum_j x[j+b]^2
        for (int j0 = 1; j0 <= kYinWin; j0 += kLPT) {
#pragma unroll
            for (int s = 0; s < kLPT; ++s) {
                const int j = j0 + s;
                const bool in = j <= kYinWin;
                const float xj = in ? xs[j] : 0.0f;
#pragma unroll
                for (int i = 0; i < kLPT; ++i) acc[i] = fmaf(xj, ring[(s + i) % kLPT], acc[i]);
                if (in) e_b = fmaf(ring[s], ring[s], e_b);
                ring[s] = xs[b + j + kLPT];
            }
        }
        // ---- d[tau] = (E0 + E[tau]) - 2 acf[tau] in float32, with librosa's 1e-6 dead zones
        float e0 = __shfl_sync(0xffffffffu, e_b, 0);
        if (fabsf(e0) < 1e-6f) e0 = 0.0f;
        float e_tau = e_b;
        float dl[kLPT];
#pragma unroll
        for (int i = 0; i < kLPT; ++i) {
            const int tau = b + i;
            float a = acc[i];
            if (fabsf(a) < 1e-6f) a = 0.0f;
            float e = e_tau;
            if (fabsf(e) < 1e-6f) e = 0.0f;
            dl[i] = __fsub_rn(__fadd_rn(e0, e), __fmul_rn(2.0f, a));
            // slide the energy window to the next lag: E[tau+1] = E[tau] + x[tau+1025]^2 - x[tau+1]^2
            const float xin = xs[tau + kYinWin + 1], xout = xs[tau + 1];
            e_tau = e_tau + xin * xin - xout * xout;
        }
        // ---- cumulative sum over tau = 1..max_period (blocked: in-lane sequential + warp scan of lane totals)
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
        offs -= run;                                                // exclusive prefix of lane totals
#pragma unroll
        for (int i = 0; i < kLPT; ++i) {
            const int tau = b + i;
            if (tau >= p.min_period && tau <= p.max_period) {
                const double cm = (double)(offs + cl[i]) / (double)tau;
                yv[tau - p.min_period] = (double)dl[i] / (cm + tiny);
            }
        }
        __syncwarp();

        // ---- first trough under the threshold, else the global minimum (first occurrence)
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
                const double a = yp + ym - 2.0 * y0;
                const double bb = (yp - ym) / 2.0;
                if (!(fabs(bb) >= fabs(a))) shift = -bb / a;
            }
            p.f0[(long long)clip * p.T + t] = (double)p.sr / ((double)(p.min_period + idx) + shift);
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
