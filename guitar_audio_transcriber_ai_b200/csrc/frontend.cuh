// File front end of Transcriber.transcribe: what sits between the decoded WAV and the slicer /
// feature kernels in the reference.
//
//   audio/slicing.py:25, audio/loading.py:85   librosa.load: soundfile decode -> float32, channel mean, resample
//   audio/slicing.py:144                        sf.write(.wav) of each clip: PCM_16 quantisation, undone by the
//                                               loader's librosa.load -> every clip goes through int16 once
//   transcribe.py:173                           librosa.resample
//
// HBM-bound streaming kernels; the resampler is a polyphase FIR with float64 accumulation (it restates
// scipy.signal.resample_poly; soxr itself is not available).
#pragma once
#include "common.cuh"

namespace gat {

// soundfile.write(float32 -> PCM_16 .wav) followed by a float32 read.  python-soundfile switches SFC_SET_CLIPPING
// on for every file it opens, so libsndfile converts with its CLIPPING routine (src/pcm.c f2les_clip_array,
// normalised input): scaled = x * 2^31 in float, >= 2^31 - 1 -> 0x7FFF, <= -2^31 -> 0x8000, otherwise
// lrintf(scaled) >> 16 - i.e. floor(x * 32768) up to the rounding at 2^-16 of a step, NOT rint(x * 32767).
// The read scales by 1 / 0x8000.
__device__ __forceinline__ int pcm16_quantize(float x) {
    const float s = x * 2147483648.0f;
    if (s >= 2147483648.0f) return 32767;        // the largest float below 2^31 is 2^31 - 128: same test as >= 2^31 - 1
    if (s <= -2147483648.0f) return -32768;
    return __float2int_rn(s) >> 16;
}

__global__ void pcm16_roundtrip_kernel(float* __restrict__ x, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        x[i] = (float)pcm16_quantize(x[i]) * (1.0f / 32768.0f);
}

// Interleaved PCM_16 frames -> mono float32: x / 32768 per channel (libsndfile read), then the float32 channel
// mean (librosa.to_mono = np.mean(axis=0): float32 sum in channel order, one division).
__global__ void pcm16_to_mono_kernel(const short* __restrict__ in, long long frames, int channels, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < frames; i += stride) {
        const short* f = in + i * channels;
        float s = (float)f[0] * (1.0f / 32768.0f);
        for (int c = 1; c < channels; ++c) s += (float)f[c] * (1.0f / 32768.0f);
        out[i] = channels == 1 ? s : s / (float)channels;
    }
}

// Interleaved float32 frames -> mono (same mean).
__global__ void f32_to_mono_kernel(const float* __restrict__ in, long long frames, int channels, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < frames; i += stride) {
        const float* f = in + i * channels;
        float s = f[0];
        for (int c = 1; c < channels; ++c) s += f[c];
        out[i] = channels == 1 ? s : s / (float)channels;
    }
}

// Rational resampling by up/down (scipy.signal.resample_poly semantics): zero-stuff by `up`, FIR `h` of odd
// length 2*half+1 (already scaled by `up`), keep every `down`-th sample, delay removed:
//   out[m] = sum_i x[i] * h[m*down - i*up + half],   0 <= i < n_in,  0 <= tap index <= 2*half.
// One thread per output sample, clips along blockIdx.y; accumulation in float64 like scipy's upfirdn with a
// float64 window, rounded to float32 once.
struct ResampleParams {
    const float* in; long long n_in;       // [N][n_in]
    float* out; long long n_out;           // [N][n_out]
    const double* h; int half;             // [2*half+1]
    int up, down;
};

__global__ void __launch_bounds__(256) resample_poly_kernel(ResampleParams p) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= p.n_out) return;
    const float* x = p.in + (long long)blockIdx.y * p.n_in;
    const long long centre = m * p.down;                  // position on the up-sampled grid
    // i*up in [centre - half, centre + half]
    long long lo = centre - p.half, hi = centre + p.half;
    long long i0 = lo <= 0 ? 0 : (lo + p.up - 1) / p.up;
    long long i1 = hi / p.up;
    if (i1 > p.n_in - 1) i1 = p.n_in - 1;
    double acc = 0.0;
    for (long long i = i0; i <= i1; ++i) acc += (double)x[i] * p.h[centre - i * p.up + p.half];
    p.out[(long long)blockIdx.y * p.n_out + m] = (float)acc;
}

}  // namespace gat
