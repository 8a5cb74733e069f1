// Feature kernels: per-clip RMS scale, fused frame+Hann+rFFT+|.|^2+mel+dB (three chains), MFCC finish.
//
//   reference call sites (paths under /root/reference/version_1/source):
//     audio/features.py:124-126   _normalize_audio_volume          -> clip_scale_kernel
//     audio/features.py:296-316   torchaudio MelSpectrogram + dB    -> stft_frames_kernel (stft2.cuh) at n_fft 2048,
//                                                                      stft_mel_kernel<float, kImage> at 512 / 1024 / 4096
//     audio/features.py:187-193   librosa.feature.mfcc + time-mean  -> stft_frames_kernel (stft2.cuh; n_fft 2048)
//     audio/slicing.py:107        librosa.onset.onset_strength      -> stft_mel_kernel<double, kSpec> (+ onset.cuh)
//
// One persistent kernel serves all three: a CTA takes (clip, chunk of frames) work items, stages the
// padded, normalised samples of the chunk in shared memory once (each sample is reused by n_fft/hop
// frames), and each warp turns one frame at a time into mel bins without leaving the SM: the spectrum
// and the mel power never touch HBM.  The mel filterbank is applied in its banded-sparse form (each FFT
// bin feeds at most two triangular filters), 2*(n_fft/2+1) MACs per frame instead of the dense
// n_mels*(n_fft/2+1).
#pragma once
#include "fft.cuh"

namespace gat {

// Banded-sparse mel filterbank in "lane slot" form.  Lane l of a warp owns up to kMaxMelsPerLane filters
// (slot q -> entry q*32 + l).  The gather loop reads four bins and four weights per step (128-bit loads).
// The host lowers each band's start to a multiple of 4 congruent to 4*(l mod 8) modulo 32 and places its
// weights at an offset with the same residue: in each 8-lane phase of a 128-bit shared load the lanes then
// touch 8 different 16-byte bank groups, for the power spectrum and for the weights alike (see
// build_sparse_fb in gat.cu for the assignment).  `len` counts groups of four bins.
struct SparseFb {
    int n_mels;
    int n_slots;         // ceil(n_mels / 32)
    int nnz;             // length of w including alignment gaps
    const int* start;    // [n_slots*32] first bin of the (padded) band; may be negative (reads land in zeros)
    const int* len;      // [n_slots*32] groups of 4 bins in the padded band (0 = empty slot)
    const int* off;      // [n_slots*32] offset of the band's first weight in w
    const int* mel;      // [n_slots*32] filter index the slot produces, -1 if empty
    const float* w;      // [nnz]
};

constexpr int kPbufLead = 32;   // zeros in front of the power spectrum so padded bands may start below bin 0

template <typename T> struct __align__(4 * sizeof(T)) Vec4 { T x, y, z, w; };

enum { kPadZero = 0, kPadReflect = 1 };
enum { kOutImage = 0, kOutSpec = 1 };

template <typename T>
struct StftMelParams {
    const float* audio;        // [N][n] float32 clips (or one long signal with N = 1)
    long long n;               // samples per clip
    int N;                     // clips
    const float* clip_scale;   // [N] divisor c = rms + 1e-9, or nullptr (no volume normalisation)
    const unsigned char* frame_gate;  // onset chain: per 512-sample block keep flag [N][gate_stride], or nullptr
    long long gate_stride;     // frame_gate entries per clip
    float sample_gate;         // onset chain: samples with |y| < sample_gate are zeroed; 0 disables both gates
    int gate_hop;              // samples per frame_gate entry (512)
    int use_async;             // 1: prefetch the next chunk's samples with cp.async into a raw landing zone
    int hop;
    int n_frames;              // frames per clip = 1 + n / hop
    int pad_mode;
    const T* window;           // [n_fft]
    const T* window_half;      // [n_fft] 0.5 * window; only read when n_fft = 4096 (window stays in global/L1 there)
    const Cpx<T>* tw;          // [n_fft/2]   W_C^(n1*k2), C = n_fft/2 (FftTables::tw)
    const Cpx<T>* w2;          // [n_fft/4]   W_n_fft^k
    SparseFb fb;
    int frames_per_cta;
    int chunks_per_clip;
    T amin;                    // 1e-10
    int power_out;             // 1: write the mel POWER (MelSpecConfig.TO_DB = False, features.py:313-316), no log
    // kOutImage: out[(clip*n_mels + m)*T + t] = 10 log10(max(amin, mel))
    // kOutSpec : out[(clip*T + t)*n_mels + m] = same, plus per-clip running max in spec_max (ordered ints)
    T* out;
    long long* spec_max;       // [N] ordered-integer encoded maxima (kOutSpec only)
};

__device__ __forceinline__ long long ordered_bits(double v) {
    long long i = __double_as_longlong(v);
    return i >= 0 ? i : i ^ 0x7fffffffffffffffLL;
}
__device__ __forceinline__ double from_ordered_bits(long long i) {
    return __longlong_as_double(i >= 0 ? i : i ^ 0x7fffffffffffffffLL);
}

__device__ __forceinline__ float db10(float x) { return 10.0f * log10f(x); }
__device__ __forceinline__ double db10(double x) { return 10.0 * log10(x); }

constexpr int kMaxMelsPerLane = 4;   // n_mels <= 128

// Shared-memory footprint of stft_mel_kernel (bytes), mirrored by the host launcher.  P = n_fft / 64.
template <typename T, int P>
__host__ __device__ inline size_t stft_mel_smem_bytes(int nwarps, int frames_per_cta, int hop, int n_mels, int nnz, bool image, bool async) {
    using G = FftGeom<P>;
    const int fcr = (frames_per_cta + G::F - 1) / G::F * G::F;          // a warp always transforms F frames together
    size_t b = 0;
    b += sizeof(FftTables<T, P>);
    if (P < 64) b += G::N * sizeof(T);                                   // window (n_fft 4096 reads it through L1)
    b += ((size_t)(fcr - 1) * hop + G::N) * sizeof(T);                   // staged samples
    if (async) b += ((size_t)(fcr - 1) * hop + G::N + 8) * sizeof(float);   // raw landing zone (async path)
    b += (size_t)nwarps * G::kXbufElems * sizeof(Cpx<T>);                // per-warp transpose buffers
    b += (size_t)4 * kMaxMelsPerLane * 32 * sizeof(int) + (size_t)nnz * sizeof(float); // sparse filterbank (lane slots)
    if (image) b += (size_t)n_mels * (frames_per_cta + 1) * sizeof(T);   // output tile
    return b + 64;
}

template <typename T, int kOut, int kThreads, int P>
__global__ void __launch_bounds__(kThreads, 1) stft_mel_kernel(StftMelParams<T> p) {
    using G = FftGeom<P>;
    constexpr int NFFT = G::N, HALF = G::N / 2, F = G::F, V = G::V;
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int lane = lane_id(), warp = warp_id();
    const int FC = p.frames_per_cta;
    const int span_len = ((FC + F - 1) / F * F - 1) * p.hop + NFFT;
    const int n_mels = p.fb.n_mels;

    unsigned char* sp = smem_raw;
    FftTables<T, P>* tab = reinterpret_cast<FftTables<T, P>*>(sp);  sp += sizeof(FftTables<T, P>);
    constexpr bool kWinSmem = P < 64;
    T* win = reinterpret_cast<T*>(sp);                        if (kWinSmem) sp += NFFT * sizeof(T);
    T* span = reinterpret_cast<T*>(sp);                       sp += (size_t)span_len * sizeof(T);
    const bool kAsync = p.use_async != 0;     // prefetch the next chunk's samples with cp.async while this one is transformed
    float* raw = reinterpret_cast<float*>(sp);                if (kAsync) sp += (size_t)(span_len + 8) * sizeof(float);
    Cpx<T>* xbuf_all = reinterpret_cast<Cpx<T>*>(sp);         sp += (size_t)nwarps * G::kXbufElems * sizeof(Cpx<T>);
    constexpr int kSlotEntries = kMaxMelsPerLane * 32;
    int* fb_start = reinterpret_cast<int*>(sp);               sp += kSlotEntries * sizeof(int);
    int* fb_len = reinterpret_cast<int*>(sp);                 sp += kSlotEntries * sizeof(int);
    int* fb_off = reinterpret_cast<int*>(sp);                 sp += kSlotEntries * sizeof(int);
    int* fb_mel = reinterpret_cast<int*>(sp);                 sp += kSlotEntries * sizeof(int);
    float* fb_w = reinterpret_cast<float*>(sp);               sp += (size_t)p.fb.nnz * sizeof(float);
    T* tile = reinterpret_cast<T*>(sp);                       // kOutImage only

    fill_fft_tables<T, P>(tab, p.tw, p.w2);
    if (kWinSmem)
        for (int i = threadIdx.x; i < NFFT; i += blockDim.x) win[i] = (T)0.5 * p.window[i];   // 1/2 of the Hermitian split, exact
    for (int i = threadIdx.x; i < span_len; i += blockDim.x) span[i] = (T)0;              // frames past a short chunk stay finite
    for (int i = threadIdx.x; i < kSlotEntries; i += blockDim.x) {
        const bool live = i < p.fb.n_slots * 32;
        fb_start[i] = live ? p.fb.start[i] : 0; fb_len[i] = live ? p.fb.len[i] : 0;
        fb_off[i] = live ? p.fb.off[i] : 0;     fb_mel[i] = live ? p.fb.mel[i] : -1;
    }
    for (int i = threadIdx.x; i < p.fb.nnz; i += blockDim.x) fb_w[i] = p.fb.w[i];
    __syncthreads();

    Cpx<T>* xbuf = xbuf_all + (size_t)warp * G::kXbufElems;
    T* pbuf = reinterpret_cast<T*>(xbuf);   // the power spectra alias the transpose buffer

    const long long n_work = (long long)p.N * p.chunks_per_clip;
    // Raw sample range a work item needs: padded index range [t0*hop, t0*hop + need) <-> samples s0 .. s0+need.
    auto issue_prefetch = [&](long long work) {
        const int clip = (int)(work / p.chunks_per_clip);
        const int t0 = (int)(work % p.chunks_per_clip) * FC;
        const int nf = min(FC, p.n_frames - t0);
        const long long s0 = (long long)t0 * p.hop - HALF;
        const long long lo = s0 < 0 ? 0 : s0;
        long long hi = s0 + (nf - 1) * p.hop + NFFT;
        hi = hi > p.n ? p.n : hi;
        const float* src = p.audio + (long long)clip * p.n + lo;
        for (int i = threadIdx.x; i < (int)(hi - lo); i += blockDim.x) cp_async4(raw + i, src + i);
        cp_async_commit();
    };
    if (kAsync && (long long)blockIdx.x < n_work) issue_prefetch(blockIdx.x);
    for (long long work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int clip = (int)(work / p.chunks_per_clip);
        const int chunk = (int)(work % p.chunks_per_clip);
        const int t0 = chunk * FC;
        const int nf = min(FC, p.n_frames - t0);
        const float* src = p.audio + (long long)clip * p.n;
        const float c = p.clip_scale ? p.clip_scale[clip] : 1.0f;
        const int need = (nf - 1) * p.hop + NFFT;
        const long long s0 = (long long)t0 * p.hop - HALF;
        const long long lo = s0 < 0 ? 0 : s0;
        long long hi = s0 + need;
        hi = hi > p.n ? p.n : hi;
        if (kAsync) { cp_async_wait_all(); __syncthreads(); }     // this chunk's raw samples have landed

        // ---- stage the chunk: centre padding, volume normalisation, (onset chain) the two gates
        constexpr bool kGates = sizeof(T) == 8;          // only the float64 onset chain gates its samples
        if (!kGates && s0 >= 0 && s0 + need <= p.n) {
            // no padding inside this chunk (all but the first and last chunks of a clip): plain indices, no branches
            const float* in = kAsync ? raw : src + s0;     // raw[0] is sample lo = s0 here
            if (p.clip_scale) {
#pragma unroll 4
                for (int i = threadIdx.x; i < need; i += blockDim.x) span[i] = (T)__fdiv_rn(in[i], c);
            } else {
#pragma unroll 4
                for (int i = threadIdx.x; i < need; i += blockDim.x) span[i] = (T)in[i];
            }
        } else {
            for (int i = threadIdx.x; i < need; i += blockDim.x) {
                long long s = s0 + i;                              // index into the un-padded clip
                bool inside = s >= 0 && s < p.n;
                if (!inside && p.pad_mode == kPadReflect) {
                    // one mirror is enough when the clip is longer than the padding (the launchers require it)
                    const long long r = s < 0 ? -s : 2 * (p.n - 1) - s;
                    s = (r >= 0 && r < p.n) ? r : reflect_index(s, p.n);
                    inside = true;
                }
                float v = 0.0f;
                if (inside) v = (kAsync && s >= lo && s < hi) ? raw[s - lo] : src[s];
                if (p.clip_scale) v = __fdiv_rn(v, c);
                if (kGates && p.sample_gate > 0.0f && inside) {
                    if (!(fabsf(v) >= p.sample_gate)) v = 0.0f;
                    if (p.frame_gate && !p.frame_gate[(long long)clip * p.gate_stride + s / p.gate_hop]) v = 0.0f;
                }
                span[i] = (T)v;
            }
        }
        __syncthreads();
        if (kAsync && work + gridDim.x < n_work) issue_prefetch(work + gridDim.x);   // overlaps the FFT phase below

        T wmax = (T)-1e300;
        for (int f0 = warp * F; f0 < nf; f0 += nwarps * F) {       // this warp transforms frames f0 .. f0+F-1 together
            Cpx<T> v[V];
            const Cpx<T>* w2v = reinterpret_cast<const Cpx<T>*>(kWinSmem ? win : p.window_half);
#pragma unroll
            for (int r = 0; r < V; ++r) {
                const int j = lane + 32 * (r % P);
                // (x[2j], x[2j+1]) of frame f0 + r/P as one aligned vector
                const Cpx<T>* x2 = reinterpret_cast<const Cpx<T>*>(span + (size_t)(f0 + r / P) * p.hop);
                const Cpx<T> xv = x2[j], wv = w2v[j];
                v[r] = Cpx<T>{xv.x * wv.x, xv.y * wv.y};
            }
            warp_rfft_power<T, P>(v, xbuf, kPbufLead, tab);
            // banded-sparse mel, one filter per (lane, slot), four bins per step; bank-conflict free by construction
#pragma unroll 1
            for (int ff = 0; ff < F; ++ff) {
                const int f = f0 + ff;
                if (f >= nf) break;
                const T* pf = pbuf + ff * G::kPbufStride + kPbufLead;
#pragma unroll
                for (int q = 0; q < kMaxMelsPerLane; ++q) {
                    const int e = q * 32 + lane;
                    const int m = fb_mel[e];
                    const int ln = fb_len[e];
                    const Vec4<T>* pb = reinterpret_cast<const Vec4<T>*>(pf + fb_start[e]);
                    const float4* w = reinterpret_cast<const float4*>(fb_w + fb_off[e]);
                    T acc = (T)0;
                    for (int k = 0; k < ln; ++k) {
                        const Vec4<T> pv = pb[k];
                        const float4 wv = w[k];
                        acc += pv.x * (T)wv.x;
                        acc += pv.y * (T)wv.y;
                        acc += pv.z * (T)wv.z;
                        acc += pv.w * (T)wv.w;
                    }
                    if (m >= 0) {
                        const T db = p.power_out ? acc : db10(acc > p.amin ? acc : p.amin);
                        if (kOut == kOutImage) {
                            tile[m * (FC + 1) + f] = db;
                        } else {
                            p.out[((long long)clip * p.n_frames + t0 + f) * n_mels + m] = db;
                            wmax = db > wmax ? db : wmax;
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (kOut == kOutSpec) {
            wmax = warp_max(wmax);
            if (lane == 0 && nf > warp * F) atomicMax(p.spec_max + clip, ordered_bits((double)wmax));
        }
        __syncthreads();
        if (kOut == kOutImage) {
            // rows of the tile to HBM: a warp (or half a warp when the chunk has <= 16 frames) per mel row
            const int lanes = nf <= 16 ? 16 : 32, rows_per_warp = 32 / lanes;
            for (int m = warp * rows_per_warp + lane / lanes; m < n_mels; m += nwarps * rows_per_warp) {
                T* dst = p.out + ((long long)clip * n_mels + m) * p.n_frames + t0;
                for (int f = lane & (lanes - 1); f < nf; f += lanes) dst[f] = tile[m * (FC + 1) + f];
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// audio/features.py:124-126 : c = sqrt(mean(y**2)) + 1e-9 in float32.  One CTA per clip.
__global__ void clip_scale_kernel(const float* __restrict__ audio, long long n, float* __restrict__ scale_out) {
    __shared__ double red[32];
    const float* y = audio + (long long)blockIdx.x * n;
    double acc = 0.0;
    // 128-bit loads over the 16-byte aligned body of the clip (clips with an odd sample count start anywhere), scalars at the ends
    const long long head = min(n, (long long)((4 - ((reinterpret_cast<unsigned long long>(y) >> 2) & 3)) & 3));
    const long long body = (n - head) / 4;
    const float4* y4 = reinterpret_cast<const float4*>(y + head);
    for (long long i = threadIdx.x; i < body; i += blockDim.x) {
        const float4 v = y4[i];
        acc += (double)__fmul_rn(v.x, v.x); acc += (double)__fmul_rn(v.y, v.y);
        acc += (double)__fmul_rn(v.z, v.z); acc += (double)__fmul_rn(v.w, v.w);
    }
    for (long long i = threadIdx.x; i < head; i += blockDim.x) acc += (double)__fmul_rn(y[i], y[i]);
    for (long long i = head + 4 * body + threadIdx.x; i < n; i += blockDim.x) acc += (double)__fmul_rn(y[i], y[i]);
    acc = warp_sum(acc);
    if (lane_id() == 0) red[warp_id()] = acc;
    __syncthreads();
    if (warp_id() == 0) {
        double v = lane_id() < (int)(blockDim.x >> 5) ? red[lane_id()] : 0.0;
        v = warp_sum(v);
        if (lane_id() == 0) {
            const float mean = (float)(v / (double)n);
            scale_out[blockIdx.x] = __fadd_rn(sqrtf(mean), 1e-9f);
        }
    }
}

// sklearn StandardScaler.transform on the float32 feature matrix (audio/features.py:145-146).  sklearn keeps
// float32 input in float32 and casts mean_/scale_ to it (``X -= xp.astype(self.mean_, X.dtype)``), so the
// result is fl32(fl32(x - fl32(mean)) / fl32(scale)).  (Older sklearn relied on numpy's mixed in-place
// rule, fl32(fl32(double(x) - mean) / scale); the two differ by at most one float32 ulp.)
__global__ void standard_scale_kernel(float* __restrict__ x, int N, int F, int ld,
                                      const double* __restrict__ mean, const double* __restrict__ scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * F) return;
    const int r = i / F, c = i - r * F;
    const float v = x[(long long)r * ld + c];
    x[(long long)r * ld + c] = __fdiv_rn(__fsub_rn(v, (float)mean[c]), (float)scale[c]);
}

}  // namespace gat
