// Frame-per-warp STFT -> mel -> dB for the float32 feature chains at n_fft = 2048 (the reference's MelSpecConfig.N_FFT and
// librosa's default), round 2 of the kernel behind
//     audio/features.py:296-316, :486-502   torchaudio MelSpectrogram + AmplitudeToDB   (image chain, HTK mel, reflect pad)
//     audio/features.py:187-193, :462-468   librosa.feature.mfcc(...).mean(axis=1)      (spec chain, Slaney mel-128, zero pad)
//
// What changed against stft_mel_kernel (features.cuh), which stays for the float64 onset chain and the other n_fft:
//   * NO block-level staging and NO block barriers.  A warp owns a frame from the first load to the last store: its 2048
//     samples come straight from global memory into the FFT's registers (32 coalesced 64-bit loads per lane; the 8x
//     overlap between neighbouring frames is served by L1/L2, HBM still sees every sample once), so warps drift apart
//     and one warp's shared-memory phases (transpose, mel gather) overlap another's butterflies.  The old kernel ran one
//     frame per warp per 16-frame chunk between three __syncthreads, all warps in the same phase at the same time: ncu
//     showed it waiting on the shared-memory pipe (951 wavefronts per frame, stalls: short scoreboard + MIO throttle).
//   * Volume normalisation costs nothing: |FFT(x / c)|^2 = |FFT(x)|^2 / c^2, so the raw samples are transformed and the
//     mel power is scaled by (1/c)^2 once per mel bin.  (Rounding differs from dividing every sample at the 1e-7
//     relative level, three orders inside the stated dB tolerance.)  No per-sample IEEE division, no staging buffer.
//   * ONE FFT, TWO filterbanks: interior MFCC frames (hop 512) see exactly the samples of every other image frame
//     (hop 256), so in the fused call the Slaney-128 bank is applied to the spectrum the image frame already has; only
//     the frames that touch the clip's ends (zero vs reflect padding: 2 + 2 of 44 at 1 s) get their own FFT.  The shared
//     frames use the image chain's window (torch.hann_window, float32), which differs from librosa's float64 Hann by
//     <= 2e-7: at -140 dB re the frame's peak, far below MFCC's 80 dB top_db clamp.
//   * The MFCC finish (top_db clamp against the clip's maximum, time mean, DCT-II) runs in the same kernel: the warp that
//     delivers a clip's last spectrum frame (global arrival counter) reads the clip's [T][128] mel-dB rows back while
//     they are still in L2 and writes the 64 coefficients.  No second kernel, no HBM read of the intermediate.
#pragma once
#include "features.cuh"

namespace gat {

struct StftFramesParams {
    const float* audio; long long n; int N;
    const float* clip_scale;       // [N] c = rms + 1e-9, read when norm_img / norm_spec
    const float* window_half;      // [2048] 0.5 * window (the Hermitian split's 1/2 folded in, exact)
    const Cpx<float>* tw; const Cpx<float>* w2;
    // ---- image chain: reflect padding, hop `hop`, out[(clip*n_mels + m)*T + t]
    int img; int hop; int T;
    SparseFb fb; int norm_img; int power_out; float amin; float* out;
    // ---- spec chain (MFCC): zero padding, hop 512
    int spec; int T2;
    int share_stride;              // image frames per spec frame (512 / hop) when the chains share FFTs, else 0
    int u_lo, u_hi;                // spec frames u in [u_lo, u_hi) touch no padding: they ride on image frame u * share_stride
    SparseFb fb2; int norm_spec;
    float* spec_scratch;           // [N][T2][128] mel dB (L2-resident between the write and the finish)
    unsigned* clip_count;          // [N] arrival counters, zeroed by the launcher
    const float* dct /* [128][n_mfcc], transposed */; int n_mfcc; float top_db; float* mfcc_out; int ld;
    // ---- work items: per clip `fa` image frames then `fb_items` spec-only frames
    int fa, fb_items;
    long long items_per_cta;
};

constexpr int kStft2P = 32;
constexpr int kStft2N = 2048;

__host__ __device__ inline size_t stft_frames_smem_bytes(int nwarps, int nnz1, int nnz2) {
    size_t b = sizeof(FftTables<float, kStft2P>) + kStft2N * sizeof(float);
    b += (size_t)2 * 4 * kMaxMelsPerLane * 32 * sizeof(int) + ((size_t)nnz1 + nnz2) * sizeof(float) + 32;
    b += (size_t)nwarps * FftGeom<kStft2P>::kXbufElems * sizeof(Cpx<float>);
    return b + 64;
}

struct FbShared { const int* start; const int* len; const int* off; const int* mel; const float* w; };

// Banded-sparse filterbank over the power spectrum `pf` (see SparseFb): lane-slot form, four bins per step, four
// independent accumulators (the gather is latency-bound on its FFMA chain otherwise).
template <typename Store>
__device__ __forceinline__ void apply_filterbank(const float* pf, const FbShared& fb, int n_slots, int lane, float scale,
                                                 float amin, bool power_out, Store&& store) {
    for (int q = 0; q < n_slots; ++q) {
        const int e = q * 32 + lane;
        const int m = fb.mel[e];
        const int ln = fb.len[e];
        const float4* pb = reinterpret_cast<const float4*>(pf + fb.start[e]);
        const float4* w = reinterpret_cast<const float4*>(fb.w + fb.off[e]);
#ifndef GAT_CPU_EMU
        float2 a01 = make_float2(0.0f, 0.0f), a23 = make_float2(0.0f, 0.0f);     // two packed FFMA2 per step
#pragma unroll 2
        for (int k = 0; k < ln; ++k) {
            const float4 pv = pb[k];
            const float4 wv = w[k];
            a01 = __ffma2_rn(make_float2(pv.x, pv.y), make_float2(wv.x, wv.y), a01);
            a23 = __ffma2_rn(make_float2(pv.z, pv.w), make_float2(wv.z, wv.w), a23);
        }
        const float acc = ((a01.x + a01.y) + (a23.x + a23.y)) * scale;
#else
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        for (int k = 0; k < ln; ++k) {
            const float4 pv = pb[k];
            const float4 wv = w[k];
            a0 = fmaf(pv.x, wv.x, a0); a1 = fmaf(pv.y, wv.y, a1);
            a2 = fmaf(pv.z, wv.z, a2); a3 = fmaf(pv.w, wv.w, a3);
        }
        const float acc = ((a0 + a1) + (a2 + a3)) * scale;
#endif
        if (m >= 0) store(m, power_out ? acc : db10(acc > amin ? acc : amin));
    }
}

// librosa.power_to_db's top_db clamp over the clip + time mean + DCT-II (ortho) rows: one warp, after the clip's last
// spectrum frame has arrived.  Same summation orders as mfcc_finish_kernel (per mel band sequential over frames, per
// coefficient sequential over bands).
__device__ __forceinline__ void mfcc_finish_warp(const StftFramesParams& p, int clip, float* scratch_smem /* >= 128 floats */) {
    const int lane = lane_id();
    const float4* rows = reinterpret_cast<const float4*>(p.spec_scratch + (long long)clip * p.T2 * 128);
    float mx = -3.0e38f;
#pragma unroll 4
    for (int t = 0; t < p.T2; ++t) {                    // L2 latency-bound: four rows in flight
#ifndef GAT_CPU_EMU
        const float4 v = __ldcg(rows + (long long)t * 32 + lane);
#else
        const float4 v = rows[(long long)t * 32 + lane];
#endif
        mx = fmaxf(fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)), mx);
    }
    mx = warp_max(mx);
    const float floor_db = mx - p.top_db;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 4
    for (int t = 0; t < p.T2; ++t) {
#ifndef GAT_CPU_EMU
        const float4 v = __ldcg(rows + (long long)t * 32 + lane);
#else
        const float4 v = rows[(long long)t * 32 + lane];
#endif
        s0 += v.x > floor_db ? v.x : floor_db; s1 += v.y > floor_db ? v.y : floor_db;
        s2 += v.z > floor_db ? v.z : floor_db; s3 += v.w > floor_db ? v.w : floor_db;
    }
    const float inv = (float)p.T2;
    __syncwarp();
    reinterpret_cast<float4*>(scratch_smem)[lane] = make_float4(s0 / inv, s1 / inv, s2 / inv, s3 / inv);
    __syncwarp();
    // p.dct is TRANSPOSED, [band j][coefficient k]: lane k reads dct[j][k], one 128-byte line per band, eight bands in
    // flight.  (With the row-major table every lane walked its own row: 32 cache lines per load instruction, and ncu had a
    // fifth of an MFCC-only launch's warp time waiting in this loop.)  Per coefficient the sum still runs over the bands
    // in increasing order.
    for (int k = lane; k < p.n_mfcc; k += 32) {
        const float* d = p.dct + k;
        float acc = 0.0f;
#pragma unroll 8
        for (int j = 0; j < 128; ++j) acc += __ldg(d + (long long)j * p.n_mfcc) * scratch_smem[j];
        p.mfcc_out[(long long)clip * p.ld + k] = acc;
    }
    __syncwarp();
}

// Real FFT of one 2048-sample frame held as 32 complex registers per lane -> power spectrum in shared memory; same
// decomposition and arithmetic as warp_rfft_power<float, 32> (fft.cuh), but the two in-lane 32-point FFTs share ONE copy
// of the butterfly code (a rolled two-trip loop).  Warps of this kernel run unsynchronised, each somewhere else in the
// frame's code: the per-frame instruction footprint has to fit the 32 KB L1.5 instruction cache (the first version,
// fully unrolled with an inlined padding path, was 166 KB and ncu's top stall was "no instruction").
template <typename T>
__device__ __forceinline__ void frame_fft_power(Cpx<T> (&v)[32], Cpx<T>* xbuf, const FftTables<T, kStft2P>* tab) {
    constexpr int C = 1024;
    const int lane = lane_id();
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        fft_dif<T, 32, 0, 32>(v);
        if (pass == 0) {
            // twiddle by W_C^(n1*k2), n1 = lane, and transpose: row k2, column n1
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int k2 = bitrev5(r);
                const Cpx<T> w = tab->tw[k2 * 32 + lane];
                xbuf[k2 * kXbufStride + lane] = (k2 == 0) ? v[r] : cmul(v[r], w);
            }
            __syncwarp();
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[n1] = xbuf[lane * kXbufStride + n1];
            __syncwarp();
        }
    }
    // v[bitrev5(k1)] = Z[32*k1 + lane]; Hermitian partner Z[C-k] sits in lane (32 - lane) % 32 at k1' = 31 - k1
    // (lane 0 pairs with itself at k1' = (32 - k1) % 32).  Bins k and C-k share E and T: |E+T|^2, |E-T|^2.
    T* pb = reinterpret_cast<T*>(xbuf);
    const int partner = (32 - lane) & 31;
    pb[lane] = (T)0;                                           // kPbufLead zeros below bin 0
    if (lane < 4) pb[kPbufLead + C + 1 + lane] = (T)0;         // tail read (times zero weights) by the vectorised mel loop
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const Cpx<T> other = v[bitrev5(31 - k1)];
        const Cpx<T> self = v[bitrev5((32 - k1) & 31)];
        const T bx = __shfl_sync(0xffffffffu, other.x, partner);
        const T by = __shfl_sync(0xffffffffu, other.y, partner);
        const Cpx<T> b = lane == 0 ? self : Cpx<T>{bx, by};
        const int k = 32 * k1 + lane;
        T p_lo, p_hi;
        split_pair_power<T>(v[bitrev5(k1)], b, tab->w2[k], p_lo, p_hi);
        pb[kPbufLead + k] = p_lo;
        pb[kPbufLead + C - k] = p_hi;
    }
    if (lane == 0) {                                           // k = C/2 pairs with itself: W^(C/2) = -i
        const Cpx<T> a = v[bitrev5(16)];
        const T er = a.x + a.x, orr = a.y + a.y;
        pb[kPbufLead + C / 2] = er * er + orr * orr;
    }
    __syncwarp();
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads, 1) stft_frames_kernel(StftFramesParams p) {
    using G = FftGeom<kStft2P>;
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = kThreads / 32;
    const int lane = lane_id(), warp = warp_id();
    unsigned char* sp = smem_raw;
    FftTables<float, kStft2P>* tab = reinterpret_cast<FftTables<float, kStft2P>*>(sp);  sp += sizeof(FftTables<float, kStft2P>);
    float* win = reinterpret_cast<float*>(sp);                 sp += kStft2N * sizeof(float);
    constexpr int kSlotEntries = kMaxMelsPerLane * 32;
    int* fbi = reinterpret_cast<int*>(sp);                     sp += (size_t)2 * 4 * kSlotEntries * sizeof(int);
    float* fbw1 = reinterpret_cast<float*>(sp);                sp += ((size_t)p.fb.nnz + 3) / 4 * 4 * sizeof(float);
    float* fbw2 = reinterpret_cast<float*>(sp);                sp += ((size_t)p.fb2.nnz + 3) / 4 * 4 * sizeof(float);
    Cpx<float>* xbuf = reinterpret_cast<Cpx<float>*>(sp) + (size_t)warp * G::kXbufElems;
    float* pbuf = reinterpret_cast<float*>(xbuf);

    fill_fft_tables<float, kStft2P>(tab, p.tw, p.w2);
    for (int i = threadIdx.x; i < kStft2N; i += kThreads) win[i] = p.window_half[i];
    for (int i = threadIdx.x; i < kSlotEntries; i += kThreads) {
        const bool l1 = p.img && i < p.fb.n_slots * 32, l2 = p.spec && i < p.fb2.n_slots * 32;
        fbi[0 * kSlotEntries + i] = l1 ? p.fb.start[i] : 0;  fbi[1 * kSlotEntries + i] = l1 ? p.fb.len[i] : 0;
        fbi[2 * kSlotEntries + i] = l1 ? p.fb.off[i] : 0;    fbi[3 * kSlotEntries + i] = l1 ? p.fb.mel[i] : -1;
        fbi[4 * kSlotEntries + i] = l2 ? p.fb2.start[i] : 0; fbi[5 * kSlotEntries + i] = l2 ? p.fb2.len[i] : 0;
        fbi[6 * kSlotEntries + i] = l2 ? p.fb2.off[i] : 0;   fbi[7 * kSlotEntries + i] = l2 ? p.fb2.mel[i] : -1;
    }
    if (p.img) for (int i = threadIdx.x; i < p.fb.nnz; i += kThreads) fbw1[i] = p.fb.w[i];
    if (p.spec) for (int i = threadIdx.x; i < p.fb2.nnz; i += kThreads) fbw2[i] = p.fb2.w[i];
    __syncthreads();          // the only block barrier: from here on every warp runs on its own
    const FbShared fb1{fbi, fbi + kSlotEntries, fbi + 2 * kSlotEntries, fbi + 3 * kSlotEntries, fbw1};
    const FbShared fb2{fbi + 4 * kSlotEntries, fbi + 5 * kSlotEntries, fbi + 6 * kSlotEntries, fbi + 7 * kSlotEntries, fbw2};
    const Cpx<float>* win2 = reinterpret_cast<const Cpx<float>*>(win);

    const int ipc = p.fa + p.fb_items;                              // items per clip
    const long long n_items = (long long)p.N * ipc;
    const long long lo = (long long)blockIdx.x * p.items_per_cta;
    long long hi = lo + p.items_per_cta;
    hi = hi > n_items ? n_items : hi;
    const bool small = n_items <= 0x7fffffffLL;                     // 32-bit item arithmetic (a 64-bit division costs ~100 instructions)
    for (long long item = lo + warp; item < hi; item += nwarps) {
        const int clip = small ? (int)((unsigned)item / (unsigned)ipc) : (int)(item / ipc);
        const int i = (int)(item - (long long)clip * ipc);
        // which frame: image frame t (reflect padding), possibly carrying spec frame u too; or a spec-only frame u (zero padding)
        bool do_img = false, do_spec = false;
        int t = 0, u = 0;
        long long s0;
        if (i < p.fa) {
            do_img = true; t = i;
            s0 = (long long)t * p.hop - kStft2N / 2;
            if (p.share_stride > 0 && t % p.share_stride == 0) {
                u = t / p.share_stride;
                do_spec = u >= p.u_lo && u < p.u_hi;
            }
        } else {
            do_spec = true;
            const int e = i - p.fa;
            u = p.share_stride > 0 ? (e < p.u_lo ? e : p.u_hi + (e - p.u_lo)) : e;
            s0 = (long long)u * 512 - kStft2N / 2;
        }
        const float* src = p.audio + (long long)clip * p.n;

        // ---- load the frame into the FFT's registers: z[j] = (x[2j], x[2j+1]) * window/2, j = lane + 32 r
        Cpx<float> v[G::V];
        const bool interior = s0 >= 0 && s0 + kStft2N <= p.n;
        const float2* g = reinterpret_cast<const float2*>(src + s0);          // generic pointer: global here, shared below
        const bool aligned = (reinterpret_cast<unsigned long long>(src + s0) & 7ull) == 0ull;
        if (!interior) {
            // the frame touches the clip's ends: stage it through this warp's scratch with a ROLLED loop - image frames
            // mirror at the ends (torch.stft reflect), spec-only frames see zeros (librosa)
            const bool reflect = do_img;
            const int n32 = (int)p.n, s32 = (int)s0;          // a clip has fewer than 2^31 samples
#pragma unroll 1
            for (int e0 = lane; e0 < kStft2N; e0 += 32 * 8) {        // eight loads in flight per lane: the loop is latency-bound
                float val[8];
#pragma unroll
                for (int u8 = 0; u8 < 8; ++u8) {
                    int s = s32 + e0 + 32 * u8;
                    bool inside = s >= 0 && s < n32;
                    if (!inside && reflect) {        // one mirror suffices: the launcher requires n > n_fft / 2
                        const int m = s < 0 ? -s : 2 * (n32 - 1) - s;
                        inside = m >= 0 && m < n32;
                        s = m;
                    }
                    val[u8] = inside ? src[s] : 0.0f;
                }
#pragma unroll
                for (int u8 = 0; u8 < 8; ++u8) pbuf[e0 + 32 * u8] = val[u8];
            }
            __syncwarp();
            g = reinterpret_cast<const float2*>(pbuf);
        }
        if (interior && !aligned) {
            // clips with an odd sample count start on odd 4-byte boundaries every other clip: 32-bit loads, same registers
            const float* g1 = src + s0 + 2 * lane;
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = Cpx<float>{g1[64 * r], g1[64 * r + 1]};
        } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) { const float2 x = g[lane + 32 * r]; v[r] = Cpx<float>{x.x, x.y}; }
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const Cpx<float> w = win2[lane + 32 * r];
#ifndef GAT_CPU_EMU
            const float2 m = __fmul2_rn(make_float2(v[r].x, v[r].y), make_float2(w.x, w.y));      // packed multiply, same products
            v[r] = Cpx<float>{m.x, m.y};
#else
            v[r].x *= w.x; v[r].y *= w.y;
#endif
        }
        __syncwarp();                                // staged samples are consumed before the transpose reuses the scratch
        frame_fft_power<float>(v, xbuf, tab);
        const float* pf = pbuf + kPbufLead;
        float ic2 = 1.0f;
        if ((do_img && p.norm_img) || (do_spec && p.norm_spec)) { const float ic = 1.0f / p.clip_scale[clip]; ic2 = ic * ic; }

        if (do_img) {
            float* dst = p.out + (long long)clip * p.fb.n_mels * p.T + t;
            const int T = p.T;
            apply_filterbank(pf, fb1, p.fb.n_slots, lane, p.norm_img ? ic2 : 1.0f, p.amin, p.power_out != 0,
                             [&](int m, float val) { dst[(long long)m * T] = val; });
        }
        if (do_spec) {
            float* dst = p.spec_scratch + ((long long)clip * p.T2 + u) * 128;
            apply_filterbank(pf, fb2, p.fb2.n_slots, lane, p.norm_spec ? ic2 : 1.0f, p.amin, false,
                             [&](int m, float val) { dst[m] = val; });
            // the warp's rows are ordered before lane 0's fence by the warp barrier, and the fence (cumulative) before the
            // arrival: one fence per frame instead of thirty-two (the pattern of a grid-wide barrier)
            __syncwarp();
            unsigned prev = 0;
            if (lane == 0) { __threadfence(); prev = atomicAdd(p.clip_count + clip, 1u); }
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == (unsigned)p.T2 - 1u) {        // the clip's last spectrum frame: finish its MFCC row
                __threadfence();
                mfcc_finish_warp(p, clip, pbuf);
            }
        }
        __syncwarp();                                 // pbuf (= xbuf) is rewritten by the next frame's transpose
    }
}

// ---------------------------------------------------------------------------------------------------
// The image chain at n_fft 1024 / 512 (MelSpecConfig.N_FFT is configurable: features.py:296-302; BASELINE config 5 sweeps
// it) in the same frame-per-warp form.  A frame folds to 512 / 256 complex points, 16 / 8 per lane, so a warp transforms
// F = 2 / 4 CONSECUTIVE frames of a clip at once (warp_rfft_power<float, P>, fft.cuh: lane = frame * P + k2 in the second
// pass) and then applies the filterbank to each spectrum in turn.  No block-level staging, no block barriers, volume
// normalisation in the power domain - everything stft_frames_kernel does for n_fft 2048, minus the MFCC chain (librosa's
// n_fft is fixed at 2048).  Replaces the round-1 chunked kernel (stft_mel_kernel) for these sizes.
template <int P>
__host__ __device__ inline size_t stft_frames_small_smem_bytes(int nwarps, int nnz) {
    return sizeof(FftTables<float, P>) + (size_t)64 * P * sizeof(float) + (size_t)4 * kMaxMelsPerLane * 32 * sizeof(int)
         + ((size_t)nnz + 3) / 4 * 4 * sizeof(float) + (size_t)nwarps * FftGeom<P>::kXbufElems * sizeof(Cpx<float>) + 64;
}

template <int kThreads, int P>
__global__ void __launch_bounds__(kThreads, 1) stft_frames_small_kernel(StftFramesParams p) {
    using G = FftGeom<P>;
    constexpr int NF = G::N, F = G::F;                       // samples per frame, frames per warp
    static_assert(G::V == 32 && F * NF == 2048, "one warp holds 2048 samples in 32 complex registers per lane");
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = kThreads / 32;
    const int lane = lane_id(), warp = warp_id();
    unsigned char* sp = smem_raw;
    FftTables<float, P>* tab = reinterpret_cast<FftTables<float, P>*>(sp);  sp += sizeof(FftTables<float, P>);
    float* win = reinterpret_cast<float*>(sp);                 sp += NF * sizeof(float);
    constexpr int kSlotEntries = kMaxMelsPerLane * 32;
    int* fbi = reinterpret_cast<int*>(sp);                     sp += (size_t)4 * kSlotEntries * sizeof(int);
    float* fbw = reinterpret_cast<float*>(sp);                 sp += ((size_t)p.fb.nnz + 3) / 4 * 4 * sizeof(float);
    Cpx<float>* xbuf = reinterpret_cast<Cpx<float>*>(sp) + (size_t)warp * G::kXbufElems;
    float* pbuf = reinterpret_cast<float*>(xbuf);

    fill_fft_tables<float, P>(tab, p.tw, p.w2);
    for (int i = threadIdx.x; i < NF; i += kThreads) win[i] = p.window_half[i];
    for (int i = threadIdx.x; i < kSlotEntries; i += kThreads) {
        const bool live = i < p.fb.n_slots * 32;
        fbi[0 * kSlotEntries + i] = live ? p.fb.start[i] : 0;  fbi[1 * kSlotEntries + i] = live ? p.fb.len[i] : 0;
        fbi[2 * kSlotEntries + i] = live ? p.fb.off[i] : 0;    fbi[3 * kSlotEntries + i] = live ? p.fb.mel[i] : -1;
    }
    for (int i = threadIdx.x; i < p.fb.nnz; i += kThreads) fbw[i] = p.fb.w[i];
    __syncthreads();          // the only block barrier
    const FbShared fb1{fbi, fbi + kSlotEntries, fbi + 2 * kSlotEntries, fbi + 3 * kSlotEntries, fbw};
    const Cpx<float>* win2 = reinterpret_cast<const Cpx<float>*>(win);

    const int gpc = (p.T + F - 1) / F;                               // groups of F frames per clip
    const long long n_items = (long long)p.N * gpc;
    const long long lo = (long long)blockIdx.x * p.items_per_cta;
    long long hi = lo + p.items_per_cta;
    hi = hi > n_items ? n_items : hi;
    const int n32 = (int)p.n;
    for (long long item = lo + warp; item < hi; item += nwarps) {
        const int clip = (int)(item / gpc);
        const int t0 = (int)(item - (long long)clip * gpc) * F;
        const float* src = p.audio + (long long)clip * p.n;
        const int s_first = t0 * p.hop - NF / 2;                     // first sample of the group's first frame
        // every frame of the group inside the clip (and a real frame): straight from global memory into the FFT's registers
        const bool interior = s_first >= 0 && s_first + (F - 1) * p.hop + NF <= n32 && t0 + F <= p.T;
        const bool aligned = (reinterpret_cast<unsigned long long>(src + s_first) & 7ull) == 0ull && (p.hop & 1) == 0;
        Cpx<float> v[32];
        if (interior && aligned) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float2* g = reinterpret_cast<const float2*>(src + s_first + f * p.hop);
#pragma unroll
                for (int r = 0; r < P; ++r) { const float2 x = g[lane + 32 * r]; v[f * P + r] = Cpx<float>{x.x, x.y}; }
            }
        } else if (interior) {                                       // clips that start on odd words: 32-bit loads
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float* g1 = src + s_first + f * p.hop + 2 * lane;
#pragma unroll
                for (int r = 0; r < P; ++r) v[f * P + r] = Cpx<float>{g1[64 * r], g1[64 * r + 1]};
            }
        } else {
            // a frame touches the clip's ends (torch.stft reflect) or lies past the last frame (zeros, never stored):
            // stage the F frames through the warp's scratch with a rolled loop
#pragma unroll 1
            for (int f = 0; f < F; ++f) {
                const int sf = s_first + f * p.hop;
                const bool real = t0 + f < p.T;
#pragma unroll 1
                for (int e0 = lane; e0 < NF; e0 += 32 * 8) {
                    float val[8];
#pragma unroll
                    for (int u8 = 0; u8 < 8; ++u8) {
                        int s = sf + e0 + 32 * u8;
                        bool inside = s >= 0 && s < n32;
                        if (!inside) {                               // one mirror suffices: the launcher requires n > n_fft / 2
                            const int m = s < 0 ? -s : 2 * (n32 - 1) - s;
                            inside = m >= 0 && m < n32;
                            s = m;
                        }
                        val[u8] = (inside && real && e0 + 32 * u8 < NF) ? src[s] : 0.0f;
                    }
#pragma unroll
                    for (int u8 = 0; u8 < 8; ++u8) if (e0 + 32 * u8 < NF) pbuf[f * NF + e0 + 32 * u8] = val[u8];
                }
            }
            __syncwarp();
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float2* g = reinterpret_cast<const float2*>(pbuf + f * NF);
#pragma unroll
                for (int r = 0; r < P; ++r) { const float2 x = g[lane + 32 * r]; v[f * P + r] = Cpx<float>{x.x, x.y}; }
            }
            __syncwarp();                            // staged samples are consumed before the transpose reuses the scratch
        }
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const Cpx<float> w = win2[lane + 32 * r];
#pragma unroll
            for (int f = 0; f < F; ++f) { v[f * P + r].x *= w.x; v[f * P + r].y *= w.y; }
        }
        warp_rfft_power<float, P>(v, xbuf, kPbufLead, tab);
        float ic2 = 1.0f;
        if (p.norm_img) { const float ic = 1.0f / p.clip_scale[clip]; ic2 = ic * ic; }
        const int T = p.T;
#pragma unroll 1
        for (int f = 0; f < F; ++f) {
            const int t = t0 + f;
            if (t < T) {
                float* dst = p.out + (long long)clip * p.fb.n_mels * T + t;
                apply_filterbank(pbuf + f * G::kPbufStride + kPbufLead, fb1, p.fb.n_slots, lane, ic2, p.amin, p.power_out != 0,
                                 [&](int m, float val) { dst[(long long)m * T] = val; });
            }
        }
        __syncwarp();                                 // pbuf (= xbuf) is rewritten by the next group
    }
}

// ---------------------------------------------------------------------------------------------------
// The float64 mel spectrogram of the onset chain (audio/slicing.py:107, librosa.onset.onset_strength on the float64 gated
// signal) in the same frame-per-warp form: zero centre padding, any even hop, the two gates of sliceNsave fused into the
// loads (sample gate |y| >= g, per-block frame gate), Slaney mel-128, 10 log10, per-signal running maximum.
// The chunked kernel it replaces (stft_mel_kernel<double>) ran six warps in lock step between block barriers and was the
// largest kernel of the hour-long configuration; FP64 arithmetic is latency-bound at this occupancy (255 registers), so
// what helps is two more warps (no staging buffers) and warps that do not wait for each other.
struct OnsetFramesParams {
    const float* audio; long long n; int N;          // [N][n] signals
    const unsigned char* frame_gate;                 // [N][gate_stride] keep flags per gate_hop samples, or nullptr
    long long gate_stride; int gate_hop; float sample_gate;   // sample_gate == 0 disables both gates
    int gate_shift;                                  // log2(gate_hop) when it is a power of two (512), else -1
    int hop; int T;                                  // frames per signal = 1 + n / hop
    const double* window;                            // [2048] float64 Hann (scipy)
    const Cpx<double>* tw; const Cpx<double>* w2;
    SparseFb fb; double amin;
    double* out;                                     // [N][T][n_mels] mel dB
    long long* spec_max;                             // [N] ordered-integer encoded maxima
    long long items_per_cta;
};

__host__ __device__ inline size_t onset_frames_smem_bytes(int nwarps, int nnz) {
    return sizeof(FftTables<double, kStft2P>) + kStft2N * sizeof(double) + (size_t)4 * kMaxMelsPerLane * 32 * sizeof(int)
         + ((size_t)nnz + 3) / 4 * 4 * sizeof(float) + (size_t)nwarps * FftGeom<kStft2P>::kXbufElems * sizeof(Cpx<double>) + 64;
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads, 1) onset_frames_kernel(OnsetFramesParams p) {
    using G = FftGeom<kStft2P>;
    GAT_DYN_SMEM(smem_raw);
    const int nwarps = kThreads / 32;
    const int lane = lane_id(), warp = warp_id();
    unsigned char* sp = smem_raw;
    FftTables<double, kStft2P>* tab = reinterpret_cast<FftTables<double, kStft2P>*>(sp);  sp += sizeof(FftTables<double, kStft2P>);
    double* win = reinterpret_cast<double*>(sp);               sp += kStft2N * sizeof(double);
    constexpr int kSlotEntries = kMaxMelsPerLane * 32;
    int* fb_start = reinterpret_cast<int*>(sp);                sp += kSlotEntries * sizeof(int);
    int* fb_len = reinterpret_cast<int*>(sp);                  sp += kSlotEntries * sizeof(int);
    int* fb_off = reinterpret_cast<int*>(sp);                  sp += kSlotEntries * sizeof(int);
    int* fb_mel = reinterpret_cast<int*>(sp);                  sp += kSlotEntries * sizeof(int);
    float* fb_w = reinterpret_cast<float*>(sp);                sp += ((size_t)p.fb.nnz + 3) / 4 * 4 * sizeof(float);
    Cpx<double>* xbuf = reinterpret_cast<Cpx<double>*>(sp) + (size_t)warp * G::kXbufElems;
    double* pbuf = reinterpret_cast<double*>(xbuf);
    float* stage = reinterpret_cast<float*>(xbuf);             // 2048 floats of the warp's scratch while a padded frame is staged

    fill_fft_tables<double, kStft2P>(tab, p.tw, p.w2);
    for (int i = threadIdx.x; i < kStft2N; i += kThreads) win[i] = 0.5 * p.window[i];     // 1/2 of the Hermitian split, exact
    for (int i = threadIdx.x; i < kSlotEntries; i += kThreads) {
        const bool live = i < p.fb.n_slots * 32;
        fb_start[i] = live ? p.fb.start[i] : 0; fb_len[i] = live ? p.fb.len[i] : 0;
        fb_off[i] = live ? p.fb.off[i] : 0;     fb_mel[i] = live ? p.fb.mel[i] : -1;
    }
    for (int i = threadIdx.x; i < p.fb.nnz; i += kThreads) fb_w[i] = p.fb.w[i];
    __syncthreads();
    const Cpx<double>* win2 = reinterpret_cast<const Cpx<double>*>(win);
    const bool gated = p.sample_gate > 0.0f;

    const long long n_items = (long long)p.N * p.T;
    const long long lo = (long long)blockIdx.x * p.items_per_cta;
    long long hi = lo + p.items_per_cta;
    hi = hi > n_items ? n_items : hi;
    const bool small = n_items <= 0x7fffffffLL;
    for (long long item = lo + warp; item < hi; item += nwarps) {
        const int clip = small ? (int)((unsigned)item / (unsigned)p.T) : (int)(item / p.T);
        const int t = (int)(item - (long long)clip * p.T);
        const long long s0 = (long long)t * p.hop - kStft2N / 2;
        const float* src = p.audio + (long long)clip * p.n;
        const unsigned char* fg = p.frame_gate ? p.frame_gate + (long long)clip * p.gate_stride : nullptr;
        // apply_db_threshold then apply_rms_threshold (slicing.py:30-39, :78-91); positions inside a signal fit 32 bits, and
        // the block index is a shift for the reference's hop of 512 (a 64-bit division per sample dominated the first version)
        auto gate = [&](float v, int s) {
            if (gated) {
                if (!(fabsf(v) >= p.sample_gate)) v = 0.0f;
                if (fg && !fg[p.gate_shift >= 0 ? (s >> p.gate_shift) : (s / p.gate_hop)]) v = 0.0f;
            }
            return v;
        };
        Cpx<double> v[G::V];
        const bool interior = s0 >= 0 && s0 + kStft2N <= p.n;
        const bool aligned = (reinterpret_cast<unsigned long long>(src + s0) & 7ull) == 0ull;
        if (interior && aligned && (!fg || (p.gate_shift >= 1))) {       // other gate hops go through the staged path below
            const float2* g = reinterpret_cast<const float2*>(src + s0);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int j = lane + 32 * r;
                const float2 x = g[j];
                const Cpx<double> w = win2[j];
                const int s = (int)s0 + 2 * j;                // both samples of the pair lie in the same gate block (hop even)
                float gx = x.x, gy = x.y;
                if (gated) {
                    if (!(fabsf(gx) >= p.sample_gate)) gx = 0.0f;
                    if (!(fabsf(gy) >= p.sample_gate)) gy = 0.0f;
                    if (fg && !fg[s >> p.gate_shift]) { gx = 0.0f; gy = 0.0f; }
                }
                v[r] = Cpx<double>{(double)gx * w.x, (double)gy * w.y};
            }
        } else {
            // padded (zeros outside the signal) or misaligned: stage the gated samples as floats through the warp's scratch
#pragma unroll 1
            for (int e0 = lane; e0 < kStft2N; e0 += 32 * 8) {        // eight loads in flight per lane
                float val[8];
#pragma unroll
                for (int u8 = 0; u8 < 8; ++u8) {
                    const long long s = s0 + e0 + 32 * u8;
                    val[u8] = (s >= 0 && s < p.n) ? src[s] : 0.0f;
                }
#pragma unroll
                for (int u8 = 0; u8 < 8; ++u8) {
                    const long long s = s0 + e0 + 32 * u8;
                    stage[e0 + 32 * u8] = (s >= 0 && s < p.n) ? gate(val[u8], (int)s) : 0.0f;
                }
            }
            __syncwarp();
            const float2* g = reinterpret_cast<const float2*>(stage);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int j = lane + 32 * r;
                const float2 x = g[j];
                const Cpx<double> w = win2[j];
                v[r] = Cpx<double>{(double)x.x * w.x, (double)x.y * w.y};
            }
            __syncwarp();
        }
        frame_fft_power<double>(v, xbuf, tab);
        const double* pf = pbuf + kPbufLead;
        // banded-sparse Slaney filterbank: four interleaved partial sums per filter (bin mod 4), folded as (a0 + a1) + (a2 + a3).
        // (One running sum was a chain of four dependent DFMAs per step at two warps per scheduler.  Widening the weights to
        // float64 in shared memory instead of converting them at every use was tried and lost: the gather is half of the
        // kernel's shared-memory wavefronts, and 32-byte weight loads made it 5 % slower.)
        double wmax = -1e300;
        double* dst = p.out + ((long long)clip * p.T + t) * p.fb.n_mels;
        for (int q = 0; q < p.fb.n_slots; ++q) {
            const int e = q * 32 + lane;
            const int m = fb_mel[e];
            const int ln = fb_len[e];
            const Vec4<double>* pb = reinterpret_cast<const Vec4<double>*>(pf + fb_start[e]);
            const float4* w = reinterpret_cast<const float4*>(fb_w + fb_off[e]);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
            for (int k = 0; k < ln; ++k) {
                const Vec4<double> pv = pb[k];
                const float4 wv = w[k];
                a0 += pv.x * (double)wv.x;
                a1 += pv.y * (double)wv.y;
                a2 += pv.z * (double)wv.z;
                a3 += pv.w * (double)wv.w;
            }
            const double acc = (a0 + a1) + (a2 + a3);
            if (m >= 0) {
                const double db = db10(acc > p.amin ? acc : p.amin);
                dst[m] = db;
                wmax = db > wmax ? db : wmax;
            }
        }
        wmax = warp_max(wmax);
        if (lane == 0) atomicMax(p.spec_max + clip, ordered_bits(wmax));
        __syncwarp();                                 // pbuf (= xbuf) is rewritten by the next frame
    }
}

}  // namespace gat
