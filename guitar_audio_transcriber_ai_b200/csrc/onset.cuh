// Whole-file segmentation (reference: audio/slicing.py:30-165): noise gate -> frame RMS gate -> spectral
// flux onset strength -> peak picking -> backtracking -> minimum separation -> fixed-length slices.
//
// The reference runs this chain in float64 (the gates multiply by a float64 mask) with several exact
// comparisons, so the arithmetic here follows its dtypes AND its summation orders where a decision hangs
// on them (frame RMS: float32 sequential; mel-band mean: float64 sequential).  The mel spectrogram itself
// comes from stft_mel_kernel<double, kOutSpec> in features.cuh.
//
// Every kernel is batched over P independent signals of equal length (gat_segment_batch: the phrases / files a
// rank owns when an hour of audio is sharded across GPUs, SURVEY 8(e) option (i)); the signal index is
// blockIdx.y (blockIdx.x for the one-CTA-per-signal kernels) and gat_segment is the P = 1 case of the same code.
#pragma once
#include "common.cuh"
#include "features.cuh"

namespace gat {

// ---- librosa.feature.rms(frame 2048, hop, reflect centre padding) on the sample-gated signal, then
//      20*log10(rms + 1e-10) in float32 (slicing.py:44-53).  librosa squares the strided (2048, T) frame
//      view in float32; numpy keeps the frame axis contiguous in the result, so np.mean over it runs
//      numpy's PAIRWISE float32 summation per frame: blocks of 128 values, eight interleaved accumulators
//      per block folded as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), block sums folded as a binary tree.  The kernel
//      reproduces that order exactly (verified bit-for-bit in tests).
//      np.log10 on float32 is glibc's log10f (not correctly rounded, libm-version dependent); we round the
//      float64 log10 instead, which can differ from it by <= 3 float32 ulp (~1e-5 dB).
struct RmsParams {
    const float* y; long long L; int T; int hop; float sample_gate;   // y[P][L]
    float* rms_db;   // [P][T]
};

constexpr int kRmsFramesPerCta = 16;

__global__ void __launch_bounds__(128) rms_db_kernel(RmsParams p) {
    // Eight lanes per frame: lane `sub` IS numpy's interleaved accumulator r[sub] - it adds samples sub, sub + 8, ... of each
    // 128-value block in order; the butterfly over the eight lanes is numpy's ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) (float
    // addition commutes, so every lane of the group ends with the same bits); the 16 block sums are then folded as numpy's
    // balanced tree (left + right).  Same result as the one-thread-per-frame formulation this replaces, 8x the parallelism:
    // 0.48 ms -> a few microseconds for a 5 s phrase, where this kernel was half of the whole call's latency.
    const int lane = lane_id();
    const int sub = lane & 7;
    const int t = (blockIdx.x * 4 + warp_id()) * 4 + (lane >> 3);
    const bool live = t < p.T;
    const float* y = p.y + (long long)blockIdx.y * p.L;
    const long long base = (long long)(live ? t : 0) * p.hop - 1024;
    const bool interior = base >= 0 && base + 2048 <= p.L;
    float bsum[16];
    if (interior) {
        // no padding inside the frame (all but the first and last two): straight indices, fully unrolled
        const float* f = y + base + sub;
#pragma unroll
        for (int blk = 0; blk < 16; ++blk) {
            float r = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float v = f[blk * 128 + i * 8];
                if (!(fabsf(v) >= p.sample_gate)) v = 0.0f;
                const float sq = __fmul_rn(v, v);
                r = i == 0 ? sq : __fadd_rn(r, sq);            // the first eight values of a block initialise the accumulators
            }
            bsum[blk] = r;
        }
    } else {
        // reflect padding (librosa.feature.rms, pad_mode "reflect"): same order, ROLLED - the mirror arithmetic must not be
        // inlined 256 times into the hot path (it was: 23 k instructions, the kernel ran 10x under its issue rate)
#pragma unroll
        for (int blk = 0; blk < 16; ++blk) bsum[blk] = 0.0f;
#pragma unroll 1
        for (int blk = 0; blk < 16; ++blk) {
            float r = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                long long s = base + blk * 128 + i * 8 + sub;
                // one mirror: the entry points require L > 1024, the padding on each side (gat_segment refuses shorter signals)
                s = s < 0 ? -s : (s >= p.L ? 2 * (p.L - 1) - s : s);
                float v = y[s];
                if (!(fabsf(v) >= p.sample_gate)) v = 0.0f;
                const float sq = __fmul_rn(v, v);
                r = i == 0 ? sq : __fadd_rn(r, sq);
            }
#pragma unroll
            for (int b = 0; b < 16; ++b) bsum[b] = b == blk ? r : bsum[b];      // register array: no dynamic indexing
        }
    }
#pragma unroll
    for (int blk = 0; blk < 16; ++blk) {
        float r = bsum[blk];
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        bsum[blk] = r;
    }
#pragma unroll
    for (int w = 1; w < 16; w <<= 1)
#pragma unroll
        for (int i = 0; i < 16; i += 2 * w) bsum[i] = __fadd_rn(bsum[i], bsum[i + w]);
    if (live && sub == 0) {
        const float rms = sqrtf(bsum[0] / 2048.0f);
        const float l = (float)log10((double)__fadd_rn(rms, 1e-10f));
        p.rms_db[(long long)blockIdx.y * p.T + t] = __fmul_rn(20.0f, l);
    }
}

// ---- scipy.ndimage.median_filter(size=5), default mode 'reflect' (edge sample repeated) (slicing.py:55)
__global__ void median5_kernel(const float* __restrict__ in, float* __restrict__ out, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    in += (long long)blockIdx.y * T; out += (long long)blockIdx.y * T;
    float v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        int i = t + k - 2;
        // 'reflect': (d c b a | a b c d | d c b a)
        while (i < 0 || i >= T) { if (i < 0) i = -i - 1; if (i >= T) i = 2 * T - 1 - i; }
        v[k] = in[i];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4 - a; ++b)
            if (v[b] > v[b + 1]) { const float x = v[b]; v[b] = v[b + 1]; v[b + 1] = x; }
    out[t] = v[2];
}

// ---- np.percentile(rms_db, 20) with linear interpolation in float32, gate = p20 + 6 dB, frame mask
//      (slicing.py:59-90).  Exact order statistics by a 4-pass 8-bit radix select; a single CTA.
struct GateParams {
    const float* db; int T;      // db[P][T]; one CTA per signal
    int k_lo;            // floor of the virtual index (T-1)*q computed in float32 on the host
    float gamma;         // its fractional part, float32
    float gate_offset;   // 6.0
    unsigned char* frame_gate;   // [P][T] 1 = keep
    float* gate_out;     // [P] gate level in dB (diagnostic)
};

__device__ __forceinline__ unsigned ordered_u32(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_u32(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// k-th smallest (0-based) of db[0..T); every thread of the CTA returns the same value.
__device__ float radix_select(const float* __restrict__ db, int T, int k, unsigned* hist /*[256] shared*/) {
    unsigned prefix = 0, mask = 0;
    int kk = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < T; i += blockDim.x) {
            const unsigned u = ordered_u32(db[i]);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        // every thread walks the 256 bins (cheap, avoids another broadcast)
        unsigned cum = 0; int bin = 0;
        for (; bin < 256; ++bin) {
            const unsigned h = hist[bin];
            if (cum + h > (unsigned)kk) break;
            cum += h;
        }
        kk -= (int)cum;
        prefix |= (unsigned)bin << shift;
        mask |= 255u << shift;
        __syncthreads();
    }
    return from_ordered_u32(prefix);
}

__global__ void __launch_bounds__(1024) rms_gate_kernel(GateParams p) {
    __shared__ unsigned hist[256];
    const float* db = p.db + (long long)blockIdx.x * p.T;
    unsigned char* frame_gate = p.frame_gate + (long long)blockIdx.x * p.T;
    const float a = radix_select(db, p.T, p.k_lo, hist);
    // the next order statistic without a second four-pass select: it is `a` again when more than k_lo + 1 values are <= a,
    // otherwise the smallest value above a (one pass: a count and a minimum over the CTA)
    float b = a;
    if (p.k_lo + 1 < p.T) {
        __shared__ unsigned s_le, s_min_gt;
        if (threadIdx.x == 0) { s_le = 0u; s_min_gt = 0xffffffffu; }
        __syncthreads();
        unsigned le = 0u, mn = 0xffffffffu;
        for (int i = threadIdx.x; i < p.T; i += blockDim.x) {
            const float v = db[i];
            if (v <= a) ++le;
            else { const unsigned u = ordered_u32(v); mn = u < mn ? u : mn; }
        }
        le = warp_sum(le);
        mn = warp_min(mn);
        if (lane_id() == 0) { atomicAdd(&s_le, le); atomicMin(&s_min_gt, mn); }
        __syncthreads();
        if (s_le <= (unsigned)(p.k_lo + 1)) b = from_ordered_u32(s_min_gt);
    }
    // numpy _lerp in float32: a + (b-a)*t, replaced by b - (b-a)*(1-t) where t >= 0.5
    const float diff = __fsub_rn(b, a);
    float q = __fadd_rn(a, __fmul_rn(diff, p.gamma));
    if (p.gamma >= 0.5f) q = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, p.gamma)));
    const float gate = __fadd_rn(q, p.gate_offset);
    if (threadIdx.x == 0) p.gate_out[blockIdx.x] = gate;
    for (int i = threadIdx.x; i < p.T; i += blockDim.x) frame_gate[i] = db[i] > gate ? 1 : 0;
}

// ---- onset strength (librosa.onset.onset_strength): S' = max(S, max(S) - 80); flux[u] = mean_m
//      relu(S'[u+1][m] - S'[u][m]); env = [0,0,0, flux...][:T].  Also tracks min / max of env.
struct FluxParams {
    const double* spec;        // [P][T][n_mels] mel dB before the top_db clamp
    const long long* spec_max; // [P] ordered bits of each signal's max
    int T, n_mels, lag_pad;    // lag_pad = 1 + 2048/(2*hop) = 3
    double top_db;
    double* env;               // [P][T]
    long long* env_minmax;     // [P][2] ordered bits: min (stored negated for atomicMax), max
};

__global__ void onset_flux_kernel(FluxParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const long long sig = blockIdx.y;
    double e = 0.0;
    if (t < p.T) {
        const int u = t - p.lag_pad;
        if (u >= 0 && u + 1 < p.T) {
            const double floor_db = from_ordered_bits(p.spec_max[sig]) - p.top_db;
            const double* s0 = p.spec + (sig * p.T + u) * p.n_mels;
            const double* s1 = s0 + p.n_mels;
            double acc = 0.0;
            for (int m = 0; m < p.n_mels; ++m) {
                const double a = s0[m] > floor_db ? s0[m] : floor_db;
                const double b = s1[m] > floor_db ? s1[m] : floor_db;
                const double d = b - a;
                acc = __dadd_rn(acc, d > 0.0 ? d : 0.0);
            }
            e = acc / (double)p.n_mels;
        }
        p.env[sig * p.T + t] = e;
    }
    double mx = t < p.T ? e : -1e300, mn = t < p.T ? e : 1e300;
    mx = warp_max(mx); mn = warp_min(mn);
    if (lane_id() == 0) {
        atomicMax(p.env_minmax + 2 * sig + 1, ordered_bits(mx));
        atomicMax(p.env_minmax + 2 * sig + 0, ordered_bits(-mn));
    }
}

// ---- onset_detect normalisation + candidate peaks (librosa.util.peak_pick's two tests)
struct PeakParams {       // all arrays carry a leading signal dimension [P]
    const double* env; const long long* env_minmax; int T;
    int pre_max, post_max, pre_avg, post_avg, wait;
    double delta;              // float32(0.07) promoted
    double* envn;              // [T] normalised envelope
    unsigned* cand;            // [ceil(T/32)] bit t%32 of word t/32 = frame t passes both peak_pick tests
    int* n_peaks; int* peaks;  // outputs of peak_select_kernel: [P], [P][T]
    int* any_nonzero;          // [P]
    int words;                 // mask words per signal = ceil(T/32) + 1
};

__global__ void peak_candidates_kernel(PeakParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const long long sig = blockIdx.y;
    const double* env = p.env + sig * p.T;
    bool ok = false;
    if (t < p.T) {
        const double mn = -from_ordered_bits(p.env_minmax[2 * sig + 0]);
        const double mx = from_ordered_bits(p.env_minmax[2 * sig + 1]);
        const double den = __dadd_rn(__dsub_rn(mx, mn), 2.2250738585072014e-308);
        auto x = [&](int i) { return __ddiv_rn(__dsub_rn(env[i], mn), den); };
        const double xt = x(t);
        p.envn[sig * p.T + t] = xt;
        if (xt != 0.0) atomicExch(p.any_nonzero + sig, 1);
        const int lo_m = t == 0 ? 0 : max(0, t - p.pre_max), hi_m = min(t + p.post_max, p.T);
        double mxw = -1e300;
        for (int i = lo_m; i < hi_m; ++i) { const double v = x(i); mxw = v > mxw ? v : mxw; }
        ok = t == 0 ? (xt >= mxw) : (xt == mxw);
        if (ok) {
            const int lo_a = t == 0 ? 0 : max(0, t - p.pre_avg), hi_a = min(t + p.post_avg, p.T);
            double acc = 0.0;
            for (int i = lo_a; i < hi_a; ++i) acc = __dadd_rn(acc, x(i));
            const double avg = __ddiv_rn(acc, (double)(hi_a - lo_a));
            ok = xt >= __dadd_rn(avg, p.delta);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);       // blockDim is a multiple of 32: word t/32 belongs to this warp
    if (lane_id() == 0 && t < p.T) p.cand[sig * p.words + (t >> 5)] = m;
}

// Sequential `wait` rule: after a peak at n the next frame examined is n + wait + 1.  One CTA per signal.  The candidate
// masks are first COMPACTED into an ordered list of frame indices (popcount + block scan, all threads), so that the only
// sequential part - the wait rule - walks the candidates (one per note) instead of scanning every mask word from global
// memory with one warp (round 1: 0.98 ms for an hour-long file, the largest serial kernel after the STFT).
constexpr int kPeakThreads = 256;

__global__ void __launch_bounds__(kPeakThreads) peak_select_kernel(PeakParams p) {
    __shared__ int warp_tot[kPeakThreads / 32];
    __shared__ int s_base;
    const long long sig = blockIdx.x;
    const unsigned* cand = p.cand + sig * p.words;
    int* peaks = p.peaks + sig * p.T;
    const int lane = lane_id(), warp = warp_id();
    if (p.any_nonzero[sig] == 0) { if (threadIdx.x == 0) p.n_peaks[sig] = 0; return; }
    const int n_words = (p.T + 31) >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    // ---- pass 1: ordered list of candidate frames in peaks[0 .. n_cand)
    for (int w0 = 0; w0 < n_words; w0 += kPeakThreads) {
        const int w = w0 + threadIdx.x;
        unsigned m = w < n_words ? cand[w] : 0u;
        if (w == n_words - 1 && (p.T & 31)) m &= (1u << (p.T & 31)) - 1u;      // frames past T do not exist
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = s_base;
        for (int q = 0; q < warp; ++q) before += warp_tot[q];
        int pos = before + incl - cnt;
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1;
            peaks[pos++] = (w << 5) + b;
        }
        __syncthreads();
        if (threadIdx.x == kPeakThreads - 1) s_base = before + incl;
        __syncthreads();
    }
    // ---- pass 2: the wait rule over the candidates, in place (the output index never overtakes the input index).  The list
    //      is staged through shared memory tile by tile so that the one sequential thread never waits on a global load.
    __shared__ int tile[2048];
    __shared__ int s_count, s_next;
    const int n_cand = s_base;
    if (threadIdx.x == 0) { s_count = 0; s_next = 0; }
    __syncthreads();
    for (int i0 = 0; i0 < n_cand; i0 += 2048) {
        const int nt = min(2048, n_cand - i0);
        for (int i = threadIdx.x; i < nt; i += kPeakThreads) tile[i] = peaks[i0 + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            int count = s_count, next_ok = s_next;
            for (int i = 0; i < nt; ++i) {
                const int n = tile[i];
                if (n >= next_ok) { peaks[count++] = n; next_ok = n + p.wait + 1; }
            }
            s_count = count; s_next = next_ok;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) p.n_peaks[sig] = s_count;
}

// ---- onset_backtrack + frames_to_samples + greedy minimum separation + slice table (slicing.py:109-136,
//      :153-161).  K is small (one entry per note), so a single thread does the sequential parts.
struct SliceParams {      // all arrays carry a leading signal dimension [P]
    const double* envn; int T;
    const int* n_peaks; const int* peaks;      // [P], [P][T]
    int hop; long long L;
    long long min_sep_samples;   // int(min_sep * sr)
    long long skip;              // int(attack_skip_sec * sr)
    long long length;            // int(length_sec * sr)
    int max_onsets;
    int* n_onsets; long long* onsets;          // [P], [P][max_onsets] filtered onset sample positions
    long long* frames_bt;                      // [P][T] backtracked frames (first n_peaks valid; diagnostic)
    long long* table;                          // [P][max_onsets][3] start, end, valid
};

__global__ void backtrack_kernel(SliceParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const long long sig = blockIdx.y;
    if (i >= p.n_peaks[sig]) return;
    const double* envn = p.envn + sig * p.T;
    int f = p.peaks[sig * p.T + i];
    // nearest local minimum at or before the event; index 0 always qualifies (fix_frames pads it)
    while (f > 0) {
        if (f <= p.T - 2 && envn[f] <= envn[f - 1] && envn[f] < envn[f + 1]) break;
        --f;
    }
    p.frames_bt[sig * p.T + i] = f;
}

// Greedy minimum separation is sequential in the kept onset, but cheap once the inputs sit in shared memory:
// the CTA stages the backtracked frames tile by tile, thread 0 runs the greedy rule, then all threads build
// the slice table in parallel.
constexpr int kSepTile = 4096;

__global__ void __launch_bounds__(1024) minsep_table_kernel(SliceParams p) {
    __shared__ long long tile[kSepTile];
    __shared__ int s_k;
    __shared__ long long s_last;
    const long long sig = blockIdx.x;
    const int n = p.n_peaks[sig];
    const long long* frames_bt = p.frames_bt + sig * p.T;
    long long* onsets = p.onsets + sig * p.max_onsets;
    long long* table = p.table + sig * p.max_onsets * 3;
    if (threadIdx.x == 0) { s_k = 0; s_last = -999999; }
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += kSepTile) {
        const int nt = min(kSepTile, n - i0);
        for (int i = threadIdx.x; i < nt; i += blockDim.x) tile[i] = frames_bt[i0 + i] * p.hop;
        __syncthreads();
        if (threadIdx.x == 0) {
            int k = s_k; long long last = s_last;
            for (int i = 0; i < nt; ++i) {
                const long long s = tile[i];
                if (s - last >= p.min_sep_samples && k < p.max_onsets) { onsets[k++] = s; last = s; }
            }
            s_k = k; s_last = last;
        }
        __syncthreads();
    }
    const int k = s_k;
    if (threadIdx.x == 0) p.n_onsets[sig] = k;
    __threadfence_block();
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const long long next = i + 1 < k ? onsets[i + 1] : onsets[k - 1];
        const long long start = onsets[i] + p.skip;
        const long long end = (start + p.length < next) ? start + p.length : next;
        const bool empty = start >= p.L || end > p.L;
        table[3 * i + 0] = start;
        table[3 * i + 1] = end;
        table[3 * i + 2] = empty ? 0 : 1;
    }
}

// ---- is_slice_loud_enough (slicing.py:96-100) on the zero-padded fixed-length clip, compaction, gather.
//      Kept clips are compacted in (signal, onset) order into ONE list: per-signal counts (slice_compact_kernel),
//      an exclusive scan over the signals (slice_base_kernel), then the gather writes rows base[signal] + k.
struct GatherParams {
    const float* y; long long L;             // y[P][L]
    const int* n_onsets; const long long* table;   // [P], [P][max_onsets][3]
    long long length;
    float min_rms_db;            // -37
    unsigned char* keep;         // [P][max_onsets]
    int* dest;                   // [P][max_onsets] position among the signal's kept clips, -1 = dropped
    int* n_clips;                // [P] kept clips per signal, then [P] = total over all signals
    long long* base;             // [P] first row of each signal in the compacted list
    float* clips;                // [max_clips][length]
    long long* clip_table;       // [max_clips][table_cols]: (signal,) onset index, start, end
    int max_onsets;
    int table_cols;              // 3 (gat_segment) or 4 (gat_segment_batch: signal index first)
    long long max_clips;         // rows the caller provided; rows past it are counted but not written
    int P;
};

__global__ void slice_loudness_kernel(GatherParams p) {
    __shared__ double red[8];
    const int i = blockIdx.x;
    const long long sig = blockIdx.y;
    if (i >= p.n_onsets[sig]) return;
    const long long* table = p.table + sig * p.max_onsets * 3;
    const float* y = p.y + sig * p.L;
    const long long start = table[3 * i], end = table[3 * i + 1];
    const bool valid = table[3 * i + 2] != 0;
    double acc = 0.0;
    if (valid)   // python slicing y[start:end] with end < start is empty; the pad makes it all zeros
        for (long long s = start + threadIdx.x; s < end; s += blockDim.x) { const float v = y[s]; acc += (double)__fmul_rn(v, v); }
    acc = warp_sum(acc);
    if (lane_id() == 0) red[warp_id()] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        bool keep = false;
        if (valid) {
            const float rms = sqrtf((float)(s / (double)p.length));
            const float db = __fmul_rn(20.0f, (float)log10((double)__fadd_rn(rms, 1e-10f)));
            keep = db > p.min_rms_db;
        }
        p.keep[sig * p.max_onsets + i] = keep ? 1 : 0;
    }
}

// Order-preserving compaction of one signal's kept clips: one CTA per signal, chunked block scan of the keep flags.
__global__ void __launch_bounds__(1024) slice_compact_kernel(GatherParams p) {
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    const long long sig = blockIdx.x;
    const int n = p.n_onsets[sig];
    const unsigned char* keep_s = p.keep + sig * p.max_onsets;
    int* dest = p.dest + sig * p.max_onsets;
    const int lane = lane_id(), warp = warp_id();
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += (int)blockDim.x) {
        const int i = i0 + threadIdx.x;
        const int keep = i < n ? (int)keep_s[i] : 0;
        int incl = keep;                                    // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (i < n) dest[i] = keep ? before + incl - keep : -1;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_base = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) p.n_clips[sig] = s_base;
}

// Exclusive scan of the per-signal clip counts (one CTA; P is the number of signals of one call), total in n_clips[P].
__global__ void __launch_bounds__(1024) slice_base_kernel(GatherParams p) {
    __shared__ long long warp_tot[32];
    __shared__ long long s_base;
    const int lane = lane_id(), warp = warp_id();
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < p.P; i0 += (int)blockDim.x) {
        const int i = i0 + threadIdx.x;
        const long long cnt = i < p.P ? (long long)p.n_clips[i] : 0;
        long long incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        long long before = s_base;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (i < p.P) p.base[i] = before + incl - cnt;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_base = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) p.n_clips[p.P] = (int)(s_base < 0x7fffffffLL ? s_base : 0x7fffffffLL);
}

__global__ void slice_gather_kernel(GatherParams p) {
    const int i = blockIdx.x;
    const long long sig = blockIdx.y;
    if (i >= p.n_onsets[sig]) return;
    const int d = p.dest[sig * p.max_onsets + i];
    if (d < 0) return;
    const long long row = p.base[sig] + d;
    if (row >= p.max_clips) return;
    const long long* table = p.table + sig * p.max_onsets * 3;
    const float* y = p.y + sig * p.L;
    const long long start = table[3 * i], end = table[3 * i + 1];
    if (threadIdx.x == 0) {
        long long* t = p.clip_table + row * p.table_cols;
        if (p.table_cols == 4) *t++ = sig;
        t[0] = i; t[1] = start; t[2] = end;
    }
    float* o = p.clips + row * p.length;
    for (long long j = threadIdx.x; j < p.length; j += blockDim.x)
        o[j] = (start + j < end) ? y[start + j] : 0.0f;
}

}  // namespace gat
