// Warp-level real FFT of 512 / 1024 / 2048 / 4096 samples, data in registers (STFT, MFCC and onset kernels).
//
// An N-point real frame is folded into C = N/2 complex points z[j] = x[2j] + i x[2j+1], C = 32 * P.  One warp
// transforms it as P x 32: every lane runs a P-point FFT in registers (over n2, stride 32), the result is
// twiddled by W_C^(n1*k2) and transposed through shared memory, every lane runs a 32-point FFT (over n1), and the
// Hermitian split X[k] = E[k] + W_N^k O[k] recovers bins 0..C (two at a time).  No block barrier is involved:
// only __syncwarp.
//   P = 32 (N = 2048, the reference's n_fft): one frame per warp, 32 complex registers per lane.
//   P = 16 / 8 (N = 1024 / 512): the warp transforms F = 32/P frames at once, so the second pass still has one
//                32-point FFT per lane (lane = frame * P + k2) and the register budget is unchanged.
//   P = 64 (N = 4096): one frame per warp, 64 complex registers per lane, two 32-point FFTs per lane in pass 2.
// T is float (feature chains) or double (onset chain, which the reference runs in float64: slicing.py:37,90
// promote the gated signal).
#pragma once
#include "common.cuh"

namespace gat {

// 2*sizeof(T)-aligned so that every complex load / store is ONE 64-bit (float) or 128-bit (double) access: with the
// natural 4-byte alignment the compiler split them into pairs of 32-bit shared-memory accesses (2-way bank conflicts
// on the stride-2 patterns, twice the LSU instructions).
template <typename T> struct __align__(2 * sizeof(T)) Cpx { T x, y; };

template <typename T> __device__ __forceinline__ Cpx<T> cadd(Cpx<T> a, Cpx<T> b) { return Cpx<T>{a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ Cpx<T> csub(Cpx<T> a, Cpx<T> b) { return Cpx<T>{a.x - b.x, a.y - b.y}; }
#ifndef GAT_CPU_EMU
// float32: a complex add / subtract is ONE packed instruction on sm_100 (FADD2 = add.rn.f32x2; the subtraction is
// fma(b, -1, a): the product is exact, so it rounds once like a - b).  Same IEEE results, half the issue slots.
__device__ __forceinline__ Cpx<float> cadd(Cpx<float> a, Cpx<float> b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return Cpx<float>{r.x, r.y};
}
__device__ __forceinline__ Cpx<float> csub(Cpx<float> a, Cpx<float> b) {
    const float2 r = __ffma2_rn(make_float2(b.x, b.y), make_float2(-1.0f, -1.0f), make_float2(a.x, a.y));
    return Cpx<float>{r.x, r.y};
}
#endif
template <typename T> __device__ __forceinline__ Cpx<T> cmul(Cpx<T> a, Cpx<T> b) {
    return Cpx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// d * (c - i s) with real constants c, s (a butterfly's twiddle)
template <typename T> __device__ __forceinline__ Cpx<T> ctwid(Cpx<T> d, T c, T s) {
    return Cpx<T>{d.x * c + d.y * s, d.y * c - d.x * s};
}
#ifndef GAT_CPU_EMU
// float32: one packed multiply by the broadcast scalar, then two FMAs - three instructions instead of four.
__device__ __forceinline__ Cpx<float> cmul(Cpx<float> a, Cpx<float> b) {
    const float2 t = __fmul2_rn(make_float2(b.x, b.y), make_float2(a.x, a.x));
    return Cpx<float>{fmaf(-a.y, b.y, t.x), fmaf(a.y, b.x, t.y)};
}
__device__ __forceinline__ Cpx<float> ctwid(Cpx<float> d, float c, float s) {
    const float2 t = __fmul2_rn(make_float2(d.x, d.y), make_float2(c, c));
    return Cpx<float>{fmaf(d.y, s, t.x), fmaf(-d.x, s, t.y)};
}
#endif

// cos(2*pi*j/64), j = 0..16
template <typename T> __host__ __device__ constexpr T cos64_tab(int j) {
    return j == 0 ? T(1.0)
         : j == 1 ? T(0.99518472667219688624483695310948)
         : j == 2 ? T(0.98078528040323044912618223613424)
         : j == 3 ? T(0.95694033573220886493579788698027)
         : j == 4 ? T(0.92387953251128675612818318939679)
         : j == 5 ? T(0.88192126434835502971275686366039)
         : j == 6 ? T(0.83146961230254523707878837761791)
         : j == 7 ? T(0.77301045336273696081090660975847)
         : j == 8 ? T(0.70710678118654752440084436210485)
         : j == 9 ? T(0.63439328416364549821517161322549)
         : j == 10 ? T(0.55557023301960222474283081394853)
         : j == 11 ? T(0.47139673682599764855638762590525)
         : j == 12 ? T(0.38268343236508977172845998403040)
         : j == 13 ? T(0.29028467725446236763619237581740)
         : j == 14 ? T(0.19509032201612826784828486847702)
         : j == 15 ? T(0.09801714032956060199419556388864)
         : T(0.0);
}
// cos / sin of 2*pi*j/64 for any j in [0, 64)
template <typename T> __host__ __device__ constexpr T cos64(int j) {
    return j <= 16 ? cos64_tab<T>(j) : j <= 32 ? -cos64_tab<T>(32 - j) : j <= 48 ? -cos64_tab<T>(j - 32) : cos64_tab<T>(64 - j);
}
template <typename T> __host__ __device__ constexpr T sin64(int j) { return cos64<T>((j + 48) & 63); }  // sin(a) = cos(a - pi/2)

__host__ __device__ constexpr int bitrev_bits(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int bitrev5(int v) { return bitrev_bits(v, 5); }
__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// One radix-2 DIF butterfly with the compile-time twiddle W_64^J = exp(-2 pi i J / 64).
template <typename T, int J>
__device__ __forceinline__ void butterfly(Cpx<T>& a, Cpx<T>& b) {
    Cpx<T> s = cadd(a, b);
    Cpx<T> d = csub(a, b);
    a = s;
    if (J == 0) {
        b = d;
    } else if (J == 16) {           // multiply by -i
        b = Cpx<T>{d.y, -d.x};
    } else if (J == 8) {            // (1 - i)/sqrt(2)
        constexpr T h = cos64_tab<T>(8);
        b = Cpx<T>{(d.x + d.y) * h, (d.y - d.x) * h};
    } else if (J == 24) {           // (-1 - i)/sqrt(2)
        constexpr T h = cos64_tab<T>(8);
        b = Cpx<T>{(d.y - d.x) * h, -(d.x + d.y) * h};
    } else {
        constexpr T c = cos64<T>(J);
        constexpr T s_ = sin64<T>(J);
        b = ctwid(d, c, s_);                                   // d * (c - i s)
    }
}

// All butterflies of the stage with block size M of the P-point FFT held in v[BASE .. BASE+P).
template <typename T, int V, int BASE, int P, int M, int B, int J>
struct StageLoop {
    __device__ static __forceinline__ void run(Cpx<T> (&v)[V]) {
        butterfly<T, J * (64 / M)>(v[BASE + B + J], v[BASE + B + J + M / 2]);
        if constexpr (J + 1 < M / 2) StageLoop<T, V, BASE, P, M, B, J + 1>::run(v);
        else if constexpr (B + M < P) StageLoop<T, V, BASE, P, M, B + M, 0>::run(v);
    }
};

// In-register P-point forward FFT on v[BASE .. BASE+P), decimation in frequency: natural order in, X[k] ends up
// in v[BASE + bitrev(k)].
template <typename T, int V, int BASE, int P>
__device__ __forceinline__ void fft_dif(Cpx<T> (&v)[V]) {
    if constexpr (P >= 64) StageLoop<T, V, BASE, P, 64, 0, 0>::run(v);
    if constexpr (P >= 32) StageLoop<T, V, BASE, P, 32, 0, 0>::run(v);
    if constexpr (P >= 16) StageLoop<T, V, BASE, P, 16, 0, 0>::run(v);
    if constexpr (P >= 8) StageLoop<T, V, BASE, P, 8, 0, 0>::run(v);
    if constexpr (P >= 4) StageLoop<T, V, BASE, P, 4, 0, 0>::run(v);
    StageLoop<T, V, BASE, P, 2, 0, 0>::run(v);
}

constexpr int kXbufStride = 33;                       // complex elements per row of the transpose buffer

template <int P>
struct FftGeom {
    static constexpr int N = 64 * P;                  // real samples per frame
    static constexpr int C = 32 * P;                  // folded complex points; bins 0..C
    static constexpr int F = P >= 32 ? 1 : 32 / P;    // frames one warp transforms together
    static constexpr int V = P * F;                   // complex registers per lane (32, or 64 for P = 64)
    static constexpr int LOGP = ilog2(P);
    static constexpr int kXbufElems = V * kXbufStride;            // complex elements of scratch per warp
    // elements of T between the power spectra of the F frames inside the scratch; chosen so that the lanes of
    // different frames (which write the same bin at the same time) land in different banks
    static constexpr int kPbufStride = P == 8 ? 520 : P == 16 ? 1040 : 2 * kXbufElems;
};

// Tables every kernel that uses warp_rfft_power keeps in shared memory (filled by fill_fft_tables).
template <typename T, int P>
struct FftTables {
    Cpx<T> tw[P * 32];    // tw[k2*32 + n1] = W_C^(n1*k2)
    Cpx<T> w2[16 * P];    // w2[k] = W_N^k, k < C/2
};

template <typename T, int P>
__device__ void fill_fft_tables(FftTables<T, P>* tab, const Cpx<T>* __restrict__ g_tw, const Cpx<T>* __restrict__ g_w2) {
    for (int i = threadIdx.x; i < P * 32; i += blockDim.x) tab->tw[i] = g_tw[i];
    for (int i = threadIdx.x; i < P * 16; i += blockDim.x) tab->w2[i] = g_w2[i];
}

// |E + T|^2 and |E - T|^2 for the bin pair (k, C-k): a = Z[k], b = Z[C-k], w = W_N^k.
template <typename T>
__device__ __forceinline__ void split_pair_power(Cpx<T> a, Cpx<T> b, Cpx<T> w, T& p_lo, T& p_hi) {
    const T er = a.x + b.x, ei = a.y - b.y;           // E = A + conj B
    const T orr = a.y + b.y, oi = b.x - a.x;          // O = -i (A - conj B)
    const T tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;
    const T pr = er + tr, pi = ei + ti, mr = er - tr, mi = ei - ti;
    p_lo = pr * pr + pi * pi;
    p_hi = mr * mr + mi * mi;
}

// Forward real FFT of F frames of N = 64*P samples by one warp, power spectra left in shared memory.
//   v[f*P + n2] in : z_f[lane + 32*n2] = (x[2j], x[2j+1]) * 0.5 * window, j = lane + 32*n2.  The caller folds
//                    the Hermitian split's factor 1/2 into the window (an exact power-of-two scaling).
//   xbuf           : per-warp scratch of kXbufElems complex values; on return, reinterpreted as T[], frame f's
//                    spectrum sits at pbuf_f = pbuf + f*kPbufStride: pbuf_f[lead + k] = |X[k]|^2 for k = 0..C,
//                    pbuf_f[0..lead) = 0 and four zeros after bin C.
// Bins k and C-k share E = Z[k] + conj Z[C-k] and T = W_N^k * O: |X[k]|^2 = |E+T|^2 and |X[C-k]|^2 = |E-T|^2,
// so each lane handles 16 (32 for P = 64) pairs instead of single bins.
template <typename T, int P>
__device__ __forceinline__ void warp_rfft_power(Cpx<T> (&v)[FftGeom<P>::V], Cpx<T>* xbuf, int lead, const FftTables<T, P>* tab) {
    using G = FftGeom<P>;
    constexpr int V = G::V, C = G::C;
    const int lane = lane_id();
    // pass 1: P-point FFTs over n2 (in registers), one per frame
    fft_dif<T, V, 0, P>(v);
    if constexpr (G::F >= 2) fft_dif<T, V, P, P>(v);
    if constexpr (G::F >= 4) { fft_dif<T, V, 2 * P, P>(v); fft_dif<T, V, 3 * P, P>(v); }
    // twiddle by W_C^(n1*k2), n1 = lane, and transpose: row (frame, k2), column n1
#pragma unroll
    for (int r = 0; r < V; ++r) {
        const int k2 = bitrev_bits(r % P, G::LOGP);
        const int row = (r / P) * P + k2;
        Cpx<T> w = tab->tw[k2 * 32 + lane];
        xbuf[row * kXbufStride + lane] = (k2 == 0) ? v[r] : cmul(v[r], w);
    }
    __syncwarp();
    T* pbuf = reinterpret_cast<T*>(xbuf);
    if constexpr (P <= 32) {
        // pass 2: lane = frame*P + k2 reads its row over n1
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) v[n1] = xbuf[lane * kXbufStride + n1];
        __syncwarp();
        fft_dif<T, V, 0, 32>(v);
        // v[bitrev5(k1)] = Z[P*k1 + k2].  The Hermitian partner Z[C-k] of k = P*k1 + k2 sits in the lane of the same
        // frame with k2' = (P - k2) % P at k1' = 31 - k1 (k2 = 0 pairs with itself: k1' = (32 - k1) % 32).
        const int k2 = lane & (P - 1);
        const int partner = (lane & ~(P - 1)) | ((P - k2) & (P - 1));
        Cpx<T> b[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const Cpx<T> other = v[bitrev5(31 - k1)];
            const Cpx<T> self = v[bitrev5((32 - k1) & 31)];
            const T bx = __shfl_sync(0xffffffffu, other.x, partner);
            const T by = __shfl_sync(0xffffffffu, other.y, partner);
            b[k1] = k2 == 0 ? self : Cpx<T>{bx, by};
        }
        // xbuf was last READ before the second FFT (with a __syncwarp after), so it can take the spectra now
        T* pb = pbuf + (lane / P) * G::kPbufStride;
        for (int i = k2; i < lead; i += P) pb[i] = (T)0;
        if (k2 < 4) pb[lead + C + 1 + k2] = (T)0;             // tail read (times zero weights) by the vectorised mel loop
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const int k = P * k1 + k2;
            T p_lo, p_hi;
            split_pair_power<T>(v[bitrev5(k1)], b[k1], tab->w2[k], p_lo, p_hi);
            pb[lead + k] = p_lo;
            pb[lead + C - k] = p_hi;
        }
        if (k2 == 0) {                                        // k = C/2 pairs with itself: W^(C/2) = -i
            const Cpx<T> a = v[bitrev5(16)];
            const T er = a.x + a.x, orr = a.y + a.y;          // E = (2 Re a, 0), O = (2 Im a, 0), T = W*O = (0, -2 Im a)
            pb[lead + C / 2] = er * er + orr * orr;
        }
    } else {
        // pass 2 (P = 64): lane reads rows k2 = lane (-> v[0..32)) and k2 = lane + 32 (-> v[32..64))
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            v[n1] = xbuf[lane * kXbufStride + n1];
            v[32 + n1] = xbuf[(lane + 32) * kXbufStride + n1];
        }
        __syncwarp();
        fft_dif<T, V, 0, 32>(v);
        fft_dif<T, V, 32, 32>(v);
        // v[bitrev5(k1)] = Z[64*k1 + lane], v[32 + bitrev5(k1)] = Z[64*k1 + 32 + lane].  Partner of (k1, k2):
        // k1' = 31 - k1, k2' = 64 - k2, which is the OTHER register half of lane 32 - lane.  Lane 0: k2 = 0 pairs
        // with its own lower half at k1' = (32 - k1) % 32, k2 = 32 with its own upper half at k1' = 31 - k1.
        const int partner = (32 - lane) & 31;
        if (lane < lead) pbuf[lane] = (T)0;
        if (lane < 4) pbuf[lead + C + 1 + lane] = (T)0;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {                     // k2 = lane
            const Cpx<T> give = v[32 + bitrev5(31 - k1)];
            const Cpx<T> self = v[bitrev5((32 - k1) & 31)];
            const T bx = __shfl_sync(0xffffffffu, give.x, partner);
            const T by = __shfl_sync(0xffffffffu, give.y, partner);
            const Cpx<T> b = lane == 0 ? self : Cpx<T>{bx, by};
            const int k = 64 * k1 + lane;
            T p_lo, p_hi;
            split_pair_power<T>(v[bitrev5(k1)], b, tab->w2[k], p_lo, p_hi);
            pbuf[lead + k] = p_lo;
            pbuf[lead + C - k] = p_hi;
        }
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {                     // k2 = lane + 32
            const Cpx<T> give = v[bitrev5(31 - k1)];
            const Cpx<T> self = v[32 + bitrev5(31 - k1)];
            const T bx = __shfl_sync(0xffffffffu, give.x, partner);
            const T by = __shfl_sync(0xffffffffu, give.y, partner);
            const Cpx<T> b = lane == 0 ? self : Cpx<T>{bx, by};
            const int k = 64 * k1 + 32 + lane;
            T p_lo, p_hi;
            split_pair_power<T>(v[32 + bitrev5(k1)], b, tab->w2[k], p_lo, p_hi);
            pbuf[lead + k] = p_lo;
            pbuf[lead + C - k] = p_hi;
        }
        if (lane == 0) {
            const Cpx<T> a = v[bitrev5(16)];
            const T er = a.x + a.x, orr = a.y + a.y;
            pbuf[lead + C / 2] = er * er + orr * orr;
        }
    }
    __syncwarp();
}

}  // namespace gat
