// Warp-level real FFT of 2048 samples, data in registers (shared by the STFT, MFCC and onset kernels).
//
// A 2048-point real frame is folded into 1024 complex points z[j] = x[2j] + i x[2j+1]; one warp
// transforms it as 32 x 32: every lane runs a 32-point FFT in registers (over n2, stride 32), the result
// is twiddled by W_1024^(n1*k2) and transposed through shared memory, every lane runs a second 32-point
// FFT (over n1), and the Hermitian split X[k] = E[k] + W_2048^k O[k] recovers bins 0..1024 (two at a time).  No block
// barrier is involved: only __syncwarp.  T is float (feature chains) or double (onset chain, which the
// reference runs in float64: slicing.py:37,90 promote the gated signal).
#pragma once
#include "common.cuh"

namespace gat {

template <typename T> struct Cpx { T x, y; };

template <typename T> __device__ __forceinline__ Cpx<T> cadd(Cpx<T> a, Cpx<T> b) { return Cpx<T>{a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ Cpx<T> csub(Cpx<T> a, Cpx<T> b) { return Cpx<T>{a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ Cpx<T> cmul(Cpx<T> a, Cpx<T> b) {
    return Cpx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}

// cos(2*pi*j/32), j = 0..8
template <typename T> __host__ __device__ constexpr T cos32_tab(int j) {
    return j == 0 ? T(1.0)
         : j == 1 ? T(0.98078528040323044912618223613424)
         : j == 2 ? T(0.92387953251128675612818318939679)
         : j == 3 ? T(0.83146961230254523707878837761791)
         : j == 4 ? T(0.70710678118654752440084436210485)
         : j == 5 ? T(0.55557023301960222474283081394853)
         : j == 6 ? T(0.38268343236508977172845998403040)
         : j == 7 ? T(0.19509032201612826784828486847702)
         : T(0.0);
}
// cos(2*pi*j/32) for any j in [0, 32)
template <typename T> __host__ __device__ constexpr T cos32(int j) {
    return j <= 8 ? cos32_tab<T>(j) : j <= 16 ? -cos32_tab<T>(16 - j) : j <= 24 ? -cos32_tab<T>(j - 16) : cos32_tab<T>(32 - j);
}
template <typename T> __host__ __device__ constexpr T sin32(int j) { return cos32<T>((j + 24) & 31); }  // sin(a) = cos(a - pi/2)

__host__ __device__ constexpr int bitrev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// One radix-2 DIF butterfly with the compile-time twiddle W_32^J = exp(-2 pi i J / 32).
template <typename T, int J>
__device__ __forceinline__ void butterfly(Cpx<T>& a, Cpx<T>& b) {
    Cpx<T> s = cadd(a, b);
    Cpx<T> d = csub(a, b);
    a = s;
    if (J == 0) {
        b = d;
    } else if (J == 8) {            // multiply by -i
        b = Cpx<T>{d.y, -d.x};
    } else if (J == 4) {            // (1 - i)/sqrt(2)
        constexpr T h = cos32_tab<T>(4);
        b = Cpx<T>{(d.x + d.y) * h, (d.y - d.x) * h};
    } else if (J == 12) {           // (-1 - i)/sqrt(2)
        constexpr T h = cos32_tab<T>(4);
        b = Cpx<T>{(d.y - d.x) * h, -(d.x + d.y) * h};
    } else {
        constexpr T c = cos32<T>(J);
        constexpr T s_ = sin32<T>(J);
        b = Cpx<T>{d.x * c + d.y * s_, d.y * c - d.x * s_};   // d * (c - i s)
    }
}

template <typename T, int M, int B, int J>
struct StageLoop {   // all butterflies of the stage with block size M, block offset B, index J
    __device__ static __forceinline__ void run(Cpx<T> (&v)[32]) {
        butterfly<T, J * (32 / M)>(v[B + J], v[B + J + M / 2]);
        if constexpr (J + 1 < M / 2) StageLoop<T, M, B, J + 1>::run(v);
        else if constexpr (B + M < 32) StageLoop<T, M, B + M, 0>::run(v);
    }
};

// In-register 32-point forward FFT, decimation in frequency: natural order in, X[k] ends up in v[bitrev5(k)].
template <typename T>
__device__ __forceinline__ void fft32_dif(Cpx<T> (&v)[32]) {
    StageLoop<T, 32, 0, 0>::run(v);
    StageLoop<T, 16, 0, 0>::run(v);
    StageLoop<T, 8, 0, 0>::run(v);
    StageLoop<T, 4, 0, 0>::run(v);
    StageLoop<T, 2, 0, 0>::run(v);
}

// Tables every kernel that uses warp_rfft2048 keeps in shared memory (filled by fill_fft_tables).
template <typename T>
struct FftTables {
    Cpx<T> tw[32 * 32];   // tw[k2*32 + n1] = W_1024^(n1*k2)
    Cpx<T> w2[1024];      // w2[k] = W_2048^k
};

template <typename T>
__device__ void fill_fft_tables(FftTables<T>* tab, const Cpx<T>* __restrict__ g_tw, const Cpx<T>* __restrict__ g_w2) {
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        tab->tw[i] = g_tw[i];
        tab->w2[i] = g_w2[i];
    }
}

constexpr int kXbufStride = 33;                       // complex elements per row of the transpose buffer
constexpr int kXbufElems = 32 * kXbufStride;          // per warp

// Forward real FFT of one 2048-sample frame by one warp, power spectrum left in shared memory.
//   v[n2]  in : z[lane + 32*n2] = (x[2j], x[2j+1]) * 0.5 * window, j = lane + 32*n2.  The caller folds the
//               Hermitian split's factor 1/2 into the window (an exact power-of-two scaling).
//   xbuf      : per-warp scratch of kXbufElems complex values; on return, reinterpreted as T[],
//               pbuf[lead + k] = |X[k]|^2 for k = 0..1024 and pbuf[0..lead) = 0.
// Bins k and 1024-k share E = Z[k] + conj Z[1024-k] and T = W_2048^k * O: |X[k]|^2 = |E+T|^2 and
// |X[1024-k]|^2 = |E-T|^2, so each lane handles 16 pairs instead of 32 single bins.
template <typename T>
__device__ __forceinline__ void warp_rfft2048_power(Cpx<T> (&v)[32], Cpx<T>* xbuf, int lead, const FftTables<T>* tab) {
    const int lane = lane_id();
    // pass 1: 32-point FFT over n2 (in registers)
    fft32_dif<T>(v);
    // twiddle by W_1024^(n1*k2), n1 = lane, and transpose: row k2, column n1
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int k2 = bitrev5(r);
        Cpx<T> w = tab->tw[k2 * 32 + lane];
        xbuf[k2 * kXbufStride + lane] = (k2 == 0) ? v[r] : cmul(v[r], w);
    }
    __syncwarp();
    // pass 2: lane = k2 reads its row over n1
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) v[n1] = xbuf[lane * kXbufStride + n1];
    __syncwarp();
    fft32_dif<T>(v);
    // v[bitrev5(k1)] = Z[32*k1 + lane].  The Hermitian partner Z[1024-k] of k = 32*k1 + lane sits in lane
    // (32 - lane) % 32 at k1' = 31 - k1 (lane 0 pairs with itself: k1' = (32 - k1) % 32): one shuffle per value.
    Cpx<T> b[16];
    const int partner = (32 - lane) & 31;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const Cpx<T> other = v[bitrev5(31 - k1)];
        const Cpx<T> self = v[bitrev5((32 - k1) & 31)];
        const T bx = __shfl_sync(0xffffffffu, other.x, partner);
        const T by = __shfl_sync(0xffffffffu, other.y, partner);
        b[k1] = lane == 0 ? self : Cpx<T>{bx, by};
    }
    // xbuf was last READ before the second FFT (with a __syncwarp after), so it can take the spectrum now
    T* pbuf = reinterpret_cast<T*>(xbuf);
    if (lane < lead) pbuf[lane] = (T)0;
    if (lane < 4) pbuf[lead + 1025 + lane] = (T)0;        // tail read (times zero weights) by the vectorised mel loop
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const int k = 32 * k1 + lane;
        const Cpx<T> a = v[bitrev5(k1)];
        const Cpx<T> w = tab->w2[k];
        const T er = a.x + b[k1].x, ei = a.y - b[k1].y;   // E = A + conj B
        const T orr = a.y + b[k1].y, oi = b[k1].x - a.x;  // O = -i (A - conj B)
        const T tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;
        const T pr = er + tr, pi = ei + ti, mr = er - tr, mi = ei - ti;
        pbuf[lead + k] = pr * pr + pi * pi;
        pbuf[lead + 1024 - k] = mr * mr + mi * mi;
    }
    if (lane == 0) {                                      // k = 512 pairs with itself: W^512 = -i
        const Cpx<T> a = v[bitrev5(16)];
        const T er = a.x + a.x, orr = a.y + a.y;          // E = (2 Re a, 0), O = (2 Im a, 0), T = W*O = (0, -2 Im a)
        pbuf[lead + 512] = er * er + orr * orr;
    }
    __syncwarp();
}

}  // namespace gat
