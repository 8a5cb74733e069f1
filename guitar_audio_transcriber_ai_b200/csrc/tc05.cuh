// Thin inline-PTX layer for the Blackwell tensor path: mbarrier, 1-D bulk TMA copies, TMEM allocation,
// tcgen05.mma (kind::tf32, operands from shared memory), tcgen05.commit / tcgen05.ld.
// sm_100a only; nothing here has a host emulation (the tcgen05 kernels are exercised on the GPU box only).
#pragma once
#ifndef GAT_CPU_EMU
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace gat {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}

// ---- 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One lane of a converged warp (elect.sync).  Code that issues tcgen05 instructions is written warp-uniformly - all
// lanes run the control flow and compute the operands - and only the issue itself is guarded by this predicate: with
// the whole block under `if (lane == 0)` ptxas cannot prove the operands uniform and wraps EVERY MMA in an
// ELECT / branch loop (seen in the SASS: ~10 extra instructions per UTCHMMA, which bound the N = 64 layer).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- tensor memory
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {         // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): 8-row x 16-byte core matrices;
// inside a core matrix rows are 16 B apart; `sbo` = bytes between 8-row groups, `lbo` = bytes between the
// two 16-byte K chunks one MMA consumes.  Any 16-byte aligned start address is legal, which is what lets a
// 3x3 tap be a plain address offset into a pixel-major plane.
__device__ __forceinline__ uint64_t smem_desc_kmajor_noswizzle(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version for sm_100
    return d;                                     // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// Instruction descriptor for kind::tf32: FP32 accumulate, TF32 x TF32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same with BF16 operands (kind::f16, a/b format 1 = BF16, FP32 accumulate): K = 16 per MMA, i.e. two 16-byte chunks
// of eight channels.  Used for the two correction passes of the split product (see conv_tc.cuh).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// FP16 x FP16 (format 0).  The two operands of one MMA must share a format: mixing F16 and BF16 is an illegal
// instruction on sm_100a (tests/gpu_probe/tc_probe_hybrid.cu).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_16bit(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on `bar` once every tcgen05 operation issued so far by this thread has completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait: issue several, then tmem_wait_ld() once.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TF32 split of an fp32 value: hi keeps the top 19 bits (sign, exponent, 10 mantissa bits), lo = x - hi is
// exact in fp32.  hi*hi + lo*hi + hi*lo recovers an fp32-faithful product on the tf32 tensor path.
__host__ __device__ inline float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
#else
    uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r;
#endif
}

// BF16 (round to nearest even) of an fp32 value, as raw bits; host and device give identical results.
__host__ __device__ inline unsigned short bf16_bits(float x) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4);
#endif
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
// Split of an fp32 value for the conv tensor path: hf = FP16(x) (round to nearest; clamped so it never overflows) and
// the remainder lo = x - hf, stored as BF16 (activations) or FP16 (weights, which are pre-scaled by a power of two so
// that their remainders sit in FP16's normal range).  x*w ~= hf*wf + lb*wb + hf*wl: FP16, BF16 and FP16 MMAs - the two
// operands of one MMA must share a format, which is why the weights carry a BF16 copy of their hi part (wb).
constexpr float kF16Max = 65504.0f;
__host__ inline void split16_weight_host(float w, unsigned short& hf, unsigned short& hb, unsigned short& lf) {
    const float c = w > kF16Max ? kF16Max : (w < -kF16Max ? -kF16Max : w);
    const __half h = __float2half_rn(c);
    const float hv = __half2float(h);
    hf = __half_as_ushort(h); hb = bf16_bits(hv); lf = __half_as_ushort(__float2half_rn(w - hv));
}
// Eight channels -> the two 16-byte chunks of an activation: hf (FP16) and lb (BF16 of the remainder).
__device__ __forceinline__ void split16x8(const float (&o)[8], uint4& hf, uint4& lb) {
    uint32_t f[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {       // packed conversions: two values per instruction
        const float x0 = fminf(fmaxf(o[2 * e], -kF16Max), kF16Max), x1 = fminf(fmaxf(o[2 * e + 1], -kF16Max), kF16Max);
        const __half2 fp = __floats2half2_rn(x0, x1);
        const float2 fv = __half22float2(fp);
        const __nv_bfloat162 lp = __floats2bfloat162_rn(o[2 * e] - fv.x, o[2 * e + 1] - fv.y);
        f[e] = *reinterpret_cast<const uint32_t*>(&fp);
        l[e] = *reinterpret_cast<const uint32_t*>(&lp);
    }
    hf = make_uint4(f[0], f[1], f[2], f[3]);
    lb = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace tc
}  // namespace gat
#endif  // GAT_CPU_EMU
