// conv2 / conv3 of the CNN (training/cnn_trainer.py:52-76: Conv2d 3x3 pad 1 -> BatchNorm -> LeakyReLU ->
// MaxPool2) as an implicit GEMM on the Blackwell tensor cores (tcgen05.mma kind::f16, accumulators in TMEM).
//
//   D[pixel (M = 128)][c_out (N)] += A[pixel][k] * B[c_out][k],   k = (tap, c_in)
//
// * Activations live in HBM as "chunk planes": [clip][c_in/8][padded pixel][8 x 16 bit], two arrays (hf, lb).
//   With that layout and the no-swizzle K-major operand format, the rows an MMA reads for tap (ky,kx) are the SAME shared-memory planes addressed
//   (ky*Wp + kx) * 16 bytes further: im2col costs nothing and every load is a contiguous 1-D bulk (TMA) copy.
// * Split product: x = hf + lo with hf = FP16(x), for activations (lo kept as BF16: lb) and for weights (pre-scaled by
//   a power of two per layer so that lo sits in FP16's normal range: wf, wl in FP16, plus wb = BF16(wf)).
//   x*w ~= hf*wf (FP16 MMA) + lb*wb (BF16 MMA) + hf*wl (FP16 MMA): three kind::f16 MMAs of K = 16 per 16 channels into
//   one FP32 TMEM tile (conv2: TWO MMAs - hf * {wf, wl} stacked along N - into two column halves the epilogue adds,
//   see FUSE below) where 3xTF32 needs six, 4 bytes per stored activation instead of 8, and the error of three TF32
//   passes (tests/gpu_probe/tc_probe_hybrid.cu: 4.9e-6 on |x| ~ 7; one TF32 pass: 5e-3).  The reference runs the CNN in
//   fp32.  (A and B of one MMA must share a format - mixing F16 and BF16 is an illegal instruction - hence wb.)
// * One CTA per SM, warp-specialised: warp 0 = weight producer, warp 1 = MMA issuer (one thread), then one (conv2) or
//   two (conv3) groups of four epilogue warps, last warp = activation producer.  The 32 input channels of a K block are staged as two HALVES of
//   16 channels with their own full/empty barriers and the MMAs run half-major (half 0: 9 taps, half 1: 9 taps),
//   so the next group's half 0 streams in from HBM while this group's half 1 is being multiplied: the
//   activation buffer is single (it fills shared memory) yet its load latency is hidden.  A work item is a GROUP of three consecutive 128-pixel tiles covering R whole image rows, so
//   every 2x2 pooling window is inside the group: the epilogue moves 32 channels at a time TMEM -> registers ->
//   shared staging, pools, adds the BN-folded bias, applies LeakyReLU and writes the next layer's planes
//   (already split hi/lo) or the dense NHWC tensor the classifier head reads.
// * Wide images (long clips) are cut into column blocks of `cw` output columns: the group's rows are then staged
//   as R+2 separate row segments of `seg` = cw+2 pixels and a tap is (ky*seg + kx) pixels further.  When one
//   block spans the full width (seg = W+2) the rows are contiguous in HBM and each plane is ONE bulk copy.
#pragma once
#ifndef GAT_CPU_EMU
#include "common.cuh"
#include "tc05.cuh"

namespace gat {

constexpr int kTcTiles = 3;                       // 128-pixel tiles per group
constexpr int kTcGroupPix = 128 * kTcTiles;
// threads of a CTA: weight producer + MMA issuer + EPI groups of four epilogue warps + activation producer
__host__ __device__ constexpr int conv_tc_threads(int epi) { return 32 * (3 + 4 * epi); }
// Two groups of four epilogue warps work on alternate 32-channel blocks with their own staging tiles: conv3 has a single
// accumulator set (TMEM), so its epilogue is exposed, and conv2's epilogue was 93 % busy once the MMAs issued at speed.
__host__ __device__ constexpr int conv_tc_epi_groups(int cout) { return cout >= 64 ? 2 : 1; }
constexpr int kTcStageStride = 33;                // floats per pooled pixel of the mode-2 maps (32 channels + 1 pad)
// Epilogue staging tile of one 32-channel block, CHANNEL-major: [32 channels][kTcStagePlane floats].  A channel's plane holds
// the group's pixels row by row with an EVEN row stride (seg rounded up) and one float of lead, so the two horizontal
// neighbours of every 2x2 pooling window are one 8-byte aligned float2: a lane per pooled pixel reads them with LDS.64 at
// unit stride - no bank conflicts.  (Round 1 staged pixel-major with a 33-float stride; the pooling reads, two pixels apart,
// were 2-way conflicted: ncu counted 90 M conflict wavefronts of 115 M in conv2, on the shared-memory port the tensor core
// fetches its operands through.)
constexpr int kTcStagePlane = kTcGroupPix + 20;   // 384 pixels + one float of padding per row (R <= kTcMaxRows) + lead, even
constexpr int kTcMaxRows = kTcStagePlane - kTcGroupPix - 2;      // image rows per group the staged planes have room for
constexpr int kTcStageFloats = 32 * kTcStagePlane;
constexpr int kTcPooledPix = kTcGroupPix / 4;     // pooled pixels of one group (epilogue mode 2 keeps them in shared memory)

struct ConvTcParams {
    const unsigned short* in_hf;                  // [clip][CIN/8][Hp*Wp][8 fp16]   FP16(x)
    const unsigned short* in_lb;                  // [clip][CIN/8][Hp*Wp][8 bf16]   BF16(x - hf)
    const unsigned short* w;                      // [CIN/32][half][9 taps] stages of {wf | wb | wl}, each [2 chunks][COUT][8 x 16 bit]
    float w_unscale;                              // 2^-S: the weights were multiplied by 2^S before the split
    const float* bias;                            // [COUT] (BatchNorm folded)
    int n_clips, H, W;                            // conv input size without the border; Hp = H+2, Wp = W+2
    int R;                                        // image rows per group (even, R*seg <= 384)
    int seg;                                      // staged pixels per row = cw + 2 (= W + 2 when one block spans the width)
    int cw;                                       // output columns per column block (even unless there is one block)
    int col_blocks;
    int groups_per_clip;                          // row blocks * col_blocks
    int out_planes;                               // 1: next layer's planes (hf / lb); 0: dense fp32 NHWC into out_dense;
                                                  // 2: AdaptiveAvgPool2d((4,4)) of the pooled map -> FC1 operand planes (one group per clip)
    float* out_dense; unsigned short* out_hf; unsigned short* out_lb;
    float* feat_hi; float* feat_lo;               // mode 2: [COUT*4 chunks][feat_rows][4] TF32 hi / fp32 remainder (fc_tc.cuh)
    long long feat_rows; long long clip0;         // mode 2: padded row count of those planes, index of this launch's first clip
    float slope;
    long long* debug;                             // optional [grid][8] cycle counters (profiling builds), else nullptr
};

__host__ __device__ inline int conv_tc_plane_pixels(int seg) { return kTcGroupPix + 2 * seg + 2; }

template <int COUT>
__host__ __device__ inline size_t conv_tc_smem_bytes(int seg, int nstage) {
    constexpr int EPI = conv_tc_epi_groups(COUT);
    return (size_t)8 * conv_tc_plane_pixels(seg) * 16              // A: 4 + 4 chunk planes (hf, lb) of a 32-channel K block
         + (size_t)nstage * 6 * COUT * 16                          // weight ring (one stage = one tap of one K half)
         + (size_t)EPI * kTcStageFloats * 4                        // epilogue staging, one tile per epilogue group
         + (COUT == 128 ? (size_t)EPI * kTcPooledPix * kTcStageStride * 4 : 0)   // pooled maps of the last conv layer (mode 2)
         + 256;                                                    // barriers, tmem slot, alignment
}

// TILES: 128-pixel tiles per group that are actually multiplied (3, or 2 when a clip's image fits 256 pixels: half-second
// clips in conv3 fill 208 of a group's 384 pixels, so the third tile was a third of the MMAs for nothing - and with two
// tiles TMEM holds TWO accumulator sets again).  The shared-memory layout is that of three tiles either way.
template <int CIN, int COUT, int NSTAGE, int TILES = kTcTiles>
__global__ void __launch_bounds__(conv_tc_threads(conv_tc_epi_groups(COUT)), 1) conv_tc_kernel(ConvTcParams p) {
    using namespace tc;
    constexpr int EPI = conv_tc_epi_groups(COUT);          // epilogue groups (4 warps each), channel blocks interleaved between them
    constexpr int kAWarp = 2 + 4 * EPI;                    // the activation producer is the last warp
    constexpr int NKB = CIN / 32;
    constexpr uint32_t W_STAGE = 6 * COUT * 16;            // wf | wb | wl, each 2 chunks x COUT x 16 B (FUSE: {wf, wl} | wb)
    // conv2 (N = 64) is bound by the tensor core's operand fetch, not its math: an M128 N64 K16 MMA reads 4 KB of pixels and
    // 2 KB of weights for 32 clocks of work.  hf*wf and hf*wl read the SAME pixel tile, so with the two weight parts stacked
    // along N they are ONE N = 128 MMA into 128 accumulator columns per tile ([0, 64): hf*wf + lb*wb, [64, 128): hf*wl;
    // the epilogue adds the halves): 14 KB of operand fetch per (tap, 16 channels) instead of 18 KB, two MMAs instead of three.
    constexpr bool FUSE = COUT == 64;
    constexpr int DCOLS = FUSE ? 2 * COUT : COUT;          // accumulator columns per 128-pixel tile
    constexpr int ACC = (2 * TILES * DCOLS <= 512) ? 2 : 1;        // accumulator sets in TMEM
    constexpr uint32_t TMEM_COLS = ACC * TILES * DCOLS <= 256 ? 256 : 512;
    extern __shared__ __align__(128) unsigned char smem[];
    const int Wp = p.W + 2, Hp = p.H + 2;
    const int seg = p.seg;
    const int Pg = conv_tc_plane_pixels(seg);
    const bool contiguous = p.col_blocks == 1 && seg == Wp;
    const uint32_t plane = (uint32_t)Pg * 16;
    unsigned char* a_buf = smem;                                   // planes 0-3: hf chunks, 4-7: lb chunks
    unsigned char* w_buf = a_buf + (size_t)8 * plane;
    float* staging = reinterpret_cast<float*>(w_buf + (size_t)NSTAGE * W_STAGE);
    float* pooled = staging + (size_t)EPI * kTcStageFloats;   // only present (and used) when COUT == 128
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(staging) + (size_t)EPI * kTcStageFloats * 4
                                                 + (COUT == 128 ? (size_t)EPI * kTcPooledPix * kTcStageStride * 4 : 0));
    // acc_empty is PER TILE (set * kTcTiles + tile): with a single accumulator set the next group's MMAs on a tile start as
    // soon as the epilogue has that tile in registers, not when the whole group has been drained
    uint64_t* a_full = bars + 0; uint64_t* a_empty = bars + 2; uint64_t* acc_full = bars + 4; uint64_t* acc_empty = bars + 6;
    uint64_t* w_full = bars + 6 + 2 * kTcTiles; uint64_t* w_empty = w_full + NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + NSTAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < 2; ++s) mbar_init(acc_full + s, 1);
        for (int s = 0; s < 2 * kTcTiles; ++s) mbar_init(acc_empty + s, 4 * EPI);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_slot;

    const int n_work = p.n_clips * p.groups_per_clip;
    const long long plane_pix = (long long)Hp * Wp;

    if (warp == 0) {
        // ===================================================== weight producer (one thread)
        if (lane == 0) {
            uint32_t use = 0;
            for (int work = blockIdx.x; work < n_work; work += gridDim.x)
                for (int kh = 0; kh < 2 * NKB; ++kh)
                    for (int tap = 0; tap < 9; ++tap, ++use) {
                        const uint32_t st = use % NSTAGE;
                        mbar_wait(w_empty + st, ((use / NSTAGE) & 1) ^ 1);
                        mbar_expect_tx(w_full + st, W_STAGE);
                        bulk_g2s(w_buf + (size_t)st * W_STAGE, p.w + (size_t)(kh * 9 + tap) * (W_STAGE / 2), W_STAGE, w_full + st);
                    }
        }
    } else if (warp == kAWarp) {
        // ===================================================== activation producer (32 lanes issue the copies)
        uint32_t it = 0;
        const uint32_t row_bytes = (uint32_t)seg * 16;
        const int rows = p.R + 2;
        for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
            const int clip = work / p.groups_per_clip, gi = work - clip * p.groups_per_clip;
            const int rb = gi / p.col_blocks, cb = gi - rb * p.col_blocks;
            // first staged pixel: row y0-1, column xs-2 (one pixel of slack so tap offsets are never negative)
            const long long q_start = (long long)(rb * p.R) * Wp + (long long)cb * p.cw - 1;
            for (int kb = 0; kb < NKB; ++kb, ++it) {
                for (int half = 0; half < 2; ++half) {
                    if (lane == 0) {
                        mbar_wait(a_empty + half, (it & 1) ^ 1);
                        mbar_expect_tx(a_full + half, contiguous ? 4 * plane : 4 * (uint32_t)rows * row_bytes);
                    }
                    __syncwarp();
                    // the 4 planes of this K half: chunks 2*half, 2*half+1 of hf and lb
                    auto plane_src = [&](int q, int& slot) -> const unsigned char* {
                        const int arr = q >> 1, c16 = half * 2 + (q & 1);
                        slot = arr * 4 + c16;
                        const unsigned short* src = arr == 0 ? p.in_hf : p.in_lb;
                        return reinterpret_cast<const unsigned char*>(src) + ((long long)clip * (CIN / 8) + kb * 4 + c16) * plane_pix * 16;
                    };
                    if (contiguous) {
                        if (lane < 4) {
                            int slot;
                            const unsigned char* src = plane_src(lane, slot);
                            bulk_g2s(a_buf + (size_t)slot * plane, src + q_start * 16, plane, a_full + half);
                        }
                    } else {                               // R+2 row segments per plane, each seg pixels from column xs-1;
                        for (int i = lane; i < 4 * rows; i += 32) {        // slot 0 of the plane stays the unused slack pixel
                            const int q = i / rows, a = i - q * rows;
                            int slot;
                            const unsigned char* src = plane_src(q, slot);
                            bulk_g2s(a_buf + (size_t)slot * plane + (size_t)(1 + a * seg) * 16,
                                     src + (q_start + 1 + (long long)a * Wp) * 16, row_bytes, a_full + half);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (the warp runs the loop, one elected lane issues)
        {
            const bool leader = elect_one();
            const uint32_t idesc_h = idesc_f16(128, DCOLS), idesc_b = idesc_bf16(128, COUT);
            // Descriptors differ only in their start address: keep the low words as integers and add offsets.
            // low word = (addr >> 4) | (LBO >> 4) << 16 ; high word = (SBO >> 4) | version 1 at bit 46.
            constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
            const uint32_t a_lbo = (plane >> 4) << 16, b_lbo = ((uint32_t)(COUT * 16) >> 4) << 16;
            const uint32_t b_lbo2 = ((uint32_t)(2 * COUT * 16) >> 4) << 16;      // FUSE: a K chunk of {wf, wl} is 2 * COUT rows
            const uint32_t a_hf = (smem_u32(a_buf) >> 4) | a_lbo;
            const uint32_t a_lb = a_hf + ((4 * plane) >> 4);
            const uint32_t a_step = (2 * plane) >> 4;                     // two 16-byte K chunks (16 channels) per MMA
            auto desc = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
            uint32_t it = 0, use = 0, wi = 0;
            long long t_acc = 0, t_a = 0, t_w = 0, t0 = 0, t_begin = clock64();
            for (int work = blockIdx.x; work < n_work; work += gridDim.x, ++wi) {
                const uint32_t as = wi % ACC;
                const uint32_t d_base = tmem + as * (uint32_t)(TILES * DCOLS);
                uint32_t accumulate = 0;
                for (int kb = 0; kb < NKB; ++kb, ++it) {
                    for (int half = 0; half < 2; ++half) {
                        t0 = clock64();
                        mbar_wait(a_full + half, it & 1);
                        t_a += clock64() - t0;
                        for (int tap = 0; tap < 9; ++tap, ++use) {
                            const uint32_t st = use % NSTAGE;
                            t0 = clock64();
                            mbar_wait(w_full + st, (use / NSTAGE) & 1);
                            t_w += clock64() - t0;
                            fence_after_thread_sync();
                            const uint32_t row_off = (uint32_t)((tap / 3) * seg + (tap % 3));         // in 16-byte units
                            const uint32_t w_base = (smem_u32(w_buf) + st * W_STAGE) >> 4;
                            const uint32_t w_hf = w_base | (FUSE ? b_lbo2 : b_lbo);                  // FUSE: {wf, wl}, N = 128
                            const uint32_t w_hb = (w_base | b_lbo) + ((FUSE ? 4 : 2) * COUT * 16 >> 4);
                            const uint32_t w_lf = w_hf + ((4 * COUT * 16) >> 4);                     // unused when FUSE
                            const uint32_t a_off = row_off + (uint32_t)half * a_step;
#pragma unroll
                            for (int g = 0; g < TILES; ++g) {
                                const uint32_t d = d_base + (uint32_t)(g * DCOLS);
                                if (accumulate == 0) {                  // first MMA of the group on this tile: its accumulators must be drained
                                    t0 = clock64();
                                    mbar_wait(acc_empty + as * kTcTiles + g, ((wi / ACC) & 1) ^ 1);
                                    t_acc += clock64() - t0;
                                    fence_after_thread_sync();
                                }
                                if (leader) {
                                    mma_16bit(d, desc(a_hf + a_off + g * 128), desc(w_hf), idesc_h, accumulate);   // hf * wf (FUSE: and hf * wl)  (FP16)
                                    mma_16bit(d, desc(a_lb + a_off + g * 128), desc(w_hb), idesc_b, 1u);           // lb * wb   (BF16)
                                    if constexpr (!FUSE)
                                        mma_16bit(d, desc(a_hf + a_off + g * 128), desc(w_lf), idesc_h, 1u);       // hf * wl   (FP16)
                                }
                            }
                            accumulate = 1;
                            if (leader) mma_commit(w_empty + st);       // weights of this stage are free once those MMAs retire
                            __syncwarp();
                        }
                        if (leader) mma_commit(a_empty + half);         // ... and so is this half of the activation buffer
                    }
                }
                if (leader) mma_commit(acc_full + as);                  // accumulators of the group are complete
            }
            if (p.debug && leader) {
                long long* d = p.debug + (long long)blockIdx.x * 8;
                d[0] = clock64() - t_begin; d[1] = t_acc; d[2] = t_a; d[3] = t_w;
            }
        }
    } else {
        // ===================================================== epilogue (4 warps = 128 threads)
        const int quarter = warp & 3;                       // the TMEM lanes this warp may read: 32*quarter ..
        const int eg = (warp - 2) >> 2;                     // epilogue group: handles channel blocks eg, eg + EPI, ...
        const int et = ((warp - 2) & 3) * 32 + lane;        // thread index inside the group
        float* const stage_g = staging + (size_t)eg * kTcStageFloats;
        const int RS = (seg + 1) & ~1;                      // even row stride of the staged planes
        float* const pooled_g = pooled + (size_t)eg * kTcPooledPix * kTcStageStride;
        constexpr int kLastCb = COUT / 32 - 1;
        const int Hpool = p.H / 2, Wpool = p.W / 2;
        const int Wp_out = Wpool + 2;
        const long long Pout = (long long)(Hpool + 2) * Wp_out;
        uint32_t wi = 0;
        long long e_wait = 0, e_total = clock64();
        for (int work = blockIdx.x; work < n_work; work += gridDim.x, ++wi) {
            const int clip = work / p.groups_per_clip, gi = work - clip * p.groups_per_clip;
            const int rb = gi / p.col_blocks, colb = gi - rb * p.col_blocks;
            int prow = Hpool - rb * (p.R / 2);              // pooled rows produced by this group
            prow = prow < p.R / 2 ? prow : p.R / 2;
            int pcol = Wpool - colb * (p.cw / 2);             // pooled columns produced by this group
            pcol = pcol < p.cw / 2 ? pcol : p.cw / 2;
            const uint32_t as = wi % ACC;
            const uint32_t t_acc = tmem + as * (uint32_t)(TILES * DCOLS);
            const long long e0 = clock64();
            mbar_wait(acc_full + as, (wi / ACC) & 1);
            e_wait += clock64() - e0;
            fence_after_thread_sync();
            for (int cb = eg; cb < COUT / 32; cb += EPI) {
#pragma unroll
                for (int g = 0; g < TILES; ++g) {
                    float v[32];
                    if constexpr (FUSE) {                                 // the two halves of the split product, one wait
                        uint32_t r0[32], r1[32];
                        const uint32_t ta = t_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * DCOLS + cb * 32);
                        tmem_ld32_issue(ta, r0);
                        tmem_ld32_issue(ta + COUT, r1);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                    } else {
                        tmem_ld32(t_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * COUT + cb * 32), v);
                    }
                    const int pixel = g * 128 + quarter * 32 + lane;      // this thread's accumulator row
                    const int prow_ = pixel / seg;
                    float* dst = stage_g + prow_ * RS + (pixel - prow_ * seg) + 1;
                    if (cb + EPI > kLastCb) {               // this group's last block of tile g is in registers: (with the others) the
                        fence_before_thread_sync();         // next group's MMAs on this tile may start
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + as * kTcTiles + g);
                    }
                    if (prow_ < p.R) {                                    // tile slots past the group's R rows hold nothing
#pragma unroll
                        for (int j = 0; j < 32; ++j) dst[j * kTcStagePlane] = v[j];
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
                const int n_items = prow * pcol * 4;            // one item = one pooled pixel x 8 channels
                for (int item = et; item < n_items; item += 128) {
                    const int pxl = item % pcol;
                    const int rest = item / pcol;
                    const int r = rest % prow, ch8 = rest / prow;
                    const int px = colb * (p.cw / 2) + pxl;
                    // window = columns 1 + 2 pxl, 2 + 2 pxl of rows 2r, 2r + 1: two aligned float2 per channel
                    const float* s00 = stage_g + (size_t)(ch8 * 8) * kTcStagePlane + 2 * r * RS + 2 + 2 * pxl;
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 top = *reinterpret_cast<const float2*>(s00 + e * kTcStagePlane);
                        const float2 bot = *reinterpret_cast<const float2*>(s00 + e * kTcStagePlane + RS);
                        const float m = fmaxf(fmaxf(top.x, top.y), fmaxf(bot.x, bot.y));
                        const float z = m * p.w_unscale + __ldg(p.bias + cb * 32 + ch8 * 8 + e);    // exact power-of-two rescale
                        o[e] = z > 0.0f ? z : z * p.slope;
                    }
                    const int Y = rb * (p.R / 2) + r;
                    if (p.out_planes == 1) {
                        const long long pix = (long long)(Y + 1) * Wp_out + px + 1;
                        const long long c8 = (long long)clip * (COUT / 8) + cb * 4 + ch8;
                        uint4 hf, lb;
                        split16x8(o, hf, lb);
                        *reinterpret_cast<uint4*>(p.out_hf + (c8 * Pout + pix) * 8) = hf;
                        *reinterpret_cast<uint4*>(p.out_lb + (c8 * Pout + pix) * 8) = lb;
                    } else if (p.out_planes == 2) {
                        float* dst = pooled_g + (size_t)(Y * Wpool + px) * kTcStageStride + ch8 * 8;
#pragma unroll
                        for (int e = 0; e < 8; ++e) dst[e] = o[e];
                    } else {
                        float* dst = p.out_dense + (((long long)clip * Hpool + Y) * Wpool + px) * COUT + cb * 32 + ch8 * 8;
                        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
                if (p.out_planes == 2) {
                    // AdaptiveAvgPool2d((4,4)) (cnn_trainer.py:105) of the 32 channels just pooled: thread = (window row i, channel);
                    // same window bounds and summation order as avgpool_planes_kernel, output straight into FC1's operand planes
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
                    const int i = et >> 5, ch = et & 31;
                    const int y0 = (i * Hpool) / 4, y1 = ((i + 1) * Hpool + 3) / 4;
                    float a4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int x0 = (j * Wpool) / 4, x1 = ((j + 1) * Wpool + 3) / 4;
                        float sum = 0.0f;
                        for (int y = y0; y < y1; ++y)
                            for (int x = x0; x < x1; ++x) sum += pooled_g[(size_t)(y * Wpool + x) * kTcStageStride + ch];
                        a4[j] = sum / (float)((y1 - y0) * (x1 - x0));
                    }
                    const long long off = (((long long)((cb * 32 + ch) * 4 + i)) * p.feat_rows + p.clip0 + clip) * 4;
                    const float4 hi = make_float4(tf32_hi(a4[0]), tf32_hi(a4[1]), tf32_hi(a4[2]), tf32_hi(a4[3]));
                    *reinterpret_cast<float4*>(p.feat_hi + off) = hi;
                    *reinterpret_cast<float4*>(p.feat_lo + off) = make_float4(a4[0] - hi.x, a4[1] - hi.y, a4[2] - hi.z, a4[3] - hi.w);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
            }
        }
        if (p.debug && et == 0 && eg == 0) {
            long long* d = p.debug + (long long)blockIdx.x * 8;
            d[4] = clock64() - e_total; d[5] = e_wait;
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// conv1 (C_in = 1, CUDA cores) writing conv2's operand directly: chunk planes, hi/lo split.
struct Conv1PlanesParams {
    const float* in; int N, H, W;
    const float* w; const float* bias;    // [9][32], [32]
    unsigned short* out_hf;               // [clip][4][(H/2+2)*(W/2+2)][8 fp16]  FP16(x)
    unsigned short* out_lb;               // [clip][4][...][8 bf16]              BF16(x - hf)
    float slope;
};

__global__ void __launch_bounds__(256, 4) conv1_pool_planes_kernel(Conv1PlanesParams p) {
    __shared__ __align__(16) float ws[9 * 32];
    __shared__ float bs[32];
    for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) ws[i] = p.w[i];
    for (int i = threadIdx.x; i < 32; i += blockDim.x) bs[i] = p.bias[i];
    __syncthreads();
    const int Hq = p.H / 2, Wq = p.W / 2;
    const int tiles = ceil_div(Hq * Wq, (int)blockDim.x);
    const int clip = blockIdx.x / tiles;
    const int q = (blockIdx.x - clip * tiles) * blockDim.x + threadIdx.x;
    if (q >= Hq * Wq) return;
    const int py = q / Wq, px = q - py * Wq;
    const float* img = p.in + (long long)clip * p.H * p.W;
    float patch[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int y = 2 * py - 1 + a, x = 2 * px - 1 + b;
            patch[a][b] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? img[y * p.W + x] : 0.0f;
        }
    const long long Pout = (long long)(Hq + 2) * (Wq + 2);
    const long long pix = (long long)(py + 1) * (Wq + 2) + px + 1;
#pragma unroll
    for (int ch8 = 0; ch8 < 4; ++ch8) {
        // eight channels x four pre-pool positions; the weights of a tap come as two 128-bit broadcast loads.  The
        // accumulators are channel PAIRS updated with the packed FFMA2 of sm_100 (fma.rn.f32x2: two IEEE fp32 FMAs per
        // issued instruction - the kernel is issue-bound, not FMA-pipe-bound); every lane of a pair sees exactly the
        // fmaf(pv, w, acc) it saw before.
        float2 acc[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) acc[e][q4] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 w0 = *reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * 32 + ch8 * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * 32 + ch8 * 8 + 4);
                const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const float pv = patch[a + ky][b + kx];
                        const float2 pv2 = make_float2(pv, pv);
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[e][a * 2 + b] = __ffma2_rn(pv2, wv[e], acc[e][a * 2 + b]);
                    }
            }
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a0 = (e & 1) ? acc[e >> 1][0].y : acc[e >> 1][0].x, a1 = (e & 1) ? acc[e >> 1][1].y : acc[e >> 1][1].x;
            const float a2 = (e & 1) ? acc[e >> 1][2].y : acc[e >> 1][2].x, a3 = (e & 1) ? acc[e >> 1][3].y : acc[e >> 1][3].x;
            const float best = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
            const float z = best + bs[ch8 * 8 + e];
            o[e] = z > 0.0f ? z : z * p.slope;
        }
        const long long c8 = (long long)clip * 4 + ch8;
        uint4 hf, lb;
        tc::split16x8(o, hf, lb);
        *reinterpret_cast<uint4*>(p.out_hf + (c8 * Pout + pix) * 8) = hf;
        *reinterpret_cast<uint4*>(p.out_lb + (c8 * Pout + pix) * 8) = lb;
    }
}

}  // namespace gat
#endif  // GAT_CPU_EMU
