// libgat.so - host side of the C ABI in include/gat.h: context, tables, weight upload, kernel launches.
// Built by nvcc for sm_100a.  (tests/emu compiles this same file with g++ against a host emulation of the
// CUDA execution model to debug kernel logic in a GPU-less container; that build is never shipped.)
#ifdef GAT_CPU_EMU
#include "cpu_emu.h"
using std::max;
using std::min;
#endif

#include "common.cuh"
#include "conv_tc.cuh"
#include "fc_tc.cuh"
#include "features.cuh"
#include "fft.cuh"
#include "frontend.cuh"
#include "infer.cuh"
#include "onset.cuh"
#include "stft2.cuh"
#include "yin.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gat.h"

using namespace gat;

namespace {

thread_local std::string g_error;

int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}

#define GAT_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return fail("cudaMalloc(%zu) failed", want); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <typename T>
int upload(DevBuf& b, const T* host, size_t count) {
    if (b.ensure(count * sizeof(T))) return 1;
    GAT_CUDA(cudaMemcpy(b.p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

struct SparseFbDev {
    DevBuf start, len, off, mel, w;
    int n_mels = 0, n_slots = 0, nnz = 0;
    SparseFb view() const {
        return SparseFb{n_mels, n_slots, nnz, start.as<int>(), len.as<int>(), off.as<int>(), mel.as<int>(), w.as<float>()};
    }
};

// dense[filter m][bin f] (given strides) -> lane-slot banded form for stft_mel_kernel.
// Slot q holds filters 32q .. 32q+31.  A filter given to lane l has its band start lowered to the nearest
// value congruent to 4*(l mod 8) modulo 32 (cost: up to 31 zero-weight bins) and its length rounded up to a
// multiple of 4, so the kernel can use conflict-free 128-bit loads.  Long filters choose first and prefer
// lanes whose earlier slots are short, which keeps the per-lane totals (the warp's loop count) balanced.
int build_sparse_fb(SparseFbDev& out, const float* dense, int n_mels, int n_freqs, long long stride_m, long long stride_f) {
    const int n_slots = (n_mels + 31) / 32;
    std::vector<int> first(n_mels), length(n_mels);
    for (int m = 0; m < n_mels; ++m) {
        int lo = -1, hi = -1;
        for (int f = 0; f < n_freqs; ++f)
            if (dense[m * stride_m + f * stride_f] != 0.0f) { if (lo < 0) lo = f; hi = f; }
        first[m] = lo < 0 ? 0 : lo;
        length[m] = lo < 0 ? 0 : hi - lo + 1;
    }
    std::vector<int> start(n_slots * 32, 0), len(n_slots * 32, 0), off(n_slots * 32, 0), mel(n_slots * 32, -1);
    std::vector<int> lane_load(32, 0);
    std::vector<float> w;
    auto lower = [](int f, int lane) { return ((f - 4 * (lane % 8)) % 32 + 32) % 32; };
    for (int q = 0; q < n_slots; ++q) {
        std::vector<int> ids;
        for (int m = 32 * q; m < n_mels && m < 32 * q + 32; ++m) ids.push_back(m);
        std::sort(ids.begin(), ids.end(), [&](int a, int b) { return length[a] > length[b]; });
        std::vector<char> taken(32, 0);
        for (int m : ids) {
            int best_lane = -1, best_cost = 1 << 30;
            for (int lane = 0; lane < 32; ++lane) {
                if (taken[lane]) continue;
                const int cost = lane_load[lane] + (length[m] + lower(first[m], lane) + 3) / 4;
                if (cost < best_cost) { best_cost = cost; best_lane = lane; }
            }
            const int d = lower(first[m], best_lane);
            const int e = q * 32 + best_lane;
            const int groups = length[m] ? (length[m] + d + 3) / 4 : 0;
            taken[best_lane] = 1;
            lane_load[best_lane] += groups;
            start[e] = first[m] - d;
            len[e] = groups;
            mel[e] = m;
            while ((int)(w.size() % 32) != 4 * (best_lane % 8)) w.push_back(0.0f);
            off[e] = (int)w.size();
            for (int k = 0; k < 4 * groups; ++k) {
                const int f = start[e] + k;
                w.push_back(f >= first[m] && f < first[m] + length[m] ? dense[m * stride_m + (long long)f * stride_f] : 0.0f);
            }
        }
    }
    if (w.empty()) w.push_back(0.0f);
    out.n_mels = n_mels; out.n_slots = n_slots; out.nnz = (int)w.size();
    if (upload(out.start, start.data(), start.size()) || upload(out.len, len.data(), len.size()) ||
        upload(out.off, off.data(), off.size()) || upload(out.mel, mel.data(), mel.size()) ||
        upload(out.w, w.data(), w.size())) return 1;
    return 0;
}

}  // namespace

struct ProfRec { const char* name; cudaEvent_t e0, e1; };

// kernels launched through a function pointer are reported under a readable name
static thread_local const char* g_kernel_alias = nullptr;   // per thread: contexts on different threads must not rename each other's launches
static inline const char* kernel_name(const char* expr) {
    const char* n = g_kernel_alias ? g_kernel_alias : expr;
    g_kernel_alias = nullptr;
    return n;
}
#define KNAME(s) (g_kernel_alias = (s))

struct gat_ctx {
    bool profiling = false;
    std::vector<ProfRec> prof;
    int device = 0;
    int num_sms = 1;
    gat_config cfg{};
    int64_t launches = 0;
    // tables
    DevBuf tw32, w2_32, tw64, w2_64, tw_mel, w2_mel, win_mel, win_mel_half, win_mfcc_half, win64, dct;
    SparseFbDev fb_mel, fb_mfcc;
    // models
    DevBuf mlp_params; int mlp_dims[kMlpMaxLayers + 1] = {0}; int mlp_n_linear = 0; int mlp_n_params = 0;
    DevBuf conv_w[3], conv_b[3], fc1_w, fc1_b, fc2_w, fc2_b;
    float conv_w_unscale[3] = {1.0f, 1.0f, 1.0f};   // 2^-S of the pre-scaled tensor-core weights
    DevBuf conv_w_tc[3];   // conv2/conv3 weights in the tensor-core operand layout (wf | wb | wl stages, see conv_tc.cuh)
    DevBuf fc1_w_tc, feat_planes, hid, tc_debug_buf;
    bool tc_debug = false;
    int conv_pass_mult = 28;  // clips per conv pass = conv_pass_mult * num_sms; measured on B200: long passes win (round 1: 5.84 ms at 1,
                              // 5.21 at 14, 5.15 at 28 per 4096 clips; round 2: 3.08 at 16, 3.05 at 28) - per-launch head/tail costs outweigh
                              // keeping activations inside the L2.  (Round 2 also tried running conv1 of pass k+1 on a side stream beside
                              // conv2 / conv3 of pass k, with the tensor-core kernels capped at 128 registers so that a conv1 CTA fits on the
                              // same SM: conv1 overlapped, but conv2 slowed from 0.68 to 0.84 ms - its epilogue warps need the issue slots
                              // conv1 takes - and the step did not move: 3.05-3.08 ms at every pass size.  Not kept.)
    int conv_ch[4] = {0, 0, 0, 0}; int hidden = 0, classes = 0; bool cnn_loaded = false;
    DevBuf scaler_mean, scaler_scale; int scaler_n = 0;
    float w_mlp = 0.2f, w_cnn = 0.8f;
    // scratch
    DevBuf clip_scale, spec, spec_max, clip_count, f0, act1, act2, act3, hz_tmp, logits_cnn, logits_mlp;
    long long act_shape[3] = {0, 0, 0};   // (chunk, H, W) the zero borders of act1/act2 were prepared for
    DevBuf seg_small, seg_rms, seg_rms_med, seg_gate, seg_env, seg_envn, seg_cand, seg_peaks, seg_frames, seg_table,
           seg_keep, seg_dest, seg_base, seg_counts;
    // end-to-end staging
    DevBuf e2e_audio[2], e2e_pcm[2], e2e_mel, e2e_mfcc, e2e_probs, e2e_mlp_probs, e2e_cnn_probs, e2e_index, e2e_conf;
    cudaStream_t e2e_stream[2] = {nullptr, nullptr};
    bool e2e_streams = false;
    int host_chunk_mult[3] = {2, 8, 0};   // smallest / largest chunk of the host entry points (x num_sms) and their order (gat_set_host_chunks)
};

extern "C" const char* gat_last_error(void) { return g_error.c_str(); }
extern "C" int gat_version(void) { return 100; }
extern "C" int64_t gat_launch_count(const gat_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int32_t gat_num_classes(const gat_ctx* ctx) { return ctx ? (ctx->classes ? ctx->classes : ctx->mlp_dims[ctx->mlp_n_linear]) : 0; }
extern "C" int32_t gat_mel_frames(const gat_ctx* ctx, int64_t n) { return ctx ? (int32_t)(1 + n / ctx->cfg.mel_hop) : 0; }

// Every kernel goes through LAUNCH: it counts launches (bench.py's gpu_launches) and, when profiling is
// switched on with gat_profile_begin, brackets the launch with CUDA events on the launching stream.
#define LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                      \
        cudaEvent_t pe0_ = nullptr, pe1_ = nullptr;                                           \
        if ((ctx)->profiling) {                                                               \
            GAT_CUDA(cudaEventCreate(&pe0_)); GAT_CUDA(cudaEventCreate(&pe1_));               \
            GAT_CUDA(cudaEventRecord(pe0_, (cudaStream_t)(stream)));                          \
        }                                                                                     \
        GAT_LAUNCH(kernel, grid, block, smem, (cudaStream_t)(stream), __VA_ARGS__);           \
        ++(ctx)->launches;                                                                    \
        GAT_CUDA(cudaGetLastError());                                                         \
        const char* kn_ = kernel_name(#kernel);                                               \
        if ((ctx)->profiling) {                                                               \
            GAT_CUDA(cudaEventRecord(pe1_, (cudaStream_t)(stream)));                          \
            (ctx)->prof.push_back(ProfRec{kn_, pe0_, pe1_});                                  \
        }                                                                                     \
    } while (0)

// Diagnostics: cycle counters of the conv_tc pipeline roles (last launch of each layer), 2 x 148 x 8 int64.
extern "C" int gat_debug_tc_counters(gat_ctx* c, long long* out_host, int64_t n) {
    if (!c) return fail("gat_debug_tc_counters: null ctx");
    if (!out_host) {   // enable
        if (c->tc_debug_buf.ensure(2 * 148 * 8 * 8)) return 1;
        GAT_CUDA(cudaMemset(c->tc_debug_buf.p, 0, 2 * 148 * 8 * 8));
        c->tc_debug = true;
        return 0;
    }
    GAT_CUDA(cudaDeviceSynchronize());
    GAT_CUDA(cudaMemcpy(out_host, c->tc_debug_buf.p, (size_t)(n < 2 * 148 * 8 ? n : 2 * 148 * 8) * 8, cudaMemcpyDeviceToHost));
    return 0;
}

// FP32-FMA peak of this device, measured: the denominator SURVEY.md 8(d) asks the FFT-bound stages to be reported
// against (MEASURED_PEAKS.json has no FP32 figure).  16 independent FMA chains per thread, 1024 threads x 2 CTAs per SM.
#ifndef GAT_CPU_EMU
__global__ void __launch_bounds__(1024, 2) fma_peak_kernel(float* out, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;     // keeps the chains alive
}
#endif
extern "C" int gat_debug_fma_peak(gat_ctx* c, int32_t iters, float* tflops_host) {
    if (!c || !tflops_host || iters < 1) return fail("gat_debug_fma_peak: bad argument");
#ifdef GAT_CPU_EMU
    return fail("gat_debug_fma_peak: needs the CUDA build");
#else
    if (c->tc_debug_buf.ensure(2 * 148 * 8 * 8)) return 1;
    cudaEvent_t e0, e1;
    GAT_CUDA(cudaEventCreate(&e0)); GAT_CUDA(cudaEventCreate(&e1));
    const int grid = 2 * c->num_sms;
    fma_peak_kernel<<<grid, 1024>>>((float*)c->tc_debug_buf.p, iters, 0.999f, 0.001f);   // warm-up
    GAT_CUDA(cudaEventRecord(e0));
    fma_peak_kernel<<<grid, 1024>>>((float*)c->tc_debug_buf.p, iters, 0.999f, 0.001f);
    GAT_CUDA(cudaEventRecord(e1));
    GAT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.0f;
    GAT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops_host = (float)(2.0 * 128.0 * iters * 1024.0 * grid / (ms * 1e-3) * 1e-12);
    return 0;
#endif
}

extern "C" int gat_set_conv_pass(gat_ctx* c, int32_t mult) {
    if (!c || mult < 1 || mult > 64) return fail("gat_set_conv_pass: bad argument");
    c->conv_pass_mult = mult;
    return 0;
}

extern "C" int gat_set_host_chunks(gat_ctx* c, int32_t min_mult, int32_t max_mult, int32_t order) {
    if (!c || min_mult < 1 || max_mult < min_mult || max_mult > 64 || order < 0 || order > 2)
        return fail("gat_set_host_chunks: bad argument");
    c->host_chunk_mult[0] = min_mult; c->host_chunk_mult[1] = max_mult; c->host_chunk_mult[2] = order;
    return 0;
}

extern "C" int gat_profile_begin(gat_ctx* c) {
    if (!c) return fail("gat_profile_begin: null ctx");
    c->profiling = true;
    return 0;
}

// Stops profiling, waits for the recorded launches and writes lines "name launches total_ms\n" into buf.
extern "C" int gat_profile_end(gat_ctx* c, char* buf, int64_t cap) {
    if (!c || !buf || cap < 1) return fail("gat_profile_end: bad argument");
    c->profiling = false;
    std::vector<std::string> names; std::vector<double> ms; std::vector<long long> cnt;
    for (ProfRec& r : c->prof) {
        float t = 0.0f;
#ifndef GAT_CPU_EMU
        GAT_CUDA(cudaEventSynchronize(r.e1));
        GAT_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
#endif
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
        size_t i = 0;
        for (; i < names.size(); ++i) if (names[i] == r.name) break;
        if (i == names.size()) { names.push_back(r.name); ms.push_back(0.0); cnt.push_back(0); }
        ms[i] += t; cnt[i] += 1;
    }
    c->prof.clear();
    std::string out;
    for (size_t i = 0; i < names.size(); ++i) {
        char line[256];
        snprintf(line, sizeof(line), "%s %lld %.6f\n", names[i].c_str(), cnt[i], ms[i]);
        out += line;
    }
    if ((int64_t)out.size() + 1 > cap) return fail("gat_profile_end: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

extern "C" int gat_ctx_create(const gat_config* cfg, int device, gat_ctx** out) {
    if (!cfg || !out) return fail("gat_ctx_create: null argument");
    if (cfg->mel_n_fft != 512 && cfg->mel_n_fft != 1024 && cfg->mel_n_fft != 2048 && cfg->mel_n_fft != 4096)
        return fail("gat_ctx_create: mel_n_fft=%d unsupported (512, 1024, 2048 or 4096)", cfg->mel_n_fft);
    if (cfg->mel_n_mels < 1 || cfg->mel_n_mels > 32 * kMaxMelsPerLane || cfg->mfcc_n_mels != 128)
        return fail("gat_ctx_create: unsupported mel sizes (mel %d, mfcc %d)", cfg->mel_n_mels, cfg->mfcc_n_mels);
    if (cfg->mel_hop < 1 || (cfg->mel_hop & 1)) return fail("gat_ctx_create: mel_hop must be even");
    if (!cfg->mel_window || !cfg->mel_fb || !cfg->stft_window || !cfg->mfcc_fb || !cfg->dct)
        return fail("gat_ctx_create: missing table pointer");
    GAT_CUDA(cudaSetDevice(device));
    gat_ctx* c = new gat_ctx();
    c->device = device;
    c->cfg = *cfg;
    c->cfg.mel_window = nullptr; c->cfg.mel_fb = nullptr; c->cfg.stft_window = nullptr; c->cfg.mfcc_fb = nullptr; c->cfg.dct = nullptr;
#ifdef GAT_CPU_EMU
    c->num_sms = 2;
#else
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
#endif
    // FFT twiddles (FftTables layout for an n_fft-point frame), computed in long double and rounded once
    auto fft_tables = [](int n_fft, std::vector<Cpx<double>>& tw, std::vector<Cpx<double>>& w2) {
        const int P = n_fft / 64, C = n_fft / 2;
        const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
        tw.resize((size_t)P * 32); w2.resize((size_t)P * 16);
        for (int k2 = 0; k2 < P; ++k2)
            for (int n1 = 0; n1 < 32; ++n1) {
                const long double a = -two_pi * (long double)(n1 * k2) / (long double)C;
                tw[k2 * 32 + n1] = Cpx<double>{(double)cosl(a), (double)sinl(a)};
            }
        for (int k = 0; k < P * 16; ++k) {
            const long double a = -two_pi * (long double)k / (long double)n_fft;
            w2[k] = Cpx<double>{(double)cosl(a), (double)sinl(a)};
        }
    };
    auto to_float = [](const std::vector<Cpx<double>>& a) {
        std::vector<Cpx<float>> f(a.size());
        for (size_t i = 0; i < a.size(); ++i) f[i] = Cpx<float>{(float)a[i].x, (float)a[i].y};
        return f;
    };
    std::vector<Cpx<double>> tw, w2, tw_m, w2_m;
    fft_tables(2048, tw, w2);
    fft_tables(cfg->mel_n_fft, tw_m, w2_m);
    const std::vector<Cpx<float>> twf = to_float(tw), w2f = to_float(w2), twmf = to_float(tw_m), w2mf = to_float(w2_m);
    std::vector<float> win_mfcc(2048);
    for (int i = 0; i < 2048; ++i) win_mfcc[i] = (float)cfg->stft_window[i];
    int rc = 0;
    rc |= upload(c->tw64, tw.data(), tw.size()); rc |= upload(c->w2_64, w2.data(), w2.size());
    rc |= upload(c->tw32, twf.data(), twf.size()); rc |= upload(c->w2_32, w2f.data(), w2f.size());
    rc |= upload(c->tw_mel, twmf.data(), twmf.size()); rc |= upload(c->w2_mel, w2mf.data(), w2mf.size());
    rc |= upload(c->win_mel, cfg->mel_window, (size_t)cfg->mel_n_fft);
    std::vector<float> win_half(cfg->mel_n_fft);
    for (int i = 0; i < cfg->mel_n_fft; ++i) win_half[i] = 0.5f * cfg->mel_window[i];
    rc |= upload(c->win_mel_half, win_half.data(), win_half.size());
    for (float& w : win_mfcc) w *= 0.5f;       // the Hermitian split's 1/2, exact
    rc |= upload(c->win_mfcc_half, win_mfcc.data(), 2048);
    rc |= upload(c->win64, cfg->stft_window, 2048);
    {   // DCT-II rows transposed to [mel band][coefficient]: the finish reads a band's coefficients with ONE coalesced load
        std::vector<float> dct_t((size_t)cfg->mfcc_n_mfcc * cfg->mfcc_n_mels);
        for (int k = 0; k < cfg->mfcc_n_mfcc; ++k)
            for (int j = 0; j < cfg->mfcc_n_mels; ++j) dct_t[(size_t)j * cfg->mfcc_n_mfcc + k] = cfg->dct[(size_t)k * cfg->mfcc_n_mels + j];
        rc |= upload(c->dct, dct_t.data(), dct_t.size());
    }
    rc |= build_sparse_fb(c->fb_mel, cfg->mel_fb, cfg->mel_n_mels, cfg->mel_n_fft / 2 + 1, 1, cfg->mel_n_mels);
    rc |= build_sparse_fb(c->fb_mfcc, cfg->mfcc_fb, cfg->mfcc_n_mels, 1025, 1025, 1);
    if (rc) { gat_ctx_destroy(c); return 1; }
    *out = c;
    return 0;
}

extern "C" void gat_ctx_destroy(gat_ctx* c) {
    if (!c) return;
    DevBuf* all[] = {&c->tw_mel, &c->w2_mel, &c->win_mel_half, &c->win_mfcc_half, &c->clip_count, &c->tw32, &c->w2_32, &c->tw64, &c->w2_64, &c->win_mel, &c->win64, &c->dct,
                     &c->fb_mel.start, &c->fb_mel.len, &c->fb_mel.off, &c->fb_mel.mel, &c->fb_mel.w,
                     &c->fb_mfcc.start, &c->fb_mfcc.len, &c->fb_mfcc.off, &c->fb_mfcc.mel, &c->fb_mfcc.w,
                     &c->mlp_params, &c->conv_w_tc[1], &c->conv_w_tc[2], &c->fc1_w_tc, &c->feat_planes, &c->hid, &c->tc_debug_buf, &c->conv_w[0], &c->conv_w[1], &c->conv_w[2], &c->conv_b[0], &c->conv_b[1], &c->conv_b[2],
                     &c->fc1_w, &c->fc1_b, &c->fc2_w, &c->fc2_b, &c->scaler_mean, &c->scaler_scale,
                     &c->clip_scale, &c->spec, &c->spec_max, &c->f0, &c->act1, &c->act2, &c->act3, &c->hz_tmp, &c->logits_cnn, &c->logits_mlp,
                     &c->seg_small, &c->seg_rms, &c->seg_rms_med, &c->seg_gate, &c->seg_env, &c->seg_envn, &c->seg_cand,
                     &c->seg_peaks, &c->seg_frames, &c->seg_table, &c->seg_keep, &c->seg_dest, &c->seg_base, &c->seg_counts,
                     &c->e2e_audio[0], &c->e2e_audio[1], &c->e2e_pcm[0], &c->e2e_pcm[1], &c->e2e_mel, &c->e2e_mfcc, &c->e2e_probs, &c->e2e_mlp_probs,
                     &c->e2e_cnn_probs, &c->e2e_index, &c->e2e_conf};
    for (DevBuf* b : all) b->release();
    if (c->e2e_streams) { cudaStreamDestroy(c->e2e_stream[0]); cudaStreamDestroy(c->e2e_stream[1]); }
    delete c;
}

extern "C" int gat_load_mlp(gat_ctx* c, const int32_t* dims, int32_t n_linear, const float* params, int64_t n_params) {
    if (!c || !dims || !params) return fail("gat_load_mlp: null argument");
    if (n_linear < 1 || n_linear > kMlpMaxLayers) return fail("gat_load_mlp: %d linear layers unsupported", n_linear);
    int64_t expect = 0;
    for (int l = 0; l < n_linear; ++l) {
        if (dims[l] < 1 || dims[l] > kMlpMaxWidth || dims[l + 1] < 1 || dims[l + 1] > kMlpMaxWidth)
            return fail("gat_load_mlp: layer width out of range (max %d)", kMlpMaxWidth);
        expect += (int64_t)dims[l] * dims[l + 1] + dims[l + 1] + (l + 1 < n_linear ? 2 * dims[l + 1] : 0);
    }
    if (expect != n_params) return fail("gat_load_mlp: expected %lld parameters, got %lld", (long long)expect, (long long)n_params);
    if (dims[n_linear] > 64) return fail("gat_load_mlp: more than 64 classes unsupported");
    if (c->cnn_loaded && c->classes != dims[n_linear]) return fail("gat_load_mlp: class count differs from the CNN's");
    if ((size_t)n_params * 4 + 8 * 2 * kMlpMaxWidth * 4 > 200 * 1024) return fail("gat_load_mlp: parameters exceed shared memory");
    if (upload(c->mlp_params, params, (size_t)n_params)) return 1;
    for (int l = 0; l <= n_linear; ++l) c->mlp_dims[l] = dims[l];
    c->mlp_n_linear = n_linear;
    c->mlp_n_params = (int)n_params;
    return 0;
}

extern "C" int gat_load_cnn(gat_ctx* c, int32_t n_conv, const int32_t* ch, const float* const* conv_w, const float* const* conv_b,
                            int32_t hidden, int32_t classes, const float* fc1_w, const float* fc1_b,
                            const float* fc2_w, const float* fc2_b) {
    if (!c || !ch || !conv_w || !conv_b || !fc1_w || !fc1_b || !fc2_w || !fc2_b) return fail("gat_load_cnn: null argument");
    if (n_conv != 3 || ch[0] != 1 || ch[1] != 32 || ch[2] != 64 || ch[3] != 128)
        return fail("gat_load_cnn: only the reference architecture 1->32->64->128 (3x3, pool 2) is implemented");
    if (hidden < 1 || hidden > 256 || classes < 1 || classes > 64) return fail("gat_load_cnn: hidden/classes out of range");
    if (c->mlp_n_linear && c->mlp_dims[c->mlp_n_linear] != classes) return fail("gat_load_cnn: class count differs from the MLP's");
    for (int i = 0; i < 3; ++i) {
        if (upload(c->conv_w[i], conv_w[i], (size_t)9 * ch[i] * ch[i + 1])) return 1;
        if (upload(c->conv_b[i], conv_b[i], (size_t)ch[i + 1])) return 1;
    }
#ifndef GAT_CPU_EMU
    for (int i = 1; i < 3; ++i) {
        // The weights of a layer are multiplied by 2^S (exact; undone in the epilogue) so that the largest is about 2^14:
        // the FP16 remainders wl of all but vanishing weights then sit in FP16's normal range.
        // One contiguous blob per pipeline stage (K block, half, tap), 6*c_out*16 bytes: wf | wb | wl, each
        // [2 chunks][c_out][8 x 16 bit] = FP16(w'), BF16(FP16(w')), FP16(w' - FP16(w')) with w' = w * 2^S.
        const int cin = ch[i], cout = ch[i + 1];
        float wmax = 0.0f;
        for (size_t k = 0; k < (size_t)9 * cin * cout; ++k) wmax = fmaxf(wmax, fabsf(conv_w[i][k]));
        int S = 0;
        if (wmax > 0.0f) { int e; frexpf(wmax, &e); S = 14 - e; }          // wmax * 2^S in [2^13, 2^14)
        S = S > 24 ? 24 : (S < -24 ? -24 : S);
        c->conv_w_unscale[i] = ldexpf(1.0f, -S);
        const size_t stage = (size_t)6 * cout * 8;                     // 16-bit elements
        std::vector<unsigned short> t((size_t)(cin / 16) * 9 * stage, 0);
        for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < cin; ++ci)
                for (int oc = 0; oc < cout; ++oc) {
                    const float w = ldexpf(conv_w[i][((size_t)tap * cin + ci) * cout + oc], S);
                    const int kh = ci / 16, c = ci % 16;                   // kh = K block * 2 + half
                    unsigned short* st = t.data() + (size_t)(kh * 9 + tap) * stage;
                    const size_t idx = ((size_t)(c / 8) * cout + oc) * 8 + (c % 8);
                    if (cout == 64) {      // conv2: {wf, wl} stacked along N per K chunk, then wb (conv_tc.cuh, FUSE)
                        const size_t cmb = ((size_t)(c / 8) * 2 * cout + oc) * 8 + (c % 8);
                        tc::split16_weight_host(w, st[cmb], st[(size_t)4 * cout * 8 + idx], st[cmb + (size_t)cout * 8]);
                    } else {
                        tc::split16_weight_host(w, st[idx], st[(size_t)2 * cout * 8 + idx], st[(size_t)4 * cout * 8 + idx]);
                    }
                }
        if (upload(c->conv_w_tc[i], t.data(), t.size())) return 1;
    }
    {   // FC1: [k/32][hi|lo][8 chunks][hidden][4]
        if (hidden != 256) return fail("gat_load_cnn: the tensor-core head is built for hidden = 256 (got %d)", hidden);
        const int K = ch[3] * 16;
        std::vector<float> t((size_t)(K / 32) * 2 * 8 * hidden * 4);
        for (int k = 0; k < K; ++k)
            for (int oc = 0; oc < hidden; ++oc) {
                const float w = fc1_w[(size_t)k * hidden + oc];
                const float hi = tc::tf32_hi(w);
                const size_t base = ((size_t)(k / 32) * 2) * 8 * hidden * 4;
                const size_t idx = ((size_t)((k % 32) / 4) * hidden + oc) * 4 + (k % 4);
                t[base + idx] = hi;
                t[base + (size_t)8 * hidden * 4 + idx] = w - hi;
            }
        if (upload(c->fc1_w_tc, t.data(), t.size())) return 1;
    }
#endif
    if (upload(c->fc1_w, fc1_w, (size_t)ch[3] * 16 * hidden) || upload(c->fc1_b, fc1_b, hidden) ||
        upload(c->fc2_w, fc2_w, (size_t)hidden * classes) || upload(c->fc2_b, fc2_b, classes)) return 1;
    for (int i = 0; i < 4; ++i) c->conv_ch[i] = ch[i];
    c->hidden = hidden; c->classes = classes; c->cnn_loaded = true;
    return 0;
}

extern "C" int gat_set_scaler(gat_ctx* c, const double* mean, const double* scale, int32_t n) {
    if (!c) return fail("gat_set_scaler: null ctx");
    if (n <= 0) { c->scaler_n = 0; return 0; }
    if (upload(c->scaler_mean, mean, n) || upload(c->scaler_scale, scale, n)) return 1;
    c->scaler_n = n;
    return 0;
}

extern "C" int gat_set_ensemble_weights(gat_ctx* c, float w_mlp, float w_cnn) {
    if (!c) return fail("gat_set_ensemble_weights: null ctx");
    c->w_mlp = w_mlp; c->w_cnn = w_cnn;
    return 0;
}

// ------------------------------------------------------------------------------------------------- features
namespace {

int launch_clip_scale(gat_ctx* c, const float* audio, int64_t N, int64_t n, void* stream) {
    if (c->clip_scale.ensure((size_t)N * sizeof(float))) return 1;
    LAUNCH(c, clip_scale_kernel, (unsigned)N, 256, 0, stream, audio, (long long)n, c->clip_scale.as<float>());
    return 0;
}

template <typename T, int kOut, int kThreads, int P>
int launch_stft_mel(gat_ctx* c, StftMelParams<T> p, void* stream) {
    using G = FftGeom<P>;
    const int nwarps = kThreads / 32;
    const int per_round = nwarps * G::F;                     // frames one pass of the CTA's warps transforms
    auto smem_for = [&](int fc, bool async) { return stft_mel_smem_bytes<T, P>(nwarps, fc, p.hop, p.fb.n_mels, p.fb.nnz, kOut == kOutImage, async); };
    int fc = (int)(8192 / p.hop) + 1;
    fc = fc > 32 ? 32 : fc;
    if (sizeof(T) == 8) fc = fc > 8 ? 8 : fc;
    if (fc > per_round) fc = fc / per_round * per_round;     // whole rounds
    else if (sizeof(T) == 4) fc = per_round;                 // at least one frame group per warp
    while (fc > per_round && smem_for(fc, false) > 227 * 1024) fc -= per_round;
    while (fc > 1 && smem_for(fc, false) > 227 * 1024) fc = (fc + 1) / 2;
    // Asynchronous prefetch needs a second (raw) copy of the chunk in shared memory: use it when at least one round
    // of frames still fits (CNN chain at hop 256: 16 frames), otherwise stage synchronously with the larger chunk.
    bool async = false;
    if (sizeof(T) == 4) {
        int fa = fc;
        while (fa > per_round && smem_for(fa, true) > 227 * 1024) fa -= per_round;
        if (fa >= per_round && smem_for(fa, true) <= 227 * 1024) { async = true; fc = fa; }
    }
    p.use_async = async ? 1 : 0;
    if (fc > p.n_frames) fc = p.n_frames;
    p.frames_per_cta = fc;
    p.chunks_per_clip = (p.n_frames + fc - 1) / fc;
    const size_t smem = smem_for(fc, async);
    if (smem > 227 * 1024) return fail("stft_mel: %zu bytes of shared memory needed (n_fft %d, hop %d)", smem, G::N, p.hop);
    auto kfn = stft_mel_kernel<T, kOut, kThreads, P>;
    GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long work = (long long)p.N * p.chunks_per_clip;
    const unsigned grid = (unsigned)(work < c->num_sms ? work : c->num_sms);
    KNAME(sizeof(T) == 8 ? "stft_mel_f64_spec" : (kOut == kOutImage ? "stft_mel_f32_image" : "stft_mel_f32_spec"));
    LAUNCH(c, kfn, grid, kThreads, smem, stream, p);
    return 0;
}

// Frame-per-warp kernel (csrc/stft2.cuh) for the float chains at n_fft 2048: image only, MFCC only, or both off one FFT.
int launch_stft_frames(gat_ctx* c, StftFramesParams p, void* stream) {
    const long long n_items = (long long)p.N * (p.fa + p.fb_items);
    if (n_items <= 0) return 0;
    const int nnz1 = p.img ? p.fb.nnz : 0, nnz2 = p.spec ? p.fb2.nnz : 0;
    if (!p.img) p.fb.nnz = 0;
    if (!p.spec) p.fb2.nnz = 0;
    const size_t budget = 227 * 1024;
    const int threads = stft_frames_smem_bytes(20, nnz1 + 3, nnz2 + 3) <= budget ? 640 : 512;
    const size_t smem = stft_frames_smem_bytes(threads / 32, nnz1 + 3, nnz2 + 3);
    if (smem > budget) return fail("stft_frames: %zu bytes of shared memory needed", smem);
    // contiguous item ranges per CTA; enough CTAs that a short call (a single note) still spreads over the SMs
    const int nwarps = threads / 32;
    long long ctas = (n_items + nwarps - 1) / nwarps;
    ctas = ctas < c->num_sms ? ctas : c->num_sms;
    p.items_per_cta = (n_items + ctas - 1) / ctas;
    ctas = (n_items + p.items_per_cta - 1) / p.items_per_cta;
    KNAME(p.img && p.spec ? "stft_frames_dual" : (p.img ? "stft_frames_image" : "stft_frames_mfcc"));
    if (threads == 640) {
        auto kfn = stft_frames_kernel<640>;
        GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(c, kfn, (unsigned)ctas, 640, smem, stream, p);
    } else {
        auto kfn = stft_frames_kernel<512>;
        GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(c, kfn, (unsigned)ctas, 512, smem, stream, p);
    }
    return 0;
}

// Frame-per-warp image chain at n_fft 1024 / 512 (csrc/stft2.cuh, stft_frames_small_kernel): F = 2 / 4 frames per warp.
template <int P>
int launch_stft_frames_small(gat_ctx* c, StftFramesParams p, void* stream) {
    constexpr int F = FftGeom<P>::F;
    const long long n_items = (long long)p.N * ((p.T + F - 1) / F);
    if (n_items <= 0) return 0;
    const int nwarps = 20;
    const size_t smem = stft_frames_small_smem_bytes<P>(nwarps, p.fb.nnz + 3);
    if (smem > (size_t)227 * 1024) return fail("stft_frames_small: %zu bytes of shared memory needed", smem);
    long long ctas = (n_items + nwarps - 1) / nwarps;
    ctas = ctas < c->num_sms ? ctas : c->num_sms;
    p.items_per_cta = (n_items + ctas - 1) / ctas;
    ctas = (n_items + p.items_per_cta - 1) / p.items_per_cta;
    auto kfn = stft_frames_small_kernel<640, P>;
    GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KNAME("stft_frames_image");
    LAUNCH(c, kfn, (unsigned)ctas, 640, smem, stream, p);
    return 0;
}

// Fills the chain-independent fields and the two chains' halves of StftFramesParams.
void stft_frames_common(gat_ctx* c, StftFramesParams& p, const float* audio, int64_t N, int64_t n, bool dual_or_img) {
    p.audio = audio; p.n = n; p.N = (int)N; p.clip_scale = c->clip_scale.as<float>();
    p.window_half = dual_or_img ? c->win_mel_half.as<float>() : c->win_mfcc_half.as<float>();
    p.tw = c->tw32.as<Cpx<float>>(); p.w2 = c->w2_32.as<Cpx<float>>();
    p.amin = 1e-10f;
}
void stft_frames_image(gat_ctx* c, StftFramesParams& p, int64_t n, bool normalize, bool power_out, float* out) {
    p.img = 1; p.hop = c->cfg.mel_hop; p.T = (int)(1 + n / c->cfg.mel_hop); p.fb = c->fb_mel.view();
    p.norm_img = normalize ? 1 : 0; p.power_out = power_out ? 1 : 0; p.out = out; p.fa = p.T;
}
int stft_frames_spec(gat_ctx* c, StftFramesParams& p, int64_t N, int64_t n, bool normalize, float* out, int ld, void* stream) {
    const int T2 = (int)(1 + n / 512);
    if (c->spec.ensure((size_t)N * T2 * 128 * sizeof(float)) || c->clip_count.ensure((size_t)N * sizeof(unsigned))) return 1;
    GAT_CUDA(cudaMemsetAsync(c->clip_count.p, 0, (size_t)N * sizeof(unsigned), (cudaStream_t)stream));
    p.spec = 1; p.T2 = T2; p.fb2 = c->fb_mfcc.view(); p.norm_spec = normalize ? 1 : 0;
    p.spec_scratch = c->spec.as<float>(); p.clip_count = c->clip_count.as<unsigned>();
    p.dct = c->dct.as<float>(); p.n_mfcc = c->cfg.mfcc_n_mfcc; p.top_db = 80.0f; p.mfcc_out = out; p.ld = ld;
    p.share_stride = 0; p.u_lo = 0; p.u_hi = 0; p.fb_items = T2;
    return 0;
}

int run_melspec(gat_ctx* c, const float* audio, int64_t N, int64_t n, bool normalize, bool scale_ready, float* out, void* stream,
                bool power_out = false) {
    const int n_fft = c->cfg.mel_n_fft;
    if (n <= n_fft / 2) return fail("melspec: clips of %lld samples are too short for reflect padding of %d", (long long)n, n_fft / 2);
    if (normalize && !scale_ready && launch_clip_scale(c, audio, N, n, stream)) return 1;
#ifndef GAT_CPU_EMU_OLD_STFT
    if (n_fft == kStft2N) {      // the reference's N_FFT: frame-per-warp kernel
        StftFramesParams q{};
        stft_frames_common(c, q, audio, N, n, true);
        stft_frames_image(c, q, n, normalize, power_out, out);
        return launch_stft_frames(c, q, stream);
    }
#endif
    static const bool chunked = getenv("GAT_STFT_CHUNKED") != nullptr;      // A/B switch: the round-1 kernel for n_fft 512 / 1024
    if ((n_fft == 1024 || n_fft == 512) && !chunked && n < 0x7fff0000LL) {
        StftFramesParams q{};
        q.audio = audio; q.n = n; q.N = (int)N; q.clip_scale = c->clip_scale.as<float>();
        q.window_half = c->win_mel_half.as<float>(); q.tw = c->tw_mel.as<Cpx<float>>(); q.w2 = c->w2_mel.as<Cpx<float>>();
        q.amin = 1e-10f;
        stft_frames_image(c, q, n, normalize, power_out, out);
        return n_fft == 1024 ? launch_stft_frames_small<16>(c, q, stream) : launch_stft_frames_small<8>(c, q, stream);
    }
    StftMelParams<float> p{};
    p.audio = audio; p.n = n; p.N = (int)N; p.clip_scale = normalize ? c->clip_scale.as<float>() : nullptr;
    p.frame_gate = nullptr; p.sample_gate = 0.0f; p.gate_hop = 512;
    p.hop = c->cfg.mel_hop; p.n_frames = (int)(1 + n / c->cfg.mel_hop); p.pad_mode = kPadReflect;
    p.window = c->win_mel.as<float>(); p.window_half = c->win_mel_half.as<float>();
    p.tw = c->tw_mel.as<Cpx<float>>(); p.w2 = c->w2_mel.as<Cpx<float>>();
    p.fb = c->fb_mel.view(); p.amin = 1e-10f; p.out = out; p.spec_max = nullptr; p.power_out = power_out ? 1 : 0;
    switch (n_fft) {     // MelSpecConfig.N_FFT is configurable (features.py:296-302; BASELINE config 5 sweeps it)
        case 512:  return launch_stft_mel<float, kOutImage, 512, 8>(c, p, stream);
        case 1024: return launch_stft_mel<float, kOutImage, 512, 16>(c, p, stream);
        case 2048: return launch_stft_mel<float, kOutImage, 512, 32>(c, p, stream);
        case 4096: return launch_stft_mel<float, kOutImage, 256, 64>(c, p, stream);
    }
    return fail("melspec: n_fft %d unsupported", n_fft);
}

int run_mfcc(gat_ctx* c, const float* audio, int64_t N, int64_t n, bool normalize, bool scale_ready, float* out, int ld, void* stream) {
    if (normalize && !scale_ready && launch_clip_scale(c, audio, N, n, stream)) return 1;
    StftFramesParams q{};
    stft_frames_common(c, q, audio, N, n, false);
    if (stft_frames_spec(c, q, N, n, normalize, out, ld, stream)) return 1;
    return launch_stft_frames(c, q, stream);
}

// Image + MFCC off ONE FFT per image frame (gat_transcribe_clips): possible when both chains transform 2048-sample
// frames and the MFCC hop (512) is a multiple of the image hop.
bool can_fuse_chains(const gat_ctx* c) { return c->cfg.mel_n_fft == kStft2N && c->cfg.mel_hop <= 512 && 512 % c->cfg.mel_hop == 0; }

int run_dual(gat_ctx* c, const float* audio, int64_t N, int64_t n, bool norm_mel, bool norm_mfcc, float* mel, float* mfcc, int ld, void* stream) {
    if (n <= kStft2N / 2) return fail("melspec: clips of %lld samples are too short for reflect padding of %d", (long long)n, kStft2N / 2);
    StftFramesParams q{};
    stft_frames_common(c, q, audio, N, n, true);
    stft_frames_image(c, q, n, norm_mel, false, mel);
    if (stft_frames_spec(c, q, N, n, norm_mfcc, mfcc, ld, stream)) return 1;
    // spec frames [u_lo, u_hi) touch no padding and ride on image frame u * stride; the others get their own zero-padded FFT
    q.share_stride = 512 / c->cfg.mel_hop;
    const int u_lo = 2;                                                    // 512 u >= 1024
    long long u_hi = n >= kStft2N ? (n - kStft2N / 2) / 512 + 1 : 0;       // 512 u + 1024 <= n
    u_hi = u_hi > q.T2 ? q.T2 : u_hi;
    q.u_lo = u_lo < q.T2 ? u_lo : q.T2;
    q.u_hi = (int)(u_hi < q.u_lo ? q.u_lo : u_hi);
    q.fb_items = q.u_lo + (q.T2 - q.u_hi);
    return launch_stft_frames(c, q, stream);
}

template <int kLPT>
int launch_yin(gat_ctx* c, const YinParams& p, void* stream) {
    int threads = 384;                        // as many warps (at most 12) as 227 KB of shared memory hold: 10 at 33 lags per lane
    while (threads > 32 && (size_t)(threads / 32) * yin_smem_per_warp<kLPT>() + 64 > (size_t)227 * 1024) threads -= 32;
    const size_t smem = (size_t)(threads / 32) * yin_smem_per_warp<kLPT>() + 64;
    auto kfn = yin_kernel<kLPT>;
    GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long work = (long long)p.N * ((p.T + p.seg_frames - 1) / p.seg_frames);
    const long long ctas = (work + threads / 32 - 1) / (threads / 32);
    const unsigned grid = (unsigned)(ctas < c->num_sms ? ctas : c->num_sms);
    KNAME("yin_kernel");
    LAUNCH(c, kfn, grid, threads, smem, stream, p);
    return 0;
}

// packed-FMA variant (yin_pair_kernel): lags interleaved by parity across the two half-warps
template <int kLPT>
int launch_yin_pair(gat_ctx* c, const YinParams& p, void* stream) {
    const int threads = 512;
    const size_t per_warp = (yin_pair_smem_per_warp<kLPT>() + 15) / 16 * 16;
    const size_t smem = (size_t)(threads / 32) * per_warp + 64;
    auto kfn = yin_pair_kernel<kLPT>;
    GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long work = (long long)p.N * ((p.T + p.seg_frames - 1) / p.seg_frames);
    const long long ctas = (work + threads / 32 - 1) / (threads / 32);
    const unsigned grid = (unsigned)(ctas < c->num_sms ? ctas : c->num_sms);
    KNAME("yin_kernel");
    LAUNCH(c, kfn, grid, threads, smem, stream, p);
    return 0;
}

// block-FFT variant (yin_fft_kernel): one forward transform per 512-sample block, one inverse per frame pair
template <int kLPT>
int launch_yin_fft(gat_ctx* c, YinParams p, int64_t N, void* stream) {
    constexpr int nwarps = yin_fft_warps<kLPT>();
    const size_t smem = yin_fft_smem_bytes<kLPT>();
    auto kfn = yin_fft_kernel<kLPT>;
    GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Segment length: a segment of s frames costs (s + 1) forward and ceil(s / 2) inverse transforms; warps take segments
    // round-robin.  Minimise the transforms of the busiest warp (long segments amortise the extra block, short ones balance).
    // Only EVEN lengths (or the whole clip): frames are then always paired as (2m, 2m + 1) whatever the batch size, so a
    // clip's result does not depend on the batch it is in (the host entry point's chunks must equal the resident call).
    const long long warps = (long long)c->num_sms * nwarps;
    int best = p.T; long long best_cost = -1;
    for (int s = 2; s <= p.T + 1; s += 2) {
        const int len = s < p.T ? s : p.T;
        const int n_seg = (p.T + len - 1) / len;
        const long long per_warp = ((long long)N * n_seg + warps - 1) / warps;
        const long long cost = per_warp * ((len + 1) + (len + 1) / 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = len; }
    }
    if (const char* e = getenv("GAT_YIN_SEG")) { const int s = atoi(e); if (s >= 2) best = s < p.T ? (s & ~1) : p.T; }    // test hook: force a segment length
    p.seg_frames = best;
    p.tw = c->tw32.as<Cpx<float>>();
    const long long work = (long long)N * ((p.T + p.seg_frames - 1) / p.seg_frames);
    const long long ctas = (work + nwarps - 1) / nwarps;
    const unsigned grid = (unsigned)(ctas < c->num_sms ? ctas : c->num_sms);
    KNAME("yin_kernel");
    LAUNCH(c, kfn, grid, 32 * nwarps, smem, stream, p);
    return 0;
}

int run_yin(gat_ctx* c, const float* audio, int64_t N, int64_t n, bool normalize, bool scale_ready, double* hz, double* f0_frames,
            float* feat, int ld, int col, void* stream) {
    const int T = (int)(1 + n / 512);
    if (normalize && !scale_ready && launch_clip_scale(c, audio, N, n, stream)) return 1;
    double* f0 = f0_frames;
    if (!f0) { if (c->f0.ensure((size_t)N * T * sizeof(double))) return 1; f0 = c->f0.as<double>(); }
    if (!hz) { if (c->hz_tmp.ensure((size_t)N * sizeof(double))) return 1; hz = c->hz_tmp.as<double>(); }
    YinParams p{};
    p.audio = audio; p.n = n; p.N = (int)N; p.clip_scale = normalize ? c->clip_scale.as<float>() : nullptr;
    p.T = T; p.hop = 512; p.sr = c->cfg.sample_rate;
    p.min_period = (int)floor((double)c->cfg.sample_rate / c->cfg.yin_fmax);
    const int mp = (int)ceil((double)c->cfg.sample_rate / c->cfg.yin_fmin);
    p.max_period = mp < kYinFrame - kYinWin - 1 ? mp : kYinFrame - kYinWin - 1;
    if (p.min_period < 1 || p.max_period <= p.min_period + 1) return fail("yin: period range [%d, %d] unusable", p.min_period, p.max_period);
    p.trough_threshold = c->cfg.yin_trough_threshold; p.f0 = f0;
    // A warp walks `seg_frames` frames of a clip with seg_frames + 1 blocks of work (12/11 blocks per frame at 11).  With
    // few clips (a single note) shorter segments trade redundant blocks for parallelism: latency, not throughput.
    const long long warps = (long long)c->num_sms * 16;
    int seg = 11;
    while (seg > 1 && (long long)N * ((T + seg - 1) / seg) < warps) seg = (seg + 1) / 2;
    p.seg_frames = T < seg ? T : seg;
    const int lags = p.max_period + 1;
    int rc;
    static const bool direct_form = getenv("GAT_YIN_DIRECT") != nullptr;      // A/B switch for the profiling tools
    if (!direct_form && lags <= 7 * 32 && n < 0x7fff0000LL) rc = launch_yin_fft<7>(c, p, N, stream);
    else if (!direct_form && lags <= 14 * 32 && n < 0x7fff0000LL) rc = launch_yin_fft<14>(c, p, N, stream);
    else if (!direct_form && lags <= 16 * 32 && n < 0x7fff0000LL) rc = launch_yin_fft<16>(c, p, N, stream);
    else if (lags <= 7 * 32) rc = launch_yin_pair<7>(c, p, stream);
    else if (lags <= 14 * 32) rc = launch_yin_pair<14>(c, p, stream);     // sr 22050: 442 lags -> 448 computed (15 per lane would be 480)
    else if (lags <= 15 * 32) rc = launch_yin_pair<15>(c, p, stream);
    else rc = launch_yin<33>(c, p, stream);
    if (rc) return 1;
    YinMedianParams m{f0, (int)N, T, hz, feat, ld, col};
    LAUNCH(c, yin_median_kernel, (unsigned)((N + 3) / 4), 128, 0, stream, m);
    return 0;
}

}  // namespace

extern "C" int gat_melspec_db(gat_ctx* c, const float* audio, int64_t N, int64_t n, int32_t normalize, float* out, void* stream) {
    if (!c || !audio || !out) return fail("gat_melspec_db: null argument");
    if (N <= 0) return 0;
    return run_melspec(c, audio, N, n, (normalize & GAT_MEL_NORMALIZE) != 0, false, out, stream, (normalize & GAT_MEL_POWER) != 0);
}

extern "C" int gat_mfcc_features(gat_ctx* c, const float* audio, int64_t N, int64_t n, int32_t normalize, int32_t add_pitch,
                                 int32_t yin_on_normalized, int32_t apply_scaler, float* out, int32_t ld, double* yin_hz, void* stream) {
    if (!c || !audio || !out) return fail("gat_mfcc_features: null argument");
    if (N <= 0) return 0;
    const int F = c->cfg.mfcc_n_mfcc + (add_pitch ? 1 : 0);
    if (ld < F) return fail("gat_mfcc_features: ld=%d < %d feature columns", ld, F);
    if (n < 1) return fail("gat_mfcc_features: empty clips");
    if (apply_scaler && c->scaler_n != F) return fail("gat_mfcc_features: scaler has %d columns, features have %d", c->scaler_n, F);
    if (run_mfcc(c, audio, N, n, normalize != 0, false, out, ld, stream)) return 1;
    if (add_pitch) {
        const bool yn = normalize && yin_on_normalized;
        if (run_yin(c, audio, N, n, yn, normalize != 0, yin_hz, nullptr, out, ld, c->cfg.mfcc_n_mfcc, stream)) return 1;
    }
    if (apply_scaler) {
        if (c->scaler_n != F) return fail("gat_mfcc_features: scaler has %d columns, features have %d", c->scaler_n, F);
        const long long tot = (long long)N * F;
        LAUNCH(c, standard_scale_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, out, (int)N, F, ld,
               c->scaler_mean.as<double>(), c->scaler_scale.as<double>());
    }
    return 0;
}

extern "C" int gat_yin(gat_ctx* c, const float* audio, int64_t N, int64_t n, int32_t normalize, double* hz, double* f0_frames, void* stream) {
    if (!c || !audio || !hz) return fail("gat_yin: null argument");
    if (N <= 0) return 0;
    return run_yin(c, audio, N, n, normalize != 0, false, hz, f0_frames, nullptr, 0, 0, stream);
}

// ------------------------------------------------------------------------------------------------- inference
namespace {

#ifdef GAT_CPU_EMU
#include "emu_run_cnn.inc"   // tests/emu: CUDA-core stand-in for the tensor-core CNN (host-emulation build only)
#else
// conv1 on CUDA cores (C_in = 1), conv2/conv3 as tcgen05 implicit GEMMs (csrc/conv_tc.cuh), head once per batch.
// Clips go through the convs in passes of conv_pass_mult * num_sms (default 28 x 148 = 4144): act1 + act2 are hf / lb
// chunk planes, 4 bytes per element (~0.3 MB per clip at T = 87); measured, long passes beat keeping them inside L2.
constexpr size_t kActGuard = 65536;   // bytes before/after the plane buffers: halo reads of edge groups stay in bounds

// Tiling of one conv layer: column blocks of `cw` output columns (one block spanning the width when the staged
// planes fit in shared memory), R image rows per group so that R * seg <= 384 tile pixels.
struct ConvTiling { int seg, cw, col_blocks, R, groups_per_clip; size_t smem; int tiles; };

// Picks the column blocking that needs the fewest 384-pixel groups per clip: nb column blocks of cw output
// columns (cw even when nb > 1 so that 2x2 pool pairs stay inside a block), seg = cw + 2 staged pixels per row,
// R rows per group.  Narrower blocks than shared memory allows often waste less of a group (W = 86: two blocks
// of 44 columns x 8 rows fill 90 % of a group, one 64-column block x 6 rows + a 24-column remainder only 60 %).
template <int COUT>
ConvTiling conv_tc_tiling(int H, int W, int nstage, bool allow_two_tiles = false) {
    ConvTiling best{};
    long long best_cost = 0;
    const size_t budget = 227 * 1024;
    const int hmax = 2 * (H / 2);
    // groups of 3 tiles (384 pixels) or, where offered, of 2: the cost of a clip is its number of 128-pixel tiles multiplied
    for (int tiles = kTcTiles; tiles >= (allow_two_tiles ? 2 : kTcTiles); --tiles)
        for (int nb = 1; nb <= (W + 1) / 2; ++nb) {
            int cw = (W + nb - 1) / nb;
            if (nb > 1 && (cw & 1)) ++cw;
            const int seg = cw + 2;
            const int group_pix = 128 * tiles;
            if (2 * seg > group_pix || conv_tc_smem_bytes<COUT>(seg, nstage) > budget) continue;
            ConvTiling t{};
            t.seg = seg; t.cw = cw; t.col_blocks = (W + cw - 1) / cw; t.tiles = tiles;
            int R = 2 * (group_pix / (2 * seg));
            R = R < kTcMaxRows ? R : kTcMaxRows;                          // the epilogue's staged planes pad every row by one float
            t.R = R < hmax ? R : hmax;
            if (t.R < 2) continue;
            t.groups_per_clip = ceil_div(H / 2, t.R / 2) * t.col_blocks;
            t.smem = conv_tc_smem_bytes<COUT>(seg, nstage);
            const long long cost = (long long)t.groups_per_clip * tiles;
            if (best.groups_per_clip == 0 || cost < best_cost) { best = t; best_cost = cost; }
        }
    return best;
}

int run_cnn(gat_ctx* c, const float* mel, int64_t N, int T, float* cnn_probs, float* cnn_logits, void* stream) {
    if (!c->cnn_loaded) return fail("infer: no CNN loaded (gat_load_cnn)");
    const int H0 = c->cfg.mel_n_mels, W0 = T;
    const int H1 = H0 / 2, W1 = W0 / 2, H2 = H1 / 2, W2 = W1 / 2, H3 = H2 / 2, W3 = W2 / 2;
    if (H3 < 1 || W3 < 1) return fail("infer: mel image %dx%d too small for three 2x2 pools", H0, W0);
    const ConvTiling t2 = conv_tc_tiling<64>(H1, W1, 8), t3 = conv_tc_tiling<128>(H2, W2, 3, true);
    if (t2.R < 2 || t3.R < 2 || t2.smem > 227 * 1024 || t3.smem > 227 * 1024)
        return fail("infer: no tensor-core conv tiling for a mel image of %d x %d", H0, W0);
    const size_t smem2 = t2.smem, smem3 = t3.smem;
    const long long per_pass = (long long)c->num_sms * c->conv_pass_mult;
    const long long chunk = N < per_pass ? N : per_pass;
    const size_t P1 = (size_t)(H1 + 2) * (W1 + 2), P2 = (size_t)(H2 + 2) * (W2 + 2);
    // The plane buffers are laid out for `cap` clips: the largest pass seen at this image size.  A smaller call reuses
    // them as they are (their zero borders stay valid: the kernels only ever write interiors), so the host entry
    // points can mix short and long chunks without re-zeroing hundreds of megabytes at every change of size.
    const bool same_geom = c->act_shape[1] == H0 && c->act_shape[2] == W0;
    const long long cap = same_geom && c->act_shape[0] >= chunk ? c->act_shape[0] : chunk;
    const size_t a1 = (size_t)cap * 4 * P1 * 16, a2 = (size_t)cap * 8 * P2 * 16;   // bytes of ONE of the two arrays (hf, lb)
    const size_t a3 = (size_t)N * H3 * W3 * 128 * 4;
    const bool fresh = !same_geom || cap != c->act_shape[0] || c->act1.cap < 2 * a1 + 2 * kActGuard || c->act2.cap < 2 * a2 + 2 * kActGuard;
    if (c->act1.ensure(2 * a1 + 2 * kActGuard) || c->act2.ensure(2 * a2 + 2 * kActGuard) || c->act3.ensure(a3)) return 1;
    if (fresh) {   // zero borders (and guards) once per geometry
        GAT_CUDA(cudaMemsetAsync(c->act1.p, 0, 2 * a1 + 2 * kActGuard, (cudaStream_t)stream));
        GAT_CUDA(cudaMemsetAsync(c->act2.p, 0, 2 * a2 + 2 * kActGuard, (cudaStream_t)stream));
        c->act_shape[0] = cap; c->act_shape[1] = H0; c->act_shape[2] = W0;
    }
    // each activation buffer: [hf | lb] chunk-plane arrays of a bytes each, between two guard bands
    auto arr = [](DevBuf& b, size_t a, int k) { return reinterpret_cast<unsigned short*>(b.as<unsigned char>() + kActGuard + (size_t)k * a); };
    unsigned short *act1_hf = arr(c->act1, a1, 0), *act1_lb = arr(c->act1, a1, 1);
    unsigned short *act2_hf = arr(c->act2, a2, 0), *act2_lb = arr(c->act2, a2, 1);
    // FC1 operand planes (written by conv3's epilogue when a clip is one group, else by avgpool_planes_kernel)
    const long long rows_pad = (N + 127) / 128 * 128;
    const size_t feat_bytes = (size_t)512 * rows_pad * 16;
    if (c->feat_planes.ensure(2 * feat_bytes) || c->hid.ensure((size_t)N * 256 * 4)) return 1;
    float* feat_hi = c->feat_planes.as<float>();
    float* feat_lo = reinterpret_cast<float*>(c->feat_planes.as<unsigned char>() + feat_bytes);
    const bool fuse_avgpool = t3.groups_per_clip == 1 && H3 * W3 <= kTcPooledPix;
    auto k2 = conv_tc_kernel<32, 64, 8>;
    auto k3 = t3.tiles == 2 ? conv_tc_kernel<64, 128, 3, 2> : conv_tc_kernel<64, 128, 3, 3>;   // two tiles: short clips (<= 256 staged pixels)
    GAT_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    GAT_CUDA(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    for (long long c0 = 0; c0 < N; c0 += chunk) {
        const int nc = (int)(N - c0 < chunk ? N - c0 : chunk);
        Conv1PlanesParams p1{mel + c0 * H0 * W0, nc, H0, W0, c->conv_w[0].as<float>(), c->conv_b[0].as<float>(), act1_hf, act1_lb, 0.01f};
        LAUNCH(c, conv1_pool_planes_kernel, (unsigned)(nc * ceil_div(H1 * W1, 256)), 256, 0, stream, p1);
        ConvTcParams p2{act1_hf, act1_lb, c->conv_w_tc[1].as<unsigned short>(), c->conv_w_unscale[1], c->conv_b[1].as<float>(), nc, H1, W1, t2.R,
                        t2.seg, t2.cw, t2.col_blocks, t2.groups_per_clip, 1, nullptr, act2_hf, act2_lb, nullptr, nullptr, 0, 0, 0.01f,
                        c->tc_debug ? c->tc_debug_buf.as<long long>() : nullptr};
        const int work2 = nc * p2.groups_per_clip;
        KNAME("conv2_tc_32_64");
        LAUNCH(c, k2, (unsigned)(work2 < c->num_sms ? work2 : c->num_sms), conv_tc_threads(conv_tc_epi_groups(64)), smem2, stream, p2);
        ConvTcParams p3{act2_hf, act2_lb, c->conv_w_tc[2].as<unsigned short>(), c->conv_w_unscale[2], c->conv_b[2].as<float>(), nc, H2, W2, t3.R,
                        t3.seg, t3.cw, t3.col_blocks, t3.groups_per_clip, fuse_avgpool ? 2 : 0,
                        c->act3.as<float>() + (size_t)c0 * H3 * W3 * 128, nullptr, nullptr, feat_hi, feat_lo, rows_pad, c0, 0.01f,
                        c->tc_debug ? c->tc_debug_buf.as<long long>() + 148 * 8 : nullptr};
        const int work3 = nc * p3.groups_per_clip;
        KNAME("conv3_tc_64_128");
        LAUNCH(c, k3, (unsigned)(work3 < c->num_sms ? work3 : c->num_sms), conv_tc_threads(conv_tc_epi_groups(128)), smem3, stream, p3);
    }
    // head: (adaptive average pool ->) FC1 (tcgen05) -> FC2 + softmax
    if (!fuse_avgpool) {
        AvgPoolPlanesParams pa{c->act3.as<float>(), (int)N, H3, W3, 128, feat_hi, feat_lo, rows_pad};
        LAUNCH(c, avgpool_planes_kernel, (unsigned)((N * 4 * 128 + 255) / 256), 256, 0, stream, pa);
    }
    auto kf = fc_tc_kernel<256>;
    const size_t fc_smem = fc_tc_smem_bytes<256>();
    GAT_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fc_smem));
    // K is always cut into the same 8 slices, whatever the batch: the summation order of a clip's hidden vector - and so
    // its probabilities, bit for bit - must not depend on how many other clips are in the call (tests check that).
    // 32 row blocks x 8 slices at 4096 clips; 8 CTAs instead of one walking all of K for a single note.
    const int row_blocks = (int)(rows_pad / 128);
    const int k_splits = 8;
    if (c->hid.ensure((size_t)k_splits * N * 256 * 4)) return 1;
    FcTcParams pf{feat_hi, feat_lo, rows_pad, c->fc1_w_tc.as<float>(), (int)N, 2048, k_splits, c->hid.as<float>()};
    KNAME("fc1_tc_2048_256");
    LAUNCH(c, kf, dim3((unsigned)row_blocks, (unsigned)k_splits), 192, fc_smem, stream, pf);
    Fc2Params p2f{c->hid.as<float>(), (int)N, 256, k_splits, c->fc1_b.as<float>(), 0.01f, c->fc2_w.as<float>(), c->fc2_b.as<float>(),
                  c->classes, cnn_logits, cnn_probs};
    const size_t fc2_smem = (size_t)((256 * c->classes + 3) & ~3) * 4 + (size_t)8 * 256 * 4 + 64;
    GAT_CUDA(cudaFuncSetAttribute(fc2_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fc2_smem));
    const long long ctas2 = (N + 7) / 8;
    LAUNCH(c, fc2_softmax_kernel, (unsigned)(ctas2 < 2 * c->num_sms ? ctas2 : 2 * c->num_sms), 256, fc2_smem, stream, p2f);
    return 0;
}
#endif

int run_mlp_ensemble(gat_ctx* c, const float* mfcc, int ld, int64_t N, const float* cnn_probs, float* probs, float* mlp_probs,
                     float* mlp_logits, int64_t* index, float* conf, void* stream) {
    if (!c->mlp_n_linear) return fail("infer: no MLP loaded (gat_load_mlp)");
    MlpParams p{};
    p.x = mfcc; p.N = (int)N; p.ld = ld; p.params = c->mlp_params.as<float>(); p.n_params = c->mlp_n_params;
    p.n_linear = c->mlp_n_linear;
    for (int l = 0; l <= c->mlp_n_linear; ++l) p.dims[l] = c->mlp_dims[l];
    p.slope = 0.1f; p.ln_eps = 1e-5f; p.cnn_probs = cnn_probs; p.w_mlp = c->w_mlp; p.w_cnn = c->w_cnn;
    p.mlp_logits = mlp_logits; p.mlp_probs = mlp_probs; p.probs = probs; p.index = (long long*)index; p.conf = conf;
    const size_t smem = ((size_t)c->mlp_n_params + 8 * 2 * kMlpMaxWidth) * sizeof(float) + 64;
    GAT_CUDA(cudaFuncSetAttribute(mlp_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ctas = (N + 7) / 8;
    LAUNCH(c, mlp_ensemble_kernel, (unsigned)(ctas < c->num_sms ? ctas : c->num_sms), 256, smem, stream, p);
    return 0;
}

}  // namespace

// argmax / confidence when only the CNN runs (GAT_FLAG_SKIP_MLP): probs == cnn_probs
namespace gat {
__global__ void argmax_kernel(const float* __restrict__ probs, int N, int classes, long long* __restrict__ index, float* __restrict__ conf) {
    const int clip = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (clip >= N) return;
    const int lane = lane_id();
    float bv = -1.0f; int bi = 0x7fffffff;
    for (int k = lane; k < classes; k += 32) {
        const float v = probs[(long long)clip * classes + k];
        if (v > bv) { bv = v; bi = k; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { index[clip] = bi; conf[clip] = bv; }
}
}  // namespace gat

extern "C" int gat_infer(gat_ctx* c, const float* mfcc, int32_t ld, const float* mel, int64_t N, int32_t T, float* probs,
                         float* mlp_probs, float* cnn_probs, int64_t* index, float* conf, float* mlp_logits, float* cnn_logits,
                         void* stream) {
    if (!c || !mfcc || !mel || !probs || !mlp_probs || !cnn_probs || !index || !conf) return fail("gat_infer: null argument");
    if (N <= 0) return 0;
    if (ld < c->mlp_dims[0]) return fail("gat_infer: ld=%d < %d MLP inputs", ld, c->mlp_dims[0]);
    const int classes = c->classes;
    float* cl = cnn_logits; float* ml = mlp_logits;
    // logits are optional for the caller but the kernels always write them: use ctx scratch
    if (!cl) { if (c->logits_cnn.ensure((size_t)N * classes * sizeof(float))) return 1; cl = c->logits_cnn.as<float>(); }
    if (!ml) { if (c->logits_mlp.ensure((size_t)N * classes * sizeof(float))) return 1; ml = c->logits_mlp.as<float>(); }
    if (run_cnn(c, mel, N, T, cnn_probs, cl, stream)) return 1;
    return run_mlp_ensemble(c, mfcc, ld, N, cnn_probs, probs, mlp_probs, ml, index, conf, stream);
}

// ------------------------------------------------------------------------------------------------- file front end
extern "C" int gat_pcm16_roundtrip(gat_ctx* c, float* audio, int64_t count, void* stream) {
    if (!c || (!audio && count > 0) || count < 0) return fail("gat_pcm16_roundtrip: bad argument");
    if (count == 0) return 0;
    const long long blocks = (count + 255) / 256;
    const unsigned grid = (unsigned)(blocks < 8LL * c->num_sms ? blocks : 8LL * c->num_sms);
    LAUNCH(c, pcm16_roundtrip_kernel, grid, 256, 0, stream, audio, (long long)count);
    return 0;
}

extern "C" int gat_decode_mono(gat_ctx* c, const void* frames_dev, int32_t sample_format, int64_t frames, int32_t channels,
                               float* out, void* stream) {
    if (!c || !frames_dev || !out || frames < 1 || channels < 1) return fail("gat_decode_mono: bad argument");
    const long long blocks = (frames + 255) / 256;
    const unsigned grid = (unsigned)(blocks < 8LL * c->num_sms ? blocks : 8LL * c->num_sms);
    if (sample_format == GAT_SAMPLE_PCM16) {
        LAUNCH(c, pcm16_to_mono_kernel, grid, 256, 0, stream, (const short*)frames_dev, (long long)frames, (int)channels, out);
    } else if (sample_format == GAT_SAMPLE_FLOAT32) {
        LAUNCH(c, f32_to_mono_kernel, grid, 256, 0, stream, (const float*)frames_dev, (long long)frames, (int)channels, out);
    } else {
        return fail("gat_decode_mono: sample_format %d unknown", sample_format);
    }
    return 0;
}

extern "C" int gat_resample(gat_ctx* c, const float* in, int64_t N, int64_t n_in, int32_t up, int32_t down,
                            const double* taps_dev, int32_t half_len, float* out, int64_t n_out, void* stream) {
    if (!c || !in || !out || !taps_dev || N < 1 || n_in < 1 || up < 1 || down < 1 || half_len < 0)
        return fail("gat_resample: bad argument");
    if (n_out != (n_in * up + down - 1) / down) return fail("gat_resample: n_out must be ceil(n_in*up/down) = %lld", (long long)((n_in * up + down - 1) / down));
    if (N > 65535) return fail("gat_resample: at most 65535 signals per call");
    ResampleParams p{in, (long long)n_in, out, (long long)n_out, taps_dev, (int)half_len, (int)up, (int)down};
    dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)N);
    LAUNCH(c, resample_poly_kernel, grid, 256, 0, stream, p);
    return 0;
}

// ------------------------------------------------------------------------------------------------- segmentation
namespace {

// Per-signal scalars shared by the segmentation kernels, in seg_small: spec max [P] (8 B) | env min, max [P][2] (16 B) |
// n_peaks [P] | any_nonzero [P] | gate value [P].
struct SegScalars { long long* spec_max; long long* env_minmax; int* n_peaks; int* any_nonzero; float* gate_val; };

int seg_scalars(gat_ctx* c, int64_t P, cudaStream_t st, SegScalars* s) {
    if (c->seg_small.ensure((size_t)P * 36 + 64)) return 1;
    unsigned char* small = c->seg_small.as<unsigned char>();
    s->spec_max = reinterpret_cast<long long*>(small);
    s->env_minmax = reinterpret_cast<long long*>(small + (size_t)P * 8);
    s->n_peaks = reinterpret_cast<int*>(small + (size_t)P * 24);
    s->any_nonzero = reinterpret_cast<int*>(small + (size_t)P * 28);
    s->gate_val = reinterpret_cast<float*>(small + (size_t)P * 32);
    GAT_CUDA(cudaMemsetAsync(small, 0x80, (size_t)P * 24, st));
    GAT_CUDA(cudaMemsetAsync(small + (size_t)P * 24, 0, (size_t)P * 12, st));
    return 0;
}

// AudioSlicer.detect_onsets (slicing.py:106-122) on P signals y[P][L]: float64 STFT -> Slaney mel-128 -> dB -> flux
// envelope -> normalise, pick peaks -> backtrack -> frames * hop -> greedy minimum separation.  `gated` applies the
// sample gate and the frame gate (seg_gate) of sliceNsave while the samples are staged.  Also fills the slice table.
// Per-signal outputs: onsets[P][max_onsets], n_onsets[P].
int onset_chain(gat_ctx* c, const float* y, int64_t P, int64_t L, const gat_slicer_params* sp, bool gated, const SegScalars& sc,
                int32_t max_onsets, int64_t* onsets, int32_t* n_onsets, double* env_out, int64_t* frames_out,
                int32_t* n_frames_out, cudaStream_t st) {
    const int To = (int)(1 + L / sp->onset_hop);      // onset frames per signal
    const int words = To / 32 + 2;
    const size_t PT = (size_t)P * To;
    if (c->spec.ensure(PT * 128 * sizeof(double)) || c->seg_env.ensure(PT * 8) || c->seg_envn.ensure(PT * 8) ||
        c->seg_cand.ensure((size_t)P * words * 4) || c->seg_peaks.ensure(PT * 4) || c->seg_frames.ensure(PT * 8) ||
        c->seg_table.ensure((size_t)P * max_onsets * 3 * 8)) return 1;
    {   // float64 STFT -> Slaney mel-128 -> dB, one warp per frame (csrc/stft2.cuh)
        OnsetFramesParams q{};
        q.audio = y; q.n = L; q.N = (int)P;
        q.frame_gate = gated ? c->seg_gate.as<unsigned char>() : nullptr;
        q.gate_stride = 1 + L / sp->rms_hop; q.gate_hop = sp->rms_hop; q.sample_gate = gated ? sp->sample_gate : 0.0f;
        q.gate_shift = -1;
        for (int sh = 0; sh < 30; ++sh) if ((1 << sh) == sp->rms_hop) q.gate_shift = sh;
        q.hop = sp->onset_hop; q.T = To;
        q.window = c->win64.as<double>(); q.tw = c->tw64.as<Cpx<double>>(); q.w2 = c->w2_64.as<Cpx<double>>();
        q.fb = c->fb_mfcc.view(); q.amin = 1e-10; q.out = c->spec.as<double>(); q.spec_max = sc.spec_max;
        constexpr int kOnsetThreads = 256;
        const size_t smem = onset_frames_smem_bytes(kOnsetThreads / 32, q.fb.nnz + 3);
        if (smem > 227 * 1024) return fail("onset_frames: %zu bytes of shared memory needed", smem);
        const long long n_items = (long long)P * To;
        long long ctas = (n_items + kOnsetThreads / 32 - 1) / (kOnsetThreads / 32);
        ctas = ctas < c->num_sms ? ctas : c->num_sms;
        q.items_per_cta = (n_items + ctas - 1) / ctas;
        ctas = (n_items + q.items_per_cta - 1) / q.items_per_cta;
        auto kfn = onset_frames_kernel<kOnsetThreads>;
        GAT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KNAME("stft_mel_f64_spec");
        LAUNCH(c, kfn, (unsigned)ctas, kOnsetThreads, smem, st, q);
    }

    // flux envelope -> normalise + candidate peaks -> sequential wait rule
    const dim3 frames_grid((unsigned)ceil_div(To, 128), (unsigned)P);
    FluxParams fp{c->spec.as<double>(), sc.spec_max, To, 128, 1 + 2048 / (2 * sp->onset_hop), 80.0, c->seg_env.as<double>(), sc.env_minmax};
    LAUNCH(c, onset_flux_kernel, frames_grid, 128, 0, st, fp);
    PeakParams pp{c->seg_env.as<double>(), sc.env_minmax, To, sp->pre_max, sp->post_max, sp->pre_avg, sp->post_avg, sp->wait,
                  (double)sp->delta, c->seg_envn.as<double>(), c->seg_cand.as<unsigned>(), sc.n_peaks, c->seg_peaks.as<int>(), sc.any_nonzero, words};
    LAUNCH(c, peak_candidates_kernel, frames_grid, 128, 0, st, pp);
    LAUNCH(c, peak_select_kernel, (unsigned)P, kPeakThreads, 0, st, pp);
    if (env_out) GAT_CUDA(cudaMemcpyAsync(env_out, c->seg_envn.p, PT * 8, cudaMemcpyDeviceToDevice, st));

    // backtrack, min separation, slice table
    SliceParams s{c->seg_envn.as<double>(), To, sc.n_peaks, c->seg_peaks.as<int>(), sp->onset_hop, (long long)L,
                  (long long)sp->min_sep_samples, (long long)sp->attack_skip, (long long)sp->clip_len, max_onsets,
                  n_onsets, (long long*)onsets, c->seg_frames.as<long long>(), c->seg_table.as<long long>()};
    LAUNCH(c, backtrack_kernel, frames_grid, 128, 0, st, s);
    LAUNCH(c, minsep_table_kernel, (unsigned)P, 1024, 0, st, s);
    if (frames_out && n_frames_out) {      // diagnostics of the single-signal entry point
        const size_t nb = (size_t)(max_onsets < To ? max_onsets : To) * 8;
        GAT_CUDA(cudaMemcpyAsync(frames_out, c->seg_frames.p, nb, cudaMemcpyDeviceToDevice, st));
        GAT_CUDA(cudaMemcpyAsync(n_frames_out, sc.n_peaks, 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// AudioSlicer.sliceNsave minus file I/O for P signals of L samples: gates -> onsets -> slice table -> loudness test ->
// one compacted clip list in (signal, onset) order.  gat_segment is P = 1 with 3-column table rows.
int segment_impl(gat_ctx* c, const float* y, int64_t P, int64_t L, const gat_slicer_params* sp, int32_t max_onsets,
                 int64_t* onsets, int32_t* n_onsets, float* clips, int64_t max_clips, int64_t* clip_table, int table_cols,
                 int32_t* n_clips /* [P + 1] */, float* rms_db_out, double* env_out, int64_t* frames_out, int32_t* n_frames_out,
                 cudaStream_t st) {
    const int T = (int)(1 + L / sp->rms_hop);         // rms frames per signal
    const size_t PT = (size_t)P * T;
    if (c->seg_rms.ensure(PT * 4) || c->seg_rms_med.ensure(PT * 4) || c->seg_gate.ensure(PT) ||
        c->seg_keep.ensure((size_t)P * max_onsets) || c->seg_dest.ensure((size_t)P * max_onsets * 4) ||
        c->seg_base.ensure((size_t)P * 8)) return 1;
    SegScalars sc{};
    if (seg_scalars(c, P, st, &sc)) return 1;

    // 1-3: sample gate (fused into the loads) -> frame RMS dB -> median-5 -> p20 + 6 dB frame gate
    RmsParams rp{y, (long long)L, T, sp->rms_hop, sp->sample_gate, c->seg_rms.as<float>()};
    LAUNCH(c, rms_db_kernel, dim3((unsigned)ceil_div(T, kRmsFramesPerCta), (unsigned)P), 128, 0, st, rp);
    LAUNCH(c, median5_kernel, dim3((unsigned)ceil_div(T, 256), (unsigned)P), 256, 0, st, c->seg_rms.as<float>(), c->seg_rms_med.as<float>(), T);
    GateParams gp{c->seg_rms_med.as<float>(), T, sp->p20_k, sp->p20_gamma, sp->gate_offset_db, c->seg_gate.as<unsigned char>(), sc.gate_val};
    LAUNCH(c, rms_gate_kernel, (unsigned)P, T >= 4096 ? 1024 : 256, 0, st, gp);
    if (rms_db_out) GAT_CUDA(cudaMemcpyAsync(rms_db_out, c->seg_rms_med.p, PT * 4, cudaMemcpyDeviceToDevice, st));

    // 4-9: onsets of the doubly gated signals
    if (onset_chain(c, y, P, L, sp, true, sc, max_onsets, onsets, n_onsets, env_out, frames_out, n_frames_out, st)) return 1;

    // 10-11: loudness test, compaction, gather
    GatherParams g{y, (long long)L, n_onsets, c->seg_table.as<long long>(), (long long)sp->clip_len, sp->min_slice_rms_db,
                   c->seg_keep.as<unsigned char>(), c->seg_dest.as<int>(), n_clips, c->seg_base.as<long long>(), clips,
                   (long long*)clip_table, max_onsets, table_cols, (long long)max_clips, (int)P};
    const dim3 onsets_grid((unsigned)max_onsets, (unsigned)P);
    LAUNCH(c, slice_loudness_kernel, onsets_grid, 256, 0, st, g);
    LAUNCH(c, slice_compact_kernel, (unsigned)P, max_onsets > 256 ? 1024 : 256, 0, st, g);
    LAUNCH(c, slice_base_kernel, 1, 1024, 0, st, g);
    LAUNCH(c, slice_gather_kernel, onsets_grid, 256, 0, st, g);
    return 0;
}

}  // namespace

extern "C" int gat_detect_onsets(gat_ctx* c, const float* y, int64_t L, const gat_slicer_params* sp, int32_t max_onsets,
                                 int64_t* onsets, int32_t* n_onsets, void* stream) {
    if (!c || !y || !sp || !onsets || !n_onsets) return fail("gat_detect_onsets: null argument");
    if (L < 1) return fail("gat_detect_onsets: empty signal");
    if (max_onsets < 1) return fail("gat_detect_onsets: max_onsets must be positive");
    if (sp->onset_hop < 2 || (sp->onset_hop & 1) || sp->onset_hop > 2048) return fail("gat_detect_onsets: hop %d unsupported (even, 2..2048)", sp->onset_hop);
    cudaStream_t st = (cudaStream_t)stream;
    SegScalars sc{};
    if (seg_scalars(c, 1, st, &sc)) return 1;
    return onset_chain(c, y, 1, L, sp, false, sc, max_onsets, onsets, n_onsets, nullptr, nullptr, nullptr, st);
}

extern "C" int gat_segment(gat_ctx* c, const float* y, int64_t L, const gat_slicer_params* sp, int32_t max_onsets,
                           int64_t* onsets, int32_t* n_onsets, float* clips, int64_t* clip_table, int32_t* n_clips,
                           float* rms_db_out, double* env_out, int64_t* frames_out, int32_t* n_frames_out, void* stream) {
    if (!c || !y || !sp || !onsets || !n_onsets || !clips || !clip_table || !n_clips) return fail("gat_segment: null argument");
    if (L <= 1024) return fail("gat_segment: signal of %lld samples is too short (reflect padding needs > 1024)", (long long)L);
    if (max_onsets < 1) return fail("gat_segment: max_onsets must be positive");
    if (sp->rms_hop < 1 || sp->onset_hop != 512) return fail("gat_segment: onset hop %d unsupported (the reference always uses 512)", sp->onset_hop);
    cudaStream_t st = (cudaStream_t)stream;
    if (c->seg_counts.ensure(8)) return 1;     // [kept clips of the signal, total]: the caller's n_clips is one int
    if (segment_impl(c, y, 1, L, sp, max_onsets, onsets, n_onsets, clips, max_onsets, clip_table, 3, c->seg_counts.as<int32_t>(),
                     rms_db_out, env_out, frames_out, n_frames_out, st)) return 1;
    GAT_CUDA(cudaMemcpyAsync(n_clips, c->seg_counts.p, 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int gat_segment_batch(gat_ctx* c, const float* y, int64_t P, int64_t L, const gat_slicer_params* sp, int32_t max_onsets,
                                 int64_t* onsets, int32_t* n_onsets, float* clips, int64_t max_clips, int64_t* clip_table,
                                 int32_t* n_clips, void* stream) {
    if (!c || !y || !sp || !onsets || !n_onsets || !clips || !clip_table || !n_clips) return fail("gat_segment_batch: null argument");
    if (P < 1) return 0;
    if (P > 65535) return fail("gat_segment_batch: at most 65535 signals per call (got %lld)", (long long)P);
    if (L <= 1024) return fail("gat_segment_batch: signals of %lld samples are too short (reflect padding needs > 1024)", (long long)L);
    if (max_onsets < 1 || max_onsets > 65535 || max_clips < 1) return fail("gat_segment_batch: max_onsets / max_clips out of range");
    if (sp->rms_hop < 1 || sp->onset_hop != 512) return fail("gat_segment_batch: onset hop %d unsupported (the reference always uses 512)", sp->onset_hop);
    return segment_impl(c, y, P, L, sp, max_onsets, onsets, n_onsets, clips, max_clips, clip_table, 4, n_clips,
                        nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------- end to end
extern "C" int gat_transcribe_clips(gat_ctx* c, const float* audio, int64_t N, int64_t n, int32_t flags, float* probs,
                                    float* mlp_probs, float* cnn_probs, int64_t* index, float* conf, float* mfcc, float* mel,
                                    double* yin_hz, void* stream) {
    if (!c || !audio || !probs || !index || !conf) return fail("gat_transcribe_clips: null argument");
    if (N <= 0) return 0;
    const bool skip_mlp = (flags & GAT_FLAG_SKIP_MLP) != 0;
    const bool add_pitch = (flags & GAT_FLAG_NO_PITCH) == 0;                 // MFCCConfig.ADD_PITCH_FEATURES
    const bool norm_mfcc = (flags & GAT_FLAG_NO_NORMALIZE_MFCC) == 0;        // MFCCConfig.NORMALIZE_AUDIO_VOLUME
    const bool norm_mel = (flags & GAT_FLAG_NO_NORMALIZE_MEL) == 0;          // MelSpecConfig.NORMALIZE_AUDIO_VOLUME
    const int T = (int)(1 + n / c->cfg.mel_hop);
    const int classes = c->classes;
    const int F = c->cfg.mfcc_n_mfcc + (add_pitch ? 1 : 0);
    if (!c->cnn_loaded) return fail("gat_transcribe_clips: no CNN loaded");
    if (!skip_mlp && (!mlp_probs || !cnn_probs)) return fail("gat_transcribe_clips: mlp_probs/cnn_probs required unless GAT_FLAG_SKIP_MLP");
    if (!skip_mlp) {
        // the MLP reads mlp_dims[0] floats per row of the [N][F] feature matrix: a mismatch would make rows bleed into each other
        if (!c->mlp_n_linear) return fail("gat_transcribe_clips: no MLP loaded (gat_load_mlp)");
        if (c->mlp_dims[0] != F)
            return fail("gat_transcribe_clips: the MLP takes %d inputs but the feature rows have %d columns (n_mfcc %d%s)",
                        c->mlp_dims[0], F, c->cfg.mfcc_n_mfcc, add_pitch ? " + pitch" : "");
        if ((flags & GAT_FLAG_APPLY_SCALER) && c->scaler_n != F)
            return fail("gat_transcribe_clips: scaler has %d columns, features have %d", c->scaler_n, F);
    }
    if (!mel) { if (c->e2e_mel.ensure((size_t)N * c->cfg.mel_n_mels * T * 4)) return 1; mel = c->e2e_mel.as<float>(); }
    if (!skip_mlp && !mfcc) { if (c->e2e_mfcc.ensure((size_t)N * F * 4)) return 1; mfcc = c->e2e_mfcc.as<float>(); }
    if (c->logits_cnn.ensure((size_t)N * classes * 4) || c->logits_mlp.ensure((size_t)N * classes * 4)) return 1;
    float* cnn_logits = c->logits_cnn.as<float>();
    float* mlp_logits = c->logits_mlp.as<float>();
    // one RMS pass serves all three chains (the reference recomputes it per chain: features.py:185,311,460,497)
    const bool any_norm = norm_mel || (!skip_mlp && norm_mfcc);
    if (any_norm && launch_clip_scale(c, audio, N, n, stream)) return 1;
    const bool fused = !skip_mlp && can_fuse_chains(c);
    if (fused) { if (run_dual(c, audio, N, n, norm_mel, norm_mfcc, mel, mfcc, F, stream)) return 1; }
    else if (run_melspec(c, audio, N, n, norm_mel, true, mel, stream)) return 1;
    if (skip_mlp) {
        if (run_cnn(c, mel, N, T, probs, cnn_logits, stream)) return 1;
        if (cnn_probs && cnn_probs != probs)
            GAT_CUDA(cudaMemcpyAsync(cnn_probs, probs, (size_t)N * classes * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        LAUNCH(c, argmax_kernel, (unsigned)((N + 7) / 8), 256, 0, stream, probs, (int)N, classes, (long long*)index, conf);
        return 0;
    }
    if (!fused && run_mfcc(c, audio, N, n, norm_mfcc, true, mfcc, F, stream)) return 1;
    if (add_pitch || yin_hz) {
        // features.py:473 hands YIN the (normalised, when the MFCC chain normalises) clip; :201 the raw one
        const bool yn = norm_mfcc && (flags & GAT_FLAG_YIN_ON_NORMALIZED) != 0;
        if (run_yin(c, audio, N, n, yn, true, yin_hz, nullptr, add_pitch ? mfcc : nullptr, F, c->cfg.mfcc_n_mfcc, stream)) return 1;
    }
    if (flags & GAT_FLAG_APPLY_SCALER) {
        const long long tot = (long long)N * F;
        LAUNCH(c, standard_scale_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, mfcc, (int)N, F, F,
               c->scaler_mean.as<double>(), c->scaler_scale.as<double>());
    }
    if (run_cnn(c, mel, N, T, cnn_probs, cnn_logits, stream)) return 1;
    return run_mlp_ensemble(c, mfcc, F, N, cnn_probs, probs, mlp_probs, mlp_logits, index, conf, stream);
}

namespace {
// Host clips -> labels.  sample_bytes = 4: float32 clips; 2: PCM_16 clips (x / 32768 on the device, libsndfile's
// scaling), which halves the bytes that cross the host link - the link is what bounds this call.
int transcribe_host(gat_ctx* c, const void* audio_host_v, int sample_bytes, int64_t N, int64_t n, int32_t flags,
                    int64_t* index_host, float* conf_host, float* probs_host) {
    const unsigned char* audio_host = static_cast<const unsigned char*>(audio_host_v);
    if (!c->e2e_streams) {
        GAT_CUDA(cudaStreamCreateWithFlags(&c->e2e_stream[0], cudaStreamNonBlocking));
        GAT_CUDA(cudaStreamCreateWithFlags(&c->e2e_stream[1], cudaStreamNonBlocking));
        c->e2e_streams = true;
    }
    const int classes = c->classes;
    // Chunk schedule.  Chunk k's kernels start when its copy has landed, and the copy of chunk k+1 runs beside them.
    //   float32 clips: the link is the bound (361 MB at the 55.6 GB/s profiles/r02_h2d_probe_*.json measure = 6.50 ms against
    //   ~3 ms of kernels), so the call ends one chunk's kernels after the LAST byte lands: sizes HALVE towards the end
    //   (..., 8, 4, 2, 1 x SMs) - each chunk's kernels finish under the next, shorter copy and the tail is the smallest chunk.
    //   PCM_16 clips: half the bytes, the kernels are the bound, the call ends sum(kernels) after the FIRST chunk lands:
    //   sizes DOUBLE from the start (1, 2, 4, 8 x SMs, ...) so the kernels start early and then run on efficient batches.
    // Round 1 used equal chunks of 4 x SMs (7.20 / 5.17 ms); long chunks at the end of a copy-bound call cost their whole
    // kernel time as tail (measured: 7.45 ms with 8 x SMs in the middle), hence the geometric schedules: 7.11 / 4.53 ms at
    // 2..8 x SMs.  What is left above the copy itself (6.60 ms measured inside this call) is the last chunk's kernel chain
    // (nine launches, ~0.25 ms however few clips) plus the result copies; the PCM_16 call pays ~1.4 ms of small-batch
    // inefficiency over the 3.05 ms the same kernels take on one resident 4096-clip batch.
    std::vector<int64_t> sizes;
    {
        const int64_t sm = c->num_sms;
        const int64_t lo = c->host_chunk_mult[0] * sm, cap = c->host_chunk_mult[1] * sm;
        int order = c->host_chunk_mult[2];
        if (order == 0) order = sample_bytes == 2 ? 1 : 2;                 // 1: increasing, 2: decreasing
        int64_t left = N, sz = lo;
        while (left > 0) {
            const int64_t take = left < sz + lo / 2 ? left : sz;            // fold a small remainder into the last chunk
            sizes.push_back(take);
            left -= take;
            sz = sz * 2 > cap ? cap : sz * 2;
        }
        if (order == 2) std::reverse(sizes.begin(), sizes.end());
    }
    int64_t chunk = 0;
    for (int64_t v : sizes) chunk = v > chunk ? v : chunk;
    if (sample_bytes == 2 && (c->e2e_pcm[0].ensure((size_t)chunk * n * 2) || c->e2e_pcm[1].ensure((size_t)chunk * n * 2))) return 1;
    if (c->e2e_audio[0].ensure((size_t)chunk * n * 4) || c->e2e_audio[1].ensure((size_t)chunk * n * 4) ||
        c->e2e_probs.ensure((size_t)N * classes * 4) || c->e2e_mlp_probs.ensure((size_t)N * classes * 4) ||
        c->e2e_cnn_probs.ensure((size_t)N * classes * 4) || c->e2e_index.ensure((size_t)N * 8) || c->e2e_conf.ensure((size_t)N * 4)) return 1;
    // The scratch buffers inside the ctx are shared, so the kernels of consecutive chunks are serialised on
    // stream 0; stream 1 only carries the host-to-device copy of the NEXT chunk, which is what overlaps.
    cudaEvent_t copied[2], consumed[2];
    for (int i = 0; i < 2; ++i) {
        GAT_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        GAT_CUDA(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
    }
    int rc = 0;
    int64_t c0 = 0;
    for (size_t k = 0; k < sizes.size() && !rc; c0 += sizes[k], ++k) {
        const int b = (int)(k & 1);
        const int64_t nc = sizes[k];
        if (k >= 2) GAT_CUDA(cudaStreamWaitEvent(c->e2e_stream[1], consumed[b], 0));
        void* landing = sample_bytes == 2 ? c->e2e_pcm[b].p : c->e2e_audio[b].p;
        cudaEvent_t pc0 = nullptr, pc1 = nullptr;
        if (c->profiling) {          // the copies appear in gat_profile_end's table as "h2d_copy"
            GAT_CUDA(cudaEventCreate(&pc0)); GAT_CUDA(cudaEventCreate(&pc1));
            GAT_CUDA(cudaEventRecord(pc0, c->e2e_stream[1]));
        }
        GAT_CUDA(cudaMemcpyAsync(landing, audio_host + (size_t)c0 * n * sample_bytes, (size_t)nc * n * sample_bytes, cudaMemcpyHostToDevice, c->e2e_stream[1]));
        if (c->profiling) { GAT_CUDA(cudaEventRecord(pc1, c->e2e_stream[1])); c->prof.push_back(ProfRec{"h2d_copy", pc0, pc1}); }
        GAT_CUDA(cudaEventRecord(copied[b], c->e2e_stream[1]));
        GAT_CUDA(cudaStreamWaitEvent(c->e2e_stream[0], copied[b], 0));
        if (sample_bytes == 2) {
            const long long count = (long long)nc * n;
            const long long blocks = (count + 255) / 256;
            LAUNCH(c, pcm16_to_mono_kernel, (unsigned)(blocks < 16LL * c->num_sms ? blocks : 16LL * c->num_sms), 256, 0, c->e2e_stream[0],
                   c->e2e_pcm[b].as<short>(), count, 1, c->e2e_audio[b].as<float>());
        }
        rc = gat_transcribe_clips(c, c->e2e_audio[b].as<float>(), nc, n, flags, c->e2e_probs.as<float>() + c0 * classes,
                                  c->e2e_mlp_probs.as<float>() + c0 * classes, c->e2e_cnn_probs.as<float>() + c0 * classes,
                                  c->e2e_index.as<int64_t>() + c0, c->e2e_conf.as<float>() + c0, nullptr, nullptr, nullptr,
                                  c->e2e_stream[0]);
        GAT_CUDA(cudaEventRecord(consumed[b], c->e2e_stream[0]));
    }
    if (!rc) {
        GAT_CUDA(cudaMemcpyAsync(index_host, c->e2e_index.p, (size_t)N * 8, cudaMemcpyDeviceToHost, c->e2e_stream[0]));
        GAT_CUDA(cudaMemcpyAsync(conf_host, c->e2e_conf.p, (size_t)N * 4, cudaMemcpyDeviceToHost, c->e2e_stream[0]));
        if (probs_host) GAT_CUDA(cudaMemcpyAsync(probs_host, c->e2e_probs.p, (size_t)N * classes * 4, cudaMemcpyDeviceToHost, c->e2e_stream[0]));
    }
    GAT_CUDA(cudaStreamSynchronize(c->e2e_stream[1]));
    GAT_CUDA(cudaStreamSynchronize(c->e2e_stream[0]));
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(copied[i]); cudaEventDestroy(consumed[i]); }
    return rc;
}
}  // namespace

extern "C" int gat_transcribe_clips_host(gat_ctx* c, const float* audio_host, int64_t N, int64_t n, int32_t flags,
                                         int64_t* index_host, float* conf_host, float* probs_host) {
    if (!c || !audio_host || !index_host || !conf_host) return fail("gat_transcribe_clips_host: null argument");
    if (N <= 0) return 0;
    return transcribe_host(c, audio_host, 4, N, n, flags, index_host, conf_host, probs_host);
}

extern "C" int gat_transcribe_clips_host_pcm16(gat_ctx* c, const int16_t* audio_host, int64_t N, int64_t n, int32_t flags,
                                               int64_t* index_host, float* conf_host, float* probs_host) {
    if (!c || !audio_host || !index_host || !conf_host) return fail("gat_transcribe_clips_host_pcm16: null argument");
    if (N <= 0) return 0;
    return transcribe_host(c, audio_host, 2, N, n, flags, index_host, conf_host, probs_host);
}
