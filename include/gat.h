/* gat.h - C ABI of the B200-native guitar-audio-transcriber hot path (libgat.so).
 *
 * The reference (gkotti4/guitar-audio-transcriber-ai, version_1) is pure Python: it has no FFI, its
 * boundary is a handful of Python classes.  Each entry point below replaces the arithmetic behind one of
 * those call sites (paths relative to /root/reference/version_1/source); INTEGRATION.md shows the ctypes
 * stub a maintainer would add at each site.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; gat_last_error() returns a thread-local
 *     message.  No C++ exception crosses the boundary.
 *   - pointers named *_dev are CUDA device pointers owned by the caller; *_host are host pointers.
 *     The library never allocates caller-visible memory.  `stream` is a cudaStream_t passed as void*;
 *     *_dev entry points only enqueue work on it and return.
 *   - a gat_ctx owns the constant tables, packed weights and scratch buffers of ONE device; it is not
 *     thread-safe; use one ctx per (device, stream).
 */
#ifndef GAT_H_
#define GAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gat_ctx gat_ctx;

/* Constant tables are passed in (host pointers, copied) rather than recomputed, so that they are
 * bit-identical to what the reference's libraries build:
 *   mel_window  torch.hann_window(n_fft)                          (torchaudio MelSpectrogram)
 *   mel_fb      torchaudio.functional.melscale_fbanks(...)        [n_fft/2+1][mel_n_mels]
 *   stft_window scipy.signal.get_window("hann", 2048, fftbins=1)  (librosa.stft, float64)
 *   mfcc_fb     librosa.filters.mel(sr, 2048, n_mels=128)         [mfcc_n_mels][1025]
 *   dct         scipy.fft.dct(eye(mfcc_n_mels), type 2, ortho)[:n_mfcc]  [mfcc_n_mfcc][mfcc_n_mels]
 */
typedef struct gat_config {
    int32_t sample_rate;        /* checkpoint target_sr (config.py:29)                               */
    int32_t mel_n_fft;          /* MelSpecConfig.N_FFT: 512, 1024, 2048 (register-resident FFT) or 4096 */
    int32_t mel_hop;            /* MelSpecConfig.HOP_LENGTH                                           */
    int32_t mel_n_mels;         /* MelSpecConfig.N_MELS (<= 128)                                      */
    const float* mel_window;
    const float* mel_fb;
    int32_t mfcc_n_mels;        /* librosa default 128                                                */
    int32_t mfcc_n_mfcc;        /* MFCCConfig.N_MFCC                                                  */
    const double* stft_window;
    const float* mfcc_fb;
    const float* dct;
    double yin_fmin;            /* YinDsp defaults 50 / 1000 (dsp/yin.py:12)                          */
    double yin_fmax;
    double yin_trough_threshold;/* librosa.yin default 0.1                                            */
} gat_config;

const char* gat_last_error(void);
int gat_version(void);

/* Transcriber.__init__ / NotePredictor.load_models (transcribe.py:26-75, note_predictor.py:29-80). */
int gat_ctx_create(const gat_config* cfg, int device, gat_ctx** out);
void gat_ctx_destroy(gat_ctx* ctx);

/* MLP (training/mlp_trainer.py:32-105).  dims[0..n_linear] = layer widths; params packed per Linear as
 * W^T[in][out], b[out] and, for every Linear but the last, LayerNorm gamma[out], beta[out]. */
int gat_load_mlp(gat_ctx* ctx, const int32_t* dims, int32_t n_linear, const float* params_host, int64_t n_params);

/* CNN (training/cnn_trainer.py:30-139), eval-mode BatchNorm already folded into the convs.
 * conv_w[i]: [9][c_in][c_out]; fc1_w: [flat][hidden] with flat index c*16 + i*4 + j; fc2_w: [hidden][classes]. */
int gat_load_cnn(gat_ctx* ctx, int32_t n_conv, const int32_t* channels /* n_conv+1 */,
                 const float* const* conv_w_host, const float* const* conv_b_host,
                 int32_t hidden, int32_t classes,
                 const float* fc1_w_host, const float* fc1_b_host,
                 const float* fc2_w_host, const float* fc2_b_host);

/* sklearn StandardScaler of the MLP checkpoint (features.py:145-146); n = 0 clears it. */
int gat_set_scaler(gat_ctx* ctx, const double* mean_host, const double* scale_host, int32_t n);

/* NotePredictor.cnn_weight / mlp_weight (note_predictor.py:25-26); defaults 0.8f / 0.2f. */
int gat_set_ensemble_weights(gat_ctx* ctx, float mlp_weight, float cnn_weight);

/* ---- features ---------------------------------------------------------------------------------- */

/* MelFeatureBuilder.extract_melspec_features per clip (features.py:296-316, :486-502):
 * out_dev[N][mel_n_mels][T], T = 1 + n / mel_hop.  `normalize` is a bit set: GAT_MEL_NORMALIZE =
 * NORMALIZE_AUDIO_VOLUME (features.py:311,:497); GAT_MEL_POWER = to_db / melspec_to_db False (features.py:313-316,
 * :499-502): the mel POWER is written instead of 10 log10. */
#define GAT_MEL_NORMALIZE 1
#define GAT_MEL_POWER     2
int gat_melspec_db(gat_ctx* ctx, const float* audio_dev, int64_t N, int64_t n, int32_t normalize,
                   float* out_dev, void* stream);

/* MelFeatureBuilder.extract_mfcc_features per clip (features.py:182-208, :458-478):
 * out_dev[N][ld] <- MFCC time-mean (n_mfcc columns) [+ log10(YIN Hz) in column n_mfcc if add_pitch].
 * yin_on_normalized: 1 = memory path (features.py:473), 0 = file path (:201).
 * apply_scaler: 1 = StandardScaler.transform afterwards (file path only, features.py:145-146).
 * yin_hz_dev: optional [N] float64 median pitch (dsp/yin.py:67), may be NULL. */
int gat_mfcc_features(gat_ctx* ctx, const float* audio_dev, int64_t N, int64_t n, int32_t normalize,
                      int32_t add_pitch, int32_t yin_on_normalized, int32_t apply_scaler,
                      float* out_dev, int32_t ld, double* yin_hz_dev, void* stream);

/* YinDsp.estimate_pitch (dsp/yin.py:39-75): hz_dev[N] float64 median f0; f0_frames_dev[N][1+n/512] optional. */
int gat_yin(gat_ctx* ctx, const float* audio_dev, int64_t N, int64_t n, int32_t normalize,
            double* hz_dev, double* f0_frames_dev, void* stream);

/* ---- inference --------------------------------------------------------------------------------- */

/* NotePredictor.predict (note_predictor.py:84-135).  mfcc_dev[N][ld] float32, mel_dev[N][mel_n_mels][T].
 * Outputs (device): probs/mlp_probs/cnn_probs [N][classes], index int64[N], conf float32[N];
 * mlp_logits_dev / cnn_logits_dev optional (may be NULL). */
int gat_infer(gat_ctx* ctx, const float* mfcc_dev, int32_t ld, const float* mel_dev, int64_t N, int32_t T,
              float* probs_dev, float* mlp_probs_dev, float* cnn_probs_dev, int64_t* index_dev, float* conf_dev,
              float* mlp_logits_dev, float* cnn_logits_dev, void* stream);

/* ---- segmentation ------------------------------------------------------------------------------ */

typedef struct gat_slicer_params {   /* AudioSlicerConfig (config.py:100-107) + sliceNsave arguments */
    double min_db_threshold;    /* -32.5  apply_db_threshold                  slicing.py:30-39     */
    float sample_gate;          /* smallest float32 |y| kept by that gate, computed by the caller      */
    int32_t rms_hop;            /* 512    apply_rms_threshold hop             slicing.py:78         */
    int32_t p20_k;              /* floor((T-1)*0.2f) in float32 (np.percentile, linear)                */
    float p20_gamma;            /* fractional part, float32                                             */
    float gate_offset_db;       /* 6.0                                         slicing.py:63         */
    int32_t onset_hop;          /* 512 (detect_onsets is called without hop)   slicing.py:151        */
    int32_t pre_max, post_max, pre_avg, post_avg, wait;  /* librosa onset_detect defaults, ceil-ed     */
    float delta;                /* 0.07f                                                                */
    int64_t min_sep_samples;    /* int(MIN_SEP * sr)                           slicing.py:114        */
    int64_t attack_skip;        /* int(ATTACK_SKIP_SEC * sr)                   slicing.py:127        */
    int64_t clip_len;           /* int(length_sec * sr)                        slicing.py:126        */
    float min_slice_rms_db;     /* -37.0                                        slicing.py:96-100     */
} gat_slicer_params;

/* AudioSlicer.sliceNsave minus file I/O (slicing.py:147-165): y_dev[L] float32 mono at the target rate.
 * Outputs (device): onsets_dev int64[max_onsets], n_onsets_dev int32[1],
 *                   clips_dev float32[max_onsets][clip_len] (kept clips, compacted, zero padded),
 *                   clip_table_dev int64[max_onsets][3] = (onset index, start, end), n_clips_dev int32[1].
 * Optional diagnostics (NULL to skip): rms_db_dev float32[T] (median-filtered), env_dev float64[T]
 * (normalised onset envelope), frames_dev int64[max_onsets] (backtracked peak frames), n_frames_dev int32[1]. */
int gat_segment(gat_ctx* ctx, const float* y_dev, int64_t L, const gat_slicer_params* sp, int32_t max_onsets,
                int64_t* onsets_dev, int32_t* n_onsets_dev, float* clips_dev, int64_t* clip_table_dev,
                int32_t* n_clips_dev, float* rms_db_dev, double* env_dev, int64_t* frames_dev,
                int32_t* n_frames_dev, void* stream);

/* gat_segment for P independent signals of L samples each (y_dev[P][L]) in ONE pass of batched kernels: the phrases /
 * files a rank owns when a long recording is sharded across GPUs (SURVEY.md 8(e) option (i): every signal is sliced
 * exactly as AudioSlicer.sliceNsave would slice it as a file of its own - its own dB maximum, percentile gate,
 * envelope normalisation and min-sep scan, slicing.py:147-165).
 * Outputs (device): onsets_dev int64[P][max_onsets], n_onsets_dev int32[P];
 *                   clips_dev float32[max_clips][clip_len]: the kept clips of ALL signals, compacted in (signal, onset)
 *                   order; clip_table_dev int64[max_clips][4] = (signal, onset index, start, end);
 *                   n_clips_dev int32[P + 1]: kept clips per signal, then their total.  Clips past max_clips are counted
 *                   but not written (the caller compares n_clips_dev[P] with max_clips). */
int gat_segment_batch(gat_ctx* ctx, const float* y_dev, int64_t P, int64_t L, const gat_slicer_params* sp,
                      int32_t max_onsets, int64_t* onsets_dev, int32_t* n_onsets_dev, float* clips_dev, int64_t max_clips,
                      int64_t* clip_table_dev, int32_t* n_clips_dev, void* stream);

/* AudioSlicer.detect_onsets(y, sr, hop_len, min_sep) on its own (slicing.py:106-122; the live prototype calls it
 * with hop 1024 on the un-gated microphone buffer, prototyping/source/transcribe_live.py:94-96): no gates, any even
 * hop.  Uses sp->onset_hop, the peak-pick fields and min_sep_samples; the signal is processed in float64 (a float32
 * signal handed to librosa would be processed in float32 - same onsets unless an envelope value ties a threshold).
 * Outputs (device): onsets_dev int64[max_onsets], n_onsets_dev int32[1]. */
int gat_detect_onsets(gat_ctx* ctx, const float* y_dev, int64_t L, const gat_slicer_params* sp, int32_t max_onsets,
                      int64_t* onsets_dev, int32_t* n_onsets_dev, void* stream);

/* ---- file front end (Transcriber.transcribe) ---------------------------------------- */

#define GAT_SAMPLE_PCM16   0
#define GAT_SAMPLE_FLOAT32 1

/* sf.write(.wav) + librosa.load of a clip (audio/slicing.py:144 then audio/loading.py:85): every sample goes
 * through PCM_16 once.  python-soundfile enables SFC_SET_CLIPPING, so libsndfile's float -> PCM_16 conversion is
 * q = lrintf(x * 2^31) >> 16 (saturating at +-full scale); the read is q / 32768.  In place on `count` device floats. */
int gat_pcm16_roundtrip(gat_ctx* ctx, float* audio_dev, int64_t count, void* stream);

/* librosa.load's decode + to_mono (audio/slicing.py:25): interleaved frames of `channels` samples
 * (PCM_16 scaled by 1/32768 as libsndfile does, or float32) -> float32 channel mean, `frames` outputs. */
int gat_decode_mono(gat_ctx* ctx, const void* frames_dev, int32_t sample_format, int64_t frames, int32_t channels,
                    float* out_dev, void* stream);

/* librosa.load(sr=...) / librosa.resample (audio/loading.py:85, transcribe.py:173) for N signals of n_in
 * samples: polyphase FIR by up/down with the caller's taps (float64, length 2*half_len+1, already scaled by
 * `up`), scipy.signal.resample_poly alignment, n_out = ceil(n_in*up/down).  The reference resamples with
 * soxr_hq, which is not reproducible here: same length convention, equivalent quality, not bit-identical. */
int gat_resample(gat_ctx* ctx, const float* in_dev, int64_t N, int64_t n_in, int32_t up, int32_t down,
                 const double* taps_dev, int32_t half_len, float* out_dev, int64_t n_out, void* stream);

/* ---- end to end --------------------------------------------------------------------------------- */

#define GAT_FLAG_YIN_ON_NORMALIZED 1   /* transcribe_note path (features.py:473)                    */
#define GAT_FLAG_APPLY_SCALER      2   /* transcribe(file) path (features.py:145-146)               */
#define GAT_FLAG_SKIP_MLP          4   /* BASELINE config 2: mel-spectrogram + CNN only              */
#define GAT_FLAG_NO_PITCH          8   /* MFCCConfig.ADD_PITCH_FEATURES = False (features.py:199,:471): rows have n_mfcc columns */
#define GAT_FLAG_NO_NORMALIZE_MFCC 16  /* MFCCConfig.NORMALIZE_AUDIO_VOLUME = False (features.py:184,:459) */
#define GAT_FLAG_NO_NORMALIZE_MEL  32  /* MelSpecConfig.NORMALIZE_AUDIO_VOLUME = False (features.py:310,:496) */

/* Transcriber.transcribe_note batched over N equal-length clips already on the device
 * (transcribe.py:147-199 steps 1-2).  With GAT_FLAG_SKIP_MLP probs == cnn_probs. Optional outputs may be NULL.
 * Fails (no kernel launched) when the loaded MLP's input width differs from the feature row width
 * n_mfcc + (pitch ? 1 : 0), or when GAT_FLAG_APPLY_SCALER is set and the scaler's width differs. */
int gat_transcribe_clips(gat_ctx* ctx, const float* audio_dev, int64_t N, int64_t n, int32_t flags,
                         float* probs_dev, float* mlp_probs_dev, float* cnn_probs_dev, int64_t* index_dev,
                         float* conf_dev, float* mfcc_dev /* [N][n_mfcc (+1)] */, float* mel_dev, double* yin_hz_dev,
                         void* stream);

/* Same, HOST buffers in and out: copies the clips to the device in chunks (double-buffered against the
 * kernels), runs gat_transcribe_clips per chunk and copies index / conf / probs back.  Blocks until done.
 * audio_host should be pinned for the copies to overlap.  probs_host may be NULL. */
int gat_transcribe_clips_host(gat_ctx* ctx, const float* audio_host, int64_t N, int64_t n, int32_t flags,
                              int64_t* index_host, float* conf_host, float* probs_host);

/* Same with PCM_16 host clips (what .wav files and audio interfaces deliver): samples are scaled by 1/32768 on the
 * device exactly as libsndfile / librosa.load would on the host (audio/loading.py:85), and half as many bytes cross
 * the host link, which is the bound of the float32 variant. */
int gat_transcribe_clips_host_pcm16(gat_ctx* ctx, const int16_t* audio_host, int64_t N, int64_t n, int32_t flags,
                                    int64_t* index_host, float* conf_host, float* probs_host);

/* Per-kernel timing with CUDA events on the launching stream (bench.py's roofline figures).
 * gat_profile_end writes one line per kernel: "<kernel> <launches> <total ms>\n". */
int gat_profile_begin(gat_ctx* ctx);
int gat_profile_end(gat_ctx* ctx, char* buf, int64_t cap);

/* Clips per CNN pass = mult * (number of SMs); default 28 (4144 clips, ~1.3 GB of activation planes at T = 87). */
int gat_set_conv_pass(gat_ctx* ctx, int32_t mult);

/* Chunk schedule of the host entry points: chunk sizes grow geometrically from min_mult x SMs to max_mult x SMs.
 * order 1 = increasing (kernel-bound calls: start early, then efficient batches), 2 = decreasing (copy-bound calls: the tail
 * after the last byte is the smallest chunk), 0 = pick by sample format (PCM_16: 1, float32: 2).  Default 2 / 8 / 0. */
int gat_set_host_chunks(gat_ctx* ctx, int32_t min_mult, int32_t max_mult, int32_t order);

/* Diagnostics for the tensor-core conv pipeline: call with out_host = NULL to switch the in-kernel cycle
 * counters on; call again with a buffer of 2*148*8 int64 to read them (conv2 then conv3; per CTA:
 * MMA-thread total, wait acc_empty, wait a_full, wait w_full, epilogue total, epilogue wait acc_full). */
int gat_debug_tc_counters(gat_ctx* ctx, long long* out_host, int64_t n);

/* Measures this device's FP32-FMA peak (TFLOP/s, 2 flops per FMA) with a register-only FMA kernel: the
 * denominator for the FFT-bound stages ("the builder must measure it").  Blocks. */
int gat_debug_fma_peak(gat_ctx* ctx, int32_t iters, float* tflops_host);

/* Number of kernels launched through this ctx so far (bench.py's gpu_launches). */
int64_t gat_launch_count(const gat_ctx* ctx);
/* classes of the loaded models, frames of the mel image for n samples. */
int32_t gat_num_classes(const gat_ctx* ctx);
int32_t gat_mel_frames(const gat_ctx* ctx, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* GAT_H_ */
