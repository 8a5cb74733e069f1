"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's version_1 transcription hot path.

This is the parity ORACLE (checker), never the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  It restates the control flow of the
reference's own files (cited per function, paths relative to /root/reference/version_1/source) on top of

  * oracle/librosa_shim  - numpy/scipy restatement of the librosa calls (librosa is absent here;
                           inferred version 0.10.2.post1 / 0.11.0; PARITY UNPINNED, see its docstring)
  * torchaudio / torch / scipy / sklearn - the genuine libraries, exactly as the reference calls them.

Pinning: the reference has no tests and no golden vectors.  oracle/make_golden.py runs the
reference's unmodified files (oracle/ref_env.py) and this port on the same inputs, requires identical
results, and writes tests/golden/*.npz; tests/test_oracle_port.py re-checks the port against those
vectors wherever it runs (the GPU box has no /root/reference).
"""
from __future__ import annotations

import math
import pathlib
import sys

import numpy as np
import torch
import torch.nn.functional as F
import torchaudio as ta
from scipy.ndimage import median_filter

_HERE = pathlib.Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))
import librosa_shim as librosa  # noqa: E402

# config.py:29-30, :36-53, :100-107 - the defaults the reference falls back to
TARGET_SR = 22050
CLIP_DURATION = 0.50
MFCC_DEFAULTS = dict(N_MFCC=64, BATCH_SIZE=32, STANDARD_SCALER=True, NORMALIZE_AUDIO_VOLUME=True,
                     ADD_PITCH_FEATURES=True)
MELSPEC_DEFAULTS = dict(N_MELS=64, N_FFT=2048, HOP_LENGTH=256, BATCH_SIZE=32, NORMALIZE_AUDIO_VOLUME=True,
                        TO_DB=True)
SLICER_DEFAULTS = dict(MIN_IN_DB_THRESHOLD=-32.5, MIN_SLICE_RMS_DB=-37.0, HOP_LEN=512, MIN_SEP=0.3,
                       ATTACK_SKIP_SEC=0.1)


# ----------------------------------------------------------------------------- features
def normalize_audio_volume(y, eps=1e-9):
    """audio/features.py:124-126."""
    rms = np.sqrt(np.mean(y ** 2))
    return y / (rms + eps)


def yin_estimate_pitch(signal, target_sr, fmin=50.0, fmax=1000.0):
    """dsp/yin.py:39-75 (+ round_to_nearest_pitch :21-37)."""
    f0 = librosa.yin(signal, fmin=fmin, fmax=fmax, sr=target_sr)
    valid = f0[~np.isnan(f0)]
    if len(valid) == 0:
        return None, {"midi": None, "note_name": None, "midi_float": None}
    pitch_hz = float(np.median(valid))
    if pitch_hz is None or np.isnan(pitch_hz) or pitch_hz <= 0:
        return pitch_hz, {"midi": None, "note_name": None, "midi_float": None}
    midi_float = librosa.hz_to_midi(pitch_hz)
    midi_rounded = int(np.round(midi_float))
    return pitch_hz, {"midi": midi_rounded, "note_name": librosa.midi_to_note(midi_rounded),
                      "midi_float": float(midi_float)}


def mfcc_vector(wave, sr, n_mfcc=64, normalize=True, add_pitch=True, yin_on_normalized=False):
    """One row of the MLP feature matrix.

    File path   audio/features.py:182-208 - YIN sees the RAW ``wave`` (:201).
    Memory path audio/features.py:458-478 - YIN sees the NORMALISED ``y`` (:473).
    """
    y = wave
    if normalize:
        y = normalize_audio_volume(y)
    mfcc = librosa.feature.mfcc(y=y, sr=sr, n_mfcc=n_mfcc)
    vec = mfcc.mean(axis=1)
    if add_pitch:
        hz, _ = yin_estimate_pitch(y if yin_on_normalized else wave, sr)
        if hz is not None:
            vec = np.concatenate((vec, np.array([float(np.log10(hz))], dtype=np.float32)), axis=0)
    return vec


def melspec_image(wave, sr, n_mels=64, n_fft=2048, hop_length=256, normalize=True, to_db=True):
    """audio/features.py:296-316 and :486-502: torchaudio MelSpectrogram(power=2) + AmplitudeToDB("power")."""
    mel = ta.transforms.MelSpectrogram(sample_rate=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels,
                                       power=2.0)
    db = ta.transforms.AmplitudeToDB(stype="power")
    y = wave.astype(np.float32)
    if normalize:
        y = normalize_audio_volume(y)
    spec = mel(torch.from_numpy(y).unsqueeze(0))
    if to_db:
        spec = db(spec)
    return spec  # (1, n_mels, T)


def extract_melspec_features(wavs, sr, n_mels=128, n_fft=1024, hop_length=256, normalize=False, to_db=True):
    """audio/features.py:275-341, the BATCH path: the two torchaudio transforms are built ONCE (:296-303), then a
    per-clip loop (:308-318) and right-zero-padding to the longest spectrogram (:320-331).  (The per-note path,
    ``melspec_image`` above, rebuilds the transforms on every call as features.py:486-493 does.)
    Returns X (N, 1, n_mels, T_max) float32."""
    mel_transform = ta.transforms.MelSpectrogram(sample_rate=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels,
                                                 power=2.0)
    to_db_transform = ta.transforms.AmplitudeToDB(stype="power")
    specs = []
    for wave in wavs:
        y = wave.astype(np.float32)
        if normalize:
            y = normalize_audio_volume(y)
        spec = mel_transform(torch.from_numpy(y).unsqueeze(0))
        if to_db:
            spec = to_db_transform(spec)
        specs.append(spec)
    max_T = max(s.shape[-1] for s in specs)
    padded = [F.pad(s, (0, max_T - s.shape[-1])) if s.shape[-1] < max_T else s[..., :max_T] for s in specs]
    return torch.stack(padded, dim=0)


def extract_inference_features_from_audio(audio, target_sr=TARGET_SR, mfcc_config=None, melspec_config=None,
                                          scaler=None, melspec_to_db=True):
    """audio/features.py:441-508.  NOTE the ``scaler`` argument is accepted and never applied there."""
    mfcc_config = mfcc_config or MFCC_DEFAULTS
    melspec_config = melspec_config or MELSPEC_DEFAULTS
    y = audio.astype(np.float32)
    vec = mfcc_vector(y, target_sr, mfcc_config["N_MFCC"], mfcc_config["NORMALIZE_AUDIO_VOLUME"],
                      mfcc_config["ADD_PITCH_FEATURES"], yin_on_normalized=True)
    mfcc_features = np.vstack([vec])
    spec = melspec_image(audio, target_sr, melspec_config["N_MELS"], melspec_config["N_FFT"],
                         melspec_config["HOP_LENGTH"], melspec_config["NORMALIZE_AUDIO_VOLUME"], melspec_to_db)
    return mfcc_features, spec.cpu().numpy()[:, None, :, :]


def extract_inference_features(wavs, target_sr, mfcc_config=None, melspec_config=None, scaler=None):
    """audio/features.py:130-158 over already-loaded, equal-length clips (the loader's pad_to_max=True).

    Here the scaler IS applied (:145-146); sklearn keeps the float32 matrix in float32 (mean_/scale_ are
    cast to it), so the result is float32."""
    mfcc_config = mfcc_config or MFCC_DEFAULTS
    melspec_config = melspec_config or MELSPEC_DEFAULTS
    X = np.vstack([mfcc_vector(w, target_sr, mfcc_config["N_MFCC"], mfcc_config["NORMALIZE_AUDIO_VOLUME"],
                               mfcc_config["ADD_PITCH_FEATURES"], yin_on_normalized=False) for w in wavs])
    if scaler:
        X = scaler.transform(X)
    M = extract_melspec_features(wavs, target_sr, melspec_config["N_MELS"], melspec_config["N_FFT"],
                                 melspec_config["HOP_LENGTH"], melspec_config["NORMALIZE_AUDIO_VOLUME"])
    return X, M


# ----------------------------------------------------------------------------- models
def mlp_forward(state, x):
    """training/mlp_trainer.py:32-105: [Linear, LayerNorm, LeakyReLU(0.1), Dropout]* + Linear, eval mode."""
    keys = sorted({int(k.split(".")[1]) for k in state if k.startswith("net.")})
    lin = [k for k in keys if state[f"net.{k}.weight"].ndim == 2]
    h = x
    for i, k in enumerate(lin):
        h = F.linear(h, state[f"net.{k}.weight"], state[f"net.{k}.bias"])
        if i + 1 < len(lin):
            g = state[f"net.{k + 1}.weight"]
            h = F.layer_norm(h, (g.shape[0],), g, state[f"net.{k + 1}.bias"], eps=1e-5)
            h = F.leaky_relu(h, 0.1)
    return h


def cnn_forward(state, x, adaptive_pool=(4, 4)):
    """training/cnn_trainer.py:30-139: [Conv3x3, BN, LeakyReLU(0.01), MaxPool2, Dropout]*3, AdaptiveAvgPool,
    Flatten, Linear, LeakyReLU, Dropout, Linear; eval mode (running stats, dropout = identity)."""
    conv_idx = sorted({int(k.split(".")[1]) for k in state
                       if k.startswith("features.") and k.endswith(".weight") and state[k].ndim == 4})
    h = x
    for k in conv_idx:
        w = state[f"features.{k}.weight"]
        h = F.conv2d(h, w, state[f"features.{k}.bias"], padding=w.shape[-1] // 2)
        if f"features.{k + 1}.running_mean" in state:
            h = F.batch_norm(h, state[f"features.{k + 1}.running_mean"], state[f"features.{k + 1}.running_var"],
                             state[f"features.{k + 1}.weight"], state[f"features.{k + 1}.bias"], False, 0.0, 1e-5)
        h = F.leaky_relu(h, 0.01)
        h = F.max_pool2d(h, 2)
    h = F.adaptive_avg_pool2d(h, adaptive_pool)
    h = torch.flatten(h, 1)
    fc = sorted({int(k.split(".")[1]) for k in state if k.startswith("classifier.") and k.endswith(".weight")})
    for i, k in enumerate(fc):
        h = F.linear(h, state[f"classifier.{k}.weight"], state[f"classifier.{k}.bias"])
        if i + 1 < len(fc):
            h = F.leaky_relu(h, 0.01)
    return h


def predict(mlp_ckpt, cnn_ckpt, mfcc_features, melspec_features, cnn_weight=0.80):
    """note_predictor.py:84-135: softmax each, (1-0.8)*mlp + 0.8*cnn in float32 numpy, first-max argmax."""
    mlp_weight = 1.0 - cnn_weight
    with torch.inference_mode():
        x = torch.from_numpy(np.asarray(mfcc_features, np.float32))
        mlp_logits = mlp_forward(mlp_ckpt["model"], x)
        mlp_probs = torch.softmax(mlp_logits, dim=-1).cpu().numpy()
        m = torch.from_numpy(np.asarray(melspec_features, np.float32))
        cnn_logits = cnn_forward(cnn_ckpt["model"], m,
                                 tuple(cnn_ckpt.get("model_init_args", {}).get("adaptive_pool", (4, 4))))
        cnn_probs = torch.softmax(cnn_logits, dim=-1).cpu().numpy()
    probs = mlp_weight * mlp_probs + cnn_weight * cnn_probs
    idx = np.argmax(probs, axis=1)
    reverse_map = mlp_ckpt.get("reverse_map")
    return {
        "indices": idx,
        "labels": [reverse_map[int(i)] for i in idx],
        "confidences": probs[np.arange(len(idx)), idx],
        "probs": probs,
        "per_model_probs": {"mlp": mlp_probs, "cnn": cnn_probs},
        "logits": {"mlp": mlp_logits.numpy(), "cnn": cnn_logits.numpy()},  # oracle-only extra
    }


def fix_len(audio, target_len):
    """transcribe.py:177-184 / audio/loading.py:54-70."""
    audio = np.asarray(audio)
    if len(audio) < target_len:
        z = np.zeros(target_len, dtype=audio.dtype)
        z[:len(audio)] = audio
        return z
    return audio[:target_len]


def transcribe_note(mlp_ckpt, cnn_ckpt, audio, clip_duration=CLIP_DURATION, sr_in=TARGET_SR):
    """transcribe.py:147-199."""
    if mlp_ckpt["config"]["target_sr"] != cnn_ckpt["config"]["target_sr"]:
        raise ValueError("[Transcriber] Target SR mismatch.")
    target_sr = mlp_ckpt["config"]["target_sr"]
    audio = audio.astype(np.float32, copy=False)
    if sr_in != target_sr:
        audio = librosa.resample(audio, orig_sr=sr_in, target_sr=target_sr).astype(np.float32, copy=False)
    audio = fix_len(audio, int(clip_duration * target_sr))
    mf, ms = extract_inference_features_from_audio(
        audio, target_sr, mlp_ckpt["config"]["features"]["params"], cnn_ckpt["config"]["features"]["params"],
        mlp_ckpt.get("scaler"), melspec_to_db=True)
    return predict(mlp_ckpt, cnn_ckpt, mf, ms)


# ----------------------------------------------------------------------------- segmentation
def apply_db_threshold(y, min_db=-45.0):
    """audio/slicing.py:30-39 (result is float64)."""
    amp_db = 20 * np.log10(np.abs(y) + 1e-10)
    return y * (amp_db > min_db).astype(float)


def compute_rms_db(y, frame_len=2048, hop_len=512, smooth=True):
    """audio/slicing.py:44-56."""
    rms = librosa.feature.rms(y=y, frame_length=frame_len, hop_length=hop_len, pad_mode="reflect")[0]
    rms_db = 20 * np.log10(rms + 1e-10)
    if smooth:
        rms_db = median_filter(rms_db, size=5)
    return rms_db


def compute_dynamic_thresholds(rms_db, noise_pct=20, signal_pct=75, gate_offset_db=6.0, slice_offset_db=10.0):
    """audio/slicing.py:59-76."""
    noise_floor = np.percentile(rms_db, noise_pct)
    signal_floor = np.percentile(rms_db, signal_pct)
    gate_db = noise_floor + gate_offset_db
    slice_min_db = noise_floor + slice_offset_db
    slice_min_db = max(slice_min_db, noise_floor + 5.0)
    slice_min_db = min(slice_min_db, signal_floor - 3.0)
    return gate_db, slice_min_db, (noise_floor, signal_floor)


def apply_rms_threshold(y, hop_len=512):
    """audio/slicing.py:78-91: frame i gates samples [i*hop, (i+1)*hop)."""
    rms_db = compute_rms_db(y=y, hop_len=hop_len)
    gate_db, _, _ = compute_dynamic_thresholds(rms_db)
    mask = np.repeat(rms_db > gate_db, hop_len)[:len(y)]
    return y * mask.astype(float)


def detect_onsets(y, sr=11025, hop_len=512, min_sep=0.25):
    """audio/slicing.py:106-122."""
    env = librosa.onset.onset_strength(y=y, sr=sr, hop_length=hop_len)
    frames = librosa.onset.onset_detect(onset_envelope=env, sr=sr, hop_length=hop_len, backtrack=True)
    samples = librosa.frames_to_samples(frames, hop_length=hop_len)
    min_samples = int(min_sep * sr)
    filtered, last = [], -999999
    for s in samples:
        if s - last >= min_samples:
            filtered.append(int(s))
            last = s
    return filtered


def slice_audio(y, onset, next_onset, sr=11025, length_sec=0.5, attack_skip_sec=0.1):
    """audio/slicing.py:125-136."""
    length = int(length_sec * sr)
    start = onset + int(attack_skip_sec * sr)
    end = min(start + length, next_onset)
    if start >= len(y) or end > len(y):
        return np.zeros((0,)), (0, 0)
    clip = y[start:end]
    if len(clip) < length:
        clip = np.pad(clip, (0, length - len(clip)))
    return clip, (start / sr, end / sr)


def is_slice_loud_enough(clip, min_rms_db=-40.0):
    """audio/slicing.py:96-100 (empty clip -> NaN -> False)."""
    with np.errstate(all="ignore"):
        rms = np.sqrt(np.mean(clip ** 2))
        return bool(20 * np.log10(rms + 1e-10) > min_rms_db)


def slice_in_memory(y, sr=TARGET_SR, length_sec=CLIP_DURATION, cfg=None):
    """audio/slicing.py:147-165 (``sliceNsave``) without load_wav and without the WAV write.

    ``detect_onsets`` is called WITHOUT hop_len (:151) so the hop is 512 whatever SLICER_CONFIG says; the
    last onset's ``next_onset`` is itself (:154) so its clip is all padding and is always dropped.
    Returns (onsets, clips[K', n] float32, table[K', 3] = (onset index i, start sample, end sample)).
    """
    cfg = cfg or SLICER_DEFAULTS
    y_gated = apply_db_threshold(y=y, min_db=cfg["MIN_IN_DB_THRESHOLD"])
    y_gated = apply_rms_threshold(y_gated, hop_len=cfg["HOP_LEN"])
    onsets = detect_onsets(y=y_gated, sr=sr, min_sep=cfg["MIN_SEP"])
    clips, table = [], []
    for i, onset in enumerate(onsets):
        next_onset = onsets[i + 1] if i + 1 < len(onsets) else onsets[-1]
        clip, times = slice_audio(y=y, onset=onset, next_onset=next_onset, sr=sr, length_sec=length_sec,
                                  attack_skip_sec=cfg["ATTACK_SKIP_SEC"])
        if not is_slice_loud_enough(clip, cfg["MIN_SLICE_RMS_DB"]):
            continue
        clips.append(clip.astype(np.float32))
        table.append((i, int(round(times[0] * sr)), int(round(times[1] * sr))))
    n = int(length_sec * sr)
    clips = np.stack(clips) if clips else np.zeros((0, n), np.float32)
    return onsets, clips, np.asarray(table, dtype=np.int64).reshape(-1, 3)


def transcribe_audio(mlp_ckpt, cnn_ckpt, y, sr=TARGET_SR, clip_duration=CLIP_DURATION):
    """transcribe.py:77-144 with the disk round trip removed (no PCM_16 quantisation, no resample:
    requires sr == checkpoint target_sr; the file front end is SURVEY 8f-1)."""
    target_sr = mlp_ckpt["config"]["target_sr"]
    if target_sr != cnn_ckpt["config"]["target_sr"]:
        raise ValueError("[Transcriber] Target SR mismatch.")
    if sr != target_sr:
        raise NotImplementedError("resampling is out of scope")
    onsets, clips, table = slice_in_memory(y, sr, clip_duration)
    if len(clips) == 0:
        raise FileNotFoundError("load_audio_dataset: No audio files found.")
    X, M = extract_inference_features(list(clips), target_sr, mlp_ckpt["config"]["features"]["params"],
                                      cnn_ckpt["config"]["features"]["params"], mlp_ckpt.get("scaler"))
    result = predict(mlp_ckpt, cnn_ckpt, X, M)
    result["dsp_info"] = [yin_estimate_pitch(c, target_sr) for c in clips]
    result["onsets"] = onsets
    result["slice_table"] = table
    return result


# ----------------------------------------------------------------------------- file pipeline
def pcm16_roundtrip(clip):
    """sf.write(.wav) (slicing.py:144; python-soundfile always enables clipping, so libsndfile's PCM_16 conversion
    is lrintf(x * 2^31) >> 16 with saturation) then librosa.load (loading.py:85, libsndfile read: / 0x8000),
    restated in oracle/soundfile_standin.py and librosa_shim.load."""
    import soundfile_standin
    q = soundfile_standin.float_to_pcm16(clip)
    return q.astype(np.float32) / np.float32(32768.0)


def transcribe_file(mlp_ckpt, cnn_ckpt, audio_path, target_sr=TARGET_SR, clip_duration=CLIP_DURATION):
    """transcribe.py:77-144: load at ``target_sr`` -> slice -> clip files (PCM_16) -> AudioDatasetLoader at the
    checkpoint's rate (resample + fix_len) -> features with the scaler -> predict -> YIN on the loaded clips.
    Clips in onset order (the reference iterates os.listdir order)."""
    ckpt_sr = mlp_ckpt["config"]["target_sr"]
    if ckpt_sr != cnn_ckpt["config"]["target_sr"]:
        raise ValueError("[Transcriber] Target SR mismatch.")
    y, sr = librosa.load(str(audio_path), sr=target_sr, mono=True)          # slicing.py:25
    onsets, clips, table = slice_in_memory(y, sr, clip_duration)
    if len(clips) == 0:
        raise FileNotFoundError("load_audio_dataset: No audio files found.")
    fixed = int(ckpt_sr * clip_duration)
    wavs = []
    for c in clips:
        w = pcm16_roundtrip(c)
        w = librosa.resample(w, orig_sr=sr, target_sr=ckpt_sr)            # librosa.load(path, sr=ckpt_sr), loading.py:85
        wavs.append(fix_len(w, fixed))                                      # loading.py:86
    X, M = extract_inference_features(wavs, ckpt_sr, mlp_ckpt["config"]["features"]["params"],
                                      cnn_ckpt["config"]["features"]["params"], mlp_ckpt.get("scaler"))
    result = predict(mlp_ckpt, cnn_ckpt, X, M)
    result["dsp_info"] = [yin_estimate_pitch(w, ckpt_sr) for w in wavs]
    result["onsets"] = onsets
    result["slice_table"] = table
    result["clips"] = np.stack(wavs)
    return result
