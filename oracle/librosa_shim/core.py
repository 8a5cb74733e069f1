"""TEST INFRASTRUCTURE ONLY - restatement of librosa.core pieces (librosa 0.10.2 / 0.11.0 semantics).

Every function names the librosa routine it restates and the reference call site that reaches it.
See the package docstring for the "parity unpinned" caveat.
"""
from __future__ import annotations

import numpy as np
import scipy.signal


def tiny(x):
    """librosa.util.tiny: smallest positive normal number of x's (floating) dtype."""
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.dtype(np.float32)
    return np.finfo(dtype).tiny


def get_window_hann(n: int) -> np.ndarray:
    """librosa.filters.get_window("hann", n, fftbins=True): periodic Hann, float64."""
    return scipy.signal.get_window("hann", n, fftbins=True)


def frame(y: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """librosa.util.frame on the last axis: (frame_length, n_frames) strided view."""
    y = np.asarray(y)
    if y.shape[-1] < frame_length:
        raise ValueError(f"Input is too short (n={y.shape[-1]}) for frame_length={frame_length}")
    view = np.lib.stride_tricks.sliding_window_view(y, frame_length, axis=-1)  # (n-fl+1, fl)
    view = view[::hop_length]
    return np.moveaxis(view, -1, -2)  # (fl, T)


def stft(y, *, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
         dtype=None, pad_mode="constant"):
    """librosa.stft.  Reached from feature.melspectrogram (features.py:187,462; slicing.py:107).

    float64 window * float32 frames -> float64 FFT -> cast to complex64 (float32 input) or kept as
    complex128 (float64 input), exactly as librosa's ``fft.rfft(fft_window * y_frames)`` written into a
    ``util.dtype_r2c(y.dtype)`` matrix does.
    """
    assert window == "hann"
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    fft_window = get_window_hann(win_length)
    if win_length < n_fft:  # util.pad_center
        lpad = (n_fft - win_length) // 2
        fft_window = np.pad(fft_window, (lpad, n_fft - win_length - lpad))
    fft_window = fft_window.reshape(-1, 1)
    if center:
        y = np.pad(y, n_fft // 2, mode=pad_mode)
    y_frames = frame(y, n_fft, hop_length)
    if dtype is None:
        dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    out = np.empty((1 + n_fft // 2, y_frames.shape[-1]), dtype=dtype, order="F")
    # librosa processes blocks of columns bounded by MAX_MEM_BLOCK; blocking does not change values.
    n_columns = max(1, int(2 ** 18 // (out.shape[0] * out.itemsize)))
    for s in range(0, y_frames.shape[-1], n_columns):
        t = min(s + n_columns, y_frames.shape[-1])
        out[:, s:t] = np.fft.rfft(fft_window * y_frames[:, s:t], axis=-2)
    return out


def _spectrogram(y, n_fft, hop_length, power, win_length=None, window="hann", center=True,
                 pad_mode="constant"):
    """librosa.core.spectrum._spectrogram: ``np.abs(stft(...)) ** power``."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, center=center,
                    window=window, pad_mode=pad_mode)) ** power
    return S, n_fft


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db; the top_db clamp uses the max over the WHOLE array (clip or file)."""
    S = np.asarray(S)
    magnitude = S
    ref_value = np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def frames_to_samples(frames, *, hop_length=512, n_fft=None):
    """librosa.frames_to_samples (slicing.py:111)."""
    offset = int(n_fft // 2) if n_fft is not None else 0
    return (np.asanyarray(frames) * hop_length + offset).astype(int)


def fft_frequencies(*, sr=22050, n_fft=2048):
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def hz_to_mel(frequencies, *, htk=False):
    frequencies = np.asanyarray(frequencies)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, *, htk=False):
    mels = np.asanyarray(mels)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, *, fmin=0.0, fmax=11025.0, htk=False):
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


def hz_to_midi(frequencies):
    """librosa.hz_to_midi (yin.py:33)."""
    return 12 * (np.log2(np.asanyarray(frequencies)) - np.log2(440.0)) + 69


_SHARP_NOTES_UNICODE = ["C", "C♯", "D", "D♯", "E", "F", "F♯", "G", "G♯", "A", "A♯", "B"]
_SHARP_NOTES_ASCII = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def midi_to_note(midi, *, octave=True, cents=False, key="C:maj", unicode=True):
    """librosa.midi_to_note (yin.py:35): key C:maj spells accidentals as sharps, unicode by default."""
    if not np.isscalar(midi):
        return [midi_to_note(x, octave=octave, cents=cents, key=key, unicode=unicode) for x in midi]
    note_map = _SHARP_NOTES_UNICODE if unicode else _SHARP_NOTES_ASCII
    note_num = int(np.round(midi))
    note_cents = int(100 * np.around(midi - note_num, 2))
    note = note_map[note_num % 12]
    if octave:
        note = "{:s}{:0d}".format(note, int(note_num / 12) - 1)
    if cents:
        note = f"{note:s}{note_cents:+02d}"
    return note


# --------------------------------------------------------------------------- YIN
def _localmin(x, axis=-2):
    """librosa.util.localmin: interior x[i] < x[i-1] and x[i] <= x[i+1]; first False; last x[-1] < x[-2]."""
    xi = np.swapaxes(x, -1, axis)
    out = np.zeros(xi.shape, dtype=bool)
    out[..., 1:-1] = (xi[..., 1:-1] < xi[..., :-2]) & (xi[..., 1:-1] <= xi[..., 2:])
    out[..., -1] = xi[..., -1] < xi[..., -2]
    return np.swapaxes(out, -1, axis)


def _parabolic_interpolation(x, axis=-2):
    """librosa.core.pitch._parabolic_interpolation.

    The numba stencil is ``a = x[1] + x[-1] - 2 * x[0]; b = (x[1] - x[-1]) / 2``.  For float32 input
    numba adds/subtracts the two float32 neighbours in float32, and only then promotes to float64
    because of the int64 literal (verified against a numba build of the same stencil text in
    tests/test_oracle_shim.py); the float64 result is stored back in x's dtype.
    """
    xi = np.swapaxes(x, -1, axis)
    shifts = np.zeros(xi.shape, dtype=x.dtype)
    xm = xi[..., :-2]
    x0 = xi[..., 1:-1].astype(np.float64)
    xp = xi[..., 2:]
    a = (xp + xm).astype(np.float64) - 2 * x0
    b = (xp - xm).astype(np.float64) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(np.abs(b) >= np.abs(a), 0.0, -b / a)
    shifts[..., 1:-1] = s.astype(x.dtype)
    return np.swapaxes(shifts, -1, axis)


def _cumulative_mean_normalized_difference(y_frames, frame_length, win_length, min_period, max_period,
                                           tiny_):
    a = np.fft.rfft(y_frames, frame_length, axis=-2)
    b = np.fft.rfft(y_frames[..., win_length:0:-1, :], frame_length, axis=-2)
    acf_frames = np.fft.irfft(a * b, frame_length, axis=-2)[..., win_length:, :]
    acf_frames[np.abs(acf_frames) < 1e-6] = 0

    energy_frames = np.cumsum(y_frames ** 2, axis=-2)
    energy_frames = energy_frames[..., win_length:, :] - energy_frames[..., :-win_length, :]
    energy_frames[np.abs(energy_frames) < 1e-6] = 0

    yin_frames = energy_frames[..., :1, :] + energy_frames - 2 * acf_frames

    yin_numerator = yin_frames[..., min_period:max_period + 1, :]
    tau_range = np.arange(1, max_period + 1).reshape(-1, 1)
    cumulative_mean = np.cumsum(yin_frames[..., 1:max_period + 1, :], axis=-2) / tau_range
    yin_denominator = cumulative_mean[..., min_period - 1:max_period, :]
    return yin_numerator / (yin_denominator + tiny_)


def yin(y, *, fmin, fmax, sr=22050, frame_length=2048, win_length=None, hop_length=None,
        trough_threshold=0.1, center=True, pad_mode="constant"):
    """librosa.yin (yin.py:49-54 calls it with fmin=50, fmax=1000, sr only).

    Returns float64 f0 per frame for float32 input (int64 index + float32 shift promotes to float64).
    """
    y = np.asarray(y)
    if win_length is None:
        win_length = frame_length // 2
    if hop_length is None:
        hop_length = frame_length // 4
    if center:
        y = np.pad(y, frame_length // 2, mode=pad_mode)
    y_frames = frame(y, frame_length, hop_length)
    min_period = int(np.floor(sr / fmax))
    max_period = min(int(np.ceil(sr / fmin)), frame_length - win_length - 1)
    tiny_ = tiny(y_frames)
    yin_frames = _cumulative_mean_normalized_difference(y_frames, frame_length, win_length,
                                                        min_period, max_period, tiny_)
    parabolic_shifts = _parabolic_interpolation(yin_frames)
    is_trough = _localmin(yin_frames, axis=-2)
    is_trough[..., 0, :] = yin_frames[..., 0, :] < yin_frames[..., 1, :]
    is_threshold_trough = np.logical_and(is_trough, yin_frames < trough_threshold)
    target_shape = list(yin_frames.shape)
    target_shape[-2] = 1
    global_min = np.argmin(yin_frames, axis=-2).reshape(target_shape)
    yin_period = np.argmax(is_threshold_trough, axis=-2).reshape(target_shape)
    no_trough = np.all(~is_threshold_trough, axis=-2, keepdims=True)
    yin_period[no_trough] = global_min[no_trough]
    yin_period = (min_period + yin_period
                  + np.take_along_axis(parabolic_shifts, yin_period, axis=-2))[..., 0, :]
    return sr / yin_period


# --------------------------------------------------------------------------- file front end
def _soxr_hq_like_filter(up, down, attenuation_db=120.0, passband=0.913):
    """Kaiser-windowed sinc to soxr HQ's published specification (pass band to 0.913 of the lower Nyquist,
    stop band from that Nyquist on, ~20 bits of rejection).  soxr is absent: NOT its coefficients."""
    import scipy.signal
    q = max(up, down)
    width = (1.0 - passband) / q
    cutoff = (1.0 + passband) / (2.0 * q)
    beta = 0.1102 * (attenuation_db - 8.7)
    half = int(np.ceil((attenuation_db - 8.0) / (2.285 * np.pi * width) / 2.0))
    return scipy.signal.firwin(2 * half + 1, cutoff, window=("kaiser", beta))


def resample(y, *, orig_sr, target_sr, res_type="soxr_hq", **_):
    """librosa.resample: output length ceil(n * target/orig), float32 in -> float32 out.  The filter is the
    soxr-HQ-like design above applied with scipy.signal.resample_poly (unpinned against soxr)."""
    if orig_sr == target_sr:
        return y
    import math
    import scipy.signal
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    h = _soxr_hq_like_filter(up, down)
    out = scipy.signal.resample_poly(np.asarray(y), up, down, axis=-1, window=h)
    n_out = int(np.ceil(np.asarray(y).shape[-1] * float(target_sr) / float(orig_sr)))
    assert out.shape[-1] == n_out
    return np.ascontiguousarray(out, dtype=np.float32)


def load(path, *, sr=22050, mono=True, **_):
    """librosa.load: libsndfile decode to float32 (16/32-bit PCM, float WAV), channel mean, resample."""
    import scipy.io.wavfile
    sr_in, data = scipy.io.wavfile.read(path)
    if data.dtype == np.int16:
        data = data.astype(np.float32) / np.float32(32768.0)
    elif data.dtype == np.int32:
        data = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    data = data.astype(np.float32)
    if data.ndim == 2 and mono:
        data = np.mean(data.T, axis=0)          # librosa.to_mono on (channels, n)
    if sr is not None and sr_in != sr:
        data = resample(data, orig_sr=sr_in, target_sr=sr)
        sr_in = sr
    return data, sr_in
