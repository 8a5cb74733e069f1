"""Host-side constant tables handed to gat_ctx_create (include/gat.h ``gat_config``).

They are built with the same library calls / formulas the reference's dependencies use, so the device
kernels multiply by bit-identical constants:

  * CNN chain  - torchaudio ``MelSpectrogram``: ``torch.hann_window`` and HTK ``melscale_fbanks`` with
    ``f_max = float(sr // 2)`` and ``all_freqs = linspace(0, sr // 2, n_freqs)`` (torchaudio
    functional.py:518-587, transforms/_transforms.py:393).
  * MFCC/onset chains - librosa 0.10/0.11: float64 periodic Hann, Slaney mel filterbank on
    ``rfftfreq`` bin centres with area normalisation, orthonormal DCT-II.
  * slicer scalars - numpy's float32 percentile index arithmetic and librosa's onset_detect defaults.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft
import scipy.signal
import torch


def hann_window_f32(n_fft: int) -> np.ndarray:
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32).numpy().copy()


def hann_window_f64(n_fft: int) -> np.ndarray:
    return np.ascontiguousarray(scipy.signal.get_window("hann", n_fft, fftbins=True), dtype=np.float64)


def htk_fbanks(sample_rate: int, n_fft: int, n_mels: int) -> np.ndarray:
    """(n_freqs, n_mels) float32, the matrix torchaudio's MelScale multiplies by (norm=None, htk)."""
    n_freqs = n_fft // 2 + 1
    f_min, f_max = 0.0, float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    return np.ascontiguousarray(fb.numpy(), dtype=np.float32)


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_fb(sample_rate: int, n_fft: int = 2048, n_mels: int = 128) -> np.ndarray:
    """(n_mels, n_freqs) float32 as librosa.filters.mel(sr, n_fft, n_mels) (htk=False, norm='slaney')."""
    fmax = float(sample_rate) / 2
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float32)
    for i in range(n_mels):
        weights[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return np.ascontiguousarray(weights)


def dct_matrix(n_mfcc: int, n_mels: int = 128) -> np.ndarray:
    """(n_mfcc, n_mels) float32 rows of the orthonormal DCT-II (scipy.fft.dct(type=2, norm='ortho'))."""
    full = scipy.fft.dct(np.eye(n_mels, dtype=np.float64), type=2, norm="ortho", axis=0)
    return np.ascontiguousarray(full[:n_mfcc], dtype=np.float32)


def sample_gate_threshold(min_db: float) -> np.float32:
    """Smallest float32 amplitude a with ``20 * np.log10(a + 1e-10) > min_db`` in float32 numpy arithmetic
    (the test AudioSlicer.apply_db_threshold applies per sample, slicing.py:32-36), found by bisection on
    the float32 bit pattern.  The device gate is then the exact comparison |y| >= threshold."""
    def keeps(bits: int) -> bool:
        a = np.array([bits], dtype=np.uint32).view(np.float32)
        return bool((20 * np.log10(a + 1e-10) > min_db)[0])
    lo, hi = 0, int(np.array([4.0], dtype=np.float32).view(np.uint32)[0])
    if keeps(lo):
        return np.float32(0.0)
    if not keeps(hi):
        return np.float32(np.inf)
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if keeps(mid):
            hi = mid
        else:
            lo = mid
    return np.array([hi], dtype=np.uint32).view(np.float32)[0]


def percentile_index_f32(n: int, pct: float):
    """np.percentile(float32 array of n values, pct) 'linear': virtual index (n-1)*q in float32, its floor
    and fractional part (numpy/lib/_function_base_impl.py: percentile -> _quantile -> _lerp)."""
    q = np.true_divide(pct, np.float32(100))
    virtual = (n - 1) * q
    k = int(np.floor(virtual))
    gamma = np.float32(virtual - np.float32(k))
    k = min(max(k, 0), n - 1)
    return k, gamma


def onset_detect_params(sr: int, hop: int) -> dict:
    """librosa.onset.onset_detect defaults, then util.peak_pick's ceil-to-int."""
    vals = {
        "pre_max": 0.03 * sr // hop,
        "post_max": 0.00 * sr // hop + 1,
        "pre_avg": 0.10 * sr // hop,
        "post_avg": 0.10 * sr // hop + 1,
        "wait": 0.03 * sr // hop,
    }
    out = {k: int(np.ceil(v)) for k, v in vals.items()}
    out["delta"] = np.float32(0.07)
    return out


def resample_filter(orig_sr: int, target_sr: int, attenuation_db: float = 120.0, passband: float = 0.913):
    """Polyphase low-pass for gat_resample: (up, down, taps float64 [2*half+1] scaled by up, half).

    The reference resamples with soxr's HQ recipe through ``librosa.load`` / ``librosa.resample``
    (audio/loading.py:85, transcribe.py:173): linear phase, pass band kept to 0.913 of the lower Nyquist
    frequency, stop band from that Nyquist frequency on, about 20 bits of rejection.  soxr is not available, so
    this designs a Kaiser-windowed sinc to the same specification; the arithmetic that applies it restates
    scipy.signal.resample_poly.  Not bit-identical to soxr (file front end)."""
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    q = max(up, down)
    width = (1.0 - passband) / q                       # transition band, in units of the up-sampled grid's Nyquist
    cutoff = (1.0 + passband) / (2.0 * q)              # centre of the transition band
    beta = 0.1102 * (attenuation_db - 8.7)
    half = int(math.ceil((attenuation_db - 8.0) / (2.285 * math.pi * width) / 2.0))
    n = np.arange(-half, half + 1, dtype=np.float64)
    h = cutoff * np.sinc(cutoff * n) * np.kaiser(2 * half + 1, beta)
    h /= h.sum()
    return up, down, np.ascontiguousarray(h * up), half
