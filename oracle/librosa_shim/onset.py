"""TEST INFRASTRUCTURE ONLY - restatement of librosa.onset.{onset_strength,onset_detect,onset_backtrack}."""
import numpy as np

from . import util
from .core import power_to_db, tiny
from .feature import melspectrogram


def onset_strength(*, y=None, sr=22050, S=None, lag=1, max_size=1, detrend=False, center=True,
                   n_fft=2048, hop_length=512, **kwargs):
    """librosa.onset.onset_strength (slicing.py:107): spectral flux of the mel-dB spectrogram.

    mel-128 Slaney power spectrogram with fmax=sr/2 -> power_to_db (top_db 80, max over the whole
    signal) -> relu of first difference -> mean over mels (np.mean over axis -2) -> left-pad
    ``lag + n_fft // (2*hop)`` zeros -> trim to the frame count.
    """
    assert max_size == 1 and not detrend
    if S is None:
        kwargs.setdefault("fmax", 0.5 * sr)
        S = np.abs(melspectrogram(y=y, sr=sr, n_fft=n_fft, hop_length=hop_length, **kwargs))
        S = power_to_db(S)
    S = np.atleast_2d(S)
    ref = S
    onset_env = S[..., lag:] - ref[..., :-lag]
    onset_env = np.maximum(0.0, onset_env)
    onset_env = np.mean(onset_env, axis=-2, keepdims=True)
    pad_width = lag
    if center:
        pad_width += n_fft // (2 * hop_length)
    onset_env = np.pad(onset_env, [(0, 0), (int(pad_width), 0)], mode="constant")
    if center:
        onset_env = onset_env[..., :S.shape[-1]]
    return onset_env[0]


def onset_backtrack(events, energy):
    """librosa.onset.onset_backtrack: roll each event back to the preceding local minimum of energy."""
    minima = np.flatnonzero((energy[1:-1] <= energy[:-2]) & (energy[1:-1] < energy[2:]))
    minima = util.fix_frames(1 + minima, x_min=0)
    return minima[util.match_events_left(events, minima)]


def onset_detect(*, y=None, sr=22050, onset_envelope=None, hop_length=512, backtrack=False, energy=None,
                 units="frames", normalize=True, **kwargs):
    """librosa.onset.onset_detect (slicing.py:109: onset_envelope, sr, hop_length, backtrack=True)."""
    assert units == "frames"
    if onset_envelope is None:
        onset_envelope = onset_strength(y=y, sr=sr, hop_length=hop_length)
    if normalize:
        onset_envelope = onset_envelope - np.min(onset_envelope, keepdims=True, axis=-1)
        onset_envelope /= np.max(onset_envelope, keepdims=True, axis=-1) + tiny(onset_envelope)
    if not onset_envelope.any() or not np.all(np.isfinite(onset_envelope)):
        return np.array([], dtype=int)
    kwargs.setdefault("pre_max", 0.03 * sr // hop_length)
    kwargs.setdefault("post_max", 0.00 * sr // hop_length + 1)
    kwargs.setdefault("pre_avg", 0.10 * sr // hop_length)
    kwargs.setdefault("post_avg", 0.10 * sr // hop_length + 1)
    kwargs.setdefault("wait", 0.03 * sr // hop_length)
    kwargs.setdefault("delta", 0.07)
    onsets = util.peak_pick(onset_envelope, **kwargs)
    if backtrack:
        if energy is None:
            energy = onset_envelope
        onsets = onset_backtrack(onsets, energy)
    return onsets
