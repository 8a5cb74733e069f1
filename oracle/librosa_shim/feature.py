"""TEST INFRASTRUCTURE ONLY - restatement of librosa.feature.{melspectrogram,mfcc,rms} (0.10.2 / 0.11.0)."""
import numpy as np
import scipy.fft

from . import filters
from .core import _spectrogram, frame, power_to_db


def melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                   window="hann", center=True, pad_mode="constant", power=2.0, **kwargs):
    """librosa.feature.melspectrogram: |STFT|**power, then a dense float32 mel-basis contraction."""
    if S is None:
        S, n_fft = _spectrogram(y, n_fft, hop_length, power, win_length=win_length, window=window,
                                center=center, pad_mode=pad_mode)
    mel_basis = filters.mel(sr=sr, n_fft=n_fft, **kwargs)
    return np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)


def mfcc(*, y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0, **kwargs):
    """librosa.feature.mfcc (reference: features.py:187-191 and :462-466, n_mfcc=64, all else default)."""
    if S is None:
        S = power_to_db(melspectrogram(y=y, sr=sr, **kwargs))
    M = scipy.fft.dct(S, axis=-2, type=dct_type, norm=norm)[..., :n_mfcc, :]
    assert lifter == 0
    return M


def rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant",
        dtype=np.float32):
    """librosa.feature.rms (reference: slicing.py:45-51, pad_mode="reflect").

    ``util.abs2(x, dtype=float32)`` is ``np.square(x, dtype=float32)``: the square is taken in float32
    even for float64 input, and the mean over the frame axis accumulates in float32.
    """
    assert S is None
    y = np.asarray(y)
    if center:
        y = np.pad(y, int(frame_length // 2), mode=pad_mode)
    x = frame(y, frame_length, hop_length)
    power = np.mean(np.square(x, dtype=dtype), axis=-2, keepdims=True)
    return np.sqrt(power)
