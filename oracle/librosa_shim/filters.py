"""TEST INFRASTRUCTURE ONLY - restatement of librosa.filters.mel (librosa 0.10.2 / 0.11.0)."""
import numpy as np

from .core import fft_frequencies, mel_frequencies


def mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """Triangular mel filterbank (n_mels, 1 + n_fft//2).

    Bin centres come from ``np.fft.rfftfreq`` (top bin is exactly sr/2 - unlike torchaudio's
    ``linspace(0, sr//2, n)``); Slaney mel scale and area normalisation by default.
    """
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        raise NotImplementedError(norm)
    return weights
