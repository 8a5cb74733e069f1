"""In-memory counterpart of the reference's ``AudioDatasetLoader`` (audio/loading.py:36-105).

The disk walk / ``librosa.load`` front end is out of scope; what the hot path needs from the
loader is ``target_sr``, ``fix_len`` and ``load_audio_dataset(pad_to_max)`` over clips already in memory.
"""
from __future__ import annotations

import numpy as np


def fix_len(y: np.ndarray, fixed_len=None) -> np.ndarray:
    """loading.py:54-70: truncate or right-zero-pad to ``fixed_len`` samples."""
    if fixed_len is None:
        return y
    if len(y) > fixed_len:
        return y[:fixed_len]
    if len(y) < fixed_len:
        return np.pad(y, (0, fixed_len - len(y)), mode="constant")
    return y


class AudioDatasetLoader:
    def __init__(self, clips, target_sr: int = 11025, mono: bool = True, duration: float | None = None, labels=None):
        self.target_sr = target_sr
        self.mono = mono
        self.fixed_len = int(self.target_sr * duration) if duration is not None else None
        self._clips = [fix_len(np.asarray(c, dtype=np.float32), self.fixed_len) for c in clips]
        self._labels = list(labels) if labels is not None else ["clip"] * len(self._clips)

    def fix_len(self, y, fixed_len=None):
        return fix_len(y, fixed_len)

    def load_audio_dataset(self, pad_to_max=True):
        if len(self._clips) == 0:
            raise FileNotFoundError("load_audio_dataset: No audio files found.")
        wavs = list(self._clips)
        if pad_to_max:
            m = max(len(w) for w in wavs)
            wavs = [np.pad(w, (0, m - len(w)), mode="constant") for w in wavs]
        return wavs, [self.target_sr] * len(wavs), list(self._labels), [f"mem://{i}" for i in range(len(wavs))]
