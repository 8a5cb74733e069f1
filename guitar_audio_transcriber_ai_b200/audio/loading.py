"""In-memory counterpart of the reference's ``AudioDatasetLoader`` (audio/loading.py:36-105).

The first argument is either a list of dataset roots (the reference's signature: every ``root/<label>/*.wav``
is decoded, averaged to mono and resampled to ``target_sr`` on the GPU - SURVEY 8f-1) or a list of clips already
in memory.  Files are visited in sorted order (the reference uses ``os.listdir`` order, which is arbitrary).
"""
from __future__ import annotations

import os
from pathlib import Path

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
    def __init__(self, clips, target_sr: int = 11025, mono: bool = True, duration: float | None = None, labels=None,
                 device=None):
        self.target_sr = target_sr
        self.mono = mono
        self.fixed_len = int(self.target_sr * duration) if duration is not None else None
        self.dataset_roots = None
        self._paths = None
        if len(clips) and all(isinstance(c, (str, os.PathLike)) for c in clips):
            self.dataset_roots = [Path(c) for c in clips]
            clips, labels, self._paths = self._read_roots(device)
        self._clips = [fix_len(np.asarray(c, dtype=np.float32), self.fixed_len) for c in clips]
        self._labels = list(labels) if labels is not None else ["clip"] * len(self._clips)

    def _read_roots(self, device):
        """loading.py:72-87: root/<label>/*.wav -> librosa.load(path, sr=target_sr, mono=True)."""
        from .slicing import AudioSlicer
        slicer = AudioSlicer(device=device)
        clips, labels, paths = [], [], []
        for root in self.dataset_roots:
            for folder in sorted(os.listdir(root)):
                folder_path = os.path.join(root, folder)
                if not os.path.isdir(folder_path):
                    continue
                for fname in sorted(os.listdir(folder_path)):
                    if not fname.endswith(".wav"):
                        continue
                    path = os.path.join(folder_path, fname)
                    y, _ = slicer.load_wav(path, self.target_sr)
                    clips.append(y.cpu().numpy()); labels.append(folder); paths.append(path)
        return clips, labels, paths

    def fix_len(self, y, fixed_len=None):
        return fix_len(y, fixed_len)

    def load_audio_dataset(self, pad_to_max=True):
        if len(self._clips) == 0:
            raise FileNotFoundError("load_audio_dataset: No audio files found.")
        wavs = list(self._clips)
        if pad_to_max:
            m = max(len(w) for w in wavs)
            wavs = [np.pad(w, (0, m - len(w)), mode="constant") for w in wavs]
        paths = list(self._paths) if self._paths is not None else [f"mem://{i}" for i in range(len(wavs))]
        return wavs, [self.target_sr] * len(wavs), list(self._labels), paths
