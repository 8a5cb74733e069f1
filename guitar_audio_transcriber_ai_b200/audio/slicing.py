"""``AudioSlicer`` - drop-in for the in-memory half of the reference's audio/slicing.py:16-165.

``detect_onsets`` / ``slice_in_memory`` run the whole gate -> onset -> slice chain on the GPU
(csrc/onset.cuh); WAV decode / write (``load_wav``, ``save_clip``) is out of scope.
"""
from __future__ import annotations

import numpy as np

from ..config import CLIP_DURATION, SLICER_CONFIG, TARGET_SR
from ..dsp.yin import shared_engine


class AudioSlicer:
    def __init__(self, device=None):
        self.device = device

    def _segment(self, y, sr, length_sec, cfg=None, diagnostics=False):
        return shared_engine(sr, self.device).segment(np.asarray(y, dtype=np.float32), length_sec, cfg, diagnostics)

    def detect_onsets_gated(self, y, sr=TARGET_SR, cfg=None) -> list[int]:
        """apply_db_threshold -> apply_rms_threshold -> detect_onsets as sliceNsave chains them
        (slicing.py:148-151): onset sample positions of the raw signal ``y``."""
        return [int(v) for v in self._segment(y, sr, CLIP_DURATION, cfg)["onsets"].cpu().numpy()]

    def slice_in_memory(self, y, sr=TARGET_SR, length_sec=CLIP_DURATION, cfg=None):
        """sliceNsave without file I/O (slicing.py:147-165).
        Returns (onsets list[int], clips float32 [K', n] on the device, table int64 [K', 3] on the host:
        onset index, start sample, end sample)."""
        r = self._segment(y, sr, length_sec, cfg)
        return [int(v) for v in r["onsets"].cpu().numpy()], r["clips"], r["table"].cpu().numpy()

    @staticmethod
    def slice_audio(y, onset, next_onset, sr=11025, length_sec=0.5, attack_skip_sec=0.1):
        """slicing.py:125-136 (host indexing helper, no arithmetic)."""
        length = int(length_sec * sr)
        start = onset + int(attack_skip_sec * sr)
        end = min(start + length, next_onset)
        if start >= len(y) or end > len(y):
            return np.zeros((0,)), (0, 0)
        clip = y[start:end]
        if len(clip) < length:
            clip = np.pad(clip, (0, length - len(clip)))
        return clip, (start / sr, end / sr)
