"""``AudioSlicer`` - drop-in for the in-memory half of the reference's audio/slicing.py:16-165.

``detect_onsets`` / ``slice_in_memory`` run the whole gate -> onset -> slice chain on the GPU
(csrc/onset.cuh).  ``load_wav`` / ``save_clip`` / ``sliceNsave`` are the file front end: the
container is parsed on the host (audio/wavio.py), sample conversion, channel mean, resampling and the PCM_16
quantisation of saved clips run on the GPU (csrc/frontend.cuh).
"""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from ..config import CLIP_DURATION, SLICER_CONFIG, TARGET_SR
from ..dsp.yin import shared_engine
from . import wavio


class AudioSlicer:
    def __init__(self, device=None):
        self.device = device

    def engine(self, sr):
        return shared_engine(sr, self.device)

    def _segment(self, y, sr, length_sec, cfg=None, diagnostics=False):
        y = y if torch.is_tensor(y) else np.asarray(y, dtype=np.float32)
        return self.engine(sr).segment(y, length_sec, cfg, diagnostics)

    # ------------------------------------------------------------------ file front end
    def load_wav(self, path, sr=11025):
        """slicing.py:19-26 (librosa.load(path, sr=sr, mono=True)): (float32 mono signal ON THE DEVICE, sr)."""
        frames, sr_file = wavio.read_wav_frames(path)
        eng = self.engine(sr)
        y = eng.decode_mono(frames)
        if sr_file != sr:
            y = eng.resample(y, sr_file, sr)
        return y, sr

    @staticmethod
    def save_clip(clip, sr, out_dir, idx, onset_s, audio_name="clip"):
        """slicing.py:139-144: ``NNNN_<name>__<onset>s.wav``, PCM_16 as soundfile writes .wav files
        (libsndfile's clipping float -> short conversion, ``wavio.float_to_pcm16``)."""
        out_dir = Path(out_dir)
        out_dir.mkdir(parents=True, exist_ok=True)
        x = clip.detach().cpu().numpy() if torch.is_tensor(clip) else np.asarray(clip, dtype=np.float32)
        q = wavio.float_to_pcm16(x)
        wavio.write_wav_pcm16(out_dir / f"{idx:04d}_{audio_name}__{onset_s:.3f}s.wav", q, sr)

    def sliceNsave(self, audio_path, out_dir, target_sr=TARGET_SR, hop_len=SLICER_CONFIG.HOP_LEN, length_sec=CLIP_DURATION,
                   min_sep=SLICER_CONFIG.MIN_SEP, min_db_threshold=SLICER_CONFIG.MIN_IN_DB_THRESHOLD,
                   min_slice_rms_db=SLICER_CONFIG.MIN_SLICE_RMS_DB, attack_skip_sec=SLICER_CONFIG.ATTACK_SKIP_SEC):
        """slicing.py:147-165: load, gate, detect onsets, slice, drop quiet clips, write the rest; returns the onsets."""
        cfg = SimpleNamespace(MIN_IN_DB_THRESHOLD=min_db_threshold, MIN_SLICE_RMS_DB=min_slice_rms_db, HOP_LEN=hop_len,
                              MIN_SEP=min_sep, ATTACK_SKIP_SEC=attack_skip_sec)
        y, sr = self.load_wav(audio_path, target_sr)
        r = self._segment(y, sr, length_sec, cfg)
        onsets = [int(v) for v in r["onsets"].cpu().numpy()]
        for clip, row in zip(r["clips"].cpu().numpy(), r["table"].cpu().numpy()):
            self.save_clip(clip, sr, out_dir, int(row[0]), onsets[int(row[0])] / sr)
        return onsets

    def detect_onsets(self, y, sr=11025, hop_len=512, min_sep=0.25) -> list[int]:
        """slicing.py:106-122: onset strength -> peak picking with backtracking -> minimum separation, on ``y``
        as given (no gates).  The live prototype calls it with hop 1024 on the microphone buffer."""
        return [int(v) for v in self.engine(sr).detect_onsets(y, hop_len, min_sep).cpu().numpy()]

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
