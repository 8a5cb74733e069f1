"""``YinDsp`` - drop-in for the reference's dsp/yin.py:11-75; the pitch track comes from csrc/yin.cuh."""
from __future__ import annotations

import math

import numpy as np

_NOTES_UNICODE = ["C", "C♯", "D", "D♯", "E", "F", "F♯", "G", "G♯", "A", "A♯", "B"]

_ENGINES: dict = {}


def shared_engine(sample_rate: int, device=None, melspec_config=None, mfcc_config=None):
    """One Engine per (sample rate, device, configs): the reference rebuilds its transforms on every call
    (features.py:486-493); here the context, tables and scratch live as long as the process."""
    from ..engine import Engine
    key = (int(sample_rate), str(device), tuple(sorted((melspec_config or {}).items())),
           tuple(sorted((mfcc_config or {}).items())))
    if key not in _ENGINES:
        _ENGINES[key] = Engine(sample_rate, melspec_config, mfcc_config, device=device)
    return _ENGINES[key]


class YinDsp:
    def __init__(self, fmin: float = 50.0, fmax: float = 1000.0, device=None):
        self.fmin = fmin
        self.fmax = fmax
        self.device = device
        self._engines: dict = {}

    @staticmethod
    def round_to_nearest_pitch(hz):
        """dsp/yin.py:21-37: (midi_rounded, note name with unicode sharps as librosa.midi_to_note, midi_float)."""
        if hz is None or np.isnan(hz) or hz <= 0:
            return None, None, None
        midi_float = 12 * (np.log2(hz) - np.log2(440.0)) + 69
        midi_rounded = int(np.round(midi_float))
        note_name = "{:s}{:0d}".format(_NOTES_UNICODE[midi_rounded % 12], int(midi_rounded / 12) - 1)
        return midi_rounded, note_name, float(midi_float)

    def _engine(self, sr):
        from ..engine import Engine
        key = int(sr)
        if key not in self._engines:
            if (self.fmin, self.fmax) == (50.0, 1000.0):
                self._engines[key] = shared_engine(sr, self.device)
            else:
                self._engines[key] = Engine(sr, device=self.device, yin_fmin=self.fmin, yin_fmax=self.fmax)
        return self._engines[key]

    def estimate_pitch_batch(self, clips, target_sr):
        """[N, n] clips -> (median Hz float64[N] on host, list of note_info dicts)."""
        hz, _ = self._engine(target_sr).yin(clips, normalize=False)
        hz = hz.cpu().numpy()
        infos = []
        for v in hz:
            if math.isnan(v):
                infos.append({"midi": None, "note_name": None, "midi_float": None})
            else:
                m, name, mf = self.round_to_nearest_pitch(float(v))
                infos.append({"midi": m, "note_name": name, "midi_float": mf})
        return hz, infos

    def estimate_pitch(self, signal, target_sr):
        """dsp/yin.py:39-75: (pitch_hz | None, {"midi", "note_name", "midi_float"})."""
        hz, infos = self.estimate_pitch_batch(np.asarray(signal, dtype=np.float32)[None, :], target_sr)
        return (None if math.isnan(hz[0]) else float(hz[0])), infos[0]
