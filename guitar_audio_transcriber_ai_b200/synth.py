"""Synthetic guitar-like signals for parity tests and benchmarks.

The reference ships no audio; every BASELINE.json config is defined on these generators.  All of it is
host-side numpy and deterministic in ``seed`` so the CPU oracle and the CUDA path see identical bytes.
"""
from __future__ import annotations

import numpy as np

MIDI_LO, MIDI_HI = 40, 86  # E2 .. D6: the 47 classes of the shipped MLP checkpoint
_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def midi_to_hz(m):
    return 440.0 * 2.0 ** ((np.asarray(m, dtype=np.float64) - 69.0) / 12.0)


def midi_to_label(m: int) -> str:
    """ASCII scientific pitch notation as used by the reference's class folders ('F#3')."""
    return f"{_NAMES[int(m) % 12]}{int(m) // 12 - 1}"


def class_names() -> list[str]:
    """The 47 labels sorted as strings - the order features.py:107-112 gives the class indices."""
    return sorted(midi_to_label(m) for m in range(MIDI_LO, MIDI_HI + 1))


def note(f0: float, dur: float, sr: int, seed: int, noise: float = 1e-3, peak: float = 0.5) -> np.ndarray:
    """Decaying 8-harmonic pluck: sum_k (1/k) sin(2 pi k f0 t + phi_k) exp(-t/tau_k), tau_k = 0.6/sqrt(k)."""
    rng = np.random.default_rng(seed)
    n = int(round(dur * sr))
    t = np.arange(n, dtype=np.float64) / sr
    phases = rng.uniform(0.0, 2.0 * np.pi, size=8)
    y = np.zeros(n, dtype=np.float64)
    for k in range(1, 9):
        if k * f0 >= 0.45 * sr:
            break
        y += (1.0 / k) * np.sin(2.0 * np.pi * k * f0 * t + phases[k - 1]) * np.exp(-t / (0.6 / np.sqrt(k)))
    na = max(1, int(round(0.004 * sr)))
    y[:na] *= 0.5 * (1.0 - np.cos(np.pi * np.arange(na) / na))
    y *= peak / max(np.max(np.abs(y)), 1e-12)
    y += noise * rng.standard_normal(n)
    return y.astype(np.float32)


def random_midi(seed: int) -> int:
    return int(np.random.default_rng(10_000_019 * 7 + seed).integers(MIDI_LO, MIDI_HI + 1))


def clip_batch(n_clips: int, dur: float, sr: int, seed0: int = 0):
    """Config 2/3: ``n_clips`` independent notes, seed = seed0 + clip index. Returns (audio[N,n] f32, midi[N])."""
    n = int(round(dur * sr))
    out = np.empty((n_clips, n), dtype=np.float32)
    midi = np.empty(n_clips, dtype=np.int64)
    for i in range(n_clips):
        midi[i] = random_midi(seed0 + i)
        out[i] = note(float(midi_to_hz(midi[i])), dur, sr, seed0 + i)
    return out, midi


def phrase(seed: int, sr: int = 22050, dur: float = 5.0, n_notes: int = 10, spacing: float = 0.5,
           t0: float = 0.05, sounding: float = 0.35, fade: float = 0.1, noise: float = 1e-3):
    """Config 1: monophonic phrase, one note every ``spacing`` s, each damped after ``sounding`` s.

    The gaps matter: the reference gates frames at (20th percentile of RMS dB) + 6 dB
    (slicing.py:59-91), so a legato phrase with no quiet frames is gated away entirely and yields no
    onsets.  Damped notes leave >20 % quiet frames, as a real single-note guitar take does.
    """
    rng = np.random.default_rng(seed)
    n = int(round(dur * sr))
    y = np.zeros(n, dtype=np.float64)
    midis = rng.integers(MIDI_LO, MIDI_HI + 1, size=n_notes)
    starts = [int(round((t0 + i * spacing) * sr)) for i in range(n_notes)]
    nf = int(round(fade * sr))
    for i, (m, s) in enumerate(zip(midis, starts)):
        e = min(s + int(round(sounding * sr)), n)
        seg = note(float(midi_to_hz(m)), (e - s) / sr, sr, seed * 1000 + i, noise=0.0).astype(np.float64)
        seg = seg[: e - s]
        seg[-nf:] *= 0.5 * (1.0 + np.cos(np.pi * np.arange(nf) / nf))
        y[s:s + len(seg)] += seg
    y += noise * rng.standard_normal(n)
    return y.astype(np.float32), midis.astype(np.int64), np.asarray(starts, dtype=np.int64)


def long_audio(n_phrases: int, sr: int = 22050, seed0: int = 0):
    """Config 4: concatenated config-1 phrases (720 of them = 1 hour)."""
    parts, midis, starts = [], [], []
    off = 0
    for p in range(n_phrases):
        y, m, s = phrase(seed0 + p, sr=sr)
        parts.append(y)
        midis.append(m)
        starts.append(s + off)
        off += len(y)
    return np.concatenate(parts), np.concatenate(midis), np.concatenate(starts)


def wav_case(name: str):
    """Deterministic WAV-file test inputs for the file pipeline (Transcriber.transcribe): returns
    (int16 frames [n] or [n, channels], sample rate).  The same arrays are written to disk by the golden-vector
    generator and by the tests, so no audio needs to be stored.

      mono22050   : config-1 phrase (seed 0), mono PCM_16 at 22 050 Hz - no resampling anywhere
      stereo32000 : config-1 phrase (seed 5) synthesised at 32 000 Hz, two channels with different gains,
                    PCM_16 - exercises the channel mean and both resampling steps (32 000 -> 22 050 -> 11 025)
    """
    def q16(x):
        return np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16)
    if name == "mono22050":
        y, _, _ = phrase(0, sr=22050)
        return q16(y), 22050
    if name == "stereo32000":
        y, _, _ = phrase(5, sr=32000)
        return np.stack([q16(0.9 * y), q16(0.7 * y)], axis=1), 32000
    raise KeyError(name)
