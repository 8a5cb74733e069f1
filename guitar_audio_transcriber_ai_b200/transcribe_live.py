"""Streaming / live path - mirror of the reference's experimental prototyping/source/transcribe_live.py:41-271
(SURVEY 8f-4), minus the sound card: ``feed()`` is the body of the ``sounddevice`` callback and ``step()`` one
pass of the main loop, so the same state machine can be driven from a microphone thread, a file or a test.

    mic block --feed()--> RingBuffer (1.5 s) --step(): buffer full?--> AudioSlicer.detect_onsets(hop 1024, min_sep 0.3)
        --> slices between consecutive onsets longer than 0.3 s, padded / trimmed to CLIP_DURATION
        --> note queue --> Transcriber.transcribe_note --> (label, confidence)

Onset detection and the per-note transcription run on the GPU through the engines the Transcriber already owns (no
per-call setup: the reference rebuilds its torchaudio transforms for every note).  Deviations from the prototype,
which does not run as shipped: ``inference`` references an undefined ``audio_slicer`` name (:242) and writes a
temporary WAV nobody reads (:249-252) - both dropped; ``RingBuffer.clear_from(idx)`` pops the ``idx`` NEWEST samples
(:76-78) although the loop means to discard audio up to the handled onset - kept as is by default
(``drop_newest=True``), with the evident intent available as ``drop_newest=False``.
"""
from __future__ import annotations

import queue

import numpy as np

from .config import CLIP_DURATION, TARGET_SR
from .transcribe import Transcriber


def rms_db(x: np.ndarray, eps=1e-12):
    """transcribe_live.py:33-36."""
    r = np.sqrt(np.mean(x * x) + eps)
    return 20.0 * np.log10(r + eps)


class RingBuffer:
    """transcribe_live.py:41-78: at most ``maxlen`` float32 samples, oldest dropped first (a numpy array instead
    of a deque of Python floats)."""

    def __init__(self, maxlen: int):
        self.maxlen = int(maxlen)
        self._data = np.zeros(0, dtype=np.float32)

    def push(self, data: np.ndarray):
        self._data = np.concatenate([self._data, np.asarray(data, dtype=np.float32).reshape(-1)])[-self.maxlen:]

    def pop(self):
        self._data = self._data[:-1]

    def get_buffer(self) -> np.ndarray:
        return self._data.copy()

    def get_slice(self, i, j) -> np.ndarray:
        if len(self._data) < i or len(self._data) < j:
            return np.zeros((0,), dtype=np.float32)
        return self._data[i:j].copy()

    def is_full(self):
        return len(self._data) == self.maxlen

    def size(self):
        return len(self._data)

    def clear(self):
        self._data = np.zeros(0, dtype=np.float32)

    def clear_from(self, idx, drop_newest: bool = True):
        if idx > len(self._data):
            raise IndexError("pop from an empty deque")
        self._data = self._data[: len(self._data) - idx] if drop_newest else self._data[idx:]


class LiveTranscriber:
    def __init__(self, device=None, buffer_duration=1.5, sample_rate=TARGET_SR, channels=1, blocksize=1024,
                 transcriber: Transcriber | None = None, drop_newest: bool = True):
        self.device = device
        self.buffer_duration = buffer_duration
        self.sample_rate = sample_rate
        self.channels = channels
        self.blocksize = blocksize
        self.buffer_maxlen = int(self.buffer_duration * self.sample_rate)
        self.buffer = RingBuffer(maxlen=self.buffer_maxlen)
        self.note_q: queue.Queue = queue.Queue(maxsize=2)
        self.drop_newest = drop_newest
        self.transcriber = transcriber if transcriber is not None else Transcriber(device=device or "cuda")

    def detect_onsets(self, y):
        """transcribe_live.py:94-96: hop 1024, min_sep 0.3, no gates."""
        return self.transcriber.slicer.detect_onsets(y, self.sample_rate, hop_len=(256 * 4), min_sep=0.3)

    @staticmethod
    def slice_from(y: np.ndarray, i, j) -> np.ndarray:
        if len(y) < i or len(y) < j:
            return np.zeros((0,), dtype=np.float32)
        return np.array(y[i:j], dtype=np.float32)

    @staticmethod
    def pad_or_trim_audio(y: np.ndarray, target_dur: float, sr: int) -> np.ndarray:
        target_len = int(target_dur * sr)
        out_y = np.zeros(target_len, y.dtype)
        if len(y) > target_len:
            out_y = y[:target_len]
        elif len(y) < target_len:
            out_y = np.pad(y, (0, target_len - len(y)))
        return out_y

    # ---- the two halves of live() (transcribe_live.py:113-222)
    def feed(self, indata: np.ndarray):
        """Body of the audio callback (:117-123): first channel of a [frames, channels] block (or a 1-D block)."""
        a = np.asarray(indata)
        self.buffer.push((a[:, 0] if a.ndim == 2 else a).astype(np.float32))

    def step(self) -> list[dict]:
        """One pass of the main loop (:166-214) without the sleep: returns the results of the notes it transcribed
        (at most one per pass, as in the prototype; a full queue raises queue.Full there too)."""
        min_slice_len = 0.3 * self.sample_rate
        if self.buffer.is_full():
            buf = self.buffer.get_buffer()
            onsets = [int(o) for o in self.detect_onsets(buf)]
            h_idx = 0
            if len(onsets) == 1:
                s = self.slice_from(buf, onsets[0], -1)
                if len(s) > min_slice_len:
                    self.note_q.put_nowait(self.pad_or_trim_audio(s, CLIP_DURATION, self.sample_rate))
                    h_idx = onsets[0]
                    del onsets[:]
            while len(onsets) >= 2:
                s = self.slice_from(buf, onsets[0], onsets[1])
                if len(s) > min_slice_len:
                    self.note_q.put_nowait(self.pad_or_trim_audio(s, CLIP_DURATION, self.sample_rate))
                    h_idx = onsets[1]
                    del onsets[:2]
                else:
                    h_idx = onsets[0]
                    del onsets[:1]
            self.buffer.clear_from(h_idx + 1, self.drop_newest)
        out = []
        try:
            note = self.note_q.get_nowait()
            if note is not None and len(note) > 0:
                r = self.inference(np.array(note, dtype=np.float32, copy=False), self.sample_rate)
                if r is not None:
                    out.append(r)
        except queue.Empty:
            pass
        return out

    def inference(self, audio: np.ndarray, sr_in=TARGET_SR):
        """transcribe_live.py:226-267: one note -> transcribe_note -> printed (label, confidence)."""
        if audio is None or len(audio) == 0:
            print("[inference] No audio provided.")
            return None
        if audio.size < int(CLIP_DURATION * sr_in):
            return None
        result = self.transcriber.transcribe_note(audio, clip_duration=CLIP_DURATION, sr_in=sr_in)
        for i, (lab, conf) in enumerate(zip(result["labels"], result["confidences"])):
            print(f"{i:03d}  {lab:>4}  (conf={conf:.2f})")
        return result

    def live(self):
        """Microphone loop (:113-222); needs the ``sounddevice`` package, which this image does not have."""
        try:
            import sounddevice as sd
        except ImportError as e:
            raise RuntimeError("LiveTranscriber.live needs the sounddevice package; drive feed()/step() yourself") from e
        import time
        with sd.InputStream(samplerate=self.sample_rate, channels=self.channels, blocksize=self.blocksize,
                            callback=lambda indata, frames, t, status: self.feed(indata), dtype="float32"):
            print("Listening to mic...Press Ctrl+C to stop.")
            try:
                while True:
                    self.step()
                    time.sleep(0.1)
            except KeyboardInterrupt:
                print("Stopping live mic...")
