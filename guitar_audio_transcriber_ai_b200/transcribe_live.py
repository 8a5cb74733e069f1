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
        """transcribe_live.py:98-102: ``y[i:j]`` as float32, empty when an index lies past the end."""
        n = len(y)
        return np.zeros((0,), dtype=np.float32) if (n < i or n < j) else np.array(y[i:j], dtype=np.float32)

    @staticmethod
    def pad_or_trim_audio(y: np.ndarray, target_dur: float, sr: int) -> np.ndarray:
        """transcribe_live.py:104-113: exactly ``int(target_dur * sr)`` samples, right zero padding."""
        want = int(target_dur * sr)
        if len(y) >= want:
            return y[:want]
        out = np.zeros(want, y.dtype)
        out[:len(y)] = y
        return out

    # ---- the two halves of live() (transcribe_live.py:113-222)
    def feed(self, indata: np.ndarray):
        """Body of the audio callback (:117-123): first channel of a [frames, channels] block (or a 1-D block)."""
        block = np.asarray(indata)
        self.buffer.push((block[:, 0] if block.ndim == 2 else block).astype(np.float32))

    def _queue_note(self, piece: np.ndarray) -> bool:
        """A slice becomes a note when it is longer than 0.3 s (:158-159); queue.Full propagates as in the prototype."""
        if len(piece) <= 0.3 * self.sample_rate:
            return False
        self.note_q.put_nowait(self.pad_or_trim_audio(piece, CLIP_DURATION, self.sample_rate))
        return True

    def step(self) -> list[dict]:
        """One pass of the main loop (:166-214) without the sleep: when the ring buffer is full, cut notes between
        consecutive onsets, drop the consumed audio, then transcribe at most one queued note."""
        if self.buffer.is_full():
            snapshot = self.buffer.get_buffer()
            marks = [int(o) for o in self.detect_onsets(snapshot)]
            consumed = 0
            if len(marks) == 1:                        # a single onset: everything after it but the last sample (:176)
                if self._queue_note(self.slice_from(snapshot, marks[0], -1)):
                    consumed, marks = marks[0], []
            k = 0
            while len(marks) - k >= 2:                 # pairs of onsets; a short slice only advances by one onset (:183-192)
                if self._queue_note(self.slice_from(snapshot, marks[k], marks[k + 1])):
                    consumed = marks[k + 1]
                    k += 2
                else:
                    consumed = marks[k]
                    k += 1
            self.buffer.clear_from(consumed + 1, self.drop_newest)
        results = []
        if not self.note_q.empty():
            note = self.note_q.get_nowait()
            if note is not None and len(note) > 0:
                res = self.inference(np.array(note, dtype=np.float32, copy=False), self.sample_rate)
                if res is not None:
                    results.append(res)
        return results

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
