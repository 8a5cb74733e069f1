"""Minimal RIFF/WAVE reader and PCM_16 writer for the file front end.

The reference decodes with libsndfile through ``librosa.load`` (audio/slicing.py:25, audio/loading.py:85) and
writes clips with ``soundfile.write`` (audio/slicing.py:144, PCM_16 for ``.wav``).  Neither library is needed
for plain WAV files: this module parses the container on the host and hands the raw interleaved frames to the
GPU, which does the sample conversion and the channel mean (csrc/frontend.cuh).

Supported: PCM 8/16/24/32-bit integer and IEEE float 32/64-bit, any channel count, WAVE_FORMAT_EXTENSIBLE.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


def read_wav_frames(path) -> tuple[np.ndarray, int]:
    """Raw interleaved frames ``[frames, channels]`` and the sample rate.

    dtype int16 and float32 are returned as stored (the GPU converts them); other encodings are converted to
    float32 here with libsndfile's scaling (8-bit: (x-128)/128, 24-bit: x/2^23, 32-bit int: x/2^31)."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"[load_audio] File not found at: {path}")
    data = path.read_bytes()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, body = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        start = pos + 8
        if cid == b"fmt ":
            fmt = data[start:start + size]
        elif cid == b"data":
            body = data[start:min(start + size, len(data))]
            break
        pos = start + size + (size & 1)
    if fmt is None or body is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, channels, rate, _, block_align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == _EXTENSIBLE and len(fmt) >= 26:
        tag = struct.unpack_from("<H", fmt, 24)[0]
    if channels < 1:
        raise ValueError(f"{path}: no channels")
    width = bits // 8
    n = len(body) // (width * channels)
    raw = np.frombuffer(body, dtype=np.uint8, count=n * width * channels)
    if tag == _PCM and bits == 16:
        out = raw.view("<i2").reshape(n, channels)
    elif tag == _FLOAT and bits == 32:
        out = raw.view("<f4").reshape(n, channels)
    elif tag == _FLOAT and bits == 64:
        out = raw.view("<f8").reshape(n, channels).astype(np.float32)
    elif tag == _PCM and bits == 8:
        out = ((raw.astype(np.float32) - 128.0) / 128.0).reshape(n, channels)
    elif tag == _PCM and bits == 24:
        b = raw.reshape(n * channels, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        out = (v.astype(np.float32) / 8388608.0).reshape(n, channels)
    elif tag == _PCM and bits == 32:
        out = (raw.view("<i4").astype(np.float64) / 2147483648.0).astype(np.float32).reshape(n, channels)
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (format tag {tag}, {bits} bits)")
    return np.ascontiguousarray(out), int(rate)


def float_to_pcm16(x) -> np.ndarray:
    """What ``soundfile.write`` stores for float32 samples in a PCM_16 file.  python-soundfile enables
    SFC_SET_CLIPPING on every file, so libsndfile runs its clipping converter (src/pcm.c f2les_clip_array):
    ``scaled = x * 2^31`` in float32; ``>= 2^31 - 1`` -> 0x7FFF; ``<= -2^31`` -> -0x8000; else
    ``lrintf(scaled) >> 16`` (a floor to the 16-bit grid, not a round).  Same arithmetic as
    ``pcm16_quantize`` in csrc/frontend.cuh."""
    s = np.asarray(x, dtype=np.float32) * np.float32(2147483648.0)
    with np.errstate(invalid="ignore"):
        q = np.rint(np.clip(s, -2147483648.0, 2147483520.0)).astype(np.int64) >> 16
    q = np.where(s >= np.float32(2147483648.0), 32767, np.where(s <= np.float32(-2147483648.0), -32768, q))
    return q.astype(np.int16)


def write_wav_pcm16(path, samples_i16: np.ndarray, sample_rate: int) -> None:
    """Mono or interleaved int16 samples -> canonical 44-byte-header WAV (what sf.write produces for .wav)."""
    a = np.ascontiguousarray(samples_i16, dtype="<i2")
    channels = 1 if a.ndim == 1 else a.shape[1]
    body = a.tobytes()
    header = b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, _PCM, channels, int(sample_rate), int(sample_rate) * channels * 2, channels * 2, 16
    ) + b"data" + struct.pack("<I", len(body))
    Path(path).write_bytes(header + body)
