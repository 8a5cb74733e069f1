"""TEST INFRASTRUCTURE ONLY - stands in for ``soundfile`` (absent here) so the reference's own
``AudioSlicer.save_clip`` (audio/slicing.py:139-144) can run unmodified.

Restates what ``soundfile.write`` does to float32 data in a ``.wav`` (default subtype PCM_16).  python-soundfile
calls ``sf_command(SFC_SET_CLIPPING, SF_TRUE)`` on every file it opens, so libsndfile converts with its clipping
routine (src/pcm.c ``f2les_clip_array``, normalised): ``scaled = x * 2^31`` (float32), ``>= 2^31 - 1`` -> 0x7FFF,
``<= -2^31`` -> -0x8000, otherwise ``lrintf(scaled) >> 16`` - a floor onto the 16-bit grid, not ``rint(x * 0x7FFF)``
(that is the non-clipping routine, which python-soundfile never reaches).  The matching read is ``x / 0x8000``.
Neither libsndfile nor python-soundfile is available in this image: parity of this one conversion is unpinned
(file front end).
"""
import numpy as np
import scipy.io.wavfile


def float_to_pcm16(x):
    s = np.asarray(x, dtype=np.float32) * np.float32(2147483648.0)
    q = np.rint(np.clip(s, -2147483648.0, 2147483520.0)).astype(np.int64) >> 16
    q = np.where(s >= np.float32(2147483648.0), 32767, np.where(s <= np.float32(-2147483648.0), -32768, q))
    return q.astype(np.int16)


def write(file, data, samplerate, subtype=None, **_):
    if subtype not in (None, "PCM_16"):
        raise NotImplementedError(f"soundfile stand-in: subtype {subtype}")
    x = np.asarray(data, dtype=np.float32)
    q = float_to_pcm16(x)
    scipy.io.wavfile.write(str(file), int(samplerate), q)


def read(file, dtype="float64", always_2d=False, **_):
    sr, data = scipy.io.wavfile.read(str(file))
    if data.dtype == np.int16:
        out = data.astype(dtype) / 32768.0
    elif data.dtype == np.int32:
        out = data.astype(dtype) / 2147483648.0
    else:
        out = data.astype(dtype)
    if always_2d and out.ndim == 1:
        out = out[:, None]
    return out, sr
