"""TEST INFRASTRUCTURE ONLY - stands in for ``soundfile`` (absent here) so the reference's own
``AudioSlicer.save_clip`` (audio/slicing.py:139-144) can run unmodified.

Restates libsndfile's normalised float -> PCM_16 conversion for ``.wav`` (default subtype PCM_16;
src/pcm.c f2s_array: ``lrintf(x * 0x7FFF)``, no clipping unless SFC_SET_CLIPPING - values are clipped here,
the reference's clips never leave [-1, 1]) and the matching read (``x / 0x8000``).  libsndfile itself is not
available in this image: parity of this one conversion is unpinned (file front end).
"""
import numpy as np
import scipy.io.wavfile


def write(file, data, samplerate, subtype=None, **_):
    if subtype not in (None, "PCM_16"):
        raise NotImplementedError(f"soundfile stand-in: subtype {subtype}")
    x = np.asarray(data, dtype=np.float32)
    q = np.clip(np.rint(x * np.float32(32767.0)), -32768, 32767).astype(np.int16)
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
