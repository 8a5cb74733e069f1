"""B200-native implementation of the guitar-audio-transcriber-ai version_1 transcription hot path.

Python surface mirrors the reference (``Transcriber``, ``NotePredictor``, ``MelFeatureBuilder``,
``YinDsp``, ``AudioSlicer``); all arithmetic runs in the sm_100a CUDA extension ``csrc/libgat.so``
behind the C ABI declared in ``include/gat.h``.  There is no CPU fallback: using any compute entry
point without the extension or without a CUDA device raises.
"""
__version__ = "0.1.0"

_LAZY = {
    "Transcriber": ("transcribe", "Transcriber"),
    "NotePredictor": ("note_predictor", "NotePredictor"),
    "MelFeatureBuilder": ("audio.features", "MelFeatureBuilder"),
    "AudioSlicer": ("audio.slicing", "AudioSlicer"),
    "YinDsp": ("dsp.yin", "YinDsp"),
    "Engine": ("engine", "Engine"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
