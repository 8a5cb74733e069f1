"""TEST INFRASTRUCTURE ONLY - run the reference's own version_1 files, unmodified, in this container.

The reference is pure Python but imports modules that are absent here (librosa,
soundfile, matplotlib, tkinter).  ``install()`` registers oracle/librosa_shim as ``librosa`` plus inert
stand-ins for the others, puts /root/reference/version_1/source on sys.path and returns the reference's
modules.  Used ONLY by oracle/make_golden.py and tests that are skipped when /root/reference is absent
(it does not exist on the GPU box); it pins oracle/port.py to the reference's real control flow.
"""
from __future__ import annotations

import importlib
import os
import pathlib
import pickle
import sys
import types

REFERENCE_ROOT = pathlib.Path(os.environ.get("GAT_REFERENCE_ROOT", "/root/reference"))
V1_SOURCE = REFERENCE_ROOT / "version_1" / "source"
_ORACLE_DIR = pathlib.Path(__file__).resolve().parent


def available() -> bool:
    return (V1_SOURCE / "transcribe.py").is_file()


def _stub(name: str, **attrs) -> types.ModuleType:
    import importlib.machinery
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave like a package so "import a.b" works
    mod.__spec__ = importlib.machinery.ModuleSpec(name, None)  # other libraries probe find_spec(name)
    sys.modules[name] = mod
    return mod


def install():
    """Returns a namespace with the reference's modules (config, features, slicing, yin, ...)."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    if str(_ORACLE_DIR) not in sys.path:
        sys.path.insert(0, str(_ORACLE_DIR))
    shim = importlib.import_module("librosa_shim")
    sys.modules["librosa"] = shim
    for sub in ("feature", "onset", "util", "filters", "display", "core"):
        sys.modules[f"librosa.{sub}"] = getattr(shim, sub) if hasattr(shim, sub) else importlib.import_module(
            f"librosa_shim.{sub}")
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            mpl = _stub("matplotlib")
            mpl.pyplot = _stub("matplotlib.pyplot")
    if "soundfile" not in sys.modules:
        try:
            importlib.import_module("soundfile")
        except Exception:
            sys.modules["soundfile"] = importlib.import_module("soundfile_standin")
    try:
        importlib.import_module("tkinter")
    except Exception:
        tk = _stub("tkinter")
        tk.filedialog = _stub("tkinter.filedialog")
        tk.messagebox = _stub("tkinter.messagebox")
    if str(V1_SOURCE) not in sys.path:
        sys.path.insert(0, str(V1_SOURCE))
    ns = types.SimpleNamespace()
    ns.config = importlib.import_module("config")
    ns.loading = importlib.import_module("audio.loading")
    ns.yin = importlib.import_module("dsp.yin")
    ns.features = importlib.import_module("audio.features")
    ns.slicing = importlib.import_module("audio.slicing")
    ns.mlp_trainer = importlib.import_module("training.mlp_trainer")
    ns.cnn_trainer = importlib.import_module("training.cnn_trainer")
    ns.note_predictor = importlib.import_module("note_predictor")
    ns.transcribe = importlib.import_module("transcribe")
    return ns


class _PosixSafeUnpickler(pickle.Unpickler):
    """The shipped MLP checkpoint pickles a pathlib.WindowsPath; map it to a pure path."""

    def find_class(self, module, name):
        if module.startswith("pathlib") and name in ("WindowsPath", "PureWindowsPath"):
            return pathlib.PureWindowsPath
        return super().find_class(module, name)


class _PickleModule:
    __name__ = "pickle"
    Unpickler = _PosixSafeUnpickler

    @staticmethod
    def load(f, **kw):
        return _PosixSafeUnpickler(f, **kw).load()

    def __getattr__(self, k):
        return getattr(pickle, k)


def load_ckpt(path):
    """torch.load(weights_only=False) that survives WindowsPath on POSIX (reference: transcribe.py:58)."""
    import torch
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PickleModule())


class MemoryLoader:
    """Stands in for audio.loading.AudioDatasetLoader (loading.py:36-105) without touching disk:
    exposes ``target_sr`` and ``load_audio_dataset(pad_to_max)`` over in-memory clips."""

    def __init__(self, clips, target_sr, labels=None):
        self.clips = [c for c in clips]
        self.target_sr = target_sr
        self.labels = list(labels) if labels is not None else ["x"] * len(self.clips)

    def load_audio_dataset(self, pad_to_max=True):
        import numpy as np
        wavs = list(self.clips)
        if len(wavs) == 0:
            raise FileNotFoundError("load_audio_dataset: No audio files found.")
        if pad_to_max:
            m = max(len(w) for w in wavs)
            wavs = [np.pad(w, (0, m - len(w)), mode="constant") for w in wavs]
        return wavs, [self.target_sr] * len(wavs), list(self.labels), [f"mem://{i}" for i in range(len(wavs))]
