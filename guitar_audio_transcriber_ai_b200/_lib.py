"""ctypes binding of include/gat.h (csrc/libgat.so).

The extension is loaded through PyTorch (``torch.ops.load_library`` registers the shared object with the
process and CUDA context torch owns); symbols are then bound with ctypes.  There is NO fallback: if the
library is missing ``load()`` raises, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
CSRC = _HERE / "csrc"
LIB_PATH = CSRC / "libgat.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

GAT_FLAG_YIN_ON_NORMALIZED = 1
GAT_FLAG_APPLY_SCALER = 2
GAT_FLAG_SKIP_MLP = 4
GAT_FLAG_NO_PITCH = 8
GAT_FLAG_NO_NORMALIZE_MFCC = 16
GAT_FLAG_NO_NORMALIZE_MEL = 32
GAT_MEL_NORMALIZE = 1
GAT_MEL_POWER = 2


class GatConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32), ("mel_n_fft", C.c_int32), ("mel_hop", C.c_int32), ("mel_n_mels", C.c_int32),
        ("mel_window", C.c_void_p), ("mel_fb", C.c_void_p),
        ("mfcc_n_mels", C.c_int32), ("mfcc_n_mfcc", C.c_int32),
        ("stft_window", C.c_void_p), ("mfcc_fb", C.c_void_p), ("dct", C.c_void_p),
        ("yin_fmin", C.c_double), ("yin_fmax", C.c_double), ("yin_trough_threshold", C.c_double),
    ]


class GatSlicerParams(C.Structure):
    _fields_ = [
        ("min_db_threshold", C.c_double), ("sample_gate", C.c_float), ("rms_hop", C.c_int32),
        ("p20_k", C.c_int32), ("p20_gamma", C.c_float), ("gate_offset_db", C.c_float), ("onset_hop", C.c_int32),
        ("pre_max", C.c_int32), ("post_max", C.c_int32), ("pre_avg", C.c_int32), ("post_avg", C.c_int32),
        ("wait", C.c_int32), ("delta", C.c_float),
        ("min_sep_samples", C.c_int64), ("attack_skip", C.c_int64), ("clip_len", C.c_int64),
        ("min_slice_rms_db", C.c_float),
    ]


_P = C.c_void_p
_PROTOTYPES = {
    "gat_last_error": (C.c_char_p, []),
    "gat_version": (C.c_int, []),
    "gat_ctx_create": (C.c_int, [C.POINTER(GatConfig), C.c_int, C.POINTER(_P)]),
    "gat_ctx_destroy": (None, [_P]),
    "gat_load_mlp": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int64]),
    "gat_load_cnn": (C.c_int, [_P, C.c_int32, _P, C.POINTER(_P), C.POINTER(_P), C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "gat_set_scaler": (C.c_int, [_P, _P, _P, C.c_int32]),
    "gat_set_ensemble_weights": (C.c_int, [_P, C.c_float, C.c_float]),
    "gat_melspec_db": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "gat_mfcc_features": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P,
                                    C.c_int32, _P, _P]),
    "gat_yin": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P]),
    "gat_infer": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gat_segment": (C.c_int, [_P, _P, C.c_int64, C.POINTER(GatSlicerParams), C.c_int32, _P, _P, _P, _P, _P, _P, _P,
                              _P, _P, _P]),
    "gat_segment_batch": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.POINTER(GatSlicerParams), C.c_int32, _P, _P, _P, C.c_int64,
                                    _P, _P, _P]),
    "gat_transcribe_clips": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gat_transcribe_clips_host": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P]),
    "gat_transcribe_clips_host_pcm16": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P]),
    "gat_detect_onsets": (C.c_int, [_P, _P, C.c_int64, C.POINTER(GatSlicerParams), C.c_int32, _P, _P, _P]),
    "gat_pcm16_roundtrip": (C.c_int, [_P, _P, C.c_int64, _P]),
    "gat_decode_mono": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, _P, _P]),
    "gat_resample": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, _P, C.c_int64, _P]),
    "gat_profile_begin": (C.c_int, [_P]),
    "gat_profile_end": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "gat_debug_tc_counters": (C.c_int, [_P, _P, C.c_int64]),
    "gat_set_conv_pass": (C.c_int, [_P, C.c_int32]),
    "gat_set_host_chunks": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "gat_debug_fma_peak": (C.c_int, [_P, C.c_int32, _P]),
    "gat_launch_count": (C.c_int64, [_P]),
    "gat_num_classes": (C.c_int32, [_P]),
    "gat_mel_frames": (C.c_int32, [_P, C.c_int64]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


class GatError(RuntimeError):
    pass


class GatLib:
    """Thin typed view over the shared object."""

    def __init__(self, path):
        self.path = pathlib.Path(path)
        self.cdll = C.CDLL(str(self.path))
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(self.cdll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc: int, exc=GatError):
        if rc != 0:
            raise exc(self.gat_last_error().decode("utf-8", "replace"))


_LIB: GatLib | None = None


def build(verbose: bool = False) -> pathlib.Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... csrc/gat.cu -> csrc/libgat.so (in-tree)."""
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [_HERE.parent / "include" / "gat.h"]
    if LIB_PATH.exists() and all(LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, str(CSRC / "gat.cu"), "-o", str(LIB_PATH)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB_PATH


def load() -> GatLib:
    """Loads csrc/libgat.so through PyTorch.  Raises if the extension has not been built."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise GatError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        import torch
        torch.ops.load_library(str(LIB_PATH))
        _LIB = GatLib(LIB_PATH)
    return _LIB


def ptr(x):
    """Address of a torch tensor / numpy array / None as c_void_p."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())
