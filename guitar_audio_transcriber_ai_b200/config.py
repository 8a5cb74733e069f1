"""Configuration surface of the reference's ``config.py`` (version_1/source/config.py:10-118).

Checkpoints store ``asdict()`` copies of these dataclasses and the inference path reads them back by key
(transcribe.py:126-127, :190-191), so the class names, field names, field order and defaults below are part of the
file format.  They are declared as data (one table per class) and turned into frozen dataclasses at import time.
"""
from __future__ import annotations

from dataclasses import asdict, field, make_dataclass  # noqa: F401  (asdict is re-exported like the reference)
from pathlib import Path

CONFIG_VERSION = "1.0.0"

PROJECT_ROOT = Path(__file__).resolve().parent.parent
CHECKPOINTS_ROOT = PROJECT_ROOT / "data" / "checkpoints"
INFERENCE_ROOT = PROJECT_ROOT / "data" / "inference"
INFERENCE_OUTPUT_ROOT = INFERENCE_ROOT / "output"

TARGET_SR = 22050          # 2 x 11025 (config.py:29)
CLIP_DURATION = 0.50

_TRAINING = (("LR", float, 1e-3), ("DECAY", float, 1e-4))
_EARLY_STOP = (("MAX_CLIP_NORM", float, 1.0), ("ES_WINDOW_LEN", int, 4), ("ES_SLOPE_LIMIT", float, -0.00015))

_SPECS = {
    "MFCCConfig": (
        ("N_MFCC", int, 64), ("BATCH_SIZE", int, 32), ("STANDARD_SCALER", bool, True),
        ("NORMALIZE_AUDIO_VOLUME", bool, True), ("ADD_PITCH_FEATURES", bool, True)),
    "MelSpecConfig": (
        ("N_MELS", int, 64), ("N_FFT", int, 2048), ("HOP_LENGTH", int, 256), ("BATCH_SIZE", int, 32),
        ("NORMALIZE_AUDIO_VOLUME", bool, True), ("TO_DB", bool, True)),
    "MLPConfig": (
        ("CHECKPOINTS_DIR", Path, CHECKPOINTS_ROOT / "mlp"), ("DEFAULT_CKPT_NAME", str, f"mlp_v{CONFIG_VERSION}.ckpt"),
        ("SAVE_CHECKPOINT", bool, True), ("HIDDEN_DIM", int, 128), ("NUM_HIDDEN_LAYERS", int, 2), ("DROPOUT", float, 0.1),
        *_TRAINING, ("EPOCHS", int, 10), *_EARLY_STOP),
    "CNNConfig": (
        ("CHECKPOINTS_DIR", Path, CHECKPOINTS_ROOT / "cnn"), ("DEFAULT_CKPT_NAME", str, f"cnn_v{CONFIG_VERSION}.ckpt"),
        ("SAVE_CHECKPOINT", bool, True), ("BASE_CHANNELS", int, 32), ("NUM_BLOCKS", int, 3), ("KERNEL_SIZE", int, 3),
        ("HIDDEN_DIM", int, 256), ("DROPOUT", float, 0.1), *_TRAINING, ("EPOCHS", int, 3), *_EARLY_STOP,
        ("USE_AMP", bool, True)),
    "AudioSlicerConfig": (
        ("MIN_IN_DB_THRESHOLD", float, -32.5), ("MIN_SLICE_RMS_DB", float, -37.0), ("HOP_LEN", int, 512),
        ("MIN_SEP", float, 0.3)),
}


def _declare(name: str):
    cls = make_dataclass(name, [(f, t, field(default=d)) for f, t, d in _SPECS[name]], frozen=True)
    cls.__module__ = __name__          # so that instances pickle / unpickle under this module
    return cls


MFCCConfig = _declare("MFCCConfig")
MelSpecConfig = _declare("MelSpecConfig")
MLPConfig = _declare("MLPConfig")
CNNConfig = _declare("CNNConfig")
AudioSlicerConfig = _declare("AudioSlicerConfig")
# the reference declares this one without an annotation, so it is a class attribute, not a dataclass field
AudioSlicerConfig.ATTACK_SKIP_SEC = 0.1

MFCC_CONFIG = MFCCConfig()
MELSPEC_CONFIG = MelSpecConfig()
MLP_CONFIG = MLPConfig()
CNN_CONFIG = CNNConfig()
SLICER_CONFIG = AudioSlicerConfig()
