"""Checkpoint reading/writing in the reference's schema (model-loading layer).

Schema written by the reference's trainers (prototyping/source/training/mlp_trainer.py:444-478,
cnn_trainer.py:479-509): ``meta``, ``config{features{type,params}, model{type,params}, target_sr,
clip_length}``, ``model`` (state_dict), ``model_init_args``, ``optimizer``, ``device``, four history
lists, ``epoch``, ``reverse_map``, ``num_classes``, ``class_names`` and, for the MLP, ``scaler``.

The shipped MLP checkpoint pickles a ``pathlib.WindowsPath`` (its config holds CHECKPOINTS_DIR), which
plain ``torch.load`` cannot instantiate on POSIX; ``load_checkpoint`` remaps it inside the unpickler.
"""
from __future__ import annotations

import pathlib
import pickle
import warnings
from datetime import datetime

import numpy as np
import torch

from .config import CONFIG_VERSION


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("pathlib") and name in ("WindowsPath", "PureWindowsPath"):
            return pathlib.PureWindowsPath
        if module.startswith("pathlib") and name == "PosixPath" and not hasattr(pathlib, "PosixPath"):
            return pathlib.PurePosixPath
        return super().find_class(module, name)


class _PickleModule:
    """Duck-typed ``pickle_module`` for torch.load."""
    __name__ = "pickle"
    Unpickler = _Unpickler

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()

    def __getattr__(self, k):
        return getattr(pickle, k)


def load_checkpoint(path, map_location="cpu") -> dict:
    """``torch.load(path, map_location, weights_only=False)`` as transcribe.py:57-60 does, POSIX-safe."""
    path = pathlib.Path(path)
    if not path.is_file():
        raise FileNotFoundError(f"Missing checkpoint: {path}")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # sklearn InconsistentVersionWarning for the pickled scaler
        return torch.load(path, map_location=map_location, weights_only=False, pickle_module=_PickleModule())


def make_checkpoint(model, model_type: str, feature_params: dict, model_params: dict, target_sr: int,
                    clip_length: float, reverse_map: dict, scaler=None, histories=None) -> dict:
    """Builds a dict in the reference trainers' schema (used by the synthetic-weights generator)."""
    histories = histories or {}
    class_names = [str(reverse_map[i]) for i in range(len(reverse_map))]
    ckpt = {
        "meta": {"config_version": CONFIG_VERSION, "datetime": datetime.now().strftime("%d/%m/%Y %H:%M:%S"),
                 "model_type": model_type},
        "config": {
            "features": {"type": "mfcc" if model_type == "mlp" else "melspec", "params": dict(feature_params)},
            "model": {"type": model_type, "params": dict(model_params)},
            "target_sr": int(target_sr),
            "clip_length": float(clip_length),
        },
        "model": model.state_dict(),
        "model_init_args": dict(model.init_args),
        "optimizer": {},
        "device": "cpu",
        "train_loss_history": list(histories.get("train_loss", [])),
        "train_accuracy_history": list(histories.get("train_acc", [])),
        "val_loss_history": list(histories.get("val_loss", [])),
        "val_accuracy_history": list(histories.get("val_acc", [])),
        "epoch": int(histories.get("epoch", 0)),
        "reverse_map": {int(k): np.str_(v) for k, v in reverse_map.items()},
        "num_classes": len(reverse_map),
        "class_names": class_names,
    }
    if model_type == "mlp":
        ckpt["scaler"] = scaler
    return ckpt


# ----------------------------------------------------------------------------- weight packing for csrc/
def _f32(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().to(torch.float32).numpy())


def pack_mlp(state: dict) -> dict:
    """Flattens an MLP state_dict (keys ``net.<i>.weight|bias``) into the arrays gat_load_mlp takes.

    Layer order: Linear, LayerNorm, (LeakyReLU, Dropout)... ; 2-d weights are Linear, 1-d are LayerNorm.
    Linear weights are stored transposed ([in][out]) so a warp reads consecutive outputs.
    """
    idx = sorted({int(k.split(".")[1]) for k in state if k.startswith("net.")})
    linear = [i for i in idx if state[f"net.{i}.weight"].ndim == 2]
    dims = [int(state[f"net.{linear[0]}.weight"].shape[1])] + [int(state[f"net.{i}.weight"].shape[0]) for i in linear]
    flat = []
    for j, i in enumerate(linear):
        flat += [np.ascontiguousarray(_f32(state[f"net.{i}.weight"]).T).ravel(), _f32(state[f"net.{i}.bias"]).ravel()]
        if j + 1 < len(linear):
            flat += [_f32(state[f"net.{i + 1}.weight"]).ravel(), _f32(state[f"net.{i + 1}.bias"]).ravel()]
    return {"dims": np.asarray(dims, dtype=np.int32), "params": np.concatenate(flat).astype(np.float32)}


def pack_cnn(state: dict, bn_eps: float = 1e-5) -> dict:
    """Folds eval-mode BatchNorm into each conv (w' = w*g/sqrt(v+eps), b' = (b-m)*g/sqrt(v+eps)+beta) in
    float64 and lays the weights out as the kernels want them: conv ``[tap(ky,kx)][c_in][c_out]``,
    fully-connected ``[in][out]``.  Accepts either key family of the aliased state dict."""
    if not any(k.startswith("features.") for k in state):
        state = {k.replace("net.0.", "features.", 1).replace("net.1.", "classifier.", 1): v for k, v in state.items()}
    conv_idx = sorted({int(k.split(".")[1]) for k in state
                       if k.startswith("features.") and k.endswith(".weight") and state[k].ndim == 4})
    convs = []
    for i in conv_idx:
        w = state[f"features.{i}.weight"].detach().cpu().double().numpy()
        b = state[f"features.{i}.bias"].detach().cpu().double().numpy()
        if f"features.{i + 1}.running_mean" in state:
            g = state[f"features.{i + 1}.weight"].detach().cpu().double().numpy()
            beta = state[f"features.{i + 1}.bias"].detach().cpu().double().numpy()
            mu = state[f"features.{i + 1}.running_mean"].detach().cpu().double().numpy()
            var = state[f"features.{i + 1}.running_var"].detach().cpu().double().numpy()
            s = g / np.sqrt(var + bn_eps)
            w = w * s[:, None, None, None]
            b = (b - mu) * s + beta
        c_out, c_in, kh, kw = w.shape
        convs.append({"w": np.ascontiguousarray(w.transpose(2, 3, 1, 0).reshape(kh * kw, c_in, c_out)).astype(np.float32),
                      "b": b.astype(np.float32), "c_in": c_in, "c_out": c_out, "k": kh})
    fc_idx = sorted({int(k.split(".")[1]) for k in state if k.startswith("classifier.") and k.endswith(".weight")})
    fcs = [{"w": np.ascontiguousarray(_f32(state[f"classifier.{i}.weight"]).T), "b": _f32(state[f"classifier.{i}.bias"])}
           for i in fc_idx]
    return {"convs": convs, "fcs": fcs}
