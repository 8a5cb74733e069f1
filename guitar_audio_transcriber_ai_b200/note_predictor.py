"""``NotePredictor`` - drop-in for the reference's note_predictor.py:15-157.

``load_models`` builds the same ``MLP`` / ``CNN`` modules from the checkpoint dicts (so ``.mlp`` / ``.cnn``
and ``reverse_map`` look as they do in the reference) and uploads packed weights to the GPU context;
``predict`` runs both forward passes, the softmaxes, the 0.2/0.8 ensemble and the argmax in csrc/infer.cuh.
"""
from __future__ import annotations

import numpy as np
import torch

from .dsp.yin import shared_engine
from .training.cnn_trainer import CNN
from .training.mlp_trainer import MLP


class NotePredictor:
    def __init__(self, device=None):
        self.device = torch.device(device or "cuda")
        self.mlp = None
        self.cnn = None
        self.reverse_map = None
        self.cnn_weight = 0.80
        self.mlp_weight = (1.0 - self.cnn_weight)
        self.engine = None
        self._sr = None
        self._cfgs = (None, None)

    def bind_engine(self, engine):
        """Attach the Engine whose context receives the weights (Transcriber does this with the checkpoint's
        sample rate and feature configs)."""
        self.engine = engine
        if self.mlp is not None:
            engine.load_mlp(self.mlp.state_dict())
        if self.cnn is not None:
            engine.load_cnn(self.cnn.state_dict())

    def load_models(self, mlp_ckpt_data: dict = None, cnn_ckpt_data: dict = None):
        """note_predictor.py:29-80."""
        if mlp_ckpt_data is not None:
            self.mlp = MLP(**mlp_ckpt_data["model_init_args"])
            if "model" not in mlp_ckpt_data:
                raise KeyError("[load_models] MLP checkpoint missing 'model' field")
            self.mlp.load_state_dict(mlp_ckpt_data["model"])
            self.mlp.eval()
            if self.reverse_map is None and mlp_ckpt_data.get("reverse_map") is not None:
                self.reverse_map = mlp_ckpt_data["reverse_map"]
        if cnn_ckpt_data is not None:
            self.cnn = CNN(**cnn_ckpt_data["model_init_args"])
            if "model" not in cnn_ckpt_data:
                raise KeyError("[load_models] CNN checkpoint missing 'model' field")
            self.cnn.load_state_dict(cnn_ckpt_data["model"])
            self.cnn.eval()
        if self.engine is None:
            cfgs = [d.get("config") if d else None for d in (mlp_ckpt_data, cnn_ckpt_data)]
            sr = next((c["target_sr"] for c in cfgs if c), 22050)
            mel = cfgs[1]["features"]["params"] if cfgs[1] else None
            mf = cfgs[0]["features"]["params"] if cfgs[0] else None
            mel = {k: mel[k] for k in ("N_MELS", "N_FFT", "HOP_LENGTH")} if mel else None
            mf = {"N_MFCC": mf["N_MFCC"]} if mf else None
            self.engine = shared_engine(sr, self.device, mel, mf)
        self.bind_engine(self.engine)

    def predict(self, mfcc_features=None, melspec_features=None):
        """note_predictor.py:84-135.  Like the reference, BOTH feature sets are needed (it reads an unassigned
        local otherwise); the error raised here is the explicit form of that."""
        if mfcc_features is None and melspec_features is None:
            raise ValueError("[predict] Must provide either mfcc_features or melspec_features")
        if mfcc_features is None or melspec_features is None:
            raise UnboundLocalError("[predict] both mfcc_features and melspec_features are required "
                                    "(note_predictor.py:110 reads both branches' results)")
        if not torch.is_tensor(mfcc_features):
            mfcc_features = np.asarray(mfcc_features, np.float32)
        if not torch.is_tensor(melspec_features):
            melspec_features = np.asarray(melspec_features, np.float32)
        self.engine.set_ensemble_weights(self.mlp_weight, self.cnn_weight)
        out = self.engine.infer(mfcc_features, melspec_features)
        return self._result(out)

    def _result(self, out: dict) -> dict:
        idx = out["indices"].cpu().numpy()
        if self.reverse_map is None:
            raise RuntimeError("[predict] reverse_map is not set")
        return {
            "indices": idx,
            "labels": [self.reverse_map[int(i)] for i in idx],
            "confidences": out["confidences"].cpu().numpy(),
            "probs": out["probs"].cpu().numpy(),
            "per_model_probs": {
                "mlp": out["mlp_probs"].cpu().numpy() if out.get("mlp_probs") is not None else None,
                "cnn": out["cnn_probs"].cpu().numpy(),
            },
        }

    def predict_debug(self, test_weights, mfcc_features=None, melspec_features=None):
        """note_predictor.py:138-157: sweep the CNN weight, restore it afterwards."""
        predictions = []
        cnn_weight, mlp_weight = self.cnn_weight, self.mlp_weight
        for weight in test_weights:
            self.cnn_weight = weight
            self.mlp_weight = 1 - weight
            prediction = self.predict(mfcc_features=mfcc_features, melspec_features=melspec_features)
            predictions.append((weight, prediction))
            print("weight: ", weight)
            print(prediction["labels"], prediction["confidences"])
            print()
        self.cnn_weight, self.mlp_weight = cnn_weight, mlp_weight
        return predictions
