"""``NotePredictor`` - drop-in for the reference's note_predictor.py:15-157.

``load_models`` builds the same ``MLP`` / ``CNN`` modules from the checkpoint dicts (so ``.mlp`` / ``.cnn``
and ``reverse_map`` look as they do in the reference) and uploads packed weights to the GPU context;
``predict`` runs both forward passes, the softmaxes, the 0.2/0.8 ensemble and the argmax in csrc/infer.cuh.
"""
from __future__ import annotations

import numpy as np
import torch

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

    @staticmethod
    def _module_from(ckpt: dict, cls, tag: str):
        """Rebuild a trainer module from a checkpoint dict: ``model_init_args`` -> constructor, ``model`` -> weights."""
        module = cls(**ckpt["model_init_args"])
        if "model" not in ckpt:
            raise KeyError(f"[load_models] {tag} checkpoint missing 'model' field")
        module.load_state_dict(ckpt["model"])
        return module.eval()

    def load_models(self, mlp_ckpt_data: dict = None, cnn_ckpt_data: dict = None):
        """note_predictor.py:29-80: either checkpoint may be omitted; the label map comes from the MLP checkpoint."""
        if mlp_ckpt_data is not None:
            self.mlp = self._module_from(mlp_ckpt_data, MLP, "MLP")
            labels = mlp_ckpt_data.get("reverse_map")
            if self.reverse_map is None and labels is not None:
                self.reverse_map = labels
        if cnn_ckpt_data is not None:
            self.cnn = self._module_from(cnn_ckpt_data, CNN, "CNN")
        if self.engine is None:
            self.engine = self._engine_for(mlp_ckpt_data, cnn_ckpt_data)
        self.bind_engine(self.engine)

    def _engine_for(self, mlp_ckpt_data, cnn_ckpt_data):
        """A context at the checkpoints' sample rate and feature sizes when no Transcriber supplied one."""
        mlp_cfg = (mlp_ckpt_data or {}).get("config")
        cnn_cfg = (cnn_ckpt_data or {}).get("config")
        sr = (mlp_cfg or cnn_cfg or {"target_sr": 22050})["target_sr"]
        mel = mf = None
        if cnn_cfg:
            params = cnn_cfg["features"]["params"]
            mel = {k: params[k] for k in ("N_MELS", "N_FFT", "HOP_LENGTH")}
        if mlp_cfg:
            mf = {"N_MFCC": mlp_cfg["features"]["params"]["N_MFCC"]}
        # A predictor owns its context: the weights live in the gat_ctx, so two predictors (an A/B comparison of
        # checkpoints, say) must not share one.  The stateless helpers (AudioSlicer, YinDsp, MelFeatureBuilder) keep
        # using the process-wide table-only engines of dsp.yin.shared_engine.
        from .engine import Engine
        return Engine(sr, mel, mf, device=self.device)

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
        """note_predictor.py:138-157: one prediction per candidate CNN weight (MLP weight = 1 - w), printed as the
        reference prints them; the configured weights are restored afterwards."""
        saved = (self.cnn_weight, self.mlp_weight)
        sweep = []
        try:
            for w in test_weights:
                self.cnn_weight, self.mlp_weight = w, 1 - w
                res = self.predict(mfcc_features=mfcc_features, melspec_features=melspec_features)
                sweep.append((w, res))
                print("weight: ", w)
                print(res["labels"], res["confidences"])
                print()
        finally:
            self.cnn_weight, self.mlp_weight = saved
        return sweep
