"""``MelFeatureBuilder`` - drop-in for the inference half of the reference's audio/features.py.

Same signatures and return types as features.py:130-158 (``extract_inference_features``) and :441-508
(``extract_inference_features_from_audio``); the arithmetic runs in csrc/features.cuh + csrc/yin.cuh.
Training-set builders and reports (features.py:24-102, :221-272, :343-435) are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from ..config import MFCCConfig, MelSpecConfig, TARGET_SR, asdict
from ..dsp.yin import shared_engine


class MelFeatureBuilder:
    def __init__(self, device=None):
        self.device = device

    def _engine(self, sr, mfcc_config, melspec_config):
        mel = {k: melspec_config[k] for k in ("N_MELS", "N_FFT", "HOP_LENGTH")}
        mf = {"N_MFCC": mfcc_config["N_MFCC"]}
        return shared_engine(sr, self.device, mel, mf)

    def _normalize_audio_volume(self, y, eps=1e-9):
        """features.py:124-126 (host helper kept for API parity; the device path fuses it into the loads)."""
        rms = np.sqrt(np.mean(y ** 2))
        return y / (rms + eps)

    # ---- batched device entry points (additions; the reference loops clip by clip in Python)
    def extract_mfcc_features_batch(self, clips, sr, n_mfcc=64, normalize_audio_volume=True, add_pitch_features=True,
                                    yin_on_normalized=False, scaler=None, melspec_config=None):
        eng = self._engine(sr, {"N_MFCC": n_mfcc}, melspec_config or asdict(MelSpecConfig()))
        if scaler is not None:
            eng.set_scaler(scaler)
        feats, hz = eng.mfcc_features(clips, normalize_audio_volume, add_pitch_features, yin_on_normalized, scaler is not None)
        return feats, hz

    def extract_melspec_features_batch(self, clips, sr, n_mels=64, n_fft=2048, hop_length=256, normalize_audio_volume=True):
        eng = self._engine(sr, asdict(MFCCConfig()), {"N_MELS": n_mels, "N_FFT": n_fft, "HOP_LENGTH": hop_length})
        return eng.melspec_db(clips, normalize_audio_volume)

    # ---- reference API
    def extract_inference_features(self, audio_loader, mfcc_config=None, melspec_config=None, scaler=None):
        """features.py:130-158: (mfcc float32 [N, 65] - sklearn's StandardScaler keeps float32 input in
        float32 -, melspec torch.Tensor [N, 1, n_mels, T])."""
        if mfcc_config is None:
            mfcc_config = asdict(MFCCConfig())
        if melspec_config is None:
            melspec_config = asdict(MelSpecConfig())
        wavs, _, _, _ = audio_loader.load_audio_dataset(pad_to_max=True)
        clips = np.stack([np.asarray(w, dtype=np.float32) for w in wavs])
        sr = audio_loader.target_sr
        eng = self._engine(sr, mfcc_config, melspec_config)
        dev = eng._clips(clips)
        if scaler:
            eng.set_scaler(scaler)
        feats, _ = eng.mfcc_features(dev, mfcc_config["NORMALIZE_AUDIO_VOLUME"], mfcc_config["ADD_PITCH_FEATURES"],
                                     yin_on_normalized=False, apply_scaler=bool(scaler))
        mel = eng.melspec_db(dev, melspec_config["NORMALIZE_AUDIO_VOLUME"])
        return feats.cpu().numpy(), mel.cpu()

    def extract_inference_features_from_audio(self, audio, target_sr=TARGET_SR, mfcc_config=None, melspec_config=None,
                                              scaler=None, melspec_to_db=True):
        """features.py:441-508: ((1, 65) float32, (1, 1, n_mels, T) float32 numpy).  As in the reference the
        ``scaler`` argument is accepted and NOT applied on this path, and YIN sees the normalised audio."""
        if mfcc_config is None:
            mfcc_config = asdict(MFCCConfig())
        if melspec_config is None:
            melspec_config = asdict(MelSpecConfig())
        if not melspec_to_db:
            raise NotImplementedError("melspec_to_db=False is not implemented (the reference always passes True)")
        eng = self._engine(target_sr, mfcc_config, melspec_config)
        dev = eng._clips(np.asarray(audio, dtype=np.float32))
        feats, _ = eng.mfcc_features(dev, mfcc_config["NORMALIZE_AUDIO_VOLUME"], mfcc_config["ADD_PITCH_FEATURES"],
                                     yin_on_normalized=True, apply_scaler=False)
        mel = eng.melspec_db(dev, melspec_config["NORMALIZE_AUDIO_VOLUME"])
        return feats.cpu().numpy(), mel.cpu().numpy()
