"""``MelFeatureBuilder`` - drop-in for the reference's audio/features.py.

Same signatures and return types as features.py:130-158 (``extract_inference_features``), :441-508
(``extract_inference_features_from_audio``) and the training-set builders: :162-219
``extract_mfcc_features``, :275-341 ``extract_melspec_features``, :221-272 / :343-435 the train/val DataLoader
builders.  The arithmetic runs batched in csrc/features.cuh + csrc/yin.cuh (the reference loops clip by clip);
label encoding, the stratified split and ``StandardScaler.fit`` are the same sklearn calls on the host.
``generate_feature_report`` (:24-102, a JSON summary) is not mirrored.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

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

    # ---- shared helpers (features.py:107-122)
    def _encode_labels_to_ints(self, labels):
        classes = sorted(set(labels))
        label_to_idx = {c: i for i, c in enumerate(classes)}
        idx_to_label = {i: c for i, c in enumerate(classes)}
        return [label_to_idx[l] for l in labels], len(classes), idx_to_label

    def _create_tensor_dataset(self, X, y):
        X_tensor = X if isinstance(X, torch.Tensor) else torch.tensor(X, dtype=torch.float32)
        return TensorDataset(X_tensor, torch.tensor(y, dtype=torch.long))

    def _labels(self, labels):
        y = np.array(labels, dtype=str)
        y_encoded, num_classes, reverse_map = self._encode_labels_to_ints(y)
        return np.array(y_encoded, dtype=int), num_classes, reverse_map

    # ---- training-set feature extraction (features.py:162-219, :275-341), one batched pass each
    def extract_mfcc_features(self, audio_loader, n_mfcc=13, normalize_audio_volume: bool = False,
                              add_pitch_features: bool = True):
        """-> (X float32 (N, n_mfcc [+1 = log10(YIN Hz)]), y_encoded int (N,), num_classes, reverse_map).
        YIN runs on the raw clip even when the MFCCs use the normalised one (features.py:201)."""
        wavs, _, labels, _ = audio_loader.load_audio_dataset(pad_to_max=True)
        clips = np.stack([np.asarray(w, dtype=np.float32) for w in wavs])
        eng = self._engine(audio_loader.target_sr, {"N_MFCC": n_mfcc}, asdict(MelSpecConfig()))
        feats, hz = eng.mfcc_features(eng._clips(clips), normalize_audio_volume, add_pitch_features,
                                      yin_on_normalized=False, apply_scaler=False)
        X = feats.cpu().numpy()
        if add_pitch_features and bool(torch.isnan(hz).any()):
            # the reference appends the pitch feature only when YIN returns one, then np.vstack fails on ragged rows
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly "
                             "(YIN found no pitch for at least one clip)")
        y_encoded, num_classes, reverse_map = self._labels(labels)
        print(f"Extracted MFCC features for {len(X)} samples.")
        return X, y_encoded, num_classes, reverse_map

    def extract_melspec_features(self, audio_loader, n_mels: int = 128, n_fft: int = 1024, hop_length: int = 256,
                                 normalize_audio_volume: bool = False, to_db: bool = True):
        """-> (X torch float32 (N, 1, n_mels, T), y_encoded, num_classes, reverse_map)."""
        wavs, _, labels, _ = audio_loader.load_audio_dataset(pad_to_max=True)
        clips = np.stack([np.asarray(w, dtype=np.float32) for w in wavs])
        eng = self._engine(audio_loader.target_sr, asdict(MFCCConfig()), {"N_MELS": n_mels, "N_FFT": n_fft, "HOP_LENGTH": hop_length})
        X = eng.melspec_db(eng._clips(clips), normalize_audio_volume, to_db).cpu()
        y_encoded, num_classes, reverse_map = self._labels(labels)
        print(f"Extracted Mel-spectrogram features for {X.shape[0]} samples. X shape: {tuple(X.shape)}")
        return X, y_encoded, num_classes, reverse_map

    def build_mfcc_train_val_dataloaders(self, audio_loader, n_mfcc=13, batch_size: int = 32, val_size: float = 0.2,
                                         shuffle_train: bool = True, shuffle_val: bool = False, normalize_audio_volume=False,
                                         standard_scaler: bool = True, seed: int = 42, num_workers: int = 0,
                                         pin_memory: bool = True, drop_last: bool = False):
        """features.py:221-272: stratified split, StandardScaler fitted on the training part."""
        from sklearn.model_selection import train_test_split
        from sklearn.preprocessing import StandardScaler
        X, y_encoded, num_classes, reverse_map = self.extract_mfcc_features(audio_loader, n_mfcc, normalize_audio_volume)
        X_tr, X_val, y_tr, y_val = train_test_split(X, y_encoded, test_size=val_size, stratify=y_encoded, random_state=seed)
        scaler = None
        if standard_scaler:
            scaler = StandardScaler().fit(X_tr)
            X_tr, X_val = scaler.transform(X_tr), scaler.transform(X_val)
            self.scaler = scaler
        dl_tr = DataLoader(self._create_tensor_dataset(X_tr, y_tr), batch_size=batch_size, shuffle=shuffle_train,
                           num_workers=num_workers, pin_memory=pin_memory, drop_last=drop_last)
        dl_val = DataLoader(self._create_tensor_dataset(X_val, y_val), batch_size=batch_size, shuffle=shuffle_val,
                            num_workers=num_workers, pin_memory=pin_memory, drop_last=False)
        return dl_tr, dl_val, X, y_encoded, num_classes, reverse_map, scaler

    def build_melspec_dataloader(self, audio_loader, n_mels: int = 128, n_fft: int = 1024, hop_length: int = 256,
                                 batch_size: int = 32, shuffle: bool = True, normalize_audio_volume: bool = False):
        """features.py:343-364."""
        X, y_encoded, num_classes, reverse_map = self.extract_melspec_features(
            audio_loader=audio_loader, n_mels=n_mels, n_fft=n_fft, hop_length=hop_length,
            normalize_audio_volume=normalize_audio_volume)
        return DataLoader(self._create_tensor_dataset(X, y_encoded), batch_size=batch_size, shuffle=shuffle), num_classes, reverse_map

    def build_melspec_train_val_dataloaders(self, audio_loader, n_mels: int = 128, n_fft: int = 1024, hop_length: int = 256,
                                            batch_size: int = 32, val_size: float = 0.2, shuffle_train: bool = True,
                                            shuffle_val: bool = False, normalize_audio_volume: bool = False, seed: int = 42,
                                            num_workers: int = 0, pin_memory: bool = True, drop_last: bool = False):
        """features.py:366-435: stratified split by index, no scaler."""
        from sklearn.model_selection import train_test_split
        X, y_encoded, num_classes, reverse_map = self.extract_melspec_features(
            audio_loader=audio_loader, n_mels=n_mels, n_fft=n_fft, hop_length=hop_length,
            normalize_audio_volume=normalize_audio_volume)
        idx_tr, idx_val, y_tr, y_val = train_test_split(np.arange(len(y_encoded)), y_encoded, test_size=val_size,
                                                        stratify=y_encoded, random_state=seed)
        dl_tr = DataLoader(self._create_tensor_dataset(X[idx_tr], y_tr), batch_size=batch_size, shuffle=shuffle_train,
                           num_workers=num_workers, pin_memory=pin_memory, drop_last=drop_last)
        dl_val = DataLoader(self._create_tensor_dataset(X[idx_val], y_val), batch_size=batch_size, shuffle=shuffle_val,
                            num_workers=num_workers, pin_memory=pin_memory, drop_last=False)
        return dl_tr, dl_val, X, y_encoded, num_classes, reverse_map

    # ---- batched device entry points (additions; the reference loops clip by clip in Python)
    def extract_mfcc_features_batch(self, clips, sr, n_mfcc=64, normalize_audio_volume=True, add_pitch_features=True,
                                    yin_on_normalized=False, scaler=None, melspec_config=None):
        eng = self._engine(sr, {"N_MFCC": n_mfcc}, melspec_config or asdict(MelSpecConfig()))
        return self._mfcc_with_scaler(eng, clips, normalize_audio_volume, add_pitch_features, yin_on_normalized, scaler)

    @staticmethod
    def _mfcc_with_scaler(eng, clips, normalize, add_pitch, yin_on_normalized, scaler):
        """The engines handed out by shared_engine are shared process-wide and must stay stateless: a scaler is
        installed for this call only."""
        if scaler is None:
            return eng.mfcc_features(clips, normalize, add_pitch, yin_on_normalized, False)
        eng.set_scaler(scaler)
        try:
            return eng.mfcc_features(clips, normalize, add_pitch, yin_on_normalized, True)
        finally:
            eng.set_scaler(None)

    def extract_melspec_features_batch(self, clips, sr, n_mels=64, n_fft=2048, hop_length=256, normalize_audio_volume=True,
                                       to_db=True):
        eng = self._engine(sr, asdict(MFCCConfig()), {"N_MELS": n_mels, "N_FFT": n_fft, "HOP_LENGTH": hop_length})
        return eng.melspec_db(clips, normalize_audio_volume, to_db)

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
        feats, _ = self._mfcc_with_scaler(eng, dev, mfcc_config["NORMALIZE_AUDIO_VOLUME"], mfcc_config["ADD_PITCH_FEATURES"],
                                          False, scaler if scaler else None)
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
        eng = self._engine(target_sr, mfcc_config, melspec_config)
        dev = eng._clips(np.asarray(audio, dtype=np.float32))
        feats, _ = eng.mfcc_features(dev, mfcc_config["NORMALIZE_AUDIO_VOLUME"], mfcc_config["ADD_PITCH_FEATURES"],
                                     yin_on_normalized=True, apply_scaler=False)
        mel = eng.melspec_db(dev, melspec_config["NORMALIZE_AUDIO_VOLUME"], bool(melspec_to_db))
        return feats.cpu().numpy(), mel.cpu().numpy()
