"""``Transcriber`` - drop-in for the reference's transcribe.py:25-199 (in-memory paths).

``transcribe_note`` keeps the reference signature and result dict; ``transcribe_notes`` (batch of clips)
and ``transcribe_audio`` (whole signal in memory: segmentation + features + ensemble + YIN) are the
batched additions.  ``transcribe(audio_path)`` is the file pipeline: WAV decode, channel mean and
resampling on the GPU, slicing, the PCM_16 round trip the reference's clip files go through, features with the
checkpoint's scaler, ensemble, per-clip YIN.  Resampling restates soxr_hq's specification, not its bits.
"""
from __future__ import annotations

from datetime import datetime
from pathlib import Path

import numpy as np
import torch

from .audio.features import MelFeatureBuilder
from .audio.slicing import AudioSlicer
from .checkpoint import load_checkpoint
from .config import CLIP_DURATION, CNN_CONFIG, INFERENCE_OUTPUT_ROOT, MLP_CONFIG, SLICER_CONFIG, TARGET_SR
from .dsp.yin import YinDsp
from .engine import Engine
from .note_predictor import NotePredictor


class Transcriber:
    def __init__(self, mlp_ckpt=None, cnn_ckpt=None, mlp_root=None, cnn_root=None, device: str = "cuda"):
        self.device = torch.device(device)
        mlp_root = Path(mlp_root) if mlp_root else MLP_CONFIG.CHECKPOINTS_DIR
        cnn_root = Path(cnn_root) if cnn_root else CNN_CONFIG.CHECKPOINTS_DIR
        mlp_path = mlp_root / (Path(mlp_ckpt) if mlp_ckpt else Path(MLP_CONFIG.DEFAULT_CKPT_NAME))
        cnn_path = cnn_root / (Path(cnn_ckpt) if cnn_ckpt else Path(CNN_CONFIG.DEFAULT_CKPT_NAME))
        if not mlp_path.is_file():
            raise FileNotFoundError(f"[Transcriber] Missing MLP checkpoint: {mlp_path}")
        if not cnn_path.is_file():
            raise FileNotFoundError(f"[Transcriber] Missing CNN checkpoint: {cnn_path}")
        self.model_ckpts = {"mlp": load_checkpoint(mlp_path), "cnn": load_checkpoint(cnn_path)}
        self.model_configs = {"mlp": self.model_ckpts["mlp"].get("config"), "cnn": self.model_ckpts["cnn"].get("config")}
        if not self.model_configs["mlp"] or not self.model_configs["cnn"]:
            raise ValueError("[Transcriber] Checkpoints missing 'config' field.")

        mel = self.model_configs["cnn"]["features"]["params"]
        mf = self.model_configs["mlp"]["features"]["params"]
        self.engine = Engine(self.model_configs["mlp"]["target_sr"],
                             {k: mel[k] for k in ("N_MELS", "N_FFT", "HOP_LENGTH")}, {"N_MFCC": mf["N_MFCC"]},
                             device=self.device)
        self.slicer = AudioSlicer(device=self.device)
        self.feature_builder = MelFeatureBuilder(device=self.device)
        self.predictor = NotePredictor(device=self.device)
        self.predictor.engine = self.engine
        self.predictor.load_models(self.model_ckpts["mlp"], self.model_ckpts["cnn"])
        scaler = self.model_ckpts["mlp"].get("scaler")
        if scaler is not None:
            self.engine.set_scaler(scaler)

    # ------------------------------------------------------------------ helpers
    def _target_sr(self) -> int:
        if self.model_configs["mlp"]["target_sr"] != self.model_configs["cnn"]["target_sr"]:
            raise ValueError("[Transcriber] Target SR mismatch.")
        return self.model_configs["mlp"]["target_sr"]

    def _feature_flags(self) -> dict:
        """The checkpoints' feature switches (features.py:184-185,:199,:310-311,:458-460,:471,:496-497) as keyword
        arguments of Engine.transcribe_clips."""
        mf = self.model_configs["mlp"]["features"]["params"]
        mel = self.model_configs["cnn"]["features"]["params"]
        return {"add_pitch": bool(mf.get("ADD_PITCH_FEATURES", True)),
                "normalize_mfcc": bool(mf.get("NORMALIZE_AUDIO_VOLUME", True)),
                "normalize_mel": bool(mel.get("NORMALIZE_AUDIO_VOLUME", True))}

    # ------------------------------------------------------------------ reference API
    def transcribe_note(self, audio: np.ndarray, clip_duration: float = CLIP_DURATION, sr_in: int = TARGET_SR) -> dict:
        """transcribe.py:147-199 for one clip."""
        return self.transcribe_notes(np.asarray(audio)[None, :], clip_duration, sr_in)

    def transcribe(self, audio_path, out_root=INFERENCE_OUTPUT_ROOT, audio_name: str = "transcribe_audio",
                   target_sr: int = TARGET_SR, clip_duration: float = CLIP_DURATION, save_clips: bool = True) -> dict:
        """transcribe.py:77-144: slice the file at ``target_sr``, pass every clip through PCM_16 (the reference
        writes clip .wav files and loads them back), bring the clips to the checkpoint's rate, extract features
        WITH the checkpoint's scaler, predict, and run YIN on the raw clips (``dsp_info``).

        ``save_clips`` keeps the reference's side effect (``out_root/<name>_<timestamp>/<name>/NNNN_clip__T.TTTs.wav``);
        the clips are never read back, the device copy is already what a reload would return.  Clips come back in
        onset order (the reference's order is whatever ``os.listdir`` yields)."""
        ckpt_sr = self._target_sr()
        slicer_engine = self.engine if int(target_sr) == int(ckpt_sr) else self.slicer.engine(target_sr)
        y, _ = self.slicer.load_wav(audio_path, target_sr)
        seg = slicer_engine.segment(y, clip_duration, SLICER_CONFIG)
        clips = seg["clips"]
        if save_clips:                                            # float clips -> PCM_16 files, as save_clip does
            out_dir = Path(out_root) / f"{audio_name}_{datetime.now().strftime('%m-%d_%H-%M-%S')}" / audio_name
            out_dir.mkdir(exist_ok=True, parents=True)
            table = seg["table"].cpu().numpy()
            onsets = seg["onsets"].cpu().numpy()
            for clip, row in zip(clips.cpu().numpy(), table):
                self.slicer.save_clip(clip, target_sr, out_dir, int(row[0]), float(onsets[int(row[0])]) / target_sr)
        if clips.shape[0]:
            slicer_engine.pcm16_roundtrip_(clips)                 # what loading those files back returns
        if clips.shape[0] == 0:
            raise FileNotFoundError("load_audio_dataset: No audio files found.")
        if int(target_sr) != int(ckpt_sr):                       # AudioDatasetLoader: librosa.load(sr=ckpt_sr) + fix_len
            clips = self.engine.resample(clips.to(self.engine.device), target_sr, ckpt_sr)
        fixed = int(ckpt_sr * clip_duration)
        if clips.shape[1] > fixed:
            clips = clips[:, :fixed].contiguous()
        elif clips.shape[1] < fixed:
            clips = torch.nn.functional.pad(clips, (0, fixed - clips.shape[1]))
        return self._predict_sliced(clips, seg)

    @staticmethod
    def _dsp_info(hz) -> list:
        """transcribe.py:139-143: (median YIN Hz, {"midi", "note_name", "midi_float"}) per clip; YinDsp.round_to_nearest_pitch
        (dsp/yin.py:21-37) evaluated over the whole array at once (same numpy ufuncs, same values)."""
        from .dsp.yin import _NOTES_UNICODE
        hz = np.asarray(hz, dtype=np.float64).reshape(-1)
        ok = ~np.isnan(hz) & (hz > 0)
        with np.errstate(all="ignore"):
            midi_float = 12 * (np.log2(hz) - np.log2(440.0)) + 69
        rounded = np.where(ok, np.round(midi_float), 0).astype(np.int64)
        info = []
        for v, good, mf, mr in zip(hz.tolist(), ok.tolist(), midi_float.tolist(), rounded.tolist()):
            if not good:
                info.append((v, {"midi": None, "note_name": None, "midi_float": None}))
            else:
                info.append((v, {"midi": mr, "note_name": "{:s}{:0d}".format(_NOTES_UNICODE[mr % 12], int(mr / 12) - 1), "midi_float": mf}))
        return info

    def _ensemble_sliced(self, clips) -> dict:
        """File-path features (scaler applied, YIN on the raw clips: features.py:145-146,:201) + ensemble, on device."""
        self.engine.set_ensemble_weights(self.predictor.mlp_weight, self.predictor.cnn_weight)
        return self.engine.transcribe_clips(clips, yin_on_normalized=False, apply_scaler=self.engine.has_scaler,
                                            return_features=True, **self._feature_flags())

    def _predict_sliced(self, clips, seg) -> dict:
        out = self._ensemble_sliced(clips)
        result = self.predictor._result(out)
        result["dsp_info"] = self._dsp_info(out["yin_hz"].cpu().numpy())
        result["onsets"] = [int(v) for v in seg["onsets"].cpu().numpy()]
        result["slice_table"] = seg["table"].cpu().numpy()
        return result

    # ------------------------------------------------------------------ batched additions
    def transcribe_notes_device(self, audio, clip_duration: float = CLIP_DURATION, sr_in: int = TARGET_SR) -> dict:
        """``transcribe_notes`` that leaves the result on the device (Engine.transcribe_clips' dict of tensors)."""
        target_sr = self._target_sr()
        target_len = int(clip_duration * target_sr)
        a = self.engine._clips(audio if torch.is_tensor(audio) else np.asarray(audio, dtype=np.float32))
        if sr_in != target_sr:                                   # transcribe.py:172-173
            a = self.engine.resample(a, sr_in, target_sr)
        if a.shape[1] < target_len:
            a = torch.nn.functional.pad(a, (0, target_len - a.shape[1]))
        elif a.shape[1] > target_len:
            a = a[:, :target_len].contiguous()
        self.engine.set_ensemble_weights(self.predictor.mlp_weight, self.predictor.cnn_weight)
        return self.engine.transcribe_clips(a, yin_on_normalized=True, apply_scaler=False, **self._feature_flags())

    def transcribe_notes(self, audio, clip_duration: float = CLIP_DURATION, sr_in: int = TARGET_SR) -> dict:
        """N clips at once ([N, n] array or device tensor); each clip follows transcribe_note exactly:
        pad/trim to int(clip_duration*target_sr), features WITHOUT the scaler, YIN on the normalised audio."""
        return self.predictor._result(self.transcribe_notes_device(audio, clip_duration, sr_in))

    def transcribe_audio(self, y, sr: int | None = None, clip_duration: float = CLIP_DURATION) -> dict:
        """transcribe.py:77-144 from memory: slice -> features (scaler applied, YIN on raw clips) -> predict ->
        per-clip YIN ``dsp_info``; plus ``onsets`` and ``slice_table``."""
        target_sr = self._target_sr()
        yt = torch.as_tensor(np.asarray(y, dtype=np.float32) if not torch.is_tensor(y) else y)
        if sr is not None and sr != target_sr:
            yt = self.engine.resample(yt.reshape(-1), sr, target_sr)
        seg = self.engine.segment(yt, clip_duration, SLICER_CONFIG)
        clips = seg["clips"]
        if clips.shape[0] == 0:
            raise FileNotFoundError("load_audio_dataset: No audio files found.")
        return self._predict_sliced(clips, seg)

    # ------------------------------------------------------------------ sharded across the GPUs of one box
    def transcribe_notes_sharded(self, audio, clip_duration: float = CLIP_DURATION, sr_in: int = TARGET_SR, group=None) -> dict:
        """``transcribe_notes`` with the clips sharded over the ranks of ``group`` (one process per GPU): every rank
        passes the SAME ``[N, n]`` batch, runs the pipeline on its contiguous block and receives everybody's labels
        through one all-gather of per-clip records.  See parallel.transcribe_notes_sharded."""
        from . import parallel
        return parallel.transcribe_notes_sharded(self, audio, clip_duration, sr_in, group)

    def transcribe_phrases_sharded(self, phrases, clip_duration: float = CLIP_DURATION, group=None) -> dict:
        """A long recording given as P independent equal-length signals ``[P, L]`` (phrases / files): rank r slices
        and transcribes its block of signals in batched kernels, then the slice tables and labels are all-gathered
        (SURVEY 8(e) option (i)).  See parallel.transcribe_phrases_sharded."""
        from . import parallel
        return parallel.transcribe_phrases_sharded(self, phrases, clip_duration, group)

    def transcribe_audio_sharded(self, y, sr: int | None = None, clip_duration: float = CLIP_DURATION, group=None) -> dict:
        """``transcribe_audio`` for ONE contiguous signal: the whole-file onset chain has global dependencies and runs
        (redundantly, identically) on every rank; the sliced clips are then sharded (SURVEY 8(e) option (ii)).
        See parallel.transcribe_audio_sharded."""
        from . import parallel
        return parallel.transcribe_audio_sharded(self, y, sr, clip_duration, group)
