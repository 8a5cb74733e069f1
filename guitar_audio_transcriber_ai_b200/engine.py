"""Device pipeline: one ``Engine`` = one gat_ctx on one GPU, driving the kernels in csrc/ through the C ABI.

PyTorch is plumbing here (device memory, streams, ``torch.distributed`` for the label all-gather); every
number comes out of libgat.so.  Batched entry points take ``[N, n]`` float32 clips resident on the GPU and
return device tensors; nothing is synchronised unless the caller asks for host arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, tables
from .checkpoint import pack_cnn, pack_mlp
from .config import MELSPEC_CONFIG, MFCC_CONFIG, SLICER_CONFIG, asdict


class Engine:
    def __init__(self, sample_rate: int, melspec_config: dict | None = None, mfcc_config: dict | None = None,
                 device=None, yin_fmin: float = 50.0, yin_fmax: float = 1000.0):
        self.lib = _lib.load()
        if device is None:
            device = "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda" and not getattr(self.lib, "_host_emulation", False):
            raise _lib.GatError("guitar_audio_transcriber_ai_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.sample_rate = int(sample_rate)
        self.melspec_config = dict(melspec_config or asdict(MELSPEC_CONFIG))
        self.mfcc_config = dict(mfcc_config or asdict(MFCC_CONFIG))
        self.n_fft = int(self.melspec_config["N_FFT"])
        self.hop = int(self.melspec_config["HOP_LENGTH"])
        self.n_mels = int(self.melspec_config["N_MELS"])
        self.n_mfcc = int(self.mfcc_config["N_MFCC"])
        self.num_classes = 0
        self.has_scaler = False

        self._tables = {
            "mel_window": tables.hann_window_f32(self.n_fft),
            "mel_fb": tables.htk_fbanks(self.sample_rate, self.n_fft, self.n_mels),
            "stft_window": tables.hann_window_f64(2048),
            "mfcc_fb": tables.slaney_mel_fb(self.sample_rate, 2048, 128),
            "dct": tables.dct_matrix(self.n_mfcc, 128),
        }
        cfg = _lib.GatConfig(
            sample_rate=self.sample_rate, mel_n_fft=self.n_fft, mel_hop=self.hop, mel_n_mels=self.n_mels,
            mel_window=self._tables["mel_window"].ctypes.data, mel_fb=self._tables["mel_fb"].ctypes.data,
            mfcc_n_mels=128, mfcc_n_mfcc=self.n_mfcc,
            stft_window=self._tables["stft_window"].ctypes.data, mfcc_fb=self._tables["mfcc_fb"].ctypes.data,
            dct=self._tables["dct"].ctypes.data,
            yin_fmin=float(yin_fmin), yin_fmax=float(yin_fmax), yin_trough_threshold=0.1)
        handle = C.c_void_p()
        with self._on_device():
            self.lib.check(self.lib.gat_ctx_create(C.byref(cfg), self.device.index or 0, C.byref(handle)), ValueError)
        self._ctx = handle
        self._resample_cache = {}
        self._slicer_cache = {}
        self._host_out = None
        self.mlp_in = 0

    # ------------------------------------------------------------------ plumbing
    def _on_device(self):
        return torch.cuda.device(self.device) if self.device.type == "cuda" else _Null()

    def _stream(self):
        if self.device.type != "cuda":
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _clips(self, audio) -> torch.Tensor:
        t = torch.as_tensor(audio)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.dim() != 2:
            raise ValueError("audio must be [n] or [N, n]")
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.gat_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.gat_launch_count(self._ctx))

    def profile_begin(self):
        self.lib.check(self.lib.gat_profile_begin(self._ctx))

    def profile_end(self) -> dict:
        """{kernel name: (launches, total milliseconds)} for everything launched since profile_begin."""
        buf = C.create_string_buffer(1 << 16)
        self.lib.check(self.lib.gat_profile_end(self._ctx, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out

    def fma_peak_tflops(self, iters: int = 4096) -> float:
        """Measured FP32-FMA peak of this GPU (TFLOP/s): the roof the FFT-bound stages are reported against."""
        v = C.c_float(0.0)
        self.lib.check(self.lib.gat_debug_fma_peak(self._ctx, iters, C.byref(v)))
        return float(v.value)

    def mel_frames(self, n: int) -> int:
        return 1 + n // self.hop

    # ------------------------------------------------------------------ models
    def load_mlp(self, state_dict: dict):
        packed = pack_mlp(state_dict)
        dims, params = packed["dims"], packed["params"]
        self.lib.check(self.lib.gat_load_mlp(self._ctx, _lib.ptr(dims), len(dims) - 1, _lib.ptr(params), params.size), ValueError)
        self.num_classes = int(dims[-1])
        self.mlp_in = int(dims[0])
        self._host_out = None

    def load_cnn(self, state_dict: dict):
        packed = pack_cnn(state_dict)
        convs, fcs = packed["convs"], packed["fcs"]
        if len(fcs) != 2:
            raise ValueError("CNN classifier must be Linear-LeakyReLU-Linear (hidden_dim > 0)")
        channels = np.asarray([convs[0]["c_in"]] + [c["c_out"] for c in convs], dtype=np.int32)
        if any(c["k"] != 3 for c in convs):
            raise ValueError("only 3x3 convolutions are implemented")
        wp = (C.c_void_p * len(convs))(*[c["w"].ctypes.data for c in convs])
        bp = (C.c_void_p * len(convs))(*[c["b"].ctypes.data for c in convs])
        hidden, classes = fcs[0]["w"].shape[1], fcs[1]["w"].shape[1]
        self._cnn_keepalive = (convs, fcs)
        self.lib.check(self.lib.gat_load_cnn(self._ctx, len(convs), _lib.ptr(channels), wp, bp, hidden, classes,
                                             _lib.ptr(fcs[0]["w"]), _lib.ptr(fcs[0]["b"]),
                                             _lib.ptr(fcs[1]["w"]), _lib.ptr(fcs[1]["b"])), ValueError)
        self.num_classes = int(classes)
        self._host_out = None

    def set_scaler(self, scaler):
        if scaler is None:
            self.lib.check(self.lib.gat_set_scaler(self._ctx, None, None, 0))
            self.has_scaler = False
            return
        mean = np.ascontiguousarray(scaler.mean_, dtype=np.float64)
        scale = np.ascontiguousarray(scaler.scale_, dtype=np.float64)
        self.lib.check(self.lib.gat_set_scaler(self._ctx, _lib.ptr(mean), _lib.ptr(scale), mean.size))
        self.has_scaler = True

    def set_ensemble_weights(self, mlp_weight: float, cnn_weight: float):
        self.lib.check(self.lib.gat_set_ensemble_weights(self._ctx, np.float32(mlp_weight), np.float32(cnn_weight)))

    # ------------------------------------------------------------------ features
    def melspec_db(self, audio, normalize: bool = True, to_db: bool = True) -> torch.Tensor:
        """[N, n] -> [N, 1, n_mels, T] (features.py:296-331); ``to_db=False`` returns the mel power (:313-316)."""
        a = self._clips(audio)
        N, n = a.shape
        out = self._empty((N, 1, self.n_mels, self.mel_frames(n)), torch.float32)
        mode = (_lib.GAT_MEL_NORMALIZE if normalize else 0) | (0 if to_db else _lib.GAT_MEL_POWER)
        with self._on_device():
            self.lib.check(self.lib.gat_melspec_db(self._ctx, _lib.ptr(a), N, n, mode, _lib.ptr(out), self._stream()), ValueError)
        return out

    def mfcc_features(self, audio, normalize=True, add_pitch=True, yin_on_normalized=False, apply_scaler=False):
        """[N, n] -> ([N, n_mfcc(+1)] float32, yin_hz [N] float64) (features.py:182-208 / :458-478)."""
        a = self._clips(audio)
        N, n = a.shape
        F = self.n_mfcc + (1 if add_pitch else 0)
        out = self._empty((N, F), torch.float32)
        hz = self._empty((N,), torch.float64)
        with self._on_device():
            self.lib.check(self.lib.gat_mfcc_features(self._ctx, _lib.ptr(a), N, n, int(normalize), int(add_pitch),
                                                      int(yin_on_normalized), int(apply_scaler), _lib.ptr(out), F,
                                                      _lib.ptr(hz), self._stream()), ValueError)
        return out, hz

    def yin(self, audio, normalize: bool = False):
        """[N, n] -> (median f0 [N] float64, frame f0 [N, 1 + n//512] float64) (dsp/yin.py:49-67)."""
        a = self._clips(audio)
        N, n = a.shape
        hz = self._empty((N,), torch.float64)
        f0 = self._empty((N, 1 + n // 512), torch.float64)
        with self._on_device():
            self.lib.check(self.lib.gat_yin(self._ctx, _lib.ptr(a), N, n, int(normalize), _lib.ptr(hz), _lib.ptr(f0), self._stream()), ValueError)
        return hz, f0

    # ------------------------------------------------------------------ inference
    def infer(self, mfcc, mel) -> dict:
        """note_predictor.py:84-135 on device tensors; returns device tensors (plus logits)."""
        x = torch.as_tensor(mfcc).to(device=self.device, dtype=torch.float32).contiguous()
        m = torch.as_tensor(mel).to(device=self.device, dtype=torch.float32).contiguous()
        if m.dim() == 4:
            if m.shape[1] != 1:
                raise ValueError("melspec features must have one channel")
            m = m[:, 0]
        N, H, T = m.shape
        if x.shape[0] != N:
            raise ValueError("mfcc and melspec batch sizes differ")
        if H != self.n_mels:
            raise ValueError(f"melspec has {H} mel bands, the engine was built for {self.n_mels}")
        K = self.num_classes
        out = {k: self._empty((N, K), torch.float32) for k in ("probs", "mlp_probs", "cnn_probs", "mlp_logits", "cnn_logits")}
        out["indices"] = self._empty((N,), torch.int64)
        out["confidences"] = self._empty((N,), torch.float32)
        with self._on_device():
            self.lib.check(self.lib.gat_infer(self._ctx, _lib.ptr(x), x.shape[1], _lib.ptr(m), N, T, _lib.ptr(out["probs"]),
                                              _lib.ptr(out["mlp_probs"]), _lib.ptr(out["cnn_probs"]), _lib.ptr(out["indices"]),
                                              _lib.ptr(out["confidences"]), _lib.ptr(out["mlp_logits"]),
                                              _lib.ptr(out["cnn_logits"]), self._stream()), ValueError)
        return out

    @staticmethod
    def _flags(yin_on_normalized, apply_scaler, skip_mlp, add_pitch=True, normalize_mfcc=True, normalize_mel=True) -> int:
        return ((_lib.GAT_FLAG_YIN_ON_NORMALIZED if yin_on_normalized else 0) | (_lib.GAT_FLAG_APPLY_SCALER if apply_scaler else 0) |
                (_lib.GAT_FLAG_SKIP_MLP if skip_mlp else 0) | (0 if add_pitch else _lib.GAT_FLAG_NO_PITCH) |
                (0 if normalize_mfcc else _lib.GAT_FLAG_NO_NORMALIZE_MFCC) | (0 if normalize_mel else _lib.GAT_FLAG_NO_NORMALIZE_MEL))

    def _check_mlp_width(self, skip_mlp: bool, add_pitch: bool):
        """The MLP reads ``mlp_in`` floats per feature row; the C side refuses a mismatch too (gat_transcribe_clips)."""
        if skip_mlp:
            return
        F = self.n_mfcc + (1 if add_pitch else 0)
        if not self.mlp_in:
            raise ValueError("no MLP loaded (Engine.load_mlp)")
        if self.mlp_in != F:
            raise ValueError(f"the loaded MLP takes {self.mlp_in} inputs but this engine builds {F} feature columns "
                             f"(N_MFCC {self.n_mfcc}{' + pitch' if add_pitch else ''})")

    def transcribe_clips(self, audio, yin_on_normalized=True, apply_scaler=False, skip_mlp=False,
                         return_features=False, add_pitch=True, normalize_mfcc=True, normalize_mel=True) -> dict:
        """Features + ensemble for N equal-length clips in one call (transcribe.py:186-197 batched).
        ``add_pitch`` / ``normalize_*`` are the checkpoints' ADD_PITCH_FEATURES / NORMALIZE_AUDIO_VOLUME switches."""
        a = self._clips(audio)
        N, n = a.shape
        K = self.num_classes
        self._check_mlp_width(skip_mlp, add_pitch)
        flags = self._flags(yin_on_normalized, apply_scaler, skip_mlp, add_pitch, normalize_mfcc, normalize_mel)
        out = {"probs": self._empty((N, K), torch.float32), "cnn_probs": self._empty((N, K), torch.float32),
               "indices": self._empty((N,), torch.int64), "confidences": self._empty((N,), torch.float32)}
        out["mlp_probs"] = None if skip_mlp else self._empty((N, K), torch.float32)
        mfcc = mel = hz = None
        if return_features:
            mel = self._empty((N, 1, self.n_mels, self.mel_frames(n)), torch.float32)
            if not skip_mlp:
                mfcc = self._empty((N, self.n_mfcc + (1 if add_pitch else 0)), torch.float32)
                hz = self._empty((N,), torch.float64)
        with self._on_device():
            self.lib.check(self.lib.gat_transcribe_clips(
                self._ctx, _lib.ptr(a), N, n, flags, _lib.ptr(out["probs"]), _lib.ptr(out["mlp_probs"]),
                _lib.ptr(out["cnn_probs"]), _lib.ptr(out["indices"]), _lib.ptr(out["confidences"]),
                _lib.ptr(mfcc), _lib.ptr(mel), _lib.ptr(hz), self._stream()), ValueError)
        out.update(mfcc=mfcc, mel=mel, yin_hz=hz)
        return out

    def transcribe_clips_host(self, audio_host, yin_on_normalized=True, apply_scaler=False, skip_mlp=False,
                              want_probs=True, add_pitch=True, normalize_mfcc=True, normalize_mel=True) -> dict:
        """Host buffers in (float32 clips, or int16 = PCM_16 clips scaled by 1/32768 on the device), host buffers
        out: H2D (chunked, overlapped), kernels, D2H inside the call.  The returned arrays are copies: the pinned
        landing buffers are reused by the next call."""
        if isinstance(audio_host, torch.Tensor):
            if audio_host.device.type != "cpu" or audio_host.dtype not in (torch.float32, torch.int16) or not audio_host.is_contiguous():
                raise ValueError("audio_host must be a contiguous float32 or int16 (PCM_16) CPU tensor (pinned for overlap)")
            pcm16 = audio_host.dtype == torch.int16
        else:
            pcm16 = isinstance(audio_host, np.ndarray) and audio_host.dtype == np.int16
            audio_host = np.ascontiguousarray(audio_host, dtype=np.int16 if pcm16 else np.float32)
        N, n = audio_host.shape
        src = _lib.ptr(audio_host)
        entry = self.lib.gat_transcribe_clips_host_pcm16 if pcm16 else self.lib.gat_transcribe_clips_host
        K = self.num_classes
        self._check_mlp_width(skip_mlp, add_pitch)
        flags = self._flags(yin_on_normalized, apply_scaler, skip_mlp, add_pitch, normalize_mfcc, normalize_mel)
        if self._host_out is None or self._host_out["probs"].shape != (N, K):
            pin = self.device.type == "cuda"
            self._host_out = {"indices": torch.empty(N, dtype=torch.int64, pin_memory=pin),
                              "confidences": torch.empty(N, dtype=torch.float32, pin_memory=pin),
                              "probs": torch.empty((N, K), dtype=torch.float32, pin_memory=pin)}
        ho = self._host_out
        with self._on_device():
            self.lib.check(entry(
                self._ctx, src, N, n, flags, _lib.ptr(ho["indices"]), _lib.ptr(ho["confidences"]),
                _lib.ptr(ho["probs"]) if want_probs else None), ValueError)
        return {"indices": ho["indices"].numpy().copy(), "confidences": ho["confidences"].numpy().copy(),
                "probs": ho["probs"].numpy().copy() if want_probs else None,
                "h2d_bytes": N * n * (2 if pcm16 else 4), "d2h_bytes": N * 12 + (N * K * 4 if want_probs else 0)}

    # ------------------------------------------------------------------ segmentation
    # ------------------------------------------------------------------ file front end
    def decode_mono(self, frames) -> torch.Tensor:
        """librosa.load's decode + to_mono on the device: interleaved [frames, channels] int16 or float32."""
        a = torch.as_tensor(np.array(frames, copy=True) if isinstance(frames, np.ndarray) and not frames.flags.writeable else frames)
        if a.dim() == 1:
            a = a.unsqueeze(1)
        if a.dtype == torch.int16:
            fmt = 0
        elif a.dtype == torch.float32:
            fmt = 1
        else:
            raise ValueError("decode_mono: int16 or float32 frames expected")
        a = a.to(self.device).contiguous()
        out = self._empty((a.shape[0],), torch.float32)
        with self._on_device():
            self.lib.check(self.lib.gat_decode_mono(self._ctx, _lib.ptr(a), fmt, a.shape[0], a.shape[1], _lib.ptr(out), self._stream()))
        return out

    def pcm16_roundtrip_(self, audio: torch.Tensor) -> torch.Tensor:
        """sf.write(PCM_16) followed by librosa.load, in place on a float32 device tensor."""
        if audio.dtype != torch.float32 or not audio.is_contiguous() or audio.device != self.device:
            raise ValueError("pcm16_roundtrip_: contiguous float32 tensor on the engine's device expected")
        with self._on_device():
            self.lib.check(self.lib.gat_pcm16_roundtrip(self._ctx, _lib.ptr(audio), audio.numel(), self._stream()))
        return audio

    def resample(self, audio, orig_sr: int, target_sr: int) -> torch.Tensor:
        """librosa.load(sr=target_sr) / librosa.resample for [n] or [N, n] signals; output length ceil(n*ratio)."""
        x = self._clips(audio)
        squeeze = torch.as_tensor(audio).dim() == 1
        if int(orig_sr) == int(target_sr):
            return x[0] if squeeze else x
        key = (int(orig_sr), int(target_sr))
        if key not in self._resample_cache:
            up, down, taps, half = tables.resample_filter(*key)
            self._resample_cache[key] = (up, down, torch.from_numpy(taps).to(self.device), half)
        up, down, taps, half = self._resample_cache[key]
        n_in = x.shape[1]
        n_out = (n_in * up + down - 1) // down
        out = self._empty((x.shape[0], n_out), torch.float32)
        with self._on_device():
            for i in range(0, x.shape[0], 65535):
                xi = x[i:i + 65535]
                self.lib.check(self.lib.gat_resample(self._ctx, _lib.ptr(xi), xi.shape[0], n_in, up, down, _lib.ptr(taps), half,
                                                     _lib.ptr(out[i:i + 65535]), n_out, self._stream()))
        return out[0] if squeeze else out

    def slicer_params(self, L: int, length_sec: float, cfg=None) -> _lib.GatSlicerParams:
        cfg = cfg or SLICER_CONFIG
        key = (int(L), float(length_sec), float(cfg.MIN_IN_DB_THRESHOLD), float(cfg.MIN_SLICE_RMS_DB), int(cfg.HOP_LEN),
               float(cfg.MIN_SEP), float(cfg.ATTACK_SKIP_SEC))
        hit = self._slicer_cache.get(key)
        if hit is None:
            if len(self._slicer_cache) > 256:
                self._slicer_cache.clear()
            hit = self._slicer_cache[key] = self._slicer_params(L, length_sec, cfg)
        return hit

    def _slicer_params(self, L: int, length_sec: float, cfg) -> _lib.GatSlicerParams:
        sr = self.sample_rate
        rms_hop = int(cfg.HOP_LEN)
        T = 1 + L // rms_hop
        k, gamma = tables.percentile_index_f32(T, 20)
        od = tables.onset_detect_params(sr, 512)
        return _lib.GatSlicerParams(
            min_db_threshold=float(cfg.MIN_IN_DB_THRESHOLD),
            sample_gate=float(tables.sample_gate_threshold(float(cfg.MIN_IN_DB_THRESHOLD))),
            rms_hop=rms_hop, p20_k=k, p20_gamma=float(gamma), gate_offset_db=6.0, onset_hop=512,
            pre_max=od["pre_max"], post_max=od["post_max"], pre_avg=od["pre_avg"], post_avg=od["post_avg"],
            wait=od["wait"], delta=float(od["delta"]),
            min_sep_samples=int(cfg.MIN_SEP * sr), attack_skip=int(cfg.ATTACK_SKIP_SEC * sr),
            clip_len=int(length_sec * sr), min_slice_rms_db=float(cfg.MIN_SLICE_RMS_DB))

    def detect_onsets(self, y, hop_len: int = 512, min_sep: float = 0.25) -> torch.Tensor:
        """AudioSlicer.detect_onsets(y, sr, hop_len, min_sep) (slicing.py:106-122) on an un-gated signal:
        int64 onset sample positions on the device."""
        yt = torch.as_tensor(y).to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        L = yt.numel()
        od = tables.onset_detect_params(self.sample_rate, int(hop_len))
        sp = _lib.GatSlicerParams(
            min_db_threshold=0.0, sample_gate=0.0, rms_hop=512, p20_k=0, p20_gamma=0.0, gate_offset_db=0.0, onset_hop=int(hop_len),
            pre_max=od["pre_max"], post_max=od["post_max"], pre_avg=od["pre_avg"], post_avg=od["post_avg"],
            wait=od["wait"], delta=float(od["delta"]), min_sep_samples=int(min_sep * self.sample_rate),
            attack_skip=0, clip_len=1, min_slice_rms_db=0.0)
        max_onsets = max(2, 1 + L // int(hop_len))
        onsets, n = self._empty((max_onsets,), torch.int64), self._empty((1,), torch.int32)
        with self._on_device():
            self.lib.check(self.lib.gat_detect_onsets(self._ctx, _lib.ptr(yt), L, C.byref(sp), max_onsets, _lib.ptr(onsets),
                                                      _lib.ptr(n), self._stream()), ValueError)
        return onsets[: int(n.item())]

    def segment(self, y, length_sec: float, cfg=None, diagnostics: bool = False) -> dict:
        """AudioSlicer.sliceNsave without file I/O (slicing.py:147-165) on one mono signal at the target rate."""
        yt = torch.as_tensor(y).to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        L = yt.numel()
        sp = self.slicer_params(L, length_sec, cfg)
        max_onsets = max(2, L // max(1, sp.min_sep_samples) + 2)
        To = 1 + L // 512
        out = {
            "onsets": self._empty((max_onsets,), torch.int64), "n_onsets": self._empty((1,), torch.int32),
            "clips": self._empty((max_onsets, sp.clip_len), torch.float32),
            "table": self._empty((max_onsets, 3), torch.int64), "n_clips": self._empty((1,), torch.int32),
        }
        diag = {}
        if diagnostics:
            diag = {"rms_db": self._empty((1 + L // sp.rms_hop,), torch.float32), "env": self._empty((To,), torch.float64),
                    "frames": self._empty((min(max_onsets, To),), torch.int64), "n_frames": self._empty((1,), torch.int32)}
        with self._on_device():
            self.lib.check(self.lib.gat_segment(
                self._ctx, _lib.ptr(yt), L, C.byref(sp), max_onsets, _lib.ptr(out["onsets"]), _lib.ptr(out["n_onsets"]),
                _lib.ptr(out["clips"]), _lib.ptr(out["table"]), _lib.ptr(out["n_clips"]),
                _lib.ptr(diag.get("rms_db")), _lib.ptr(diag.get("env")), _lib.ptr(diag.get("frames")),
                _lib.ptr(diag.get("n_frames")), self._stream()), ValueError)
        k = int(out["n_onsets"].item())
        m = int(out["n_clips"].item())
        res = {"onsets": out["onsets"][:k], "clips": out["clips"][:m], "table": out["table"][:m], "params": sp}
        if diagnostics:
            nf = int(diag["n_frames"].item())
            res.update(rms_db=diag["rms_db"], env=diag["env"], frames=diag["frames"][:nf])
        return res

    def segment_batch(self, signals, length_sec: float, cfg=None) -> dict:
        """``segment`` for P independent equal-length signals ``[P, L]`` in one pass of batched kernels
        (gat_segment_batch): each signal is sliced exactly as AudioSlicer.sliceNsave slices a file of its own.
        Returns device tensors: onsets [P, max_onsets] (first n_onsets[p] valid), n_onsets [P], clips [K, n] in
        (signal, onset) order, table [K, 4] = (signal, onset index, start, end), n_clips [P]."""
        Y = torch.as_tensor(signals).to(device=self.device, dtype=torch.float32).contiguous()
        if Y.dim() != 2:
            raise ValueError("segment_batch: signals must be [P, L]")
        P, L = Y.shape
        sp = self.slicer_params(L, length_sec, cfg)
        max_onsets = max(2, L // max(1, sp.min_sep_samples) + 2)
        out = {"onsets": self._empty((P, max_onsets), torch.int64), "n_onsets": self._empty((P,), torch.int32),
               "n_clips": self._empty((P + 1,), torch.int32)}
        if P == 0:
            return {"onsets": out["onsets"], "n_onsets": out["n_onsets"], "clips": self._empty((0, sp.clip_len), torch.float32),
                    "table": self._empty((0, 4), torch.int64), "n_clips": out["n_clips"][:0], "params": sp}
        # K onsets yield at most K - 1 clips (the last onset never does, slicing.py:154)
        max_clips = P * (max_onsets - 1)
        clips = self._empty((max_clips, sp.clip_len), torch.float32)
        table = self._empty((max_clips, 4), torch.int64)
        with self._on_device():
            for p0 in range(0, P, 65535):
                if p0:
                    raise ValueError("segment_batch: at most 65535 signals per call")
                self.lib.check(self.lib.gat_segment_batch(
                    self._ctx, _lib.ptr(Y), P, L, C.byref(sp), max_onsets, _lib.ptr(out["onsets"]), _lib.ptr(out["n_onsets"]),
                    _lib.ptr(clips), max_clips, _lib.ptr(table), _lib.ptr(out["n_clips"]), self._stream()), ValueError)
        k = int(out["n_clips"][P].item())
        return {"onsets": out["onsets"], "n_onsets": out["n_onsets"], "clips": clips[:k], "table": table[:k],
                "n_clips": out["n_clips"][:P], "params": sp}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
