#!/usr/bin/env python
"""Benchmark of the transcription hot path (BASELINE.json metric: audio-seconds transcribed per second).

    python bench.py [--config C] [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path
    python bench.py --impl reference [--config C] [--steps K] [--warmup W]  # the reference's CPU path (oracle port)

``--config`` picks the BASELINE.json configuration (1-based position in ``configs``; default 2 = the one the metric
is quoted on at N = 1):

  1  one 5 s phrase, full pipeline host to host (segmentation -> features -> ensemble -> YIN)     [replicas]
  2  4096 x 1 s note clips per GPU, mel spectrogram + CNN only                                    [weak]
  3  the same 4096 clips per GPU, MFCC + YIN + MLP + CNN ensemble; MLP / YIN / truth agreement    [weak]
  4  one hour of audio (720 phrases), onset detection + slicing + ensemble, sharded over the GPUs [strong]
  5  sweep N in {1e4, 1e5, 1e6} clips x {0.5, 1, 2, 4} s x n_fft {1024, 2048, 4096}, clips made on
     the device, sharded over the GPUs, with a host-CPU column                                    [strong]

A "step" is one pass of the path over the configuration's batch.  ``value`` times it with the inputs resident in
HBM (CUDA events, max over ranks); ``e2e`` times the call that takes HOST buffers (pinned), host->device and
device->host copies inside the timed region.  Work shards across ranks with no data-path collective; every step ends
with the all-gather of the per-clip records (labels, confidences, slice tables).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SR = 22050
CLIPS_PER_GPU = 4096
CLIP_SECONDS = 1.0
PHRASES_PER_HOUR = 720
PHRASE_SECONDS = 5.0
CKPT = ROOT / "tests" / "golden" / "ckpt"
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
WORKLOADS = {
    1: "configs[0]: one 5 s mono phrase (10 plucked notes), sr 22050: onset slicing -> mel + MFCC + YIN -> CNN + MLP ensemble -> labels, host to host",
    2: "configs[1]: 4096 x 1 s note clips per GPU, sr 22050, mel-spectrogram (n_fft 2048, hop 256, 64 mels, dB) + CNN, synthetic 8-harmonic plucks",
    3: "configs[2]: 4096 x 1 s note clips per GPU, sr 22050, MFCC-64 + log10(YIN) -> MLP, mel -> CNN, 0.2/0.8 ensemble; MLP vs YIN vs truth agreement",
    4: "configs[3]: 1 hour of synthetic multi-note audio (720 x 5 s phrases, sr 22050): onset detection + slicing + ensemble + YIN, sharded over the GPUs",
    5: "configs[4]: sweep N in {1e4,1e5,1e6} clips x {0.5,1,2,4} s x n_fft {1024,2048,4096} (hop 256, 64 mels), mel + CNN, clips generated on the device, sharded over the GPUs, label all-gather",
}


def make_clips(n_clips: int, seed0: int, seconds: float = CLIP_SECONDS):
    from guitar_audio_transcriber_ai_b200 import synth
    return synth.clip_batch(n_clips, seconds, SR, seed0=seed0)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions run (a 20-step resident region
    lasts ~55 ms, so the sampler stays on through the end-to-end regions as well: same kernels, same load)."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "clocks_event_reasons.hw_slowdown", "clocks_event_reasons.hw_thermal_slowdown",
              "clocks_event_reasons.sw_thermal_slowdown", "clocks_event_reasons.sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", "--query-gpu=" + ",".join(self.FIELDS), "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ============================================================================================== CPU reference arm
# The reference's own path for each configuration on the host cores, through oracle/port.py (the reference's control
# flow over genuine torchaudio / torch and the librosa restatement).  Used by --impl reference and by the
# cpu_baseline leg (rank 0, N = 1) of the CUDA arm, where its labels double as a parity guard.
def _port():
    sys.path.insert(0, str(ROOT / "oracle"))
    import port
    return port


def cpu_mel_cnn(clips: np.ndarray, cnn_ck, n_fft: int = 2048):
    """cfg 2 / 5: MelFeatureBuilder.extract_melspec_features (features.py:275-341: transforms built once, per-clip loop,
    pad + stack) followed by ONE batched CNN forward + softmax + argmax (note_predictor.py:102-123)."""
    port = _port()
    with torch.inference_mode():
        X = port.extract_melspec_features(list(clips), SR, n_mels=64, n_fft=n_fft, hop_length=256, normalize=True)
        probs = torch.softmax(port.cnn_forward(cnn_ck["model"], X), dim=-1).numpy()
    return np.argmax(probs, axis=1)


def cpu_full_ensemble(clips: np.ndarray, mlp_ck, cnn_ck):
    """cfg 3: the in-memory API per note (transcribe.py:147-199): MFCC + YIN features WITHOUT the scaler, mel image,
    both models, the ensemble.  Returns the label indices."""
    port = _port()
    idx = []
    for c in clips:
        r = port.transcribe_note(mlp_ck, cnn_ck, c, CLIP_SECONDS, SR)
        idx.append(int(r["indices"][0]))
    return np.asarray(idx)


def cpu_phrases(phrases, mlp_ck, cnn_ck):
    """cfg 1 / 4: Transcriber.transcribe from memory per phrase (transcribe.py:77-144 without the disk round trip)."""
    port = _port()
    out = []
    for y in phrases:
        out.append(port.transcribe_audio(mlp_ck, cnn_ck, y, SR, 0.5))
    return out


def load_ckpts():
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    return load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt"), load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")


def cpu_sample_inputs(cfg: int, args):
    """Bounded sample of configuration ``cfg`` for the CPU arm: (inputs, audio seconds per step, description).
    Synthesis happens here, outside every timed region."""
    from guitar_audio_transcriber_ai_b200 import synth
    if cfg == 1:
        return [synth.phrase(0, sr=SR)[0]], PHRASE_SECONDS, "the 5 s phrase itself (oracle port: whole pipeline)"
    if cfg == 2:
        n = args.cpu_sample or 256
        return make_clips(n, 0)[0], n * CLIP_SECONDS, (f"{n} of the 4096 clips per step (oracle port: torchaudio mel loop with the transforms built "
                                                       "once, as features.py:296-318, + one batched torch CNN forward)")
    if cfg == 3:
        n = args.cpu_sample or 64
        return make_clips(n, 0)[0], n * CLIP_SECONDS, (f"{n} of the 4096 clips per step (oracle port: transcribe_note per clip - librosa-restated "
                                                       "MFCC + YIN, torchaudio mel, torch MLP + CNN)")
    if cfg == 4:
        n = args.cpu_sample or 4
        return [synth.phrase(s, sr=SR)[0] for s in range(n)], n * PHRASE_SECONDS, (f"{n} of the 720 phrases per step (oracle port: gates, onsets, "
                                                                                   "slices, features, ensemble, YIN per phrase)")
    if cfg == 5:
        n = args.cpu_sample or 16
        grid = {(n_fft, dur): make_clips(n, 0, dur)[0] for n_fft in (1024, 2048, 4096) for dur in (0.5, 1.0, 2.0, 4.0)}
        return grid, sum(n * dur for (_, dur) in grid), f"{n} clips at each of the 12 (n_fft, duration) points per step (oracle port: torchaudio mel + torch CNN)"
    raise SystemExit(f"unknown config {cfg}")


def cpu_sample_step(cfg: int, inputs, mlp_ck, cnn_ck):
    """One CPU step over the prepared sample; returns the payload (labels / results / per-point rates)."""
    if cfg in (1, 4):
        return cpu_phrases(inputs, mlp_ck, cnn_ck)
    if cfg == 2:
        return cpu_mel_cnn(inputs, cnn_ck)
    if cfg == 3:
        return cpu_full_ensemble(inputs, mlp_ck, cnn_ck)
    rates = {}
    for (n_fft, dur), clips in inputs.items():
        t0 = time.perf_counter()
        cpu_mel_cnn(clips, cnn_ck, n_fft)
        rates[(n_fft, dur)] = len(clips) * dur / (time.perf_counter() - t0)
    return rates


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mlp_ck, cnn_ck = load_ckpts()
    inputs, audio_s, desc = cpu_sample_inputs(args.config, args)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_sample_step(args.config, inputs, mlp_ck, cnn_ck)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample_step(args.config, inputs, mlp_ck, cnn_ck)
    dt = time.perf_counter() - t0
    value = audio_s * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.config in (4, 5) else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOADS[args.config], "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ============================================================================================== CUDA arm
def bind_to_gpu_numa_node(local_rank: int) -> str:
    """One process per GPU: pin this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers
    are allocated (first-touch places them on that NUMA node).  Best effort: returns what happened for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() or 64) // 64 + 16)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        if not use:
            return f"gpu-local cpus not in this process's cpuset ({len(allowed)} cpus allowed)"
        os.sched_setaffinity(0, use)
        return f"bound to {len(use)} gpu-local cpus"
    except Exception as e:          # NVML missing, old driver, ...
        return f"not bound ({type(e).__name__})"


class Env:
    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                             "(use --impl reference for the host baseline)")
        import torch.distributed as dist
        self.dist = dist
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.device)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def stft_flops_per_frame(n_fft: int) -> float:
    """EXECUTED algorithmic work of one frame: real FFT (2.5 N log2 N) + the banded-sparse filterbank (each of the
    n_fft/2+1 bins feeds two triangular filters: 2 MACs = 4 flops per bin).  The dense-equivalent mel GEMM
    (2 * bins * n_mels) is NOT what the kernel executes and is not counted."""
    return 2.5 * n_fft * math.log2(n_fft) + 4.0 * (n_fft // 2 + 1)


def work_table(n_clips: int, n: int, T: int, n_fft: int = 2048, full: bool = False, T512: int | None = None, lags: int = 442):
    """kernel -> (contract bound, algorithmic bytes or flops per STEP on this rank, executed fp32 flops or None)."""
    H1, W1 = 32, T // 2
    H2, W2 = 16, W1 // 2
    T512 = T512 if T512 is not None else 1 + n // 512
    t = {
        "clip_scale_kernel": ("hbm", n_clips * (4 * n + 4), None),
        "stft_mel_f32_image": ("hbm", n_clips * (4 * n + 4 * 64 * T), n_clips * T * stft_flops_per_frame(n_fft)),
        "stft_frames_image": ("hbm", n_clips * (4 * n + 4 * 64 * T), n_clips * T * stft_flops_per_frame(n_fft)),
        "conv1_pool_planes_kernel": ("hbm", n_clips * (4 * 64 * T + 4 * H1 * W1 * 32), None),
        "conv2_tc_32_64": ("tensor", n_clips * 2.0 * H1 * W1 * 64 * 32 * 9, None),
        "conv12_tc_1_32_64": ("tensor", n_clips * 2.0 * H1 * W1 * 64 * 32 * 9, None),
        "conv3_tc_64_128": ("tensor", n_clips * 2.0 * H2 * W2 * 128 * 64 * 9, None),
        "avgpool_planes_kernel": ("hbm", n_clips * (4 * 8 * (T // 8) * 128 + 2 * 4 * 2048), None),
        "fc1_tc_2048_256": ("tensor", n_clips * 2.0 * 2048 * 256, None),
        "fc2_softmax_kernel": ("hbm", n_clips * (4 * 256 + 8 * 47), None),
        "argmax_kernel": ("hbm", n_clips * (4 * 47 + 12), None),
    }
    if full:
        t.update({
            "stft_mel_f32_spec": ("hbm", n_clips * (4 * n + 4 * 65), n_clips * T512 * stft_flops_per_frame(2048)),
            "mfcc_finish_kernel": ("hbm", n_clips * (4 * 128 * T512 + 4 * 64), None),
            # one FFT per image frame + 4 edge frames of the MFCC chain; the Slaney bank rides on every other image frame
            "stft_frames_dual": ("hbm", n_clips * (4 * n + 4 * 64 * T + 4 * 65), n_clips * ((T + 4) * stft_flops_per_frame(n_fft) + T512 * 4.0 * 1025)),
            "stft_frames_mfcc": ("hbm", n_clips * (4 * n + 4 * 65), n_clips * T512 * stft_flops_per_frame(2048)),
            # EXECUTED work of the block-FFT form (csrc/yin.cuh, yin_fft_kernel): per clip one forward complex 1024-point
            # transform per 512-sample block (T512 frames + one extra block per segment of ~12 frames), one inverse per frame
            # pair, 5 N log2 N = 51 200 FLOP each, plus the spectrum product (12 FLOP per bin and block).  (The direct form
            # this replaced executed 2 * 512 * 448 FLOP per block, ten times as much; SURVEY 8(d)'s naive per-frame form twice that.)
            "yin_kernel": ("hbm", n_clips * (4 * n + 8 * T512),
                           n_clips * ((T512 + -(-T512 // 12) + -(-T512 // 2)) * 51200.0 + (T512 + -(-T512 // 12)) * 1024 * 12.0)),
            "yin_median_kernel": ("hbm", n_clips * (8 * T512 + 12), None),
            "mlp_ensemble_kernel": ("hbm", n_clips * (4 * 65 + 4 * 47 * 3 + 12), None),
        })
    return t


def kernel_rows(prof: dict, steps: int, work: dict, peaks: dict, fma_peak: float | None):
    rows = []
    total = sum(ms for _, ms in prof.values()) or 1.0
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        per_step_ms = ms / steps
        bound, amount, flops = work.get(name, ("hbm", 0.0, None))
        if bound == "hbm":
            ach = amount / (per_step_ms * 1e-3) / 1e9 if per_step_ms > 0 else 0.0
            peak, unit = peaks["hbm_gbs"], "GB/s"
        else:
            ach = amount / (per_step_ms * 1e-3) / 1e12 if per_step_ms > 0 else 0.0
            peak, unit = peaks["tflops_sustained"], "TFLOP/s"
        row = {"kernel": name, "launches_per_step": cnt / steps, "ms_per_step": per_step_ms, "share": ms / total,
               "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak if peak else None}
        if flops and fma_peak and per_step_ms > 0:
            f = flops / (per_step_ms * 1e-3) / 1e12
            row["fp32"] = {"achieved": f, "peak": fma_peak, "unit": "TFLOP/s", "frac": f / fma_peak,
                           "what": "executed algorithmic FP32 work (rFFT 2.5 N log2 N + 2 MAC per bin sparse mel; YIN: block-FFT difference form, 1.6 complex 1024-point FFTs per frame) vs the FP32-FMA peak measured live"}
        rows.append(row)
    return rows


def roofline_from(rows, peaks):
    top = rows[0]
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        traffic = json.loads(tfile.read_text()).get(top["kernel"])
    r = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
         "frac": top["frac"], "traffic": traffic,
         "peak_source": peaks["source"] + (" (sustained bf16)" if top["bound"] == "tensor" else " (copy)"),
         "share_of_step": top["share"]}
    if "fp32" in top:
        r["fp32"] = top["fp32"]
    return r


# ---------------------------------------------------------------------------------------------- configurations
class ClipConfig:
    """cfg 2 (mel + CNN) and cfg 3 (full ensemble) on 4096 one-second clips per GPU, weak scaling."""
    scaling = "weak"

    def __init__(self, env: Env, full: bool):
        from guitar_audio_transcriber_ai_b200 import parallel
        from guitar_audio_transcriber_ai_b200.engine import Engine
        self.env, self.full, self.parallel = env, full, parallel
        self.mlp_ck, self.cnn_ck = load_ckpts()
        self.eng = Engine(SR, device=env.device)
        self.eng.load_cnn(self.cnn_ck["model"]); self.eng.load_mlp(self.mlp_ck["model"])
        self.n_total = CLIPS_PER_GPU * env.world
        self.lo, self.hi = parallel.shard_bounds(self.n_total, env.world, env.rank)
        clips, self.midi = make_clips(self.hi - self.lo, seed0=self.lo)
        self.host = torch.from_numpy(clips).pin_memory()
        self.dev = self.host.to(env.device)
        self.n = self.host.shape[1]
        self.T = self.eng.mel_frames(self.n)
        self.engines = [self.eng]
        self.last = None

    def audio_seconds(self):
        return self.n_total * CLIP_SECONDS

    def step_resident(self):
        out = self.eng.transcribe_clips(self.dev, skip_mlp=not self.full, yin_on_normalized=True)
        rec = self.parallel.pack_records(out["indices"], out["confidences"])
        self.last = self.parallel.all_gather_records(rec, self.n_total)
        return self.last

    def step_host(self, buf=None):
        out = self.eng.transcribe_clips_host(self.host if buf is None else buf, skip_mlp=not self.full, want_probs=False)
        if self.env.world > 1:
            rec = self.parallel.pack_records(torch.from_numpy(out["indices"]).to(self.env.device),
                                             torch.from_numpy(out["confidences"]).to(self.env.device))
            self.parallel.all_gather_records(rec, self.n_total)
        return out["h2d_bytes"], out["d2h_bytes"]

    def profile_step(self):
        self.eng.transcribe_clips(self.dev, skip_mlp=not self.full, yin_on_normalized=True)

    def work(self):
        return work_table(self.hi - self.lo, self.n, self.T, full=self.full)

    def checksum(self):
        return int(self.last[:, 0].sum().item())

    def config_extra(self):
        d = {"clips_per_gpu": CLIPS_PER_GPU, "clip_seconds": CLIP_SECONDS, "sample_rate": SR,
             "parallelism": f"clip-sharded x{self.env.world}, label all-gather",
             "l2": "inputs (361 MB per GPU) exceed the 126 MB L2"}
        if self.full and self.env.rank == 0:
            d["agreement"] = self.agreement()
        return d

    def agreement(self):
        """cfg 3's check: MLP (MFCC + pitch features) vs the YIN DSP baseline vs the synthesised truth, on this rank's clips."""
        from guitar_audio_transcriber_ai_b200 import synth
        out = self.eng.transcribe_clips(self.dev, yin_on_normalized=True, return_features=True)
        names = list(self.mlp_ck["reverse_map"][i] for i in range(len(self.mlp_ck["reverse_map"])))
        label_midi = np.array([next(m for m in range(synth.MIDI_LO, synth.MIDI_HI + 1) if synth.midi_to_label(m) == str(nm)) for nm in names])
        truth = self.midi
        ens = label_midi[out["indices"].cpu().numpy()]
        mlp = label_midi[out["mlp_probs"].argmax(dim=1).cpu().numpy()]
        cnn = label_midi[out["cnn_probs"].argmax(dim=1).cpu().numpy()]
        hz = out["yin_hz"].cpu().numpy()
        yin = np.round(12 * np.log2(hz / 440.0) + 69).astype(np.int64)
        return {"clips": int(len(truth)), "ensemble_vs_truth": float((ens == truth).mean()), "mlp_vs_truth": float((mlp == truth).mean()),
                "cnn_vs_truth": float((cnn == truth).mean()), "yin_vs_truth": float((yin == truth).mean()),
                "mlp_vs_yin": float((mlp == yin).mean()), "ensemble_vs_yin": float((ens == yin).mean()),
                "yin_octave_errors": float((np.abs(yin - truth) % 12 == 0)[yin != truth].mean()) if (yin != truth).any() else 0.0}

    def cpu_guard(self, args):
        sample = args.cpu_sample or (64 if self.full else 1024)
        sub = self.host[:sample].numpy()
        fn = (lambda x: cpu_full_ensemble(x, self.mlp_ck, self.cnn_ck)) if self.full else (lambda x: cpu_mel_cnn(x, self.cnn_ck))
        fn(sub[:8])
        t0 = time.perf_counter()
        labels = fn(sub)
        dt = time.perf_counter() - t0
        mism = int((labels != self.last[:sample, 0].cpu().numpy()).sum())
        what = ("transcribe_note per clip - librosa-restated MFCC + YIN, torchaudio mel, torch MLP + CNN" if self.full else
                "torchaudio mel loop with the transforms built once, as features.py:296-318, + one batched torch CNN forward")
        return {"value": sample * CLIP_SECONDS / dt, "sample": f"first {sample} of the 4096 clips, one pass, {dt:.1f} s (oracle port: {what})",
                "labels_compared": sample, "label_mismatches": mism}


class PhraseConfig:
    """cfg 4: an hour of audio as 720 five-second phrases, sharded over the ranks (strong scaling), and cfg 1: one phrase."""

    def __init__(self, env: Env, n_phrases: int, mode: str):
        from guitar_audio_transcriber_ai_b200 import Transcriber, parallel, synth
        self.env, self.parallel, self.mode = env, parallel, mode
        self.single = n_phrases == 1
        self.scaling = "weak" if self.single else "strong"
        self.tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", CKPT, CKPT, device=str(env.device))
        self.mlp_ck, self.cnn_ck = self.tr.model_ckpts["mlp"], self.tr.model_ckpts["cnn"]
        self.eng = self.tr.engine
        self.engines = [self.eng]
        self.P = n_phrases
        if self.single:                       # replicas: every rank transcribes its own phrase
            self.lo, self.hi = env.rank, env.rank + 1
        elif mode == "file":                  # one contiguous signal on every rank
            self.lo, self.hi = 0, n_phrases
        else:
            self.lo, self.hi = parallel.shard_bounds(n_phrases, env.world, env.rank)
        Y = np.stack([synth.phrase(s, sr=SR)[0] for s in range(self.lo, self.hi)]) if self.hi > self.lo else np.zeros((0, int(SR * PHRASE_SECONDS)), np.float32)
        self.L = Y.shape[1]
        self.host = torch.from_numpy(Y).pin_memory()
        # rotate over enough device copies that consecutive steps never find their input in the 126 MB L2
        copies = max(1, min(8, int(math.ceil(300e6 / max(1, Y.nbytes)))))
        self.dev = [self.host.to(env.device).clone() for _ in range(copies)]
        self.k = 0
        self.last = None

    def audio_seconds(self):
        return (self.env.world if self.single else self.P) * PHRASE_SECONDS

    def _rows(self, Y):
        if self.single:
            return self.parallel.phrases_rows_device(self.tr, Y, 0.5, signal_offset=self.lo, gather=False)
        if self.mode == "file":
            return self.parallel.audio_rows_device(self.tr, Y.reshape(-1), 0.5)
        return self.parallel.phrases_rows_device(self.tr, Y, 0.5, signal_offset=self.lo,
                                                 signals_per_rank=-(-self.P // self.env.world))

    def _valid(self):
        """The last step's gathered rows without the all-gather's padding."""
        return self.parallel.valid_rows(self.last)

    def step_resident(self):
        self.k += 1
        self.last = self._rows(self.dev[self.k % len(self.dev)])
        return self.last

    def step_host(self):
        """Host signal in (pinned), host rows out: the H2D copy of this rank's signals and the D2H copy of the gathered
        rows are inside the call."""
        Y = self.host.to(self.env.device, non_blocking=True)
        rows = self._rows(Y).cpu()
        return self.host.numel() * 4, rows.numel() * 8

    def profile_step(self):
        self._rows(self.dev[0])

    def work(self):
        n_clips = int(self._valid().shape[0]) if self.last is not None else 0
        mine = n_clips if (self.single or self.mode == "file") else max(1, n_clips // self.env.world)
        n = int(0.5 * SR)
        t = work_table(mine, n, 1 + n // 256, full=True)
        Ltot = (self.hi - self.lo) * self.L
        To = (self.hi - self.lo) * (1 + self.L // 512)
        t.update({
            "stft_mel_f64_spec": ("hbm", 4 * Ltot + 8 * 128 * To, To * stft_flops_per_frame(2048)),
            "rms_db_kernel": ("hbm", 4 * Ltot + 4 * To, None),
            "slice_gather_kernel": ("hbm", mine * 2 * 4 * n, None),
        })
        return t

    def checksum(self):
        return int(self._valid()[:, 4].sum().item())

    def config_extra(self):
        d = {"phrases": self.P, "phrase_seconds": PHRASE_SECONDS, "sample_rate": SR, "clips_found": int(self._valid().shape[0]),
             "l2": f"{len(self.dev)} device copies of the input rotated between steps (> 126 MB L2 in total)"}
        if self.single:
            d["parallelism"] = f"replicas x{self.env.world} (one phrase per GPU, no collective)"
        elif self.mode == "file":
            d["parallelism"] = (f"SURVEY 8(e) option (ii): whole-file segmentation on every rank (serial term), sliced clips sharded x{self.env.world}, "
                                "all-gather of label / slice-table rows")
        else:
            d["parallelism"] = (f"SURVEY 8(e) option (i): phrases as independent files sharded x{self.env.world} (batched segmentation + ensemble per rank), "
                                "all-gather of label / slice-table rows")
        return d

    def cpu_guard(self, args):
        from guitar_audio_transcriber_ai_b200 import synth
        sample = 1 if self.single else (args.cpu_sample or 4)
        phrases = [synth.phrase(s, sr=SR)[0] for s in range(sample)]
        t0 = time.perf_counter()
        res = cpu_phrases(phrases, self.mlp_ck, self.cnn_ck)
        dt = time.perf_counter() - t0
        rows = self._valid().cpu().numpy()
        mism = compared = 0
        table_ok = True
        if not (self.mode == "file" and not self.single):      # whole-file segmentation is not comparable phrase by phrase
            for p, r in enumerate(res):
                mine = rows[rows[:, 0] == p]
                compared += len(r["indices"])
                if len(mine) != len(r["indices"]):
                    table_ok = False
                    mism += abs(len(mine) - len(r["indices"]))
                    continue
                mism += int((mine[:, 4] != r["indices"]).sum())
                table_ok &= bool(np.array_equal(mine[:, 1:4], r["slice_table"]))
        return {"value": sample * PHRASE_SECONDS / dt,
                "sample": f"first {sample} of the {self.P} phrases, one pass, {dt:.1f} s (oracle port: gates, onsets, slices, features, ensemble, YIN per phrase)",
                "labels_compared": compared, "label_mismatches": mism, "slice_tables_equal": table_ok}


class SweepConfig:
    """cfg 5: N x duration x n_fft sweep, clips generated on the device (Philox), sharded over the ranks."""
    scaling = "strong"

    def __init__(self, env: Env, args):
        from guitar_audio_transcriber_ai_b200 import parallel
        from guitar_audio_transcriber_ai_b200.engine import Engine
        self.env, self.parallel, self.Engine = env, parallel, Engine
        self.mlp_ck, self.cnn_ck = load_ckpts()
        self.Ns = [int(float(x)) for x in args.sweep_n.split(",")]
        self.durs = [float(x) for x in args.sweep_dur.split(",")]
        self.nffts = [int(x) for x in args.sweep_nfft.split(",")]
        self.batch_bytes = args.sweep_batch_gb * 1e9
        self.engines = []
        self.by_nfft = {}
        for n_fft in self.nffts:
            e = Engine(SR, {"N_MELS": 64, "N_FFT": n_fft, "HOP_LENGTH": 256}, {"N_MFCC": 64}, device=env.device)
            e.load_cnn(self.cnn_ck["model"])
            self.by_nfft[n_fft] = e
            self.engines.append(e)
        self.eng = self.by_nfft[self.nffts[0]]
        self.points = []
        self.last = None

    def audio_seconds(self):
        return sum(N * d for N in self.Ns for d in self.durs) * len(self.nffts)

    def make_batch(self, n_clips: int, n: int, gen: torch.Generator) -> torch.Tensor:
        """Decaying 8-harmonic notes + noise, MIDI 40..86, float32 [n_clips, n] (the recipe of synth.note, on the device)."""
        dev = self.env.device
        midi = torch.randint(40, 87, (n_clips, 1), device=dev, generator=gen)
        f0 = 440.0 * torch.exp2((midi.float() - 69.0) / 12.0)
        t = torch.arange(n, device=dev, dtype=torch.float32)[None, :] / SR
        y = torch.zeros(n_clips, n, device=dev)
        phase = torch.rand(n_clips, 8, device=dev, generator=gen) * (2 * math.pi)
        for k in range(1, 9):
            y += (1.0 / k) * torch.sin(2 * math.pi * k * f0 * t + phase[:, k - 1:k]) * torch.exp(-t * (k ** 0.5) / 0.6)
        y *= 0.5 / y.abs().amax(dim=1, keepdim=True)
        y += 1e-3 * torch.randn(n_clips, n, device=dev, generator=gen)
        return y

    def run_sweep(self, record: bool):
        """One pass over the grid.  Returns the device milliseconds spent inside transcribe calls (generation excluded)."""
        env = self.env
        total_ms = 0.0
        points = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for n_fft in self.nffts:
            eng = self.by_nfft[n_fft]
            for dur in self.durs:
                n = int(SR * dur)
                per_batch = max(148, int(self.batch_bytes / (4 * n)) // 148 * 148)
                for N in self.Ns:
                    lo, hi = self.parallel.shard_bounds(N, env.world, env.rank)
                    gen = torch.Generator(device=env.device); gen.manual_seed(1234 + env.rank)
                    done, ms = lo, 0.0
                    warm = True
                    point_rec = torch.zeros((hi - lo, 4), dtype=torch.int64, device=env.device)
                    while done < hi:
                        b = min(per_batch, hi - done)
                        clips = self.make_batch(b, n, gen)
                        if warm:                       # untimed: sizes the workspaces for this batch shape
                            eng.transcribe_clips(clips, skip_mlp=True); warm = False
                        torch.cuda.synchronize(env.device)
                        e0.record()
                        out = eng.transcribe_clips(clips, skip_mlp=True)
                        point_rec[done - lo: done - lo + b] = self.parallel.pack_records(out["indices"], out["confidences"])
                        e1.record()
                        torch.cuda.synchronize(env.device)
                        ms += e0.elapsed_time(e1)
                        done += b
                        del clips
                    # the label all-gather of the point: every rank contributes its ceil(N/G) records
                    e0.record()
                    self.last = self.parallel.all_gather_records(point_rec, N)
                    e1.record()
                    torch.cuda.synchronize(env.device)
                    ms += e0.elapsed_time(e1)
                    ms = env.max_over_ranks(ms)
                    total_ms += ms
                    points.append({"n_fft": n_fft, "dur_s": dur, "N": N, "ms": round(ms, 3), "audio_s_per_s": round(N * dur / (ms * 1e-3))})
        if record:
            self.points = points
        return total_ms

    def work(self):
        return {}

    def checksum(self):
        return int(self.last[:, 0].sum().item()) if self.last is not None else 0

    def config_extra(self):
        return {"N": self.Ns, "durations_s": self.durs, "n_fft": self.nffts, "sample_rate": SR,
                "parallelism": f"clips sharded x{self.env.world}, generated on the device per rank (Philox), label all-gather per point",
                "l2": f"batches of up to {self.batch_bytes / 1e9:.0f} GB of fresh clips (>> 126 MB L2)",
                "timing": "CUDA events around the transcribe calls of every batch (generation excluded) + the all-gather, max over ranks per point"}


def run_ours(args):
    env = Env(args)
    peaks = measured_peaks()
    cfg_no = args.config
    if cfg_no in (2, 3):
        cfg = ClipConfig(env, full=(cfg_no == 3))
    elif cfg_no == 4:
        cfg = PhraseConfig(env, args.phrases, args.cfg4_mode)
    elif cfg_no == 1:
        cfg = PhraseConfig(env, 1, "phrases")
    elif cfg_no == 5:
        return run_sweep(env, SweepConfig(env, args), args, peaks)
    else:
        raise SystemExit(f"unknown --config {cfg_no}")
    warmup = max(args.warmup, 3)

    for _ in range(warmup):
        cfg.step_resident()
    env.barrier()
    launches0 = sum(e.launch_count for e in cfg.engines)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(env.local_rank)
    clocks.__enter__()
    env.barrier()
    e0.record()
    for _ in range(args.steps):
        cfg.step_resident()
    e1.record()
    env.barrier()
    launches = sum(e.launch_count for e in cfg.engines) - launches0
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    value = cfg.audio_seconds() * args.steps / (ms * 1e-3)

    # end to end through the host-buffer entry point (pinned host memory -> HBM -> kernels -> host)
    for _ in range(warmup):
        cfg.step_host()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h2d, d2h = cfg.step_host()
    torch.cuda.synchronize(env.device)
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    clocks.__exit__(None, None, None)
    h2d, d2h = env.sum_over_ranks(h2d), env.sum_over_ranks(d2h)
    e2e = {"value": cfg.audio_seconds() * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": 1e3 * e2e_s / args.steps}

    e2e16 = None
    if cfg_no == 2:
        # the same call fed PCM_16 clips (the format .wav files hold): extra information, not the headline.
        host16 = torch.clamp(torch.round(cfg.host * 32767.0), -32768, 32767).to(torch.int16).pin_memory()
        cfg.step_host(host16)
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            h16, d16 = cfg.step_host(host16)
        torch.cuda.synchronize(env.device)
        e2e16_s = env.max_over_ranks(time.perf_counter() - t0)
        e2e16 = {"value": cfg.audio_seconds() * args.steps / e2e16_s, "unit": UNIT, "h2d_bytes_per_step": int(env.sum_over_ranks(h16)),
                 "d2h_bytes_per_step": int(env.sum_over_ranks(d16)), "ms_per_step": 1e3 * e2e16_s / args.steps,
                 "note": "same C-ABI path fed int16 PCM host clips (gat_transcribe_clips_host_pcm16); extra, not the headline"}

    # per-kernel timing (separate pass: event pairs around every launch perturb the total slightly)
    prof_steps = 2
    env.barrier()
    cfg.eng.profile_begin()
    for _ in range(prof_steps):
        cfg.profile_step()
    prof = cfg.eng.profile_end()
    fma_peak = cfg.eng.fma_peak_tflops()
    rows = kernel_rows(prof, prof_steps, cfg.work(), peaks, fma_peak)
    roofline = roofline_from(rows, peaks)

    cpu_baseline = None
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_baseline = {"unit": UNIT, "cores": cores, "kind": "port", **cfg.cpu_guard(args)}

    extra = cfg.config_extra()
    if env.rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": env.world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": cfg.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[cfg_no], **extra, "host_affinity_rank0": env.numa, "labels_checksum": cfg.checksum()},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roofline, "kernels": rows,
            "cpu_baseline": cpu_baseline,
        }
        if e2e16:
            line["e2e_pcm16"] = e2e16
        print(json.dumps(line))
    if env.world > 1:
        env.dist.destroy_process_group()


def run_sweep(env: Env, cfg: SweepConfig, args, peaks):
    """cfg 5 has its own harness: the timed region is the sum of the per-batch CUDA-event intervals of one pass over the
    grid (generation of the next batch sits between them), so ``value`` = all audio seconds of the grid / that time."""
    cfg.run_sweep(record=False) if args.warmup > 0 and args.sweep_warm_pass else None
    env.barrier()
    launches0 = sum(e.launch_count for e in cfg.engines)
    total_ms = 0.0
    with ClockSampler(env.local_rank) as clocks:
        for _ in range(args.steps):
            total_ms += cfg.run_sweep(record=True)
    launches = sum(e.launch_count for e in cfg.engines) - launches0
    value = cfg.audio_seconds() * args.steps / (total_ms * 1e-3)
    cpu_rates, cpu_baseline = {}, None
    if env.rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        inputs, audio_s, desc = cpu_sample_inputs(5, args)
        cpu_mel_cnn(next(iter(inputs.values()))[:4], cfg.cnn_ck)
        t0 = time.perf_counter()
        cpu_rates = cpu_sample_step(5, inputs, cfg.mlp_ck, cfg.cnn_ck)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": audio_s / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc + f", {dt:.1f} s"}
        for p in cfg.points:
            r = cpu_rates.get((p["n_fft"], p["dur_s"]))
            if r:
                p["cpu_audio_s_per_s"] = round(r, 1)
                p["vs_cpu"] = round(p["audio_s_per_s"] / r, 1)
    # roofline of the dominant kernel at the reference's own point of the grid (n_fft 2048, 1 s), this rank's share
    eng = cfg.by_nfft.get(2048, cfg.eng)
    n = int(SR * 1.0)
    gen = torch.Generator(device=env.device); gen.manual_seed(7)
    clips = cfg.make_batch(4096, n, gen)
    eng.transcribe_clips(clips, skip_mlp=True)
    eng.profile_begin()
    for _ in range(2):
        eng.transcribe_clips(clips, skip_mlp=True)
    prof = eng.profile_end()
    rows = kernel_rows(prof, 2, work_table(4096, n, 1 + n // 256, eng.n_fft), peaks, eng.fma_peak_tflops())
    if env.rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": env.world, "steps": args.steps, "warmup": 1 if args.sweep_warm_pass else 0,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": cfg.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[5], **cfg.config_extra(), "host_affinity_rank0": env.numa, "labels_checksum": cfg.checksum(),
                       "warmup_note": "every (shape, n_fft) point runs one untimed batch first (workspace sizing)"},
            "e2e": None, "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roofline_from(rows, peaks),
            "kernels": rows, "cpu_baseline": cpu_baseline, "sweep": cfg.points,
        }))
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    # Exactly ONE line may reach stdout (the JSON record): libraries such as NCCL print banners to the C-level
    # stdout, so route fd 1 to stderr for the duration of the run and emit the record on the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (1-based); default 2 = configs[1]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="clips / phrases the CPU arm processes per step (0 = per-config default)")
    ap.add_argument("--phrases", type=int, default=PHRASES_PER_HOUR, help="cfg 4: phrases of 5 s (720 = one hour)")
    ap.add_argument("--cfg4-mode", choices=["phrases", "file"], default="phrases",
                    help="cfg 4: phrases = SURVEY 8(e) option (i), independent phrases sharded; file = option (ii), one contiguous signal")
    ap.add_argument("--sweep-n", default="1e4,1e5,1e6")
    ap.add_argument("--sweep-dur", default="0.5,1,2,4")
    ap.add_argument("--sweep-nfft", default="1024,2048,4096")
    ap.add_argument("--sweep-batch-gb", type=float, default=8.0)
    ap.add_argument("--sweep-warm-pass", action="store_true", help="cfg 5: run the whole grid once untimed first")
    args = ap.parse_args()
    if args.config == 5 and args.steps == 10:
        args.steps = 1                     # one pass over the 36-point grid is the default step count of the sweep
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
