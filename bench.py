#!/usr/bin/env python
"""Benchmark of the transcription hot path (BASELINE.json metric: audio-seconds transcribed per second).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)

Workload at N = 1 is BASELINE.json configs[1]: a batch of 4,096 one-second note clips (sr 22050), mel
spectrogram (n_fft 2048, hop 256, 64 mels, dB) + CNN, per GPU.  A "step" is one pass of that path over the
batch.  `value` times the path with the clips already resident in HBM; `e2e` times the C-ABI call that takes
HOST buffers (pinned), host->device and device->host copies inside the timed region.  Clips shard across
ranks with no data-path collective; each step ends with the all-gather of the per-clip label records.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
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
CKPT = ROOT / "tests" / "golden" / "ckpt"
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
WORKLOAD = ("configs[1]: 4096 x 1 s note clips per GPU, sr 22050, mel-spectrogram (n_fft 2048, hop 256, 64 mels, dB) "
            "+ CNN, synthetic 8-harmonic plucks")


def make_clips(n_clips: int, seed0: int) -> np.ndarray:
    from guitar_audio_transcriber_ai_b200 import synth
    clips, _ = synth.clip_batch(n_clips, CLIP_SECONDS, SR, seed0=seed0)
    return clips


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "clocks_event_reasons.hw_slowdown", "clocks_event_reasons.hw_thermal_slowdown",
              "clocks_event_reasons.sw_thermal_slowdown", "clocks_event_reasons.sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", "--query-gpu=" + ",".join(self.FIELDS), "--format=csv,noheader,nounits", "-lms", "200"],
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


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pass(clips: np.ndarray, cnn_ck):
    """The reference's path for this workload on the host: MelFeatureBuilder.extract_melspec_features' per-clip
    loop (features.py:307-331) followed by ONE batched CNN forward + softmax + argmax (note_predictor.py:102-123)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import port
    with torch.inference_mode():
        specs = [port.melspec_image(c, SR) for c in clips]
        X = torch.stack(specs, dim=0)
        probs = torch.softmax(port.cnn_forward(cnn_ck["model"], X), dim=-1).numpy()
    return np.argmax(probs, axis=1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cnn_ck = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    sample = 256
    clips = make_clips(sample, 0)
    for _ in range(args.warmup):
        cpu_reference_pass(clips[:32], cnn_ck)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(clips, cnn_ck)
    dt = time.perf_counter() - t0
    value = sample * CLIP_SECONDS * args.steps / dt
    desc = f"{sample} of the 4096 clips per step (oracle port: torchaudio mel + torch CNN, genuine libraries)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample_clips_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- CUDA arm
def kernel_table(prof: dict, steps: int, n_clips: int, T: int, peaks: dict):
    """Per-kernel averages from the profiled pass + the algorithmic work each one does per step."""
    n = int(SR * CLIP_SECONDS)
    H1, W1 = 32, T // 2
    H2, W2 = 16, W1 // 2
    work = {   # kernel -> (bound, algorithmic bytes or flops per STEP)
        "clip_scale_kernel": ("hbm", n_clips * (4 * n + 4)),
        "stft_mel_f32_image": ("hbm", n_clips * (4 * n + 4 * 64 * T)),
        "conv1_pool_planes_kernel": ("hbm", n_clips * (4 * 64 * T + 4 * H1 * W1 * 32)),       # fp32 image in, hf + lb (2 x 16 bit) planes out
        "conv2_tc_32_64": ("tensor", n_clips * 2.0 * H1 * W1 * 64 * 32 * 9),
        "conv3_tc_64_128": ("tensor", n_clips * 2.0 * H2 * W2 * 128 * 64 * 9),
        "avgpool_planes_kernel": ("hbm", n_clips * (4 * 8 * (T // 8) * 128 + 2 * 4 * 2048)),   # only launched for clips longer than one conv3 group
        "fc1_tc_2048_256": ("tensor", n_clips * 2.0 * 2048 * 256),
        "fc2_softmax_kernel": ("hbm", n_clips * (4 * 256 + 8 * 47)),
        "argmax_kernel": ("hbm", n_clips * (4 * 47 + 12)),
    }
    rows = []
    total = sum(ms for _, ms in prof.values()) or 1.0
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        per_step_ms = ms / steps
        bound, amount = work.get(name, ("hbm", 0.0))
        if bound == "hbm":
            ach = amount / (per_step_ms * 1e-3) / 1e9 if per_step_ms > 0 else 0.0
            peak, unit = peaks["hbm_gbs"], "GB/s"
        else:
            ach = amount / (per_step_ms * 1e-3) / 1e12 if per_step_ms > 0 else 0.0
            peak, unit = peaks["tflops_sustained"], "TFLOP/s"
        rows.append({"kernel": name, "launches_per_step": cnt / steps, "ms_per_step": per_step_ms, "share": ms / total,
                     "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak if peak else None})
    return rows


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """One process per GPU: pin this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers
    are allocated (first-touch places them on that NUMA node), so the host->device copies of different ranks do not
    all cross the same socket link.  Best effort: returns what happened for the JSON line."""
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


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the host baseline)")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from guitar_audio_transcriber_ai_b200 import parallel
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    from guitar_audio_transcriber_ai_b200.engine import Engine

    eng = Engine(SR, device=device)
    cnn_ck = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    mlp_ck = load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt")
    eng.load_cnn(cnn_ck["model"]); eng.load_mlp(mlp_ck["model"])
    n_total = CLIPS_PER_GPU * world
    lo, hi = parallel.shard_bounds(n_total, world, rank)
    host = torch.from_numpy(make_clips(hi - lo, seed0=lo)).pin_memory()
    dev = host.to(device, non_blocking=False)
    n = host.shape[1]
    T = eng.mel_frames(n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step_resident():
        out = eng.transcribe_clips(dev, skip_mlp=True)
        rec = parallel.pack_records(out["indices"], out["confidences"])
        return parallel.all_gather_records(rec, n_total)

    def step_host(buf=None):
        out = eng.transcribe_clips_host(host if buf is None else buf, skip_mlp=True, want_probs=False)
        if world > 1:
            rec = parallel.pack_records(torch.from_numpy(out["indices"]).to(device), torch.from_numpy(out["confidences"]).to(device))
            parallel.all_gather_records(rec, n_total)
        return out

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 1)):
        step_resident()
    barrier()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            rec = step_resident()
        e1.record()
        barrier()
    launches = eng.launch_count - launches0
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = n_total * CLIP_SECONDS * args.steps / (ms * 1e-3)
    labels_checksum = int(rec[:, 0].sum().item())

    # end to end through the host-buffer C-ABI call (pinned host memory -> HBM -> kernels -> host)
    for _ in range(max(args.warmup, 1)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ho = step_host()
    torch.cuda.synchronize(device)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_total * CLIP_SECONDS * args.steps / e2e_s

    # the same call fed PCM_16 clips (the format .wav files hold): extra information, not the headline - the float32
    # call above is copy-bound on the host link and this halves the bytes.  Labels can differ from the float32 run
    # only through the 16-bit quantisation of the input.
    host16 = torch.clamp(torch.round(host * 32767.0), -32768, 32767).to(torch.int16).pin_memory()
    step_host(host16)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ho16 = step_host(host16)
    torch.cuda.synchronize(device)
    e2e16_s = max_over_ranks(time.perf_counter() - t0)

    # per-kernel timing (separate pass: event pairs around every launch perturb the total slightly)
    peaks = measured_peaks()
    prof_steps = 2
    barrier()
    eng.profile_begin()
    for _ in range(prof_steps):
        eng.transcribe_clips(dev, skip_mlp=True)
    prof = eng.profile_end()
    rows = kernel_table(prof, prof_steps, hi - lo, T, peaks)
    top = rows[0]
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        traffic = json.loads(tfile.read_text()).get(top["kernel"])
    roofline = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
                "frac": top["frac"], "traffic": traffic, "peak_source": peaks["source"] + (" (sustained bf16)" if top["bound"] == "tensor" else " (copy)"),
                "share_of_step": top["share"]}
    if top["kernel"].startswith("stft_mel"):
        # The STFT chain is nominally HBM-bound (the contract's roof, above) but at n_fft 2048 / hop 256 it does 148 FLOP
        # per algorithmic byte: its real roof is FP32 issue.  Report that too, against the FP32-FMA peak MEASURED on
        # this device (gat_debug_fma_peak; MEASURED_PEAKS.json has no FP32 figure) - SURVEY 8(d) "reporting rule".
        fma_peak = eng.fma_peak_tflops()
        flops = (hi - lo) * T * (2.5 * 2048 * 11 + 2.0 * 1025 * 64)        # rFFT 2.5 N log2 N + dense-equivalent mel GEMM
        ach = flops / (top["ms_per_step"] * 1e-3) / 1e12
        roofline["fp32"] = {"achieved": ach, "peak": fma_peak, "unit": "TFLOP/s", "frac": ach / fma_peak,
                            "peak_source": "measured live (register-only FMA kernel)", "flops_per_frame": 2.5 * 2048 * 11 + 2.0 * 1025 * 64}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = 1024
        sub = host[:sample].numpy()
        cpu_reference_pass(sub[:16], cnn_ck)
        t0 = time.perf_counter()
        cpu_labels = cpu_reference_pass(sub, cnn_ck)
        dt = time.perf_counter() - t0
        # the same pass doubles as a parity guard: the CPU oracle's labels for these clips vs the timed CUDA path's
        mismatches = int((cpu_labels != rec[:sample, 0].cpu().numpy()).sum())
        cpu_baseline = {"value": sample * CLIP_SECONDS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {sample} of the 4096 clips, one pass, {dt:.1f} s (oracle port: per-clip torchaudio mel loop + one batched torch CNN forward)",
                        "labels_compared": sample, "label_mismatches": mismatches}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": CLIPS_PER_GPU, "clip_seconds": CLIP_SECONDS, "sample_rate": SR,
                       "parallelism": f"clip-sharded x{world}, label all-gather", "host_affinity_rank0": numa, "l2": "inputs (361 MB per GPU) exceed the 126 MB L2",
                       "labels_checksum": labels_checksum},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ho["h2d_bytes"] * world, "d2h_bytes_per_step": ho["d2h_bytes"] * world,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "e2e_pcm16": {"value": n_total * CLIP_SECONDS * args.steps / e2e16_s, "unit": UNIT, "h2d_bytes_per_step": ho16["h2d_bytes"] * world,
                          "d2h_bytes_per_step": ho16["d2h_bytes"] * world, "ms_per_step": 1e3 * e2e16_s / args.steps,
                          "note": "same C-ABI path fed int16 PCM host clips (gat_transcribe_clips_host_pcm16); extra, not the headline"},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "kernels": rows,
            "cpu_baseline": cpu_baseline,
        }))
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
