"""BASELINE cfg 5: N in {1e4, 1e5, 1e6} clips x duration x n_fft, hop 256, needs a B200.

The clips are generated ON THE DEVICE from torch's counter-based (Philox) generator, batch by batch, because the
large points do not fit anywhere as host data (1e6 x 4 s = 353 GB); only the batch being transcribed is resident.
Each batch runs the bench's hot path (normalise -> STFT -> HTK mel -> dB -> CNN -> labels, gat_transcribe_clips
with skip_mlp); timing = CUDA events around the transcribe calls only (generation excluded), summed over batches.

Per point the tool prints audio-s/s, the per-stage share, the mel chain's algorithmic GB/s vs the HBM peak and its
algorithmic FLOP/s vs the FP32-FMA peak measured on this device by gat_debug_fma_peak.
"""
import argparse, json, math, pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
from guitar_audio_transcriber_ai_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--max-n", type=float, default=1e6)
ap.add_argument("--batch-gb", type=float, default=8.0)
ap.add_argument("--n-fft", type=int, nargs="*", default=[1024, 2048, 4096])
ap.add_argument("--dur", type=float, nargs="*", default=[0.5, 1.0, 2.0, 4.0])
a = ap.parse_args()
root = pathlib.Path(__file__).resolve().parent.parent
ck = load_checkpoint(root / "tests/golden/ckpt/cnn_synth_sr22050.ckpt")
peaks_path = root / "MEASURED_PEAKS.json"
hbm_peak = 6553.0
if peaks_path.is_file():
    pk = json.loads(peaks_path.read_text())
    hbm_peak = float(pk.get("hbm_gbs", hbm_peak)) if isinstance(pk, dict) else hbm_peak
SR = 22050
dev = torch.device("cuda:0")


def make_batch(n_clips: int, n: int, gen: torch.Generator) -> torch.Tensor:
    """Decaying 8-harmonic notes + noise, MIDI 40..86, float32 [n_clips, n] (same recipe as synth.note)."""
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


rows = []
fma_peak = None
for n_fft in a.n_fft:
    eng = Engine(SR, {"N_MELS": 64, "N_FFT": n_fft, "HOP_LENGTH": 256}, {"N_MFCC": 64}, device=dev)
    eng.load_cnn(ck["model"])
    if fma_peak is None:
        fma_peak = eng.fma_peak_tflops()
    for dur in a.dur:
        n = int(SR * dur)
        T = 1 + n // 256
        per_batch = max(148, int(a.batch_gb * 1e9 / (4 * n)) // 148 * 148)
        for N in (10_000, 100_000, 1_000_000):
            if N > a.max_n:
                continue
            gen = torch.Generator(device=dev); gen.manual_seed(1234)
            done, ms, kern = 0, 0.0, {}
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            warm = True
            while done < N:
                b = min(per_batch, N - done)
                clips = make_batch(b, n, gen)
                if warm:      # untimed: sizes the workspaces for this batch shape
                    eng.transcribe_clips(clips, skip_mlp=True); warm = False
                torch.cuda.synchronize()
                e0.record(); r = eng.transcribe_clips(clips, skip_mlp=True); e1.record()
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
                if done == 0:      # per-kernel split of the first batch
                    eng.profile_begin(); eng.transcribe_clips(clips, skip_mlp=True)
                    kern = {k: v[1] for k, v in eng.profile_end().items()}
                    kb = b
                done += b
                del clips, r
            mel_ms = sum(v for k, v in kern.items() if k.startswith("stft_mel")) * (N / kb)
            cnn_ms = sum(v for k, v in kern.items() if k.startswith(("conv", "fc", "avgpool"))) * (N / kb)
            alg_bytes = N * (4 * n + 4 * 64 * T)
            alg_flops = N * T * (2.5 * n_fft * math.log2(n_fft) + 2 * (n_fft // 2 + 1) * 64)
            rows.append({"n_fft": n_fft, "dur_s": dur, "N": N, "ms": round(ms, 3), "audio_s_per_s": round(N * dur / (ms * 1e-3)),
                         "mel_ms": round(mel_ms, 3), "cnn_ms": round(cnn_ms, 3),
                         "kernels_ms_first_batch": {k: round(v, 3) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])}, "first_batch": kb,
                         "mel_GBps": round(alg_bytes / (mel_ms * 1e-3) * 1e-9, 1), "mel_hbm_frac": round(alg_bytes / (mel_ms * 1e-3) * 1e-9 / hbm_peak, 4),
                         "mel_TFLOPs": round(alg_flops / (mel_ms * 1e-3) * 1e-12, 2), "mel_fma_frac": round(alg_flops / (mel_ms * 1e-3) * 1e-12 / fma_peak, 4)})
            print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    eng.close()
print(json.dumps({"fp32_fma_peak_tflops": round(fma_peak, 2), "hbm_peak_GBps": hbm_peak, "rows": rows}, indent=1))
