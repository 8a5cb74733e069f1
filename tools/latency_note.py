"""Single-note latency of Transcriber.transcribe_note (host array in, result dict out); needs a B200."""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from guitar_audio_transcriber_ai_b200 import Transcriber, synth
ck = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/ckpt"
tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", ck, ck, device="cuda:0")
a = synth.note(220.0, 0.5, 22050, 1)
for _ in range(20): tr.transcribe_note(a, 0.5, 22050)
torch.cuda.synchronize()
ts = []
for _ in range(200):
    t0 = time.perf_counter(); r = tr.transcribe_note(a, 0.5, 22050); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"transcribe_note: median {np.median(ts):.3f} ms, p10 {np.percentile(ts,10):.3f}, p90 {np.percentile(ts,90):.3f}; label {r['labels']}")
dev = torch.from_numpy(a[None]).cuda()
eng = tr.engine
for _ in range(20): eng.transcribe_clips(dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(200): eng.transcribe_clips(dev)
e1.record(); torch.cuda.synchronize()
print(f"engine.transcribe_clips (resident, N=1): {e0.elapsed_time(e1)/200:.3f} ms per call (GPU timeline)")
eng.profile_begin(); eng.transcribe_clips(dev); prof = eng.profile_end()
print({k: round(v[1], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})
