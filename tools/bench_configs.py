"""Times the BASELINE.json configurations that are not the driver's bench line (needs a B200).

cfg1: one 5 s phrase, full pipeline (segment -> features -> ensemble -> YIN) through Transcriber.transcribe_audio
cfg3: 4096 x 1 s clips, MFCC + YIN + MLP + CNN ensemble (transcribe_clips without skip_mlp)
cfg4: N phrases concatenated (default 720 = 1 hour): whole-file segmentation + ensemble on the sliced clips
Prints one JSON object with per-kernel times from the CUDA-event profiler.
"""
import argparse, json, pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from guitar_audio_transcriber_ai_b200 import Transcriber, synth

ap = argparse.ArgumentParser(); ap.add_argument("--phrases", type=int, default=720); a = ap.parse_args()
ck = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/ckpt"
tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", ck, ck, device="cuda:0")
eng = tr.engine
out = {}

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

def profile(fn):
    eng.profile_begin(); fn(); return {k: round(v[1], 4) for k, v in sorted(eng.profile_end().items(), key=lambda kv: -kv[1][1])}

# cfg1
y, _, _ = synth.phrase(0)
t0 = time.perf_counter(); res = tr.transcribe_audio(y, 22050, 0.5); torch.cuda.synchronize(); first = time.perf_counter() - t0
t0 = time.perf_counter()
for _ in range(20): res = tr.transcribe_audio(y, 22050, 0.5)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
out["cfg1_5s_phrase"] = {"ms_per_call_host_to_host": 1e3 * dt, "audio_s_per_s": 5.0 / dt, "labels": [str(s) for s in res["labels"]]}

# cfg3
clips, midi = synth.clip_batch(4096, 1.0, 22050, 0)
dev = torch.from_numpy(clips).cuda()
ms, r = timed(lambda: eng.transcribe_clips(dev, yin_on_normalized=True))
out["cfg3_4096x1s_mfcc_yin_mlp_cnn"] = {"ms": ms, "audio_s_per_s": 4096 / (ms * 1e-3), "kernels_ms": profile(lambda: eng.transcribe_clips(dev, yin_on_normalized=True))}
hz = eng.yin(dev)[0].cpu().numpy()
yin_midi = np.round(12 * np.log2(hz / 440.0) + 69).astype(int)
names = synth.class_names(); lab = np.array([names.index(synth.midi_to_label(m)) for m in midi])
idx = r["indices"].cpu().numpy()
out["cfg3_4096x1s_mfcc_yin_mlp_cnn"].update(ensemble_vs_truth=float((idx == lab).mean()), yin_vs_truth=float((yin_midi == midi).mean()))

# cfg4
yl, ml, sl = synth.long_audio(a.phrases, 22050, 0)
ydev = torch.from_numpy(yl).cuda()
def full():
    seg = eng.segment(ydev, 0.5)
    return seg, eng.transcribe_clips(seg["clips"], yin_on_normalized=False, apply_scaler=True)
ms, (seg, r) = timed(full, reps=3, warm=1)
dur = len(yl) / 22050
out["cfg4_long_audio"] = {"audio_seconds": dur, "ms": ms, "audio_s_per_s": dur / (ms * 1e-3), "onsets": int(seg["onsets"].shape[0]),
                          "clips": int(seg["clips"].shape[0]), "kernels_ms": profile(full)}
print(json.dumps(out, indent=1))
