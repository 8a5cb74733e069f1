"""Runs the YIN kernel alone on 4096 x 1 s clips (for ncu captures; needs a B200)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from guitar_audio_transcriber_ai_b200 import synth
from guitar_audio_transcriber_ai_b200.engine import Engine
sr = int(sys.argv[1]) if len(sys.argv) > 1 else 22050
eng = Engine(sr, device="cuda:0")
clips, _ = synth.clip_batch(256, 1.0, sr, 0)
dev = torch.from_numpy(clips).cuda().repeat(16, 1)
for _ in range(2):
    hz, f0 = eng.yin(dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); eng.yin(dev); e1.record(); torch.cuda.synchronize()
print("yin ms", e0.elapsed_time(e1), "clips", dev.shape[0], "sr", sr)
