"""Small end-to-end run of every kernel (for compute-sanitizer; needs a B200)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from guitar_audio_transcriber_ai_b200 import Transcriber, synth
from guitar_audio_transcriber_ai_b200.engine import Engine
ck = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/ckpt"
tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", ck, ck, device="cuda:0")
clips, _ = synth.clip_batch(5, 0.5, 22050, 0)
print(tr.transcribe_notes(clips, 0.5, 22050)["labels"])
clips, _ = synth.clip_batch(3, 2.0, 22050, 10)
print(tr.transcribe_notes(clips, 2.0, 22050)["labels"])
y, _, _ = synth.phrase(1)
print(tr.transcribe_audio(y, 22050, 0.5)["labels"])
print(tr.engine.transcribe_clips_host(torch.from_numpy(synth.clip_batch(7, 0.5, 22050, 20)[0]).pin_memory())["indices"])
eng = Engine(22050, {"N_MELS": 64, "N_FFT": 1024, "HOP_LENGTH": 256}, device="cuda:0")
print(eng.melspec_db(clips).shape)
torch.cuda.synchronize(); print("mini_run ok")
