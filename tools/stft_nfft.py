"""Times the image chain (clip_scale + STFT -> mel -> dB) at a given n_fft on 4096 x 1 s clips (needs a B200).
GAT_STFT_CHUNKED=1 selects the round-1 chunked kernel for n_fft 512 / 1024 (A/B)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from guitar_audio_transcriber_ai_b200.engine import Engine
from guitar_audio_transcriber_ai_b200 import synth
n_fft = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = Engine(22050, {"N_MELS": 64, "N_FFT": n_fft, "HOP_LENGTH": 256}, device="cuda:0")
base, _ = synth.clip_batch(256, 1.0, 22050, 0)
dev = torch.from_numpy(base).cuda().repeat(16, 1).contiguous()
for _ in range(3):
    mel = eng.melspec_db(dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    mel = eng.melspec_db(dev)
e1.record(); torch.cuda.synchronize()
print("n_fft", n_fft, "melspec_db (clip_scale + stft image) ms:", e0.elapsed_time(e1) / 5)
