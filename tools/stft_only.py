"""Runs only the STFT -> mel kernels on the bench's 4096 x 1 s clips (for ncu captures; needs a B200)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from guitar_audio_transcriber_ai_b200.engine import Engine
from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
from guitar_audio_transcriber_ai_b200 import synth
ck = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/ckpt"
eng = Engine(22050, device="cuda:0")
eng.load_cnn(load_checkpoint(ck / "cnn_synth_sr22050.ckpt")["model"]); eng.load_mlp(load_checkpoint(ck / "mlp_synth_sr22050.ckpt")["model"])
base, _ = synth.clip_batch(256, 1.0, 22050, 0)
dev = torch.from_numpy(base).cuda().repeat(16, 1).contiguous()
for _ in range(3):
    mel = eng.melspec_db(dev)
    feats, hz = eng.mfcc_features(dev, add_pitch=False)
    out = eng.transcribe_clips(dev, yin_on_normalized=True) if "full" in sys.argv else None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); mel = eng.melspec_db(dev); e1.record(); torch.cuda.synchronize()
print("melspec_db (clip_scale + stft image) ms:", e0.elapsed_time(e1))
