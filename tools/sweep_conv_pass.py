"""Sweeps the number of clips per CNN pass (needs a B200)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from guitar_audio_transcriber_ai_b200.engine import Engine
from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
from guitar_audio_transcriber_ai_b200 import synth
ck = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/ckpt"
eng = Engine(22050, device="cuda:0")
eng.load_cnn(load_checkpoint(ck / "cnn_synth_sr22050.ckpt")["model"]); eng.load_mlp(load_checkpoint(ck / "mlp_synth_sr22050.ckpt")["model"])
clips, _ = synth.clip_batch(256, 1.0, 22050, 0)
a = torch.from_numpy(np.tile(clips, (16, 1))).cuda()
for mult in (2, 4, 7, 8, 14, 16, 28):
    eng.lib.check(eng.lib.gat_set_conv_pass(eng._ctx, mult))
    for _ in range(3): eng.transcribe_clips(a, skip_mlp=True)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(5): eng.transcribe_clips(a, skip_mlp=True)
    e1.record(); torch.cuda.synchronize()
    eng.profile_begin(); eng.transcribe_clips(a, skip_mlp=True); pr = eng.profile_end()
    print(mult, round(e0.elapsed_time(e1) / 5, 3), {k: round(v[1], 3) for k, v in pr.items() if "conv" in k})
