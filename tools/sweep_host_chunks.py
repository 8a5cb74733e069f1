"""Sweeps the chunk schedule of gat_transcribe_clips_host / _pcm16 on the bench's 4096 x 1 s clips (needs a B200)."""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from guitar_audio_transcriber_ai_b200.engine import Engine
from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
from guitar_audio_transcriber_ai_b200 import synth
ck = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/ckpt"
eng = Engine(22050, device="cuda:0")
eng.load_cnn(load_checkpoint(ck / "cnn_synth_sr22050.ckpt")["model"]); eng.load_mlp(load_checkpoint(ck / "mlp_synth_sr22050.ckpt")["model"])
clips, _ = synth.clip_batch(256, 1.0, 22050, 0)
f32 = torch.from_numpy(np.tile(clips, (16, 1))).pin_memory()
i16 = torch.clamp(torch.round(f32 * 32767.0), -32768, 32767).to(torch.int16).pin_memory()
for sched in ((4, 4, 1), (1, 8, 0), (1, 8, 1), (1, 8, 2), (2, 8, 0), (1, 16, 0), (1, 4, 0), (2, 16, 0)):
    eng.lib.check(eng.lib.gat_set_host_chunks(eng._ctx, *sched))
    row = []
    for buf in (f32, i16):
        for _ in range(3): eng.transcribe_clips_host(buf, skip_mlp=True, want_probs=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): eng.transcribe_clips_host(buf, skip_mlp=True, want_probs=False)
        torch.cuda.synchronize(); row.append(round((time.perf_counter() - t0) / 10 * 1e3, 3))
    print(sched, "f32 ms", row[0], "pcm16 ms", row[1], flush=True)
for sched in ((1, 8, 0),):
    eng.lib.check(eng.lib.gat_set_host_chunks(eng._ctx, *sched))
    for buf, name in ((f32, "f32"), (i16, "pcm16")):
        eng.transcribe_clips_host(buf, skip_mlp=True, want_probs=False)
        eng.profile_begin(); eng.transcribe_clips_host(buf, skip_mlp=True, want_probs=False); pr = eng.profile_end()
        print(sched, name, "sum kernels", round(sum(v[1] for k, v in pr.items() if k != "h2d_copy"), 3), {k: (v[0], round(v[1], 3)) for k, v in pr.items()}, flush=True)
