"""Prints the conv_tc pipeline cycle counters (diagnostic; needs a B200)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import ctypes as C
import numpy as np, torch
from guitar_audio_transcriber_ai_b200.engine import Engine
from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
from guitar_audio_transcriber_ai_b200 import synth, _lib
ck = pathlib.Path(__file__).resolve().parent.parent / "tests/golden/ckpt"
eng = Engine(22050, device="cuda:0")
eng.load_cnn(load_checkpoint(ck / "cnn_synth_sr22050.ckpt")["model"]); eng.load_mlp(load_checkpoint(ck / "mlp_synth_sr22050.ckpt")["model"])
dur = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
clips, _ = synth.clip_batch(148, dur, 22050, 0)
a = torch.from_numpy(np.tile(clips, (4, 1))).cuda()
for _ in range(3): eng.transcribe_clips(a, skip_mlp=True)
eng.lib.check(eng.lib.gat_debug_tc_counters(eng._ctx, None, 0))
eng.transcribe_clips(a, skip_mlp=True)
buf = np.zeros(2 * 148 * 8, np.int64)
eng.lib.check(eng.lib.gat_debug_tc_counters(eng._ctx, _lib.ptr(buf), buf.size))
d = buf.reshape(2, 148, 8)
for name, layer in (("conv2", d[0]), ("conv3", d[1])):
    m = layer.mean(0)
    print(f"dur {dur} {name}: mma total {m[0]:.0f} clk | wait acc_empty {m[1]:.0f} | wait a_full {m[2]:.0f} | wait w_full {m[3]:.0f} | issue+rest {m[0]-m[1]-m[2]-m[3]:.0f} || epilogue total {m[4]:.0f} | wait acc_full {m[5]:.0f} | busy {m[4]-m[5]:.0f}")
