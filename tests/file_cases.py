"""Shared checks for the file pipeline (Transcriber.transcribe): used by the emulation test (CPU) and the GPU test."""
import pathlib

import numpy as np
import scipy.io.wavfile

from conftest import CKPT, GOLD
from tolerances import PROB_ABS, YIN_CENTS, cents

CASES = {
    # golden name: (wav case, mlp ckpt, cnn ckpt, slicing target_sr, clips bit-exact?)
    "mono22050": ("mono22050", "mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", 22050, True),
    "stereo32000_ckpt11025": ("stereo32000", "mlp_v1.0.0.ckpt", "cnn_synth_sr11025.ckpt", 22050, False),
}
# resampled audio: float64 accumulation on both sides, summation order differs -> one float32 ulp of |x| <= 1
RESAMPLE_ABS = 2e-7


def write_case(tmp_path, wav_case):
    from guitar_audio_transcriber_ai_b200 import synth
    frames, sr = synth.wav_case(wav_case)
    path = pathlib.Path(tmp_path) / f"{wav_case}.wav"
    scipy.io.wavfile.write(str(path), sr, frames)
    return path


def check_transcribe_file(device, tmp_path, name, prob_tol=PROB_ABS):
    from guitar_audio_transcriber_ai_b200 import Transcriber
    wav_case, mlp, cnn, target_sr, exact = CASES[name]
    g = np.load(GOLD / "files.npz")
    path = write_case(tmp_path, wav_case)
    tr = Transcriber(mlp, cnn, CKPT, CKPT, device=device)
    out_root = pathlib.Path(tmp_path) / "out"
    res = tr.transcribe(path, out_root=out_root, audio_name="t", target_sr=target_sr, clip_duration=0.5)
    # resampling the input moves samples by ~1e-7, which may flip nothing here: onsets are checked exactly
    assert res["onsets"] == g[f"{name}_onsets"].tolist()
    assert np.array_equal(res["slice_table"], g[f"{name}_table"])
    assert [str(s) for s in res["labels"]] == [str(s) for s in g[f"{name}_labels"]]
    assert np.array_equal(res["indices"], g[f"{name}_indices"])
    assert np.abs(res["probs"] - g[f"{name}_probs"]).max() <= prob_tol
    hz = np.array([d[0] for d in res["dsp_info"]])
    assert np.all(cents(hz, g[f"{name}_yin_hz"]) <= YIN_CENTS)
    assert set(res) >= {"indices", "labels", "confidences", "probs", "per_model_probs", "dsp_info"}
    # the clip files the reference writes (slicing.py:139-144): same names, PCM_16, at the slicing rate
    files = sorted((out_root).glob("t_*/t/*.wav"))
    assert len(files) == len(res["labels"])
    onsets = g[f"{name}_onsets"]
    for f, row in zip(files, g[f"{name}_table"]):
        assert f.name == f"{int(row[0]):04d}_clip__{onsets[int(row[0])] / target_sr:.3f}s.wav"
        sr_f, data = scipy.io.wavfile.read(str(f))
        assert sr_f == target_sr and data.dtype == np.int16 and data.shape == (int(target_sr * 0.5),)
    if exact:       # no resampling: the saved clips are exactly the golden (already quantised) clips
        for f, clip in zip(files, g[f"{name}_clips"]):
            _, data = scipy.io.wavfile.read(str(f))
            assert np.array_equal(data.astype(np.float32) / np.float32(32768.0), clip)
    tr.engine.close()
    return res
