"""Kernel LOGIC checked in this GPU-less container: csrc/*.cu(h) compiled for the host against tests/emu/cpu_emu.h
(one std::thread per CUDA thread) and driven through the same C ABI and Python engine as the CUDA build.
This validates indexing / FFT decomposition / scans before GPU time is spent; the real parity gate is
tests/test_gpu_parity.py on the B200."""
import numpy as np
import pytest
import torch

from conftest import CKPT, golden_audio
from tolerances import ENV_ABS, PROB_ABS, YIN_CENTS, cents, mel_ok, mfcc_ok


@pytest.fixture(scope="module")
def emu_tr():
    if torch.cuda.is_available():
        pytest.skip("a real GPU is present: the CUDA build is tested instead")
    import emu_loader
    from guitar_audio_transcriber_ai_b200 import _lib
    saved = (_lib._LIB, _lib.load)
    emu_loader.install()
    from guitar_audio_transcriber_ai_b200 import Transcriber
    tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", CKPT, CKPT, device="cpu")
    yield tr
    tr.engine.close()
    _lib._LIB, _lib.load = saved
    from guitar_audio_transcriber_ai_b200.dsp import yin
    yin._ENGINES.clear()


def test_emu_features_and_labels(emu_tr, golden_clips_22050):
    g = golden_clips_22050
    ks = [0, 5, 13]                                                        # two 0.5 s clips and a 1 s clip
    for k in ks:
        a = golden_audio(g, k)
        mel = emu_tr.engine.melspec_db(a[None]).numpy()[0]
        feats, hz = emu_tr.engine.mfcc_features(a[None], yin_on_normalized=True)
        assert mel_ok(mel, g[f"mel_{k}"][0]) and mfcc_ok(feats.numpy()[0, :64], g[f"mfcc_{k}"][0, :64])
        assert cents(hz.numpy()[0], g[f"yin_hz_{k}"]) <= YIN_CENTS
        res = emu_tr.transcribe_note(a, float(g["durations"][k]), 22050)
        assert str(res["labels"][0]) == str(g[f"label_{k}"][0])
        assert np.abs(res["probs"] - g[f"probs_{k}"]).max() <= 5e-5


def test_emu_segmentation(emu_tr, golden_phrases):
    from guitar_audio_transcriber_ai_b200 import synth
    g = golden_phrases
    y, _, _ = synth.phrase(int(g["seeds"][1]), sr=22050)
    r = emu_tr.engine.segment(y, 0.5, diagnostics=True)
    assert r["onsets"].numpy().tolist() == g["onsets_1"].tolist()
    assert r["frames"].numpy().tolist() == g["frames_bt_1"].tolist()
    assert np.array_equal(r["table"].numpy(), g["table_1"])
    env = g["onset_env_1"]
    en = (env - env.min()) / ((env - env.min()).max() + np.finfo(np.float64).tiny)
    assert np.abs(r["env"].numpy() - en).max() <= ENV_ABS


def test_emu_segment_batch_equals_per_signal(emu_tr, golden_phrases):
    """gat_segment_batch over P phrases == gat_segment on each phrase alone == the reference's onsets / tables."""
    from guitar_audio_transcriber_ai_b200 import synth
    g = golden_phrases
    seeds = [int(s) for s in g["seeds"]]
    Y = np.stack([synth.phrase(s, sr=22050)[0] for s in seeds])
    b = emu_tr.engine.segment_batch(Y, 0.5)
    table = b["table"].numpy()
    clips = b["clips"].numpy()
    row = 0
    for p, _ in enumerate(seeds):
        k = int(b["n_onsets"][p])
        assert b["onsets"][p, :k].numpy().tolist() == g[f"onsets_{p}"].tolist()
        one = emu_tr.engine.segment(Y[p], 0.5)
        m = one["clips"].shape[0]
        assert int(b["n_clips"][p]) == m == g[f"table_{p}"].shape[0]
        assert np.all(table[row:row + m, 0] == p)
        assert np.array_equal(table[row:row + m, 1:], g[f"table_{p}"])
        assert np.array_equal(clips[row:row + m], one["clips"].numpy())
        row += m
    assert row == table.shape[0]


@pytest.mark.parametrize("seg", ["2", "6", "44"])
def test_emu_yin_block_fft_segmentations(emu_tr, seg, monkeypatch):
    """yin_fft_kernel: every position of a block in its frame pair (first / odd / even / odd-and-last), at forced segment
    lengths, frame by frame against the restated librosa.yin - and identical frames whatever the segment length."""
    import librosa_shim as L
    from guitar_audio_transcriber_ai_b200 import synth
    clips, _ = synth.clip_batch(2, 1.0, 22050, seed0=77)
    monkeypatch.setenv("GAT_YIN_SEG", seg)
    _, f0 = emu_tr.engine.yin(clips)
    monkeypatch.setenv("GAT_YIN_SEG", "12")
    _, f12 = emu_tr.engine.yin(clips)
    assert torch.equal(f0, f12)
    for i in range(2):
        ref = L.yin(clips[i], fmin=50, fmax=1000, sr=22050)
        assert f0[i].shape == ref.shape and np.max(cents(f0[i].numpy(), ref)) <= 0.5


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (512, 100)])
def test_emu_small_n_fft_frame_kernel(emu_tr, n_fft, hop):
    """stft_frames_small_kernel (two / four frames per warp) against genuine torchaudio: interior groups, groups that touch
    the reflect padding, a ragged last group, clips with an odd sample count (32-bit loads), hops that are not 256."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.engine import Engine
    eng = Engine(22050, {"N_MELS": 64, "N_FFT": n_fft, "HOP_LENGTH": hop}, device="cpu")
    try:
        clips, _ = synth.clip_batch(2, 0.3337, 22050, seed0=40 + n_fft)
        mel = eng.melspec_db(clips).numpy()
        for i in range(2):
            ref = port.melspec_image(clips[i], 22050, 64, n_fft, hop).numpy()
            assert mel[i].shape == ref.shape
            bound = 1e-4 if n_fft >= 1024 else 5e-4          # n_fft 512: see test_other_n_fft_against_torchaudio
            assert np.all(np.abs(mel[i] - ref) <= bound * np.maximum(np.abs(ref), 20.0))
    finally:
        eng.close()
