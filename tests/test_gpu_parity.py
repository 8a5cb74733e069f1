"""CUDA path vs the CPU oracle and the committed golden vectors (run with -m gpu on a B200).

Everything goes through the C ABI (libgat.so) via the package's ctypes binding; the oracle (oracle/port.py)
is only the checker.
"""
import numpy as np
import pytest
import torch

from conftest import CKPT, golden_audio
from tolerances import ENV_ABS, PROB_ABS, YIN_CENTS, cents, mel_ok, mel_ok_degenerate, mfcc_ok

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tr22():
    from guitar_audio_transcriber_ai_b200 import Transcriber
    return Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", CKPT, CKPT, device="cuda:0")


@pytest.fixture(scope="module")
def tr11():
    from guitar_audio_transcriber_ai_b200 import Transcriber
    return Transcriber("mlp_v1.0.0.ckpt", "cnn_synth_sr11025.ckpt", CKPT, CKPT, device="cuda:0")


def _by_duration(g):
    groups = {}
    for k, d in enumerate(g["durations"]):
        groups.setdefault(float(d), []).append(k)
    return groups


def test_extension_is_the_cuda_build(tr22):
    from guitar_audio_transcriber_ai_b200 import _lib
    assert _lib.load().path.name == "libgat.so"
    assert not getattr(_lib.load(), "_host_emulation", False)
    assert tr22.engine.device.type == "cuda"


@pytest.mark.parametrize("which", ["22050", "11025"])
def test_features_match_golden(which, tr22, tr11, golden_clips_22050, golden_clips_11025):
    tr, g = (tr22, golden_clips_22050) if which == "22050" else (tr11, golden_clips_11025)
    sr = int(g["sr"])
    for dur, ks in _by_duration(g).items():
        clips = np.stack([golden_audio(g, k) for k in ks])
        mel = tr.engine.melspec_db(clips).cpu().numpy()
        feats, hz = tr.engine.mfcc_features(clips, yin_on_normalized=True)
        feats, hz = feats.cpu().numpy(), hz.cpu().numpy()
        for i, k in enumerate(ks):
            assert mel[i].shape == g[f"mel_{k}"][0].shape
            assert mel_ok(mel[i], g[f"mel_{k}"][0]), (which, k, np.abs(mel[i] - g[f"mel_{k}"][0]).max())
            ref = g[f"mfcc_{k}"][0]
            assert mfcc_ok(feats[i, :64], ref[:64]), (which, k, np.abs(feats[i, :64] - ref[:64]).max())
            assert abs(feats[i, 64] - ref[64]) <= 2e-6, (which, k)
            assert cents(hz[i], g[f"yin_hz_{k}"]) <= YIN_CENTS, (which, k, hz[i], g[f"yin_hz_{k}"])


def test_yin_frames_and_notes(tr22, golden_clips_22050):
    from guitar_audio_transcriber_ai_b200 import YinDsp
    g = golden_clips_22050
    yin = YinDsp(device="cuda:0")
    for dur, ks in _by_duration(g).items():
        clips = np.stack([golden_audio(g, k) for k in ks])
        hz, f0 = tr22.engine.yin(clips, normalize=False)
        f0 = f0.cpu().numpy()
        for i, k in enumerate(ks):
            ref = g[f"yin_f0_{k}"]
            assert f0[i].shape == ref.shape
            assert np.median(cents(f0[i], ref)) <= 0.01
            assert np.max(cents(f0[i], ref)) <= 0.5, (k, np.max(cents(f0[i], ref)))
            p, info = yin.estimate_pitch(clips[i], 22050)
            assert info["midi"] == int(g[f"yin_midi_{k}"]) and info["note_name"] == str(g[f"yin_note_{k}"])


@pytest.mark.parametrize("which", ["22050", "11025"])
def test_predict_from_golden_features(which, tr22, tr11, golden_clips_22050, golden_clips_11025):
    """NotePredictor.predict on the reference's own features: isolates the CNN/MLP/ensemble kernels."""
    tr, g = (tr22, golden_clips_22050) if which == "22050" else (tr11, golden_clips_11025)
    for dur, ks in _by_duration(g).items():
        mf = np.concatenate([g[f"mfcc_{k}"] for k in ks])
        ms = np.concatenate([g[f"mel_{k}"] for k in ks])
        res = tr.predictor.predict(mf, ms)
        for i, k in enumerate(ks):
            assert int(res["indices"][i]) == int(g[f"index_{k}"][0])
            assert str(res["labels"][i]) == str(g[f"label_{k}"][0])
            assert np.abs(res["probs"][i] - g[f"probs_{k}"][0]).max() <= PROB_ABS
            assert np.abs(res["per_model_probs"]["mlp"][i] - g[f"mlp_probs_{k}"][0]).max() <= PROB_ABS
            assert np.abs(res["per_model_probs"]["cnn"][i] - g[f"cnn_probs_{k}"][0]).max() <= PROB_ABS
            assert abs(float(res["confidences"][i]) - float(g[f"conf_{k}"][0])) <= PROB_ABS


@pytest.mark.parametrize("which", ["22050", "11025"])
def test_transcribe_note_labels_exact(which, tr22, tr11, golden_clips_22050, golden_clips_11025):
    tr, g = (tr22, golden_clips_22050) if which == "22050" else (tr11, golden_clips_11025)
    sr = int(g["sr"])
    for k in range(len(g["seeds"])):
        res = tr.transcribe_note(golden_audio(g, k), clip_duration=float(g["durations"][k]), sr_in=sr)
        assert set(res) == {"indices", "labels", "confidences", "probs", "per_model_probs"}
        assert str(res["labels"][0]) == str(g[f"label_{k}"][0]), (which, k)
        assert int(res["indices"][0]) == int(g[f"index_{k}"][0])
        assert np.abs(res["probs"] - g[f"probs_{k}"]).max() <= 5e-5, (which, k, np.abs(res["probs"] - g[f"probs_{k}"]).max())
        assert res["indices"].dtype == np.int64 and res["probs"].dtype == np.float32


def test_segmentation_exact(tr22, golden_phrases):
    from guitar_audio_transcriber_ai_b200 import synth
    g = golden_phrases
    for k, seed in enumerate(g["seeds"]):
        y, _, _ = synth.phrase(int(seed), sr=22050)
        r = tr22.engine.segment(y, 0.5, diagnostics=True)
        assert r["onsets"].cpu().numpy().tolist() == g[f"onsets_{k}"].tolist()
        assert r["frames"].cpu().numpy().tolist() == g[f"frames_bt_{k}"].tolist()
        assert np.array_equal(r["table"].cpu().numpy(), g[f"table_{k}"])
        env = g[f"onset_env_{k}"]
        en = env - env.min()
        en = en / (en.max() + np.finfo(np.float64).tiny)
        assert np.abs(r["env"].cpu().numpy() - en).max() <= ENV_ABS
        assert np.abs(r["rms_db"].cpu().numpy() - g[f"rms_db_{k}"]).max() <= 5e-5


def test_long_audio_segmentation_exact(tr22):
    """Two minutes of audio (24 phrases): global dB max / percentile / min-max dependencies at a non-trivial
    scale, onsets and slice table still identical to the CPU oracle."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    y, _, _ = synth.long_audio(24, 22050, seed0=300)
    want_onsets, want_clips, want_table = port.slice_in_memory(y, 22050, 0.5)
    r = tr22.engine.segment(y, 0.5)
    assert r["onsets"].cpu().numpy().tolist() == want_onsets
    assert np.array_equal(r["table"].cpu().numpy(), want_table)
    assert np.array_equal(r["clips"].cpu().numpy(), want_clips)          # gathered samples are copies: bit-exact


def test_full_hour_segmentation_properties(tr22):
    """BASELINE config 4 at full size (720 phrases = 1 h): invariants of slicing.py:106-161 that need no oracle."""
    from guitar_audio_transcriber_ai_b200 import synth
    y, _, _ = synth.long_audio(720, 22050, seed0=0)
    yd = torch.from_numpy(y).cuda()
    r1 = tr22.engine.segment(yd, 0.5)
    r2 = tr22.engine.segment(yd, 0.5)
    on = r1["onsets"].cpu().numpy()
    tab = r1["table"].cpu().numpy()
    assert np.array_equal(on, r2["onsets"].cpu().numpy()) and torch.equal(r1["clips"], r2["clips"])     # deterministic
    assert len(on) > 5000 and np.all(on % 512 == 0)                      # frames_to_samples: multiples of the hop
    assert np.all(np.diff(on) >= int(0.3 * 22050))                       # greedy minimum separation
    assert len(tab) <= len(on) - 1                                       # the last onset never yields a clip
    skip, length = int(0.1 * 22050), int(0.5 * 22050)
    assert np.array_equal(tab[:, 1], on[tab[:, 0]] + skip)               # start = onset + attack skip
    nxt = on[np.minimum(tab[:, 0] + 1, len(on) - 1)]
    assert np.array_equal(tab[:, 2], np.minimum(tab[:, 1] + length, nxt))
    clips = r1["clips"].cpu().numpy()
    k = 1234
    assert np.array_equal(clips[k, : tab[k, 2] - tab[k, 1]], y[tab[k, 1]: tab[k, 2]]) and not clips[k, tab[k, 2] - tab[k, 1]:].any()
    rms_db = 20 * np.log10(np.sqrt((clips.astype(np.float64) ** 2).mean(1)) + 1e-10)
    assert np.all(rms_db > -37.0 - 1e-3)                                 # every kept slice passed the loudness test


def test_segmentation_at_sr_11025(tr11):
    """The shipped MLP checkpoint's rate: librosa's onset_detect defaults become pre_max 0 / wait 0 there."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    for seed in (5, 6):
        y, _, _ = synth.phrase(seed, sr=11025)
        want_onsets, want_clips, want_table = port.slice_in_memory(y, 11025, 0.5)
        r = tr11.engine.segment(y, 0.5)
        assert r["onsets"].cpu().numpy().tolist() == want_onsets
        assert np.array_equal(r["table"].cpu().numpy(), want_table)
        assert np.array_equal(r["clips"].cpu().numpy(), want_clips)


def test_c_abi_argument_errors(tr22):
    """Every entry point reports bad arguments through its return code + gat_last_error, never by crashing."""
    import ctypes as C
    lib = tr22.engine.lib
    ctx = tr22.engine._ctx
    assert lib.gat_melspec_db(ctx, None, 1, 22050, 1, None, None) != 0 and b"null" in lib.gat_last_error()
    assert lib.gat_yin(None, None, 1, 22050, 0, None, None, None) != 0
    assert lib.gat_infer(ctx, None, 65, None, 1, 44, None, None, None, None, None, None, None, None) != 0
    assert lib.gat_segment(ctx, None, 1000, None, 4, None, None, None, None, None, None, None, None, None, None) != 0
    assert lib.gat_set_conv_pass(ctx, 0) != 0 and lib.gat_set_conv_pass(ctx, 16) == 0
    bad = np.zeros(3, np.int32)
    assert lib.gat_load_mlp(ctx, bad.ctypes.data_as(C.c_void_p), 9, bad.ctypes.data_as(C.c_void_p), 3) != 0
    a = torch.zeros(1, 600, device="cuda")
    out = torch.zeros(1, 64, 3, device="cuda")
    assert lib.gat_melspec_db(ctx, C.c_void_p(a.data_ptr()), 1, 600, 1, C.c_void_p(out.data_ptr()), None) != 0
    assert b"too short" in lib.gat_last_error()
    # file front end / onset entry points
    x = torch.zeros(2, 1000, device="cuda")
    y = torch.zeros(2, 500, device="cuda")
    taps = torch.zeros(33, dtype=torch.float64, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    assert lib.gat_resample(ctx, P(x), 2, 1000, 1, 2, P(taps), 16, P(y), 499, None) != 0 and b"n_out" in lib.gat_last_error()
    assert lib.gat_resample(ctx, P(x), 2, 1000, 0, 2, P(taps), 16, P(y), 500, None) != 0
    assert lib.gat_resample(ctx, P(x), 2, 1000, 1, 2, P(taps), 16, P(y), 500, None) == 0
    assert lib.gat_decode_mono(ctx, P(x), 7, 1000, 2, P(y), None) != 0 and b"sample_format" in lib.gat_last_error()
    assert lib.gat_decode_mono(ctx, None, 0, 1000, 2, P(y), None) != 0
    assert lib.gat_pcm16_roundtrip(ctx, None, 10, None) != 0 and lib.gat_pcm16_roundtrip(ctx, None, 0, None) == 0
    sp = tr22.engine.slicer_params(22050, 0.5)
    on = torch.zeros(64, dtype=torch.int64, device="cuda")
    n = torch.zeros(1, dtype=torch.int32, device="cuda")
    sp.onset_hop = 511
    assert lib.gat_detect_onsets(ctx, P(x), 2000, C.byref(sp), 64, P(on), P(n), None) != 0 and b"hop" in lib.gat_last_error()
    sp.onset_hop = 512
    assert lib.gat_detect_onsets(ctx, P(x), 2000, C.byref(sp), 0, P(on), P(n), None) != 0
    assert lib.gat_transcribe_clips_host_pcm16(ctx, None, 4, 11025, 0, None, None, None) != 0
    torch.cuda.synchronize()


def test_transcribe_audio_matches_reference_pipeline(tr22, golden_phrases):
    from guitar_audio_transcriber_ai_b200 import synth
    g = golden_phrases
    for k, seed in enumerate(g["seeds"]):
        y, _, _ = synth.phrase(int(seed), sr=22050)
        res = tr22.transcribe_audio(y, 22050, 0.5)
        assert [str(s) for s in res["labels"]] == [str(s) for s in g[f"labels_{k}"]]
        assert res["indices"].tolist() == g[f"indices_{k}"].tolist()
        assert res["onsets"] == g[f"onsets_{k}"].tolist()
        assert np.array_equal(res["slice_table"], g[f"table_{k}"])
        assert np.abs(res["probs"] - g[f"probs_{k}"]).max() <= 5e-5
        hz = np.array([d[0] for d in res["dsp_info"]])
        assert np.max(cents(hz, g[f"yin_hz_{k}"])) <= YIN_CENTS


def test_live_oracle_on_fresh_inputs(tr22):
    """Not only the committed vectors: new seeds, CUDA vs oracle/port.py run right here on the host."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    mlp_ck, cnn_ck = load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt"), load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    clips, _ = synth.clip_batch(24, 1.0, 22050, seed0=5000)
    got = tr22.transcribe_notes(clips, 1.0, 22050)
    for i in range(len(clips)):
        want = port.transcribe_note(mlp_ck, cnn_ck, clips[i], 1.0, 22050)
        assert str(got["labels"][i]) == str(want["labels"][0])
        assert np.abs(got["probs"][i] - want["probs"][0]).max() <= 5e-5


@pytest.mark.parametrize("dur", [2.0, 4.0])
def test_long_clips_against_oracle(dur, tr22):
    """BASELINE config 5 durations: 2 s and 4 s clips (mel images 64x173 / 64x345) exercise the column-blocked conv
    tiling, multi-chunk STFT staging and longer YIN / MFCC reductions."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    mlp_ck, cnn_ck = load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt"), load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    clips, _ = synth.clip_batch(5, dur, 22050, seed0=int(dur * 1000))
    got = tr22.transcribe_notes(clips, dur, 22050)
    out = tr22.engine.transcribe_clips(clips, yin_on_normalized=True, return_features=True)
    for i in range(len(clips)):
        want = port.transcribe_note(mlp_ck, cnn_ck, clips[i], dur, 22050)
        mf, ms = port.extract_inference_features_from_audio(clips[i], 22050)
        assert mel_ok(out["mel"][i].cpu().numpy(), ms[0]) and mfcc_ok(out["mfcc"][i, :64].cpu().numpy(), mf[0, :64])
        assert abs(float(out["mfcc"][i, 64]) - mf[0, 64]) <= 2e-6
        assert str(got["labels"][i]) == str(want["labels"][0])
        assert np.abs(got["per_model_probs"]["cnn"][i] - want["per_model_probs"]["cnn"][0]).max() <= 5e-5
        assert np.abs(got["probs"][i] - want["probs"][0]).max() <= 5e-5


@pytest.mark.parametrize("n_fft", [512, 1024, 4096])
def test_other_n_fft_against_torchaudio(n_fft):
    """MelSpecConfig.N_FFT is configurable; BASELINE config 5 sweeps 1024 / 2048 / 4096."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.engine import Engine
    eng = Engine(22050, {"N_MELS": 64, "N_FFT": n_fft, "HOP_LENGTH": 256}, device="cuda:0")
    clips, _ = synth.clip_batch(6, 1.0, 22050, seed0=40 + n_fft)
    mel = eng.melspec_db(clips).cpu().numpy()
    for i in range(len(clips)):
        ref = port.melspec_image(clips[i], 22050, 64, n_fft, 256).numpy()
        assert mel[i].shape == ref.shape
        if n_fft >= 1024:
            assert mel_ok(mel[i], ref), (n_fft, i, np.abs(mel[i] - ref).max())
        else:   # 64 mel bands over 257 bins: the lowest filters weigh a fraction of ONE weak bin, where the two float32
                # FFTs' rounding noise is a larger share of the value; not a BASELINE configuration, checked loosely
            assert np.abs(mel[i] - ref).max() <= 2e-2, (n_fft, i, np.abs(mel[i] - ref).max())
    eng.close()


def test_edge_cases_against_oracle(tr22):
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    eng = tr22.engine
    sr = 22050
    silent = np.zeros(11025, np.float32)
    dc = np.full(11025, 0.25, np.float32)
    click = np.zeros(11025, np.float32); click[5000] = 1.0
    short = synth.note(330.0, 0.2, sr, 9)        # transcribe_note zero-pads it to 0.5 s
    long_ = synth.note(196.0, 0.9, sr, 10)       # ... and truncates this one
    for name, a in (("silent", silent), ("dc", dc), ("click", click)):
        mel = eng.melspec_db(a[None]).cpu().numpy()[0]
        ref = port.melspec_image(a, sr).numpy()
        assert np.all(np.isfinite(mel)) and mel_ok_degenerate(mel, ref), name
        hz, _ = eng.yin(a[None])
        ref_hz, _ = port.yin_estimate_pitch(a, sr)
        assert cents(hz.cpu().numpy()[0], ref_hz) <= YIN_CENTS, (name, hz, ref_hz)
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    mlp_ck, cnn_ck = load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt"), load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    for a in (short, long_):
        got = tr22.transcribe_note(a, 0.5, sr)
        want = port.transcribe_note(mlp_ck, cnn_ck, a, 0.5, sr)
        assert str(got["labels"][0]) == str(want["labels"][0])
        assert np.abs(got["probs"] - want["probs"]).max() <= 5e-5


def test_full_size_properties(tr22):
    """BASELINE config 2 at full size (4096 one-second clips): properties that need no oracle."""
    from guitar_audio_transcriber_ai_b200 import synth
    eng = tr22.engine
    base, _ = synth.clip_batch(64, 1.0, 22050, seed0=9000)
    reps = 4096 // 64
    gains = (0.25 + 0.5 * np.arange(reps, dtype=np.float32) / reps)
    big = torch.from_numpy(np.concatenate([base * g for g in gains])).cuda()
    out1 = eng.transcribe_clips(big, skip_mlp=True, return_features=True)
    out2 = eng.transcribe_clips(big, skip_mlp=True, return_features=True)
    torch.cuda.synchronize()
    # idempotent / deterministic
    assert torch.equal(out1["indices"], out2["indices"]) and torch.equal(out1["probs"], out2["probs"])
    assert torch.equal(out1["mel"], out2["mel"])
    # batch independence: a clip's result does not depend on its neighbours or its position
    solo = eng.transcribe_clips(big[1000:1003].contiguous(), skip_mlp=True, return_features=True)
    assert torch.equal(solo["mel"], out1["mel"][1000:1003]) and torch.equal(solo["probs"], out1["probs"][1000:1003])
    # volume normalisation makes the features gain-invariant (up to float32 rounding of y/c): same labels
    idx = out1["indices"].view(reps, 64)
    assert torch.equal(idx, idx[0:1].expand_as(idx))
    mel = out1["mel"].view(reps, 64, -1)
    assert float((mel - mel[0:1]).abs().max()) < 2e-2
    assert torch.all(torch.isfinite(out1["probs"])) and float((out1["probs"].sum(1) - 1).abs().max()) < 1e-5
    # and the first 64 agree with the CPU oracle
    import port
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    cnn_ck = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    with torch.inference_mode():
        imgs = torch.stack([port.melspec_image(base[i] * gains[0], 22050) for i in range(16)])
        probs = torch.softmax(port.cnn_forward(cnn_ck["model"], imgs), -1).numpy()
    assert np.abs(out1["probs"][:16].cpu().numpy() - probs).max() <= 5e-5
    assert out1["indices"][:16].cpu().numpy().tolist() == probs.argmax(1).tolist()


def test_host_buffer_entry_point(tr22):
    from guitar_audio_transcriber_ai_b200 import synth
    clips, _ = synth.clip_batch(700, 0.5, 22050, seed0=7000)   # > one 512-clip chunk: exercises the double buffer
    pinned = torch.from_numpy(clips).pin_memory()
    dev = tr22.engine.transcribe_clips(torch.from_numpy(clips).cuda(), yin_on_normalized=True)
    host = tr22.engine.transcribe_clips_host(pinned, yin_on_normalized=True)
    assert host["indices"].tolist() == dev["indices"].cpu().numpy().tolist()
    assert np.array_equal(host["probs"], dev["probs"].cpu().numpy())


def test_error_behaviour(tr22):
    from guitar_audio_transcriber_ai_b200 import Transcriber
    from guitar_audio_transcriber_ai_b200.engine import Engine
    with pytest.raises(FileNotFoundError):
        Transcriber("nope.ckpt", "cnn_synth_sr22050.ckpt", CKPT, CKPT, device="cuda:0")
    with pytest.raises(ValueError):
        Engine(22050, {"N_MELS": 64, "N_FFT": 1000, "HOP_LENGTH": 256}, device="cuda:0")
    with pytest.raises(ValueError):
        tr22.predictor.predict(None, None)
    with pytest.raises(UnboundLocalError):
        tr22.predictor.predict(np.zeros((1, 65), np.float32), None)
    with pytest.raises(ValueError):
        tr22.engine.melspec_db(np.zeros((1, 512), np.float32))     # too short for reflect padding


# ---------------------------------------------------------------------------------------- file front end (8f-1)
@pytest.mark.gpu
def test_front_end_kernels_against_oracle(tr22):
    """decode + channel mean and the PCM_16 round trip bit-exact; the resampler within one float32 ulp."""
    import librosa_shim
    import port
    import file_cases
    eng = tr22.engine
    rng = np.random.default_rng(3)
    st = rng.integers(-32768, 32767, (100_001, 2)).astype(np.int16)
    assert np.array_equal(eng.decode_mono(st).cpu().numpy(), np.mean((st.astype(np.float32) / np.float32(32768.0)).T, axis=0))
    x = (0.9 * rng.uniform(-1, 1, (5, 11025))).astype(np.float32)
    x[0, :5] = [1.0, -1.0, 0.5 / 32767, 1.5 / 32767, 2.5 / 32767]
    t = torch.from_numpy(x.copy()).cuda()
    eng.pcm16_roundtrip_(t)
    assert np.array_equal(t.cpu().numpy(), np.stack([port.pcm16_roundtrip(r) for r in x]))
    for a, b in ((22050, 11025), (32000, 22050), (11025, 22050), (44100, 22050), (48000, 22050)):
        out = eng.resample(x, a, b).cpu().numpy()
        want = np.stack([librosa_shim.resample(r, orig_sr=a, target_sr=b) for r in x])
        assert out.shape == want.shape and np.abs(out - want).max() <= file_cases.RESAMPLE_ABS


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mono22050", "stereo32000_ckpt11025"])
def test_transcribe_file_matches_reference(name, tmp_path):
    """Transcriber.transcribe(path) against vectors from the reference's own transcribe() (oracle/make_golden.py)."""
    import file_cases
    file_cases.check_transcribe_file("cuda:0", tmp_path, name)


@pytest.mark.gpu
def test_transcribe_notes_resamples_like_the_reference(tr22):
    """transcribe_note with sr_in != target_sr (transcribe.py:172-173): resample, then the usual path."""
    import port
    import ref_env
    from guitar_audio_transcriber_ai_b200 import synth
    mlp_ck, cnn_ck = ref_env.load_ckpt(CKPT / "mlp_synth_sr22050.ckpt"), ref_env.load_ckpt(CKPT / "cnn_synth_sr22050.ckpt")
    # 16 kHz input leaves the mel bins above 8 kHz at the float32 rounding floor of the FFT (about -140 dB re the
    # peak): there the dB image is implementation noise in ANY float32 FFT, and the CNN sees it -> looser bound.
    for seed, sr_in, tol in ((3, 44100, 5e-5), (5, 32000, 5e-5), (4, 16000, 1e-2)):
        a = synth.note(float(synth.midi_to_hz(synth.random_midi(seed))), 0.5, sr_in, seed)
        want = port.transcribe_note(mlp_ck, cnn_ck, a, 0.5, sr_in)
        got = tr22.transcribe_note(a, 0.5, sr_in)
        assert [str(s) for s in got["labels"]] == [str(s) for s in want["labels"]]
        assert np.abs(got["per_model_probs"]["mlp"] - want["per_model_probs"]["mlp"]).max() <= 5e-5
        assert np.abs(got["probs"] - want["probs"]).max() <= tol


@pytest.mark.gpu
def test_training_feature_builders(tr22):
    """features.py:162-435: batched dataset features == the oracle's per-clip loop; loaders/splits."""
    import train_cases
    train_cases.check_training_builders("cuda:0", n_classes=6, per_class=4)


@pytest.mark.gpu
def test_detect_onsets_ungated_any_hop(tr22):
    """AudioSlicer.detect_onsets(y, sr, hop_len, min_sep) on raw signals (slicing.py:106-122), hops 256 / 512 / 1024."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    for seed in range(4):
        y, _, _ = synth.phrase(seed, sr=22050)
        for hop, min_sep in ((512, 0.25), (1024, 0.3), (256, 0.3)):
            assert tr22.slicer.detect_onsets(y, 22050, hop, min_sep) == port.detect_onsets(y.astype(np.float64), 22050, hop, min_sep)
    yl, _, _ = synth.long_audio(6, 22050, 40)
    assert tr22.slicer.detect_onsets(yl, 22050, 1024, 0.3) == port.detect_onsets(yl.astype(np.float64), 22050, 1024, 0.3)


@pytest.mark.gpu
def test_live_transcriber():
    """Streaming path on the GPU: same notes, labels and probabilities as the prototype's loop on the oracle."""
    import live_cases
    live_cases.check_live("cuda:0")


@pytest.mark.gpu
def test_host_entry_point_pcm16(tr22):
    """gat_transcribe_clips_host_pcm16 == the float32 entry fed x/32768 (libsndfile's read scaling), bit for bit."""
    from guitar_audio_transcriber_ai_b200 import synth
    clips, _ = synth.clip_batch(700, 0.5, 22050, 900)           # more than one 592-clip chunk
    q = np.clip(np.rint(clips * 32767.0), -32768, 32767).astype(np.int16)
    a = tr22.engine.transcribe_clips_host(torch.from_numpy(q).pin_memory())
    ia, pa = a["indices"].copy(), a["probs"].copy()
    b = tr22.engine.transcribe_clips_host(torch.from_numpy(q.astype(np.float32) / np.float32(32768.0)).pin_memory())
    assert np.array_equal(ia, b["indices"]) and np.array_equal(pa, b["probs"])
    assert a["h2d_bytes"] * 2 == b["h2d_bytes"]


def test_context_lifecycle_releases_memory():
    """gat_ctx_create / gat_ctx_destroy: every workspace a context grew is returned to the device."""
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    from guitar_audio_transcriber_ai_b200.engine import Engine
    cnn = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")["model"]
    mlp = load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt")["model"]
    clips = torch.from_numpy(synth.clip_batch(64, 0.5, 22050, 0)[0]).cuda()
    y = torch.from_numpy(synth.phrase(0)[0]).cuda()

    def cycle():
        eng = Engine(22050, device="cuda:0")
        eng.load_cnn(cnn); eng.load_mlp(mlp)
        eng.transcribe_clips(clips)
        eng.segment(y, 0.5)
        eng.resample(clips[:4], 22050, 11025)
        torch.cuda.synchronize()
        eng.close()

    cycle()
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(10):
        cycle()
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 8 << 20, f"{(free0 - free1) / 2**20:.1f} MiB not returned after 10 create/destroy cycles"
