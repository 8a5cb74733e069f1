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


@pytest.mark.parametrize("sr", [11025, 16000, 22050, 25000, 44100])
def test_yin_block_fft_against_oracle_at_other_rates(sr):
    """The block-FFT difference function (csrc/yin.cuh, yin_fft_kernel: 7 / 14 / 16 lags per lane) and, above 25.6 kHz,
    the direct form it falls back to, frame by frame against the restated librosa.yin; odd frame counts, raw and
    volume-normalised input, and a batch big enough that segments are longer than one frame pair."""
    import librosa_shim as L
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.engine import Engine
    eng = Engine(sr, device="cuda:0")
    try:
        for dur, n_clips in ((0.61, 6), (1.0, 700)):
            clips, _ = synth.clip_batch(n_clips, dur, sr, seed0=400 + sr % 97)
            for norm in (False, True):
                hz, f0 = eng.yin(clips, normalize=norm)
                hz, f0 = hz.cpu().numpy(), f0.cpu().numpy()
                for i in range(0, n_clips, max(1, n_clips // 6)):
                    y = clips[i]
                    if norm:
                        y = (y / (np.sqrt(np.mean(y ** 2)) + 1e-9)).astype(np.float32)
                    ref = L.yin(y, fmin=50, fmax=1000, sr=sr)
                    # frames 0 and 1 integrate over the zero padding (and a synthetic note starts at a zero crossing): their
                    # difference function is an energy ramp without troughs, and which lag wins is decided by the float32
                    # rounding noise of whoever computes it (tolerances.py) - they must be finite, the rest must agree
                    assert f0[i].shape == ref.shape and np.all(np.isfinite(f0[i]))
                    assert np.max(cents(f0[i][2:], ref[2:])) <= 0.5, (sr, dur, norm, i, np.max(cents(f0[i][2:], ref[2:])))
                    assert cents(hz[i], np.median(ref)) <= YIN_CENTS
        # a clip's frames do not depend on the batch it is in (segment lengths do): the first 6 of 700 alone
        clips, _ = synth.clip_batch(700, 1.0, sr, seed0=400 + sr % 97)
        _, f_all = eng.yin(clips)
        _, f_few = eng.yin(clips[:6])
        assert torch.equal(f_all[:6], f_few)
    finally:
        eng.close()


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
                # FFTs' rounding noise is a larger share of the value.  profiles/r02_parity.json (256 clips): max 1.5e-2 dB =
                # 2.8e-4 of max(|ref|, 20 dB), all of it in mel band 0, p99 7e-6 -> bound at 5e-4 (was 2e-2 dB absolute)
            assert np.all(np.abs(mel[i] - ref) <= 5e-4 * np.maximum(np.abs(ref), 20.0)), (n_fft, i, np.abs(mel[i] - ref).max())
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
    # and the first 64 agree with the CPU oracle (the bench run compares another 1024 labels; tools/parity_report.py all 4096)
    import port
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    cnn_ck = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    with torch.inference_mode():
        imgs = port.extract_melspec_features([base[i] * gains[0] for i in range(64)], 22050, 64, 2048, 256, normalize=True)
        probs = torch.softmax(port.cnn_forward(cnn_ck["model"], imgs), -1).numpy()
    assert np.abs(out1["probs"][:64].cpu().numpy() - probs).max() <= PROB_ABS
    assert out1["indices"][:64].cpu().numpy().tolist() == probs.argmax(1).tolist()


def test_yin_full_size_properties(tr22):
    """The block-FFT YIN at BASELINE config 3's size (4096 one-second clips): properties that need no oracle."""
    from guitar_audio_transcriber_ai_b200 import synth
    eng = tr22.engine
    base, midi = synth.clip_batch(64, 1.0, 22050, seed0=9100)
    dev = torch.from_numpy(np.concatenate([base] * 64)).cuda()
    hz1, f1 = eng.yin(dev)
    hz2, f2 = eng.yin(dev)
    assert torch.equal(f1, f2) and torch.equal(hz1, hz2)                       # deterministic
    assert torch.equal(f1.view(64, 64, -1), f1[:64].unsqueeze(0).expand(64, -1, -1))   # position in the batch does not matter
    _, solo = eng.yin(dev[1000:1003].contiguous())
    assert torch.equal(solo, f1[1000:1003])                                    # nor does the batch size (segment lengths differ)
    # a power-of-two gain scales every sample, product, butterfly and sum exactly: the same f0 bit for bit, except where an
    # autocorrelation value or an energy crosses librosa's absolute 1e-6 dead zone (frames in the padding, a few lags a million)
    _, f3 = eng.yin(dev * 0.5)
    a, b = f1[:, 2:].cpu().numpy(), f3[:, 2:].cpu().numpy()
    assert np.mean(a == b) > 0.999 and np.max(cents(a, b)) < 1e-2
    # volume normalisation (the in-memory path) makes the pitch gain-invariant up to float32 rounding of y * (1 / c)
    hz_n, _ = eng.yin(dev, normalize=True)
    hz_g, _ = eng.yin(dev * 0.37, normalize=True)
    assert np.max(cents(hz_n.cpu().numpy(), hz_g.cpu().numpy())) < 1e-2
    # and it is a pitch detector: the median lands on the synthesised note for most clips (octave errors are YIN's own)
    got = np.round(12 * np.log2(hz1[:64].cpu().numpy() / 440.0) + 69).astype(int)
    assert np.mean(got == np.asarray(midi)[:64]) > 0.75


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
    # Distribution over 32 clips (profiles/r02_parity.json): per-clip max |d probs| p50 5.7e-4, max 6.6e-3, labels 32 / 32.
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


# ---------------------------------------------------------------------------------------- round 2: sharding, flags, CLI
def test_segment_batch_matches_single_signal_and_golden(tr22, golden_phrases):
    """gat_segment_batch over P phrases == gat_segment per phrase == the reference's onsets / slice tables, bit for bit;
    also at a size with more signals than SMs and with one degenerate (silent) signal in the batch."""
    from guitar_audio_transcriber_ai_b200 import synth
    g = golden_phrases
    seeds = [int(s) for s in g["seeds"]]
    eng = tr22.engine
    Y = np.stack([synth.phrase(s, sr=22050)[0] for s in seeds])
    b = eng.segment_batch(Y, 0.5)
    table, clips = b["table"].cpu().numpy(), b["clips"].cpu().numpy()
    row = 0
    for p in range(len(seeds)):
        k = int(b["n_onsets"][p])
        assert b["onsets"][p, :k].cpu().numpy().tolist() == g[f"onsets_{p}"].tolist()
        m = g[f"table_{p}"].shape[0]
        assert int(b["n_clips"][p]) == m and np.all(table[row:row + m, 0] == p)
        assert np.array_equal(table[row:row + m, 1:], g[f"table_{p}"])
        one = eng.segment(Y[p], 0.5)
        assert np.array_equal(clips[row:row + m], one["clips"].cpu().numpy())
        row += m
    assert row == table.shape[0]
    # 300 signals (> 148 SMs), signal 7 silent: every signal still equals its own single-signal run
    big = np.stack([synth.phrase(100 + s, sr=22050, dur=3.0, n_notes=6)[0] for s in range(12)] * 25)
    big[7] = 0.0
    bb = eng.segment_batch(big, 0.5)
    tb = bb["table"].cpu().numpy()
    assert int(bb["n_clips"][7]) == 0 and int(bb["n_onsets"][7]) == eng.segment(big[7], 0.5)["onsets"].shape[0]
    for p in (0, 5, 11, 12, 150, 299):
        one = eng.segment(big[p], 0.5)
        mine = tb[tb[:, 0] == p]
        assert np.array_equal(mine[:, 1:], one["table"].cpu().numpy())
        k = int(bb["n_onsets"][p])
        assert torch.equal(bb["onsets"][p, :k], one["onsets"])
    assert torch.equal(bb["clips"][: int(bb["n_clips"][0])], eng.segment(big[0], 0.5)["clips"])


def _gpu_rank_worker(rank, world, port_no, q):
    """One rank of a 2-rank run of the sharded entry points on CUDA.  With two GPUs: NCCL, one GPU per rank.  On a
    one-GPU box both ranks share cuda:0 and the records travel over gloo (parallel._all_gather_into stages them)."""
    import os
    import torch.distributed as dist
    from guitar_audio_transcriber_ai_b200 import Transcriber, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    two = torch.cuda.device_count() >= 2
    dev = f"cuda:{rank if two else 0}"
    torch.cuda.set_device(dev)
    if world > 1:
        if two:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", CKPT, CKPT, device=dev)
    Y = np.stack([synth.phrase(s, sr=22050)[0] for s in range(9)])
    a = tr.transcribe_phrases_sharded(Y, 0.5)
    clips, _ = synth.clip_batch(301, 0.5, 22050, 40)
    b = tr.transcribe_notes_sharded(clips, 0.5, 22050)
    c = tr.transcribe_audio_sharded(Y.reshape(-1), 22050, 0.5)
    keep = lambda r: {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in r.items() if k != "local_probs"}
    q.put((rank, world, keep(a), keep(b), keep(c), dist.get_backend() if world > 1 else "none"))
    if world > 1:
        dist.destroy_process_group()


def test_sharded_entry_points_two_ranks_equal_one(tr22, golden_phrases):
    """N1: transcribe_phrases_sharded / transcribe_notes_sharded / transcribe_audio_sharded with two ranks return, on
    both ranks, exactly what the single-GPU call returns (labels, confidences, slice tables, onsets, YIN), and the
    gathered onsets / tables equal the reference's golden vectors."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gpu_rank_worker, args=(r, 2, 29621, q)) for r in range(2)]
    procs.append(ctx.Process(target=_gpu_rank_worker, args=(0, 1, 29623, q)))
    [p.start() for p in procs]
    got = [q.get(timeout=900) for _ in procs]
    [p.join(120) for p in procs]
    strip = lambda r: {k: v for k, v in r.items() if k != "local_range"}
    for part in (2, 3, 4):
        ref = strip(got[0][part])
        assert len(ref["labels"]) > 0
        for g2 in got[1:]:
            assert strip(g2[part]) == ref, part
    a = got[0][2]
    g = golden_phrases
    tab = np.asarray(a["slice_table"])
    for p in range(4):                                   # phrases 0..3 are the golden seeds
        assert a["onsets"][p] == g[f"onsets_{p}"].tolist()
        assert np.array_equal(tab[tab[:, 0] == p][:, 1:], g[f"table_{p}"])
    # the two-rank run really split the work
    two = [x for x in got if x[1] == 2]
    assert sorted(x[2]["local_range"] for x in two) == [(0, 5), (5, 9)]
    assert sorted(x[3]["local_range"] for x in two) == [(0, 151), (151, 301)]


def test_feature_switches_against_oracle(tr22):
    """NORMALIZE_AUDIO_VOLUME / ADD_PITCH_FEATURES / TO_DB off (features.py:184-185,:199-206,:313-316,:458-502)."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.audio.features import MelFeatureBuilder
    eng = tr22.engine
    clips, _ = synth.clip_batch(6, 0.5, 22050, seed0=77)
    clips *= np.linspace(0.2, 1.0, 6, dtype=np.float32)[:, None]
    mel_raw = eng.melspec_db(clips, normalize=False).cpu().numpy()
    mel_pow = eng.melspec_db(clips, normalize=True, to_db=False).cpu().numpy()
    f_raw, _ = eng.mfcc_features(clips, normalize=False, add_pitch=False)
    f_raw = f_raw.cpu().numpy()
    assert f_raw.shape == (6, 64)
    for i in range(6):
        assert mel_ok(mel_raw[i], port.melspec_image(clips[i], 22050, normalize=False).numpy())
        ref_p = port.melspec_image(clips[i], 22050, normalize=True, to_db=False).numpy()
        assert np.all(np.abs(mel_pow[i] - ref_p) <= 1e-4 * np.maximum(np.abs(ref_p), 1e-5 * ref_p.max())), i
        ref_f = port.mfcc_vector(clips[i], 22050, 64, normalize=False, add_pitch=False)
        assert ref_f.shape == (64,) and mfcc_ok(f_raw[i], ref_f)
    # the reference API with the switches in the config dicts
    fb = MelFeatureBuilder(device="cuda:0")
    mf_cfg = dict(N_MFCC=64, NORMALIZE_AUDIO_VOLUME=False, ADD_PITCH_FEATURES=False)
    ms_cfg = dict(N_MELS=64, N_FFT=2048, HOP_LENGTH=256, NORMALIZE_AUDIO_VOLUME=False)
    a, m = fb.extract_inference_features_from_audio(clips[2], 22050, mf_cfg, ms_cfg, None, melspec_to_db=False)
    wa, wm = port.extract_inference_features_from_audio(clips[2], 22050, mf_cfg, ms_cfg, None, melspec_to_db=False)
    assert a.shape == wa.shape == (1, 64) and m.shape == wm.shape
    assert mfcc_ok(a[0], wa[0]) and np.all(np.abs(m - wm) <= 1e-4 * np.maximum(np.abs(wm), 1e-5 * wm.max()))
    # fused path with the switches: an MLP without the pitch column
    from guitar_audio_transcriber_ai_b200.engine import Engine
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    from guitar_audio_transcriber_ai_b200.training.mlp_trainer import MLP
    torch.manual_seed(0)
    mlp = MLP(num_features=64, hidden_dim=128, num_hidden_layers=2, num_classes=47).eval()
    cnn_ck = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    mlp_ck = {"model": mlp.state_dict(), "reverse_map": load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt")["reverse_map"]}
    e2 = Engine(22050, device="cuda:0")
    e2.load_cnn(cnn_ck["model"]); e2.load_mlp(mlp_ck["model"])
    with pytest.raises(ValueError, match="64 inputs"):
        e2.transcribe_clips(clips)                                   # 65 feature columns into a 64-input MLP: refused
    out = e2.transcribe_clips(clips, add_pitch=False, normalize_mfcc=False, normalize_mel=False, return_features=True)
    X = np.vstack([port.mfcc_vector(c, 22050, 64, normalize=False, add_pitch=False) for c in clips])
    M = port.extract_melspec_features(list(clips), 22050, 64, 2048, 256, normalize=False)
    want = port.predict(mlp_ck, cnn_ck, X, M)
    assert out["indices"].cpu().numpy().tolist() == want["indices"].tolist()
    assert np.abs(out["probs"].cpu().numpy() - want["probs"]).max() <= 5e-5
    e2.close()


def test_mlp_width_is_checked_at_the_c_abi(tr22):
    """gat_transcribe_clips must refuse an MLP whose input width differs from the feature row width."""
    import ctypes as C
    from guitar_audio_transcriber_ai_b200 import _lib
    from guitar_audio_transcriber_ai_b200.engine import Engine
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    eng = Engine(22050, mfcc_config={"N_MFCC": 32}, device="cuda:0")          # 33 feature columns
    eng.load_cnn(load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")["model"])
    eng.load_mlp(load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt")["model"])   # 65 inputs
    a = torch.zeros((2, 11025), device="cuda:0")
    K = eng.num_classes
    bufs = [torch.empty((2, K), device="cuda:0") for _ in range(3)]
    idx, conf = torch.empty(2, dtype=torch.int64, device="cuda:0"), torch.empty(2, device="cuda:0")
    rc = eng.lib.gat_transcribe_clips(eng._ctx, _lib.ptr(a), 2, 11025, 0, _lib.ptr(bufs[0]), _lib.ptr(bufs[1]), _lib.ptr(bufs[2]),
                                      _lib.ptr(idx), _lib.ptr(conf), None, None, None, None)
    assert rc != 0 and b"65 inputs" in eng.lib.gat_last_error()
    with pytest.raises(ValueError):
        eng.transcribe_clips(a)
    eng.close()


def test_two_predictors_do_not_share_weights():
    """stand-alone NotePredictors own their contexts; loading a second one leaves the first intact."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.note_predictor import NotePredictor
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    mlp_ck, cnn_ck = load_checkpoint(CKPT / "mlp_synth_sr22050.ckpt"), load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    clips, _ = synth.clip_batch(3, 0.5, 22050, seed0=5)
    feats = [port.extract_inference_features_from_audio(c, 22050) for c in clips]
    X = np.vstack([f[0] for f in feats]); M = np.concatenate([f[1] for f in feats])
    p1 = NotePredictor(device="cuda:0"); p1.load_models(mlp_ck, cnn_ck)
    before = p1.predict(X, M)
    other = dict(cnn_ck); other["model"] = {k: (v * 0.5 if v.dtype.is_floating_point else v) for k, v in cnn_ck["model"].items()}
    p2 = NotePredictor(device="cuda:0"); p2.load_models(mlp_ck, other)
    assert p2.engine is not p1.engine
    after = p1.predict(X, M)
    assert np.array_equal(before["probs"], after["probs"])
    assert not np.array_equal(p2.predict(X, M)["probs"], before["probs"])


def test_cli_on_gpu(tmp_path, capsys):
    """f3: transcribe_cli.main on a WAV file (transcribe_cli.py:16-114): console table, results file, clip files."""
    import file_cases
    from guitar_audio_transcriber_ai_b200 import transcribe_cli
    from guitar_audio_transcriber_ai_b200.audio import wavio
    from conftest import GOLD
    g = np.load(GOLD / "files.npz")
    wav = file_cases.write_case(tmp_path, "mono22050").rename(tmp_path / "take.wav")
    res = transcribe_cli.main(["--audio", str(wav), "--out", str(tmp_path / "out"), "--save_results", "--save_clips",
                               "--mlp_ckpt", "mlp_synth_sr22050.ckpt", "--cnn_ckpt", "cnn_synth_sr22050.ckpt",
                               "--mlp_root", str(CKPT), "--cnn_root", str(CKPT), "--device", "cuda:0"])
    printed = capsys.readouterr().out
    labels = [str(s) for s in g["mono22050_labels"]]
    assert [str(s) for s in res["labels"]] == labels
    lines = [ln for ln in printed.splitlines() if ln[:3].isdigit()]
    assert len(lines) == len(labels)
    for i, (ln, lab) in enumerate(zip(lines, labels)):
        assert ln.startswith(f"{i:03d}  {lab:>4}  (conf={res['confidences'][i]:.2f})  {res['dsp_info'][i][1]['note_name']}")
    txt = (tmp_path / "out" / "take_transcription.txt").read_text(encoding="utf-8").splitlines()
    assert txt[:len(labels)] == [f"{i},{lab},{res['confidences'][i]:.4f}" for i, lab in enumerate(labels)]
    assert "Full result dict:" in txt
    clips = sorted((tmp_path / "out").rglob("*_clip__*.wav")) or sorted((tmp_path / "out").rglob("*.wav"))
    assert len(clips) == len(labels)
    frames, sr = wavio.read_wav_frames(clips[0])
    assert sr == 22050 and frames.dtype == np.int16 and frames.shape[0] == 11025
    with pytest.raises(ValueError):
        transcribe_cli.main(["--audio", str(tmp_path / "out" / "take_transcription.txt")])
    with pytest.raises(FileNotFoundError):
        transcribe_cli.main(["--audio", str(tmp_path / "missing.wav")])


@pytest.mark.parametrize("n_fft", [1024, 4096])
@pytest.mark.parametrize("dur", [2.0, 4.0])
def test_sweep_corners_against_torchaudio(n_fft, dur):
    """BASELINE config 5's corners (n_fft 1024 / 4096 x 2 s / 4 s): mel image vs genuine torchaudio, CNN labels and
    probabilities vs torch."""
    import port
    from guitar_audio_transcriber_ai_b200 import synth
    from guitar_audio_transcriber_ai_b200.engine import Engine
    from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint
    cnn_ck = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    eng = Engine(22050, {"N_MELS": 64, "N_FFT": n_fft, "HOP_LENGTH": 256}, device="cuda:0")
    eng.load_cnn(cnn_ck["model"])
    clips, _ = synth.clip_batch(4, dur, 22050, seed0=int(n_fft + dur))
    out = eng.transcribe_clips(clips, skip_mlp=True, return_features=True)
    with torch.inference_mode():
        X = port.extract_melspec_features(list(clips), 22050, 64, n_fft, 256, normalize=True)
        probs = torch.softmax(port.cnn_forward(cnn_ck["model"], X), -1).numpy()
    mel = out["mel"].cpu().numpy()
    for i in range(len(clips)):
        assert mel_ok(mel[i], X[i].numpy()), (n_fft, dur, i, np.abs(mel[i] - X[i].numpy()).max())
    assert out["indices"].cpu().numpy().tolist() == probs.argmax(1).tolist()
    assert np.abs(out["probs"].cpu().numpy() - probs).max() <= 5e-5
    eng.close()
