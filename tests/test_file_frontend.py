"""File front end on CPU: WAV container parsing, the resampling filter design, the oracle's file
pipeline against the golden vectors made from the reference's own ``Transcriber.transcribe``, and the product's
``Transcriber.transcribe(path)`` through the host-emulated kernels."""
import struct

import numpy as np
import pytest
import scipy.io.wavfile
import scipy.signal
import torch

from conftest import CKPT, GOLD
import file_cases


def test_wav_reader_matches_scipy(tmp_path):
    from guitar_audio_transcriber_ai_b200.audio import wavio
    rng = np.random.default_rng(0)
    cases = {
        "i16_mono": rng.integers(-32768, 32767, 1000).astype(np.int16),
        "i16_stereo": rng.integers(-32768, 32767, (777, 2)).astype(np.int16),
        "f32_stereo": rng.standard_normal((500, 2)).astype(np.float32),
        "f64_mono": rng.standard_normal(300),
        "i32_mono": rng.integers(-2**31, 2**31 - 1, 400).astype(np.int32),
        "u8_mono": rng.integers(0, 255, 300).astype(np.uint8),
    }
    for name, data in cases.items():
        p = tmp_path / f"{name}.wav"
        scipy.io.wavfile.write(str(p), 16000, data)
        frames, sr = wavio.read_wav_frames(p)
        assert sr == 16000 and frames.shape == (data.shape[0], 1 if data.ndim == 1 else data.shape[1])
        ref = data.reshape(frames.shape)
        if data.dtype == np.int16:
            assert frames.dtype == np.int16 and np.array_equal(frames, ref)
        elif data.dtype == np.float32:
            assert frames.dtype == np.float32 and np.array_equal(frames, ref)
        elif data.dtype == np.float64:
            assert np.array_equal(frames, ref.astype(np.float32))
        elif data.dtype == np.int32:
            assert np.array_equal(frames, (ref / 2147483648.0).astype(np.float32))
        else:
            assert np.array_equal(frames, (ref.astype(np.float32) - 128.0) / 128.0)


def test_pcm16_quantiser_known_answers():
    """soundfile.write always runs libsndfile's CLIPPING float -> PCM_16 converter (python-soundfile sets
    SFC_SET_CLIPPING on every file): lrintf(x * 2^31) >> 16, saturating.  Known answers worked by hand from
    src/pcm.c f2les_clip_array; rint(x * 32767), the non-clipping routine, gives 29490 for 0.9."""
    import soundfile_standin
    from guitar_audio_transcriber_ai_b200.audio import wavio
    x = np.array([0.9, -0.9, 1.0, -1.0, 1.5, -1.5, 0.0, 0.5 / 32768, 1.0 / 32768, -0.25 / 32768, -1.0 / 32768,
                  0.99998474, 32767.5 / 32768, -32767.5 / 32768], dtype=np.float32)
    want = np.array([29491, -29492, 32767, -32768, 32767, -32768, 0, 0, 1, -1, -1, 32767, 32767, -32768], dtype=np.int16)
    assert np.array_equal(wavio.float_to_pcm16(x), want)
    assert np.array_equal(soundfile_standin.float_to_pcm16(x), want)
    rng = np.random.default_rng(11)
    r = rng.uniform(-1.2, 1.2, 100_000).astype(np.float32)
    ref = np.clip(np.floor(np.rint(r.astype(np.float64) * 2147483648.0) / 65536.0), -32768, 32767).astype(np.int16)
    assert np.array_equal(wavio.float_to_pcm16(r), ref)
    assert np.array_equal(soundfile_standin.float_to_pcm16(r), ref)


def test_wav_reader_24bit_extensible_and_errors(tmp_path):
    from guitar_audio_transcriber_ai_b200.audio import wavio
    vals = np.array([0, 1, -1, 8388607, -8388608, 123456, -654321], dtype=np.int32)
    body = b"".join(int(v & 0xFFFFFF).to_bytes(3, "little") for v in vals)
    fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 8000, 8000 * 3, 3, 24) + struct.pack("<HHI", 22, 24, 4) + struct.pack("<H", 1) + b"\x00" * 14
    blob = b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + 4 + 8 + len(body) + 1) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt \
        + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(body)) + body + b"\x00"
    p = tmp_path / "x24.wav"
    p.write_bytes(blob)
    frames, sr = wavio.read_wav_frames(p)
    assert sr == 8000 and np.array_equal(frames[:, 0], vals.astype(np.float32) / 8388608.0)
    with pytest.raises(FileNotFoundError):
        wavio.read_wav_frames(tmp_path / "missing.wav")
    (tmp_path / "junk.wav").write_bytes(b"not a wav file at all")
    with pytest.raises(ValueError):
        wavio.read_wav_frames(tmp_path / "junk.wav")


def test_wav_writer_is_readable(tmp_path):
    from guitar_audio_transcriber_ai_b200.audio import wavio
    q = np.random.default_rng(1).integers(-32768, 32767, 5000).astype(np.int16)
    wavio.write_wav_pcm16(tmp_path / "o.wav", q, 22050)
    sr, data = scipy.io.wavfile.read(str(tmp_path / "o.wav"))
    assert sr == 22050 and np.array_equal(data, q)
    assert (tmp_path / "o.wav").stat().st_size == 44 + 2 * q.size


@pytest.mark.parametrize("rates", [(22050, 11025), (32000, 22050), (44100, 22050), (11025, 22050), (48000, 11025)])
def test_resample_filter_design(rates):
    """numpy-only design in tables.py == scipy.signal.firwin with the same Kaiser specification."""
    import math
    from guitar_audio_transcriber_ai_b200 import tables
    up, down, taps, half = tables.resample_filter(*rates)
    g = math.gcd(*rates)
    assert (up, down) == (rates[1] // g, rates[0] // g) and taps.shape == (2 * half + 1,)
    q = max(up, down)
    ref = scipy.signal.firwin(2 * half + 1, (1.0 + 0.913) / (2.0 * q), window=("kaiser", 0.1102 * (120.0 - 8.7)))
    assert np.abs(taps - up * ref).max() <= 1e-12
    w, h = scipy.signal.freqz(taps / up, worN=1 << 15)
    stop = np.abs(h[w >= np.pi / q])
    passband = np.abs(h[w <= 0.913 * np.pi / q])
    assert 20 * np.log10(stop.max()) < -110.0 and np.abs(20 * np.log10(passband)).max() < 1e-3


@pytest.mark.parametrize("name", list(file_cases.CASES))
def test_oracle_file_pipeline_matches_golden(name, tmp_path):
    """oracle/port.py::transcribe_file == the reference's own Transcriber.transcribe (stored by make_golden)."""
    import port
    import ref_env
    wav_case, mlp, cnn, target_sr, _ = file_cases.CASES[name]
    g = np.load(GOLD / "files.npz")
    path = file_cases.write_case(tmp_path, wav_case)
    torch.set_num_threads(1)
    res = port.transcribe_file(ref_env.load_ckpt(CKPT / mlp), ref_env.load_ckpt(CKPT / cnn), path, target_sr, 0.5)
    assert res["onsets"] == g[f"{name}_onsets"].tolist()
    assert np.array_equal(res["clips"], g[f"{name}_clips"])
    assert [str(s) for s in res["labels"]] == [str(s) for s in g[f"{name}_labels"]]
    assert np.abs(res["probs"] - g[f"{name}_probs"]).max() <= 1e-6


@pytest.fixture()
def emu_lib():
    if torch.cuda.is_available():
        pytest.skip("a real GPU is present: the CUDA build is tested instead")
    import emu_loader
    from guitar_audio_transcriber_ai_b200 import _lib
    from guitar_audio_transcriber_ai_b200.dsp import yin
    saved = (_lib._LIB, _lib.load)
    emu_loader.install()
    yield
    for e in yin._ENGINES.values():
        e.close()
    yin._ENGINES.clear()
    _lib._LIB, _lib.load = saved


def test_emu_front_end_kernels(emu_lib):
    """decode + channel mean, PCM_16 round trip and the polyphase resampler against the oracle's restatement."""
    import librosa_shim
    import port
    from guitar_audio_transcriber_ai_b200.engine import Engine
    rng = np.random.default_rng(3)
    eng = Engine(22050, device="cpu")
    st = rng.integers(-32768, 32767, (4001, 2)).astype(np.int16)
    ref = np.mean((st.astype(np.float32) / np.float32(32768.0)).T, axis=0)
    assert np.array_equal(eng.decode_mono(st).numpy(), ref)
    f3 = rng.standard_normal((1000, 3)).astype(np.float32)
    assert np.allclose(eng.decode_mono(f3).numpy(), np.mean(f3.T, axis=0), rtol=0, atol=1e-7)
    x = (0.9 * rng.uniform(-1, 1, (3, 5000))).astype(np.float32)
    x[0, :5] = [1.0, -1.0, 0.5 / 32767, 1.5 / 32767, 2.5 / 32767]
    t = torch.from_numpy(x.copy())
    eng.pcm16_roundtrip_(t)
    assert np.array_equal(t.numpy(), np.stack([port.pcm16_roundtrip(r) for r in x]))
    for a, b in ((22050, 11025), (32000, 22050), (11025, 22050)):
        out = eng.resample(x, a, b).numpy()
        want = np.stack([librosa_shim.resample(r, orig_sr=a, target_sr=b) for r in x])
        assert out.shape == want.shape and np.abs(out - want).max() <= file_cases.RESAMPLE_ABS
    eng.close()


@pytest.mark.parametrize("name", list(file_cases.CASES))
def test_emu_transcribe_file(emu_lib, name, tmp_path):
    file_cases.check_transcribe_file("cpu", tmp_path, name, prob_tol=5e-5)


def test_emu_cli(emu_lib, tmp_path, capsys):
    """transcribe_cli.py:96-114: console table and the txt dump, through the emulated kernels."""
    from guitar_audio_transcriber_ai_b200 import transcribe_cli
    path = file_cases.write_case(tmp_path, "mono22050")
    g = np.load(GOLD / "files.npz")
    res = transcribe_cli.main(["--audio", str(path), "--out", str(tmp_path / "o"), "--save_results", "--device", "cpu",
                               "--mlp_ckpt", "mlp_synth_sr22050.ckpt", "--cnn_ckpt", "cnn_synth_sr22050.ckpt",
                               "--mlp_root", str(CKPT), "--cnn_root", str(CKPT)])
    out = capsys.readouterr().out
    labels = [str(s) for s in g["mono22050_labels"]]
    assert "Idx |  Label |  Confidence | (YIN Note Estimate)" in out
    assert f"000  {labels[0]:>4}  (conf=" in out
    txt = (tmp_path / "o" / "mono22050_transcription.txt").read_text(encoding="utf-8")
    rows = txt.split("\n\nFull result dict:\n")[0].splitlines()
    assert [r.split(",")[1] for r in rows] == labels and rows[0].startswith("0,")
    assert "'dsp_info'" in txt
    assert not list((tmp_path / "o").glob("*/*/*.wav"))       # clips are only kept with --save_clips
    with pytest.raises(ValueError):
        transcribe_cli.main(["--audio", str(tmp_path / "o" / "mono22050_transcription.txt"), "--device", "cpu"])


def test_emu_training_feature_builders(emu_lib):
    """features.py:162-435 through the emulated kernels, against the oracle port per clip."""
    import train_cases
    train_cases.check_training_builders("cpu")


def test_emu_live_transcriber(emu_lib):
    """Streaming path: feed()/step() against the prototype's loop restated on the oracle."""
    import live_cases
    live_cases.check_live("cpu")
