"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz by running the REFERENCE'S OWN FILES.

Runs /root/reference/version_1/source/*.py unmodified (oracle/ref_env.py: librosa -> oracle/librosa_shim,
genuine torchaudio/torch/scipy/sklearn) on seeded synthetic inputs, runs oracle/port.py on the same
inputs, REQUIRES the two to agree exactly, and stores the reference outputs.  Inputs are not stored:
they are regenerated from guitar_audio_transcriber_ai_b200.synth seeds.

    python oracle/make_golden.py            # needs /root/reference; not runnable on the GPU box

What the vectors pin: the port's fidelity to the reference's control flow, dtype promotions and quirks.
What they cannot pin: librosa's own arithmetic (restated in librosa_shim; parity unpinned).
"""
from __future__ import annotations

import contextlib
import pathlib
import sys
import warnings

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import port  # noqa: E402
import ref_env  # noqa: E402
from guitar_audio_transcriber_ai_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
CKPT = GOLD / "ckpt"

CLIP_CASES = {
    # name: (sr, mlp ckpt, cnn ckpt, [(seed, duration_s)])
    "clips_sr22050": (22050, "mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt",
                      [(s, 0.5) for s in range(12)] + [(100 + s, 1.0) for s in range(6)]),
    "clips_sr11025": (11025, "mlp_v1.0.0.ckpt", "cnn_synth_sr11025.ckpt", [(200 + s, 0.5) for s in range(12)]),
}
PHRASE_SEEDS = [0, 1, 2, 3]


@contextlib.contextmanager
def posix_safe_torch_load():
    """transcribe.py:58 calls torch.load(weights_only=False), which cannot build the WindowsPath pickled in
    the shipped MLP checkpoint on Linux.  Swap in the same call with a path-tolerant unpickler."""
    orig = torch.load
    torch.load = lambda path, **kw: orig(path, map_location="cpu", weights_only=False,
                                         pickle_module=ref_env._PickleModule())
    try:
        yield
    finally:
        torch.load = orig


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype or not np.array_equal(a, b, equal_nan=True):
        raise SystemExit(f"[make_golden] port != reference for {what}: shapes {a.shape}/{b.shape} "
                         f"dtypes {a.dtype}/{b.dtype} maxdiff "
                         f"{np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) if a.shape == b.shape else 'n/a'}")


def clip_cases(ns):
    for name, (sr, mlp_name, cnn_name, items) in CLIP_CASES.items():
        with posix_safe_torch_load(), contextlib.redirect_stdout(None):
            tr = ns.transcribe.Transcriber(mlp_ckpt=mlp_name, cnn_ckpt=cnn_name, mlp_root=CKPT, cnn_root=CKPT)
        mlp_ck = ref_env.load_ckpt(CKPT / mlp_name)
        cnn_ck = ref_env.load_ckpt(CKPT / cnn_name)
        out = {"sr": sr, "seeds": np.array([s for s, _ in items]), "durations": np.array([d for _, d in items])}
        for k, (seed, dur) in enumerate(items):
            midi = synth.random_midi(seed)
            audio = synth.note(float(synth.midi_to_hz(midi)), dur, sr, seed)
            with contextlib.redirect_stdout(None):
                res = tr.transcribe_note(audio, clip_duration=dur, sr_in=sr)
                mf, ms = tr.feature_builder.extract_inference_features_from_audio(
                    audio, sr, tr.model_configs["mlp"]["features"]["params"],
                    tr.model_configs["cnn"]["features"]["params"], tr.model_ckpts["mlp"].get("scaler"))
                hz, info = ns.yin.YinDsp().estimate_pitch(audio, sr)
            f0 = sys.modules["librosa"].yin(audio, fmin=50.0, fmax=1000.0, sr=sr)
            pres = port.transcribe_note(mlp_ck, cnn_ck, audio, clip_duration=dur, sr_in=sr)
            pmf, pms = port.extract_inference_features_from_audio(
                audio, sr, mlp_ck["config"]["features"]["params"], cnn_ck["config"]["features"]["params"])
            phz, pinfo = port.yin_estimate_pitch(audio, sr)
            same(mf, pmf, f"{name}[{k}] mfcc"); same(ms, pms, f"{name}[{k}] mel")
            same(res["probs"], pres["probs"], f"{name}[{k}] probs"); same(res["indices"], pres["indices"], "idx")
            assert res["labels"] == pres["labels"] and hz == phz and info == pinfo
            out[f"mfcc_{k}"] = mf
            out[f"mel_{k}"] = ms
            out[f"probs_{k}"] = res["probs"]
            out[f"mlp_probs_{k}"] = res["per_model_probs"]["mlp"]
            out[f"cnn_probs_{k}"] = res["per_model_probs"]["cnn"]
            out[f"index_{k}"] = res["indices"]
            out[f"label_{k}"] = np.array(res["labels"], dtype=str)
            out[f"conf_{k}"] = res["confidences"]
            out[f"yin_hz_{k}"] = np.float64(hz)
            out[f"yin_f0_{k}"] = f0
            out[f"yin_midi_{k}"] = np.int64(info["midi"])
            out[f"yin_note_{k}"] = np.array(info["note_name"])
            out[f"true_midi_{k}"] = np.int64(midi)
        np.savez_compressed(GOLD / f"{name}.npz", **out)
        print(f"[make_golden] wrote {name}.npz ({len(items)} clips)")


def phrase_cases(ns):
    sr = 22050
    mlp_name, cnn_name = "mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt"
    mlp_ck = ref_env.load_ckpt(CKPT / mlp_name)
    cnn_ck = ref_env.load_ckpt(CKPT / cnn_name)
    with posix_safe_torch_load(), contextlib.redirect_stdout(None):
        tr = ns.transcribe.Transcriber(mlp_ckpt=mlp_name, cnn_ckpt=cnn_name, mlp_root=CKPT, cnn_root=CKPT)
    sl = ns.slicing.AudioSlicer()
    lib = sys.modules["librosa"]
    cfg = ns.config.SLICER_CONFIG
    out = {"sr": sr, "seeds": np.array(PHRASE_SEEDS)}
    for k, seed in enumerate(PHRASE_SEEDS):
        y, midis, starts = synth.phrase(seed, sr=sr)
        # --- the reference's sliceNsave body (slicing.py:147-165) minus load_wav / save_clip
        g1 = sl.apply_db_threshold(y=y, min_db=cfg.MIN_IN_DB_THRESHOLD)
        rms_db = sl.compute_rms_db(y=g1, hop_len=cfg.HOP_LEN)
        gate_db, _, _ = sl.compute_dynamic_thresholds(rms_db)
        g2 = sl.apply_rms_threshold(g1, hop_len=cfg.HOP_LEN)
        env = lib.onset.onset_strength(y=g2, sr=sr, hop_length=512)
        frames_bt = lib.onset.onset_detect(onset_envelope=env, sr=sr, hop_length=512, backtrack=True)
        frames_raw = lib.onset.onset_detect(onset_envelope=env, sr=sr, hop_length=512, backtrack=False)
        onsets = sl.detect_onsets(y=g2, sr=sr, min_sep=cfg.MIN_SEP)
        clips, table = [], []
        for i, onset in enumerate(onsets):
            nxt = onsets[i + 1] if i + 1 < len(onsets) else onsets[-1]
            clip, times = sl.slice_audio(y=y, onset=onset, next_onset=nxt, sr=sr, length_sec=0.5,
                                         attack_skip_sec=cfg.ATTACK_SKIP_SEC)
            if not sl.is_slice_loud_enough(clip, cfg.MIN_SLICE_RMS_DB):
                continue
            clips.append(clip.astype(np.float32))
            table.append((i, int(round(times[0] * sr)), int(round(times[1] * sr))))
        # --- the reference's transcribe() steps 3-6 (transcribe.py:124-142) on those clips, from memory
        loader = ref_env.MemoryLoader(clips, sr)
        with contextlib.redirect_stdout(None):
            mf, ms = tr.feature_builder.extract_inference_features(
                loader, tr.model_configs["mlp"]["features"]["params"],
                tr.model_configs["cnn"]["features"]["params"], tr.model_ckpts["mlp"].get("scaler"))
            res = tr.predictor.predict(mf, ms)
        dsp = [ns.yin.YinDsp().estimate_pitch(c, sr) for c in clips]
        # --- port
        p_on, p_clips, p_table = port.slice_in_memory(y, sr, 0.5)
        pres = port.transcribe_audio(mlp_ck, cnn_ck, y, sr, 0.5)
        same(np.array(onsets), np.array(p_on), f"phrase[{k}] onsets")
        same(np.stack(clips), p_clips, f"phrase[{k}] clips")
        same(np.array(table, dtype=np.int64).reshape(-1, 3), p_table, f"phrase[{k}] table")
        same(res["probs"], pres["probs"], f"phrase[{k}] probs")
        assert res["labels"] == pres["labels"]
        assert [d[0] for d in dsp] == [d[0] for d in pres["dsp_info"]]
        same(port.compute_rms_db(port.apply_db_threshold(y, -32.5)), rms_db, "rms_db")
        out[f"rms_db_{k}"] = rms_db
        out[f"gate_db_{k}"] = np.asarray(gate_db)
        out[f"gate1_kept_{k}"] = np.int64(np.count_nonzero(g1))
        out[f"gate2_kept_{k}"] = np.int64(np.count_nonzero(g2))
        out[f"onset_env_{k}"] = env
        out[f"frames_raw_{k}"] = np.asarray(frames_raw, dtype=np.int64)
        out[f"frames_bt_{k}"] = np.asarray(frames_bt, dtype=np.int64)
        out[f"onsets_{k}"] = np.asarray(onsets, dtype=np.int64)
        out[f"table_{k}"] = np.asarray(table, dtype=np.int64).reshape(-1, 3)
        out[f"mfcc_{k}"] = np.asarray(mf)
        out[f"mel_{k}"] = ms.numpy()
        out[f"probs_{k}"] = res["probs"]
        out[f"indices_{k}"] = res["indices"]
        out[f"labels_{k}"] = np.array(res["labels"], dtype=str)
        out[f"yin_hz_{k}"] = np.array([d[0] for d in dsp], dtype=np.float64)
        out[f"true_midi_{k}"] = midis
        out[f"true_starts_{k}"] = starts
        print(f"[make_golden] phrase seed {seed}: {len(onsets)} onsets -> {len(clips)} clips; labels {res['labels']}")
    np.savez_compressed(GOLD / "phrases_sr22050.npz", **out)
    print("[make_golden] wrote phrases_sr22050.npz")


FILE_CASES = {
    # name: (wav case, mlp ckpt, cnn ckpt, slicing target_sr)
    "mono22050": ("mono22050", "mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", 22050),
    "stereo32000_ckpt11025": ("stereo32000", "mlp_v1.0.0.ckpt", "cnn_synth_sr11025.ckpt", 22050),
}


def file_cases(ns):
    """Transcriber.transcribe(audio_path) (transcribe.py:77-144) run VERBATIM on WAV files: librosa.load ->
    shim (decode + restated resampler), sf.write -> oracle/soundfile_standin (libsndfile PCM_16 restated).
    os.listdir is sorted for the run (the reference iterates clips in directory order, which is arbitrary)."""
    import os
    import tempfile
    import scipy.io.wavfile
    out = {}
    for name, (wav, mlp_name, cnn_name, target_sr) in FILE_CASES.items():
        frames, sr_file = synth.wav_case(wav)
        mlp_ck = ref_env.load_ckpt(CKPT / mlp_name)
        cnn_ck = ref_env.load_ckpt(CKPT / cnn_name)
        with tempfile.TemporaryDirectory() as tmp:
            path = pathlib.Path(tmp) / "in.wav"
            scipy.io.wavfile.write(str(path), sr_file, frames)
            orig_listdir = os.listdir
            os.listdir = lambda p=".": sorted(orig_listdir(p))
            try:
                with posix_safe_torch_load(), contextlib.redirect_stdout(None):
                    tr = ns.transcribe.Transcriber(mlp_ckpt=mlp_name, cnn_ckpt=cnn_name, mlp_root=CKPT, cnn_root=CKPT)
                    res = tr.transcribe(path, out_root=pathlib.Path(tmp) / "out", audio_name="t", target_sr=target_sr,
                                        clip_duration=0.5)
            finally:
                os.listdir = orig_listdir
            pres = port.transcribe_file(mlp_ck, cnn_ck, path, target_sr, 0.5)
        same(res["probs"], pres["probs"], f"file[{name}] probs")
        same(res["indices"], pres["indices"], f"file[{name}] indices")
        assert list(res["labels"]) == list(pres["labels"])
        ref_hz = np.array([np.nan if d[0] is None else d[0] for d in res["dsp_info"]], dtype=np.float64)
        port_hz = np.array([np.nan if d[0] is None else d[0] for d in pres["dsp_info"]], dtype=np.float64)
        same(ref_hz, port_hz, f"file[{name}] yin")
        out[f"{name}_probs"] = res["probs"]
        out[f"{name}_indices"] = res["indices"]
        out[f"{name}_labels"] = np.array(res["labels"], dtype=str)
        out[f"{name}_yin_hz"] = ref_hz
        out[f"{name}_onsets"] = np.asarray(pres["onsets"], dtype=np.int64)
        out[f"{name}_table"] = pres["slice_table"]
        out[f"{name}_clips"] = pres["clips"]
        print(f"[make_golden] file case {name}: {len(pres['onsets'])} onsets -> {len(res['labels'])} clips; labels {list(res['labels'])}")
    np.savez_compressed(GOLD / "files.npz", **out)
    print("[make_golden] wrote files.npz")


def main():
    warnings.simplefilter("ignore")
    torch.set_num_threads(1)  # single-threaded reductions: the vectors do not depend on the core count
    ns = ref_env.install()
    if "--files-only" not in sys.argv:
        clip_cases(ns)
        phrase_cases(ns)
    file_cases(ns)


if __name__ == "__main__":
    main()
