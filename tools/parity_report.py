"""Error DISTRIBUTION of every stage of the CUDA path against the CPU oracle over the bench's clips (needs a B200).

    python tools/parity_report.py [--clips 4096] [--librosa-clips 1024] > profiles/r02_parity.json

VERDICT r1 #7 / weak #9: the parity tests state tolerances; this records how far inside them the path actually sits.
Per stage: max / p99 / p50 of the error over ALL clips of BASELINE configs[1] (4096 x 1 s, seed = clip index):
  mel image   dB error, raw and relative to max(|ref|, 20 dB) (the tests' floor), vs genuine torchaudio
  CNN probs   |d| vs genuine torch on the oracle's own mel images; label mismatches
  MFCC, YIN   vs the librosa restatement (slower: `--librosa-clips` of them, spread over the host cores)
  ensemble    labels vs oracle; MLP / YIN / truth agreement rates (cfg 3's check)
and the two documented tolerance exceptions as distributions: n_fft 512 (2e-2 dB) and 16 kHz input (1e-2 on probs).
The oracle is only the checker here; nothing in this file is on the product path.
"""
import argparse, json, os, pathlib, sys
import numpy as np, torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle")]
import port  # noqa: E402
from guitar_audio_transcriber_ai_b200 import Transcriber, synth  # noqa: E402
from guitar_audio_transcriber_ai_b200.engine import Engine  # noqa: E402

CK = ROOT / "tests" / "golden" / "ckpt"
SR = 22050


def dist(x):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    return {"max": float(x.max()), "p99": float(np.percentile(x, 99)), "p50": float(np.percentile(x, 50)), "n": int(x.size)}


def _librosa_row(args):
    clip, = args
    vec = port.mfcc_vector(clip, SR, 64, True, True, yin_on_normalized=True)
    hz, _ = port.yin_estimate_pitch(port.normalize_audio_volume(clip), SR)
    return vec, hz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--librosa-clips", type=int, default=1024)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    clips, midi = synth.clip_batch(a.clips, 1.0, SR, 0)
    # the librosa-restated rows first, in forked workers, BEFORE this process touches CUDA
    m = min(a.librosa_clips, a.clips)
    sel = np.linspace(0, a.clips - 1, m).astype(int)
    import multiprocessing as mp
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        rows = pool.map(_librosa_row, [(clips[i],) for i in sel], chunksize=8)
    tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", CK, CK, device="cuda:0")
    mlp_ck, cnn_ck = tr.model_ckpts["mlp"], tr.model_ckpts["cnn"]
    out = tr.engine.transcribe_clips(torch.from_numpy(clips).cuda(), yin_on_normalized=True, return_features=True)
    got = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    rep = {"clips": a.clips, "workload": "BASELINE configs[1]/[2]: 4096 x 1 s clips, sr 22050, seed = clip index"}

    # ---- mel image + CNN vs genuine torchaudio / torch
    with torch.inference_mode():
        X = port.extract_melspec_features(list(clips), SR, 64, 2048, 256, normalize=True)
        cnn_probs = torch.softmax(port.cnn_forward(cnn_ck["model"], X), -1).numpy()
        cnn_probs_on_ours = torch.softmax(port.cnn_forward(cnn_ck["model"], torch.from_numpy(got["mel"])), -1).numpy()
    ref = X.numpy()
    d = np.abs(got["mel"] - ref)
    rep["mel_db"] = {"abs_dB": dist(d), "rel_to_max_ref_20dB": dist(d / np.maximum(np.abs(ref), 20.0)),
                     "rel_no_floor": dist(d / np.maximum(np.abs(ref), 1e-6)), "tolerance": "1e-4 * max(|ref|, 20 dB)",
                     "per_clip_max_rel": dist((d / np.maximum(np.abs(ref), 20.0)).reshape(a.clips, -1).max(1)),
                     "bins_within_70dB_of_clip_peak_abs_dB": dist(d[ref >= ref.reshape(a.clips, -1).max(1)[:, None, None, None] - 70.0])}
    rep["cnn_probs"] = {"abs": dist(np.abs(got["cnn_probs"] - cnn_probs)), "tolerance": 2e-5,
                        "abs_given_identical_input": dist(np.abs(got["cnn_probs"] - cnn_probs_on_ours)),
                        "label_mismatches": int((got["cnn_probs"].argmax(1) != cnn_probs.argmax(1)).sum())}

    # ---- MFCC + YIN vs the librosa restatement
    vec = np.stack([r[0] for r in rows]); hz = np.array([r[1] for r in rows], dtype=np.float64)
    dm = np.abs(got["mfcc"][sel, :64] - vec[:, :64])
    rep["mfcc"] = {"abs": dist(dm), "rel_to_max_ref_1": dist(dm / np.maximum(np.abs(vec[:, :64]), 1.0)), "tolerance": "1e-4 * max(|ref|, 1)",
                   "pitch_feature_abs": dist(np.abs(got["mfcc"][sel, 64] - vec[:, 64])), "pitch_tolerance": 2e-6, "clips": int(m)}
    rep["yin_median_cents"] = {**dist(1200.0 * np.abs(np.log2(got["yin_hz"][sel] / hz))), "tolerance": 0.05}
    X65 = vec.astype(np.float32)
    want = port.predict(mlp_ck, cnn_ck, X65, ref[sel])
    rep["ensemble"] = {"probs_abs": dist(np.abs(got["probs"][sel] - want["probs"])), "mlp_probs_abs": dist(np.abs(got["mlp_probs"][sel] - want["per_model_probs"]["mlp"])),
                       "label_mismatches": int((got["indices"][sel] != want["indices"]).sum()), "clips": int(m), "tolerance": 2e-5}

    # ---- cfg 3 agreement rates
    names = [str(mlp_ck["reverse_map"][i]) for i in range(len(mlp_ck["reverse_map"]))]
    lm = np.array([next(k for k in range(synth.MIDI_LO, synth.MIDI_HI + 1) if synth.midi_to_label(k) == nm) for nm in names])
    yin = np.round(12 * np.log2(got["yin_hz"] / 440.0) + 69).astype(np.int64)
    ens, mlp, cnn = lm[got["indices"]], lm[got["mlp_probs"].argmax(1)], lm[got["cnn_probs"].argmax(1)]
    rep["agreement"] = {"ensemble_vs_truth": float((ens == midi).mean()), "mlp_vs_truth": float((mlp == midi).mean()), "cnn_vs_truth": float((cnn == midi).mean()),
                        "yin_vs_truth": float((yin == midi).mean()), "mlp_vs_yin": float((mlp == yin).mean()), "ensemble_vs_yin": float((ens == yin).mean()),
                        "yin_misses_that_are_octaves": float(((yin - midi) % 12 == 0)[yin != midi].mean()) if (yin != midi).any() else None}

    # ---- the two documented exceptions, as distributions
    exc = {}
    for n_fft in (512, 1024, 4096):
        e = Engine(SR, {"N_MELS": 64, "N_FFT": n_fft, "HOP_LENGTH": 256}, device="cuda:0")
        sub = clips[:256]
        mel = e.melspec_db(sub).cpu().numpy()[:, 0]
        with torch.inference_mode():
            r = port.extract_melspec_features(list(sub), SR, 64, n_fft, 256, normalize=True).numpy()[:, 0]
        dd = np.abs(mel - r)
        exc[f"n_fft_{n_fft}"] = {"abs_dB": dist(dd), "rel_to_max_ref_20dB": dist(dd / np.maximum(np.abs(r), 20.0)),
                                 "abs_dB_by_mel_band_max": [float(v) for v in dd.max(axis=(0, 2))[:8]], "clips": 256}
        e.close()
    probs16, labels16 = [], 0
    for seed in range(32):
        x = synth.note(float(synth.midi_to_hz(synth.random_midi(seed))), 0.5, 16000, seed)
        w = port.transcribe_note(mlp_ck, cnn_ck, x, 0.5, 16000)
        g = tr.transcribe_note(x, 0.5, 16000)
        probs16.append(np.abs(g["probs"] - w["probs"]).max()); labels16 += int(str(g["labels"][0]) != str(w["labels"][0]))
    exc["input_16kHz_resampled"] = {"probs_abs_per_clip_max": dist(probs16), "label_mismatches": labels16, "clips": 32,
                                    "note": "includes the resampler, which restates soxr_hq's specification (unpinned)"}
    rep["exceptions"] = exc
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
