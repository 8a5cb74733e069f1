"""TEST INFRASTRUCTURE ONLY - synthesise MLP + CNN checkpoints in the reference's schema.

The reference's CNN checkpoint is a stripped large blob (/root/reference/.MISSING_LARGE_BLOBS) and the
shipped MLP checkpoint is for sr = 11025, so the BASELINE configs (sr 22050) have no weights at all.
This script briefly TRAINS both models on synthetic notes (features from the CPU oracle) so that logits
are well separated - with random weights "bit-exact labels" would hinge on near-ties.  Output goes to
tests/golden/ckpt/ and is committed; rerun with ``python oracle/make_ckpt.py``.

Schema follows prototyping/source/training/{mlp,cnn}_trainer.py ``save()``.
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import pathlib
import sys

import numpy as np
import torch
from sklearn.preprocessing import StandardScaler

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import port  # noqa: E402
from guitar_audio_transcriber_ai_b200 import synth  # noqa: E402
from guitar_audio_transcriber_ai_b200.checkpoint import make_checkpoint  # noqa: E402
from guitar_audio_transcriber_ai_b200.config import CNNConfig, MFCCConfig, MLPConfig, MelSpecConfig, asdict  # noqa: E402
from guitar_audio_transcriber_ai_b200.training.cnn_trainer import CNN  # noqa: E402
from guitar_audio_transcriber_ai_b200.training.mlp_trainer import MLP  # noqa: E402


def _make_clip(args):
    sr, midi, variant, seed = args
    f0 = float(synth.midi_to_hz(midi)) * 2.0 ** (np.random.default_rng(seed).uniform(-0.15, 0.15) / 12.0)
    n = int(0.5 * sr) if variant != 1 else int(1.0 * sr)
    if variant == 2:  # what the slicer hands over: attack skipped, short body, noise tail, zero padding
        full = synth.note(f0, 0.35, sr, seed)
        body = full[int(0.1 * sr):]
        clip = np.zeros(n, np.float32)
        clip[:len(body)] = body
        gap = int(0.15 * sr)
        clip[len(body):len(body) + gap] = 1e-3 * np.random.default_rng(seed + 1).standard_normal(gap)
    else:
        clip = synth.note(f0, n / sr, sr, seed)
    return clip


def _features(args):
    torch.set_num_threads(1)
    sr, clip = args
    vec = port.mfcc_vector(clip, sr, 64, True, True, yin_on_normalized=False)
    img = port.melspec_image(clip, sr).numpy()
    return vec, img


def build_dataset(sr, per_class, pool):
    names = synth.class_names()
    label_of = {name: i for i, name in enumerate(names)}
    jobs, labels = [], []
    seed = 1_000_000 + sr
    for m in range(synth.MIDI_LO, synth.MIDI_HI + 1):
        for j in range(per_class):
            jobs.append((sr, m, j % 3, seed))
            labels.append(label_of[synth.midi_to_label(m)])
            seed += 1
    clips = pool.map(_make_clip, jobs, chunksize=16)
    feats = pool.map(_features, [(sr, c) for c in clips], chunksize=8)
    X = np.vstack([f[0] for f in feats]).astype(np.float32)
    imgs = [torch.from_numpy(f[1]) for f in feats]
    return X, imgs, np.asarray(labels), {i: n for i, n in enumerate(names)}


def train(model, batches, epochs, lr):
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=1e-4)
    loss_fn = torch.nn.CrossEntropyLoss(label_smoothing=0.05)
    hist = {"train_loss": [], "train_acc": []}
    for ep in range(epochs):
        model.train()
        tot, correct, n = 0.0, 0, 0
        for xb, yb in batches():
            opt.zero_grad()
            out = model(xb)
            loss = loss_fn(out, yb)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            tot += float(loss) * len(yb)
            correct += int((out.argmax(1) == yb).sum())
            n += len(yb)
        hist["train_loss"].append(tot / n)
        hist["train_acc"].append(correct / n)
        print(f"  epoch {ep}: loss {tot / n:.4f} acc {correct / n:.4f}", flush=True)
    hist["epoch"] = epochs
    model.eval()
    return hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "ckpt"))
    ap.add_argument("--per-class", type=int, default=30)
    ap.add_argument("--srs", default="22050,11025")
    a = ap.parse_args()
    out = pathlib.Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    with mp.Pool(min(8, mp.cpu_count())) as pool:
        for sr in [int(s) for s in a.srs.split(",")]:
            print(f"[make_ckpt] sr={sr}: building dataset", flush=True)
            X, imgs, y, reverse_map = build_dataset(sr, a.per_class, pool)
            yt = torch.from_numpy(y)
            g = torch.Generator().manual_seed(sr)

            if sr != 11025:  # the real MLP checkpoint covers 11025
                scaler = StandardScaler().fit(X)
                Xs = torch.from_numpy(scaler.transform(X).astype(np.float32))
                mlp = MLP(num_features=X.shape[1], hidden_dim=128, num_hidden_layers=2, num_classes=47, dropout=0.1)

                def mlp_batches():
                    perm = torch.randperm(len(y), generator=g)
                    for i in range(0, len(y), 32):
                        j = perm[i:i + 32]
                        yield Xs[j], yt[j]
                print("[make_ckpt] training MLP", flush=True)
                hist = train(mlp, mlp_batches, 40, 1e-3)
                ck = make_checkpoint(mlp, "mlp", asdict(MFCCConfig()),
                                     {k: str(v) if isinstance(v, pathlib.Path) else v for k, v in asdict(MLPConfig()).items()},
                                     sr, 0.5, reverse_map, scaler=scaler, histories=hist)
                torch.save(ck, out / f"mlp_synth_sr{sr}.ckpt")

            cnn = CNN(num_classes=47)
            by_T = {}
            for i, im in enumerate(imgs):
                by_T.setdefault(im.shape[-1], []).append(i)

            def cnn_batches():
                for T, idx in by_T.items():
                    idx = torch.tensor(idx)[torch.randperm(len(idx), generator=g)]
                    for i in range(0, len(idx), 32):
                        j = idx[i:i + 32]
                        yield torch.stack([imgs[k] for k in j.tolist()]), yt[j]
            print("[make_ckpt] training CNN", flush=True)
            hist = train(cnn, cnn_batches, 6, 1e-3)
            ck = make_checkpoint(cnn, "cnn", asdict(MelSpecConfig()),
                                 {k: str(v) if isinstance(v, pathlib.Path) else v for k, v in asdict(CNNConfig()).items()},
                                 sr, 0.5, reverse_map, histories=hist)
            torch.save(ck, out / f"cnn_synth_sr{sr}.ckpt")
    print("[make_ckpt] done", flush=True)


if __name__ == "__main__":
    main()
