"""Shared checks for the training-set feature builders (features.py:162-435): emulation (CPU) and GPU tests."""
import numpy as np
import torch

from tolerances import YIN_CENTS, cents, mel_ok, mfcc_ok


def make_dataset(n_classes, per_class, sr=22050, dur=0.5):
    from guitar_audio_transcriber_ai_b200 import synth
    clips, labels = [], []
    for c in range(n_classes):
        midi = 45 + 3 * c
        for k in range(per_class):
            clips.append(synth.note(float(synth.midi_to_hz(midi)), dur, sr, 1000 + 10 * c + k))
            labels.append(synth.midi_to_label(midi))
    order = np.random.default_rng(0).permutation(len(clips))          # labels not grouped
    return [clips[i] for i in order], [labels[i] for i in order]


def check_training_builders(device, n_classes=3, per_class=2):
    import port
    from guitar_audio_transcriber_ai_b200.audio.features import MelFeatureBuilder
    from guitar_audio_transcriber_ai_b200.audio.loading import AudioDatasetLoader
    sr = 22050
    clips, labels = make_dataset(n_classes, per_class, sr)
    loader = AudioDatasetLoader(clips, target_sr=sr, duration=0.5, labels=labels)
    fb = MelFeatureBuilder(device=device)
    # reference defaults: n_mfcc 13, raw audio, pitch feature appended (features.py:162-168)
    X, y, n, rev = fb.extract_mfcc_features(loader)
    assert X.shape == (len(clips), 14) and X.dtype == np.float32 and n == n_classes
    assert [rev[int(i)] for i in y] == labels and list(rev.values()) == sorted(set(labels))
    for row, c in zip(X, clips):
        want = port.mfcc_vector(c, sr, n_mfcc=13, normalize=False, add_pitch=True, yin_on_normalized=False)
        assert mfcc_ok(row[:13], want[:13])
        assert cents(10.0 ** float(row[13]), 10.0 ** float(want[13])) <= YIN_CENTS
    # reference defaults: 128 mels, n_fft 1024, hop 256, raw audio (features.py:276-283)
    M, y2, n2, rev2 = fb.extract_melspec_features(loader)
    assert isinstance(M, torch.Tensor) and tuple(M.shape) == (len(clips), 1, 128, 1 + int(sr * 0.5) // 256)
    assert np.array_equal(y, y2) and n2 == n and rev2 == rev
    for img, c in zip(M.numpy(), clips):
        assert mel_ok(img[0], np.asarray(port.melspec_image(c, sr, n_mels=128, n_fft=1024, hop_length=256, normalize=False))[0])
    # loaders: stratified split + scaler fitted on the training part, as sklearn does it in the reference
    dl_tr, dl_val, Xa, ya, n3, rev3, scaler = fb.build_mfcc_train_val_dataloaders(loader, n_mfcc=13, batch_size=4, val_size=0.5,
                                                                                  pin_memory=False)
    assert np.array_equal(Xa, X) and len(dl_tr.dataset) + len(dl_val.dataset) == len(clips)
    tr_x = torch.cat([b[0] for b in dl_tr]).numpy()
    assert np.abs(tr_x.mean(0)).max() < 1e-4 and scaler.mean_.shape == (14,)
    assert set(torch.cat([b[1] for b in dl_val]).tolist()) == set(y.tolist())       # stratified: every class in both parts
    dl, n4, rev4 = fb.build_melspec_dataloader(loader, batch_size=4, shuffle=False)
    xb, yb = next(iter(dl))
    assert tuple(xb.shape) == (4, 1, 128, M.shape[-1]) and yb.tolist() == y[:4].tolist()
    a, b, Xm, ym, n5, rev5 = fb.build_melspec_train_val_dataloaders(loader, batch_size=4, val_size=0.5, pin_memory=False)
    assert len(a.dataset) == len(b.dataset) == len(clips) // 2 and torch.equal(Xm, M)
