"""The librosa restatement (oracle/librosa_shim) has no librosa to be checked against (parity
unpinned), so it is pinned by analytic known answers and by independent implementations that ARE installed:
torchaudio's Slaney filterbank, transformers.audio_utils' librosa-style spectrogram, scipy, and a numba
build of the stencil texts librosa compiles."""
import numpy as np
import pytest
import scipy.fft
import torch

import librosa_shim as L


def _sine(f0, sr, dur=1.0, amp=0.5):
    t = np.arange(int(sr * dur)) / sr
    return (amp * np.sin(2 * np.pi * f0 * t)).astype(np.float32)


@pytest.mark.parametrize("sr,f0", [(22050, 110.0), (22050, 440.0), (11025, 196.0), (22050, 880.0)])
def test_yin_pure_tone(sr, f0):
    f = L.yin(_sine(f0, sr), fmin=50.0, fmax=1000.0, sr=sr)
    assert f.dtype == np.float64 and f.shape == (1 + sr // 512,)
    inner = f[3:-3]
    assert np.max(1200 * np.abs(np.log2(inner / f0))) < 3.0          # within 3 cents on fully filled frames


def test_yin_silence_reports_sr_over_min_period():
    f = L.yin(np.zeros(11025, np.float32), fmin=50.0, fmax=1000.0, sr=22050)
    assert np.all(f == 22050 / 22)                                    # all-zero CMND -> argmin 0 -> min_period


def test_parabolic_interpolation_matches_numba_stencil():
    numba = pytest.importorskip("numba")

    @numba.stencil
    def _pi_stencil(x):
        a = x[1] + x[-1] - 2 * x[0]
        b = (x[1] - x[-1]) / 2
        if np.abs(b) >= np.abs(a):
            return 0
        return -b / a

    @numba.guvectorize(["void(float32[:], float32[:])", "void(float64[:], float64[:])"], "(n)->(n)", nopython=True)
    def _pi_wrapper(x, y):
        y[:] = _pi_stencil(x)

    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        x = rng.random(4000).astype(dt)
        want = np.empty_like(x)
        _pi_wrapper(x, want)
        want[0] = want[-1] = 0
        got = L.core._parabolic_interpolation(x.reshape(-1, 1))[:, 0]
        assert np.array_equal(got, want)


def test_localmin_definition():
    x = np.array([3.0, 1.0, 1.0, 2.0, 0.5, 0.5, 0.4])
    assert L.util.localmin(x).tolist() == [False, True, False, False, True, False, True]


def test_mel_filterbank_against_torchaudio_and_transformers():
    import torchaudio
    from transformers.audio_utils import mel_filter_bank
    fb = L.filters.mel(sr=22050, n_fft=2048, n_mels=128)
    assert fb.shape == (128, 1025) and fb.dtype == np.float32
    ta = torchaudio.functional.melscale_fbanks(1025, 0.0, 11025.0, 128, 22050, norm="slaney", mel_scale="slaney").numpy().T
    hf = mel_filter_bank(1025, 128, 0.0, 11025.0, 22050, norm="slaney", mel_scale="slaney").T
    assert np.abs(fb - ta).max() < 5e-7 and np.abs(fb - hf).max() < 5e-7
    # each filter is a non-negative triangle and every interior bin feeds at most two filters
    assert fb.min() >= 0 and np.max((fb > 0).sum(0)) <= 2
    # at sr 11025 librosa's bin centres reach 5512.5 Hz (rfftfreq), torchaudio's stop at 5512: they differ
    fb11 = L.filters.mel(sr=11025, n_fft=2048, n_mels=128)
    ta11 = torchaudio.functional.melscale_fbanks(1025, 0.0, 5512.5, 128, 11025, norm="slaney", mel_scale="slaney").numpy().T
    assert 1e-6 < np.abs(fb11 - ta11).max() < 1e-2


def test_melspectrogram_against_transformers_pipeline():
    from transformers.audio_utils import mel_filter_bank, spectrogram, window_function
    y = _sine(330.0, 22050, 0.5) + 0.01 * np.random.default_rng(1).standard_normal(11025).astype(np.float32)
    S = L.feature.melspectrogram(y=y, sr=22050)
    fb = mel_filter_bank(1025, 128, 0.0, 11025.0, 22050, norm="slaney", mel_scale="slaney")
    ref = spectrogram(y.astype(np.float64), window_function(2048, "hann", periodic=True), frame_length=2048, hop_length=512,
                      fft_length=2048, power=2.0, center=True, pad_mode="constant", mel_filters=fb)
    assert S.shape == ref.shape == (128, 22)
    assert np.max(np.abs(S - ref) / (np.abs(ref) + 1e-6 * ref.max())) < 2e-3
    db = L.power_to_db(S)
    assert db.max() - db.min() <= 80.0 + 1e-4 and db.dtype == np.float32


def test_mfcc_is_orthonormal_dct_of_log_mel():
    y = _sine(262.0, 22050, 0.5)
    S = L.power_to_db(L.feature.melspectrogram(y=y, sr=22050))
    M = L.feature.mfcc(y=y, sr=22050, n_mfcc=64)
    basis = scipy.fft.dct(np.eye(128), type=2, norm="ortho", axis=0)
    assert np.allclose(basis @ basis.T, np.eye(128), atol=1e-12)
    assert M.shape == (64, 22) and M.dtype == np.float32
    assert np.abs(M - (basis[:64] @ S.astype(np.float64))).max() < 1e-3


def test_rms_uses_numpy_pairwise_float32_sum():
    """csrc/onset.cuh's rms_db_kernel reproduces this order: 128-blocks, 8 interleaved lanes, binary tree."""
    rng = np.random.default_rng(2)
    y = rng.standard_normal(6000)
    got = L.feature.rms(y=y, pad_mode="reflect")[0]
    yp = np.pad(y, 1024, mode="reflect")
    col = np.square(yp[512 * 3: 512 * 3 + 2048].astype(np.float32))

    def block(b):
        r = [b[i] for i in range(8)]
        for i in range(8, 128):
            r[i % 8] = np.float32(r[i % 8] + b[i])
        return np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) + np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
    sums = [block(col[k * 128:(k + 1) * 128]) for k in range(16)]
    while len(sums) > 1:
        sums = [np.float32(sums[i] + sums[i + 1]) for i in range(0, len(sums), 2)]
    assert got.dtype == np.float32 and got[3] == np.sqrt(np.float32(sums[0] / np.float32(2048)))


def test_peak_pick_brute_force_and_backtrack():
    rng = np.random.default_rng(3)
    x = rng.random(300)
    kw = dict(pre_max=1, post_max=1, pre_avg=4, post_avg=5, delta=0.07, wait=1)
    got = L.util.peak_pick(x, **kw).tolist()
    want, n = [], 0
    while n < len(x):
        lo = max(0, n - 1)
        is_max = x[n] >= x[:1].max() if n == 0 else x[n] == x[lo:n + 1].max()
        seg = x[max(0, n - 4) if n else 0:n + 5]
        if is_max and x[n] >= seg.sum() / len(seg) + np.float64(np.float32(0.07)) - 1e-15:
            want.append(n)
            n += 2
        else:
            n += 1
    assert got == want
    env = np.array([0, 0, 0, 1, 3, 2, 1, 1, 4, 0, 0, 2, 5, 1], dtype=float)
    assert L.onset.onset_backtrack(np.array([4, 8, 12]), env).tolist() == [2, 7, 10]


def test_onset_click_train():
    sr = 22050
    y = 1e-4 * np.random.default_rng(4).standard_normal(sr * 2)
    clicks = [4000, 15000, 27000, 38000]
    for c in clicks:
        y[c:c + 400] += 0.8 * np.sin(2 * np.pi * 880 * np.arange(400) / sr) * np.exp(-np.arange(400) / 100)
    env = L.onset.onset_strength(y=y, sr=sr, hop_length=512)
    assert env.shape == (1 + len(y) // 512,) and np.all(env[:3] == 0)
    frames = L.onset.onset_detect(onset_envelope=env, sr=sr, hop_length=512, backtrack=False)
    for c in clicks:
        assert np.min(np.abs(frames * 512 - c)) <= 3 * 512 + 1024      # flux is delayed by the 3-frame compensation window
    assert L.frames_to_samples(np.array([2, 5]), hop_length=512).tolist() == [1024, 2560]


def test_note_names():
    assert L.midi_to_note(54) == "F♯3" and L.midi_to_note(40) == "E2" and L.midi_to_note(86) == "D6"
    assert abs(float(L.hz_to_midi(440.0)) - 69.0) < 1e-12
