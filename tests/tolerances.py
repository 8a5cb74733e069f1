"""Stated float32 tolerances for CUDA-vs-oracle parity (BASELINE.json: "e.g. max rel err 1e-4 on dB features").

dB features (mel image):  |d| <= 1e-4 * max(|ref|, 20 dB).  The floor matters because the CNN chain has no
    top_db clamp: bins ~70 dB under the clip's peak carry the float32 FFT's own rounding noise (both in
    torch's pocketfft and in ours), which is relative to the frame energy, not to the bin.
MFCC time-means:          |d| <= 1e-4 * max(|ref|, 1)
YIN median f0:            |cents| <= 0.05 ; log10(Hz) feature |d| <= 2e-6 ; frame-wise f0 cents <= 0.5 on
                          frames that carry signal (the reference's own float32 FFT autocorrelation is noisy at
                          that level on near-silent frames)
probabilities:            |d| <= 2e-5 ; logits |d| <= 1e-4 * max(|ref|, 1)
onset envelope (fp64):    |d| <= 1e-10
onset frames / samples, slice tables, label indices and labels: exact.
"""
import numpy as np

MEL_REL, MEL_FLOOR = 1e-4, 20.0
MFCC_REL, MFCC_FLOOR = 1e-4, 1.0
YIN_CENTS = 0.05
PROB_ABS = 2e-5
ENV_ABS = 1e-10


def mel_ok(got, ref):
    return bool(np.all(np.abs(got - ref) <= MEL_REL * np.maximum(np.abs(ref), MEL_FLOOR)))


def mfcc_ok(got, ref):
    return bool(np.all(np.abs(got - ref) <= MFCC_REL * np.maximum(np.abs(ref), MFCC_FLOOR)))


def cents(a, b):
    return 1200.0 * np.abs(np.log2(np.asarray(a, dtype=np.float64) / np.asarray(b, dtype=np.float64)))


def mel_ok_degenerate(got, ref, dyn_db=70.0):
    """For degenerate inputs (silence, DC, a single click) most mel bands hold nothing but the float32
    FFT's rounding noise, ~-150 dB under the peak and different in every FFT implementation.  Compare in dB
    only the bands within ``dyn_db`` of the clip's peak; the rest must merely sit below that line."""
    peak = ref.max()
    strong = ref >= peak - dyn_db
    ok_strong = np.all(np.abs(got[strong] - ref[strong]) <= MEL_REL * np.maximum(np.abs(ref[strong]), MEL_FLOOR))
    ok_weak = np.all(got[~strong] <= peak - dyn_db + 1.0)
    return bool(ok_strong and ok_weak)
