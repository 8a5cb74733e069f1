import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests" / "emu"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLD = ROOT / "tests" / "golden"
CKPT = GOLD / "ckpt"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_clips_22050():
    return np.load(GOLD / "clips_sr22050.npz")


@pytest.fixture(scope="session")
def golden_clips_11025():
    return np.load(GOLD / "clips_sr11025.npz")


@pytest.fixture(scope="session")
def golden_phrases():
    return np.load(GOLD / "phrases_sr22050.npz")


def golden_audio(g, k):
    """Regenerates the k-th input clip of a clips_*.npz fixture from its seed."""
    from guitar_audio_transcriber_ai_b200 import synth
    seed, dur, sr = int(g["seeds"][k]), float(g["durations"][k]), int(g["sr"])
    return synth.note(float(synth.midi_to_hz(synth.random_midi(seed))), dur, sr, seed)
