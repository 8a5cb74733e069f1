"""Host-side pieces of the product that need no GPU: tables, checkpoint layer, C ABI surface, sharding."""
import ctypes
import os
import pathlib
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import CKPT, ROOT

from guitar_audio_transcriber_ai_b200 import _lib, tables
from guitar_audio_transcriber_ai_b200.checkpoint import load_checkpoint, pack_cnn, pack_mlp
from guitar_audio_transcriber_ai_b200.training.cnn_trainer import CNN
from guitar_audio_transcriber_ai_b200.training.mlp_trainer import MLP


def test_tables_match_the_reference_libraries():
    import torchaudio
    import librosa_shim as L
    for sr in (22050, 11025):
        mine = tables.htk_fbanks(sr, 2048, 64)
        ta = torchaudio.transforms.MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=256, n_mels=64).mel_scale.fb.numpy()
        assert np.array_equal(mine, ta)                                   # bit-identical to what torchaudio multiplies by
        assert np.array_equal(tables.slaney_mel_fb(sr, 2048, 128), L.filters.mel(sr=sr, n_fft=2048, n_mels=128))
    assert np.array_equal(tables.hann_window_f32(2048), torch.hann_window(2048).numpy())
    assert np.array_equal(tables.hann_window_f64(2048), L.get_window_hann(2048))
    d = tables.dct_matrix(64, 128).astype(np.float64)
    assert np.abs(d @ d.T - np.eye(64)).max() < 1e-6


def test_sample_gate_threshold_is_the_exact_boundary():
    thr = tables.sample_gate_threshold(-32.5)
    below = np.nextafter(thr, np.float32(0))
    keep = lambda a: bool((20 * np.log10(np.array([a], np.float32) + 1e-10) > -32.5)[0])
    assert keep(thr) and not keep(below) and abs(float(thr) - 10 ** (-32.5 / 20)) < 1e-6


@pytest.mark.parametrize("n", [11, 216, 1723, 155040])
def test_percentile_index_reproduces_numpy(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    k, gamma = tables.percentile_index_f32(n, 20)
    s = np.sort(x)
    a, b = s[k], s[min(k + 1, n - 1)]
    diff = np.float32(b - a)
    q = np.float32(a + np.float32(diff * gamma))
    if gamma >= 0.5:
        q = np.float32(b - np.float32(diff * np.float32(np.float32(1) - gamma)))
    assert q == np.percentile(x, 20)


def test_onset_detect_params():
    assert tables.onset_detect_params(22050, 512) == {"pre_max": 1, "post_max": 1, "pre_avg": 4, "post_avg": 5, "wait": 1,
                                                       "delta": np.float32(0.07)}
    p = tables.onset_detect_params(11025, 512)
    assert (p["pre_max"], p["post_max"], p["pre_avg"], p["post_avg"], p["wait"]) == (0, 1, 2, 3, 0)


def test_real_mlp_checkpoint_loads_on_posix():
    ck = load_checkpoint(CKPT / "mlp_v1.0.0.ckpt")
    assert ck["config"]["target_sr"] == 11025 and ck["num_classes"] == 47 and ck["model_init_args"]["num_features"] == 65
    assert ck["reverse_map"][0] == "A#2" and len(ck["reverse_map"]) == 47
    assert ck["scaler"].mean_.shape == (65,) and ck["scaler"].mean_.dtype == np.float64
    m = MLP(**ck["model_init_args"])
    m.load_state_dict(ck["model"])                                        # strict
    packed = pack_mlp(ck["model"])
    assert packed["dims"].tolist() == [65, 128, 64, 47]
    assert packed["params"].size == 65 * 128 + 128 * 3 + 128 * 64 + 64 * 3 + 64 * 47 + 47
    with pytest.raises(FileNotFoundError):
        load_checkpoint(CKPT / "missing.ckpt")


def test_cnn_checkpoint_aliasing_and_bn_fold():
    ck = load_checkpoint(CKPT / "cnn_synth_sr22050.ckpt")
    keys = list(ck["model"].keys())
    assert len(keys) == 50 and any(k.startswith("net.0.") for k in keys) and any(k.startswith("features.") for k in keys)
    assert "use_batchnorm" not in ck["model_init_args"]
    m = CNN(**ck["model_init_args"]).eval()
    m.load_state_dict(ck["model"])
    packed = pack_cnn(ck["model"])
    assert [c["w"].shape for c in packed["convs"]] == [(9, 1, 32), (9, 32, 64), (9, 64, 128)]
    assert packed["fcs"][0]["w"].shape == (2048, 256) and packed["fcs"][1]["w"].shape == (256, 47)
    only_net = {k: v for k, v in ck["model"].items() if k.startswith("net.")}
    alt = pack_cnn(only_net)
    assert all(np.array_equal(a["w"], b["w"]) for a, b in zip(packed["convs"], alt["convs"]))
    # folded conv == conv + BatchNorm(eval) of the module
    x = torch.randn(2, 1, 64, 44)
    with torch.inference_mode():
        want = m.features[1](m.features[0](x))
        w = torch.from_numpy(packed["convs"][0]["w"]).permute(2, 1, 0).reshape(32, 1, 3, 3)
        got = torch.nn.functional.conv2d(x, w, torch.from_numpy(packed["convs"][0]["b"]), padding=1)
    assert (want - got).abs().max() < 1e-4


def test_cabi_exports_every_declared_symbol():
    header = (ROOT / "include" / "gat.h").read_text()
    declared = set(re.findall(r"\b(gat_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    path = _lib.build()
    lib = ctypes.CDLL(str(path))                                          # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    lib.gat_version.restype = ctypes.c_int
    assert lib.gat_version() >= 100
    lib.gat_last_error.restype = ctypes.c_char_p
    lib.gat_ctx_create.restype = ctypes.c_int
    assert lib.gat_ctx_create(None, 0, None) != 0 and b"null" in lib.gat_last_error()   # argument check, no compute


def test_product_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from guitar_audio_transcriber_ai_b200.engine import Engine
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from guitar_audio_transcriber_ai_b200.engine import Engine\n"
            "try:\n    Engine(22050, device='cpu')\nexcept Exception as e:\n    print(type(e).__name__, e)\n") % str(ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "GatError" in out and "no CPU fallback" in out


def test_product_never_imports_the_oracle():
    """The product path must not route through the CPU oracle (no import, no path manipulation towards it)."""
    pkg = ROOT / "guitar_audio_transcriber_ai_b200"
    bad = re.compile(r"^\s*(import|from)\s+(port|ref_env|librosa_shim|oracle|librosa)\b|sys\.path.*oracle|emu_loader|libgat_emu", re.M)
    for f in list(pkg.rglob("*.py")) + list((pkg / "csrc").glob("*.cu*")):
        assert not bad.search(f.read_text()), f


def test_shard_bounds_cover_everything():
    from guitar_audio_transcriber_ai_b200.parallel import shard_bounds
    for n in (0, 1, 7, 4096, 4097):
        for g in (1, 2, 3, 8):
            spans = [shard_bounds(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) == (-(-n // g) if n else 0)


def _gloo_worker(rank, world, port_no, n_total, q):
    import torch.distributed as dist
    from guitar_audio_transcriber_ai_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_bounds(n_total, world, rank)
    idx = torch.arange(lo, hi, dtype=torch.int64) % 47
    conf = (torch.arange(lo, hi, dtype=torch.float32) + 0.5) / n_total
    table = torch.stack([torch.arange(lo, hi), torch.arange(lo, hi) * 10, torch.arange(lo, hi) * 10 + 5], dim=1)
    rec = parallel.all_gather_records(parallel.pack_records(idx, conf, table), n_total)
    q.put((rank, rec.numpy()))
    dist.destroy_process_group()


def test_label_all_gather_world_size_2():
    import torch.multiprocessing as mp
    from guitar_audio_transcriber_ai_b200 import parallel
    n_total = 37                                                          # ragged: 19 + 18
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29611, n_total, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert np.array_equal(got[0], got[1]) and got[0].shape == (n_total, 4)
    idx, conf, start, end = parallel.unpack_records(torch.from_numpy(got[0]))
    assert idx.tolist() == [i % 47 for i in range(n_total)]
    assert torch.allclose(conf, (torch.arange(n_total, dtype=torch.float32) + 0.5) / n_total)
    assert start.tolist() == [10 * i for i in range(n_total)] and end.tolist() == [10 * i + 5 for i in range(n_total)]


def _sharded_worker(rank, world, port_no, q):
    """One rank of the sharded entry points on CPU: gloo + the host-emulated kernels (tests/emu)."""
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import emu_loader
    emu_loader.install()
    from guitar_audio_transcriber_ai_b200 import Transcriber, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    ck = ROOT / "tests" / "golden" / "ckpt"
    tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", ck, ck, device="cpu")
    Y = np.stack([synth.phrase(s, sr=22050, dur=2.0, n_notes=4)[0] for s in (0, 1, 2)])
    a = tr.transcribe_phrases_sharded(Y, 0.5)
    clips, _ = synth.clip_batch(5, 0.5, 22050, 40)
    b = tr.transcribe_notes_sharded(clips, 0.5, 22050)
    c = tr.transcribe_audio_sharded(Y[:2].reshape(-1), 22050, 0.5)
    keep = lambda r: {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in r.items() if k not in ("local_probs",)}
    q.put((rank, keep(a), keep(b), keep(c)))
    if world > 1:
        dist.destroy_process_group()


def test_sharded_entry_points_world_size_2_equal_world_size_1():
    """transcribe_phrases_sharded / transcribe_notes_sharded / transcribe_audio_sharded: both ranks of a world-size-2
    gloo run return the same result, and it equals the single-process result bit for bit (labels, confidences, slice
    tables, onsets, YIN) - the kernels run through the host emulation here, through CUDA in tests/test_gpu_parity.py."""
    if torch.cuda.is_available():
        pytest.skip("a real GPU is present: the CUDA 2-rank test covers this")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, 29613, q)) for r in range(2)]
    procs.append(ctx.Process(target=_sharded_worker, args=(0, 1, 29615, q)))
    [p.start() for p in procs]
    got = [q.get(timeout=600) for _ in procs]
    [p.join(60) for p in procs]
    strip = lambda r: {k: v for k, v in r.items() if k != "local_range"}
    for part in (1, 2, 3):
        ref = strip(got[0][part])
        assert len(ref["labels"]) > 0
        for g in got[1:]:
            assert strip(g[part]) == ref
    ranges = sorted(g[1]["local_range"] for g in got)
    assert ranges == [(0, 2), (0, 3), (2, 3)]
