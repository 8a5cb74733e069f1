"""Clip sharding across the GPUs of one box and the label all-gather.

After slicing every clip is independent, so rank r simply owns a contiguous block of ceil(N/G) clips and
runs the whole pipeline on it; weights are replicated.  The only collective in the system is one
all-gather of a fixed-width record per clip (label index, confidence, slice start, slice end), padded so
every rank contributes the same count.  ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) does it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ceil(N/G) items for ``rank`` (the last ranks may be short or empty)."""
    per = -(-n_items // world_size) if n_items > 0 else 0
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def pack_records(indices: torch.Tensor, conf: torch.Tensor, table: torch.Tensor | None = None) -> torch.Tensor:
    """[n, 4] int64 records: label index, float32 confidence bits, start sample, end sample."""
    n = indices.shape[0]
    rec = torch.zeros((n, 4), dtype=torch.int64, device=indices.device)
    rec[:, 0] = indices.to(torch.int64)
    rec[:, 1] = conf.to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    if table is not None and n:
        rec[:, 2] = table[:, -2].to(rec.device)
        rec[:, 3] = table[:, -1].to(rec.device)
    return rec


def unpack_records(rec: torch.Tensor):
    idx = rec[:, 0]
    conf = rec[:, 1].to(torch.int32).view(torch.float32)
    return idx, conf, rec[:, 2], rec[:, 3]


def _all_gather_into(out: torch.Tensor, inp: torch.Tensor, group=None) -> None:
    """``dist.all_gather_into_tensor``; the gloo backend (CPU tests, or several ranks sharing one GPU in a test) has no
    CUDA all-gather, so CUDA tensors are staged through the host for THAT backend only.  NCCL never takes this branch."""
    if inp.is_cuda and dist.get_backend(group) == "gloo":
        h_out = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(h_out, inp.cpu(), group=group)
        out.copy_(h_out)
        return
    dist.all_gather_into_tensor(out, inp, group=group)


def all_gather_records(rec: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Gathers every rank's [n_r, 4] block (rank r holds shard_bounds(n_total, G, r)) into [n_total, 4]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    per = -(-n_total // world)
    padded = torch.zeros((per, 4), dtype=torch.int64, device=rec.device)
    padded[: rec.shape[0]] = rec
    out = torch.empty((world * per, 4), dtype=torch.int64, device=rec.device)
    _all_gather_into(out, padded, group)
    return out[:n_total]


# --------------------------------------------------------------------------------------------------------------------
# Sharded entry points (one process per GPU).  Every rank calls with the SAME arguments, works on its own block and
# ends with the all-gather of fixed-width int64 rows - the only collective on the path.  Results are identical on every
# rank and identical, bit for bit, to the single-GPU call (tests/test_gpu_parity.py, tests/test_host_logic.py).
def _world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_rows(rows: torch.Tensor, cap: int, group=None) -> torch.Tensor:
    """All-gather of ``[k_r, C]`` int64 rows whose count differs per rank, WITHOUT exchanging the counts: every rank
    pads its block to the same static ``cap`` rows (a bound all ranks can compute from the call's arguments) with rows
    whose first column is -1, and ONE ``all_gather_into_tensor`` moves them.  No host synchronisation: the result
    ``[G * cap, C]`` stays on the device, padding included; ``valid_rows`` drops it."""
    world, _ = _world(group)
    if world == 1:
        return rows
    C = rows.shape[1]
    if rows.shape[0] > cap:
        raise ValueError(f"gather_rows: {rows.shape[0]} rows exceed the static bound {cap}")
    padded = torch.full((cap, C), -1, dtype=torch.int64, device=rows.device)
    padded[: rows.shape[0]] = rows
    out = torch.empty((world * cap, C), dtype=torch.int64, device=rows.device)
    _all_gather_into(out, padded, group)
    return out


def valid_rows(rows: torch.Tensor) -> torch.Tensor:
    """Drops the padding rows of ``gather_rows`` (first column < 0); rank order and row order are preserved."""
    return rows[rows[:, 0] >= 0]


def _f32_bits(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float32).contiguous().view(torch.int32).to(torch.int64)


def _f64_bits(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float64).contiguous().view(torch.int64)


def clip_rows(table4: torch.Tensor, out: dict) -> torch.Tensor:
    """[k, 7] int64 rows per clip: signal, onset index, start, end, label index, float32 confidence bits,
    float64 YIN median bits.  ``table4`` is gat_segment_batch's table (or an equivalent built by the caller)."""
    hz = out.get("yin_hz")
    hz_bits = _f64_bits(hz) if hz is not None else torch.zeros_like(out["indices"])
    return torch.cat([table4, out["indices"][:, None], _f32_bits(out["confidences"])[:, None], hz_bits[:, None]], dim=1)


def _result_from_rows(tr, rows: torch.Tensor, with_hz: bool) -> dict:
    rows = valid_rows(rows).cpu()
    idx = rows[:, 4].numpy()
    res = {
        "indices": idx,
        "labels": [tr.predictor.reverse_map[int(i)] for i in idx],
        "confidences": rows[:, 5].to(torch.int32).view(torch.float32).numpy(),
        "slice_table": rows[:, :4].numpy(),
    }
    if with_hz:
        res["dsp_info"] = tr._dsp_info(rows[:, 6].contiguous().view(torch.float64).numpy())
    return res


def transcribe_notes_sharded(tr, audio, clip_duration, sr_in, group=None) -> dict:
    """Transcriber.transcribe_notes over the ranks of ``group``: rank r transcribes clips ``shard_bounds(N, G, r)``
    of the common ``[N, n]`` batch (transcribe.py:147-199 per clip), then one all-gather of [label, confidence]
    rows.  Returns indices / labels / confidences for ALL N clips on every rank (probabilities stay per rank:
    ``local_probs`` with ``local_range``)."""
    world, rank = _world(group)
    a = audio if torch.is_tensor(audio) else torch.as_tensor(audio)
    if a.dim() == 1:
        a = a.unsqueeze(0)
    N = a.shape[0]
    lo, hi = shard_bounds(N, world, rank)
    eng = tr.engine
    rows = torch.zeros((hi - lo, 7), dtype=torch.int64, device=eng.device)
    local = None
    if hi > lo:
        local = tr.transcribe_notes_device(a[lo:hi], clip_duration, sr_in)
        rows[:, 1] = torch.arange(lo, hi, device=eng.device)
        rows[:, 4] = local["indices"]
        rows[:, 5] = _f32_bits(local["confidences"])
    res = _result_from_rows(tr, gather_rows(rows, -(-N // world) if N else 0, group), with_hz=False)
    del res["slice_table"]
    res["local_range"] = (lo, hi)
    res["local_probs"] = local["probs"].cpu().numpy() if local is not None else None
    return res


def phrases_rows_device(tr, signals, clip_duration, signal_offset: int = 0, gather: bool = True, group=None,
                        want_onsets: bool = False, signals_per_rank: int | None = None):
    """The device half of transcribe_phrases_sharded for the signals THIS rank owns (``signals[P_r, L]``, global
    index of the first one = ``signal_offset``): ONE gat_segment_batch, ONE batched ensemble + YIN over the kept
    clips, then (``gather``) ONE all-gather of the [k, 7] int64 rows, padded to the static bound
    ``signals_per_rank * (max_onsets - 1)`` (K onsets yield at most K - 1 clips) so that no count has to be exchanged.
    Returns the rows on the device, padding rows (first column -1) included - ``valid_rows`` drops them - and, with
    ``want_onsets``, the per-signal onset rows [P, 1 + max_onsets] = (count, onsets...)."""
    eng = tr.engine
    Y = signals if torch.is_tensor(signals) else torch.as_tensor(signals)
    sp = eng.slicer_params(Y.shape[1], clip_duration)
    max_onsets = max(2, Y.shape[1] // max(1, sp.min_sep_samples) + 2)
    per_rank = Y.shape[0] if signals_per_rank is None else signals_per_rank
    rows = torch.zeros((0, 7), dtype=torch.int64, device=eng.device)
    onset_rows = torch.zeros((0, 1 + max_onsets), dtype=torch.int64, device=eng.device)
    if Y.shape[0]:
        seg = eng.segment_batch(Y, clip_duration)
        if seg["clips"].shape[0]:
            out = tr._ensemble_sliced(seg["clips"])
            table = seg["table"].clone()
            table[:, 0] += signal_offset
            rows = clip_rows(table, out)
        if want_onsets:
            onset_rows = torch.cat([seg["n_onsets"].to(torch.int64)[:, None], seg["onsets"]], dim=1)
    if gather:
        rows = gather_rows(rows, per_rank * (max_onsets - 1), group)
    if not want_onsets:
        return rows
    return rows, (gather_rows(onset_rows, per_rank, group) if gather else onset_rows)


def transcribe_phrases_sharded(tr, phrases, clip_duration, group=None) -> dict:
    """SURVEY 8(e) option (i): the recording is P independent signals ``[P, L]`` at the checkpoint's rate.  Rank r
    owns signals ``shard_bounds(P, G, r)``: ONE gat_segment_batch over them (every signal sliced as a file of its
    own, slicing.py:147-165), ONE batched ensemble + YIN over the kept clips (transcribe.py:118-143), then the
    all-gather of the per-clip rows and of the per-signal onset lists.  ``slice_table`` rows are
    (signal, onset index, start sample, end sample) in (signal, onset) order."""
    world, rank = _world(group)
    Y = phrases if torch.is_tensor(phrases) else torch.as_tensor(phrases)
    if Y.dim() != 2:
        raise ValueError("transcribe_phrases_sharded: phrases must be [P, L]")
    P = Y.shape[0]
    lo, hi = shard_bounds(P, world, rank)
    rows, on = phrases_rows_device(tr, Y[lo:hi], clip_duration, signal_offset=lo, gather=True, group=group, want_onsets=True,
                                   signals_per_rank=-(-P // world) if P else 0)
    res = _result_from_rows(tr, rows, with_hz=True)
    on = valid_rows(on).cpu().numpy()
    res["onsets"] = [on[p, 1:1 + int(on[p, 0])].tolist() for p in range(P)]
    res["local_range"] = (lo, hi)
    return res


def audio_rows_device(tr, y, clip_duration, group=None, want_seg: bool = False):
    """The device half of transcribe_audio_sharded: whole-file segmentation of ``y`` (already at the checkpoint's
    rate) on THIS rank, ensemble + YIN on this rank's block of the sliced clips, all-gather of the [k, 7] rows."""
    world, rank = _world(group)
    eng = tr.engine
    seg = eng.segment(y, clip_duration)
    K = seg["clips"].shape[0]
    lo, hi = shard_bounds(K, world, rank)
    rows = torch.zeros((0, 7), dtype=torch.int64, device=eng.device)
    if hi > lo:
        out = tr._ensemble_sliced(seg["clips"][lo:hi])
        table = torch.zeros((hi - lo, 4), dtype=torch.int64, device=eng.device)
        table[:, 1:] = seg["table"][lo:hi]
        rows = clip_rows(table, out)
    rows = gather_rows(rows, -(-K // world) if K else 0, group)       # every rank knows K: the segmentation ran on all of them
    return (rows, seg, (lo, hi)) if want_seg else rows


def transcribe_audio_sharded(tr, y, sr, clip_duration, group=None) -> dict:
    """SURVEY 8(e) option (ii): ONE contiguous signal.  Whole-file onset detection has global dependencies (dB
    maximum, percentile gate, envelope min / max, the sequential ``wait`` and min-sep scans): every rank runs it on
    the full signal - same kernels, same result, no broadcast - and then transcribes only ITS block of the sliced
    clips; the per-clip rows are all-gathered.  The segmentation is the serial (Amdahl) term of this mode."""
    eng = tr.engine
    target_sr = tr._target_sr()
    yt = torch.as_tensor(y)
    if sr is not None and sr != target_sr:
        yt = eng.resample(yt.reshape(-1), sr, target_sr)
    rows, seg, rng = audio_rows_device(tr, yt, clip_duration, group, want_seg=True)
    if seg["clips"].shape[0] == 0:
        raise FileNotFoundError("load_audio_dataset: No audio files found.")
    res = _result_from_rows(tr, rows, with_hz=True)
    res["slice_table"] = res["slice_table"][:, 1:]
    res["onsets"] = [int(v) for v in seg["onsets"].cpu().numpy()]
    res["local_range"] = rng
    return res
