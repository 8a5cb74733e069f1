"""TEST INFRASTRUCTURE ONLY - restatement of librosa.util.{peak_pick,localmin,match_events,fix_frames}."""
import numpy as np

from .core import _localmin, frame, tiny  # noqa: F401


def localmin(x, *, axis=0):
    return _localmin(np.asarray(x), axis=axis)


def peak_pick(x, *, pre_max, post_max, pre_avg, post_avg, delta, wait):
    """librosa.util.peak_pick, numba gufunc version (>= 0.10.1).

    Window lengths are ``ceil``-ed to ints; ``delta`` reaches the gufunc as float32; the window mean is
    a sequential sum in x's dtype divided by the count; frame 0 only looks forward.
    """
    x = np.asarray(x)
    pre_max = int(np.ceil(pre_max))
    post_max = int(np.ceil(post_max))
    pre_avg = int(np.ceil(pre_avg))
    post_avg = int(np.ceil(post_avg))
    wait = int(np.ceil(wait))
    delta = x.dtype.type(np.float32(delta))
    n_x = x.shape[0]
    peaks = np.zeros(n_x, dtype=bool)

    def _mean(seg):  # numba's np.mean on a 1-d slice: running sum then divide
        acc = x.dtype.type(0)
        for v in seg:
            acc = acc + v
        return acc / x.dtype.type(len(seg)) if len(seg) else x.dtype.type(np.nan)

    peaks[0] = x[0] >= np.max(x[:min(post_max, n_x)])
    peaks[0] &= x[0] >= _mean(x[:min(post_avg, n_x)]) + delta
    n = wait + 1 if peaks[0] else 1
    while n < n_x:
        maxn = np.max(x[max(0, n - pre_max):min(n + post_max, n_x)])
        peaks[n] = x[n] == maxn
        if not peaks[n]:
            n += 1
            continue
        avgn = _mean(x[max(0, n - pre_avg):min(n + post_avg, n_x)])
        peaks[n] &= x[n] >= avgn + delta
        if not peaks[n]:
            n += 1
            continue
        n += wait + 1
    return np.flatnonzero(peaks)


def fix_frames(frames, *, x_min=0, x_max=None, pad=True):
    frames = np.asarray(frames)
    if pad and (x_min is not None or x_max is not None):
        frames = np.clip(frames, x_min, x_max)
    if pad:
        pad_data = []
        if x_min is not None:
            pad_data.append(x_min)
        if x_max is not None:
            pad_data.append(x_max)
        frames = np.concatenate((np.asarray(pad_data, dtype=frames.dtype), frames))
    return np.unique(frames).astype(int)


def match_events_left(events_from, events_to):
    """librosa.util.match_events(left=True, right=False): nearest target that is <= each event."""
    events_from = np.asarray(events_from)
    events_to = np.asarray(events_to)
    idx = np.searchsorted(events_to, events_from, side="right") - 1
    if np.any(idx < 0):
        raise ValueError("no target at or before some event")
    return idx
