"""Shared check for the streaming path (transcribe_live mirror): emulation (CPU) and GPU tests."""
import numpy as np
import torch

from conftest import CKPT


def live_reference_loop(blocks, sr, mlp_ck, cnn_ck):
    """The prototype's state machine (prototyping/source/transcribe_live.py:113-267) restated on the oracle:
    deque ring buffer, port.detect_onsets(hop 1024, min_sep 0.3), port.transcribe_note."""
    import collections
    import port
    ring = collections.deque(maxlen=int(1.5 * sr))
    pending, results = collections.deque(), []
    for blk in blocks:
        ring.extend(blk.tolist())
        if len(ring) == ring.maxlen:
            buf = np.array(list(ring), dtype=np.float32)
            onsets = [int(o) for o in port.detect_onsets(buf, sr, hop_len=1024, min_sep=0.3)]
            h_idx, min_len = 0, 0.3 * sr

            def note(s):
                t = int(0.5 * sr)
                return s[:t] if len(s) > t else np.pad(s, (0, t - len(s)))
            if len(onsets) == 1:
                s = buf[onsets[0]:-1]
                if len(s) > min_len:
                    pending.append(note(s)); h_idx = onsets[0]; onsets = []
            while len(onsets) >= 2:
                s = buf[onsets[0]:onsets[1]]
                if len(s) > min_len:
                    pending.append(note(s)); h_idx = onsets[1]; onsets = onsets[2:]
                else:
                    h_idx = onsets[0]; onsets = onsets[1:]
            for _ in range(h_idx + 1):
                ring.pop()
        if pending:
            results.append(port.transcribe_note(mlp_ck, cnn_ck, pending.popleft(), 0.5, sr))
    return results


def check_live(device, tol=5e-5):
    import ref_env
    from guitar_audio_transcriber_ai_b200 import Transcriber, synth
    from guitar_audio_transcriber_ai_b200.transcribe_live import LiveTranscriber, RingBuffer
    sr = 22050
    y, _, _ = synth.phrase(1, sr=sr)
    blocks = [y[i:i + 1024] for i in range(0, len(y), 1024)]
    tr = Transcriber("mlp_synth_sr22050.ckpt", "cnn_synth_sr22050.ckpt", CKPT, CKPT, device=device)
    live = LiveTranscriber(device=device, sample_rate=sr, transcriber=tr)
    got = []
    for blk in blocks:
        live.feed(blk[:, None])
        got += live.step()
    torch.set_num_threads(1)
    want = live_reference_loop(blocks, sr, ref_env.load_ckpt(CKPT / "mlp_synth_sr22050.ckpt"), ref_env.load_ckpt(CKPT / "cnn_synth_sr22050.ckpt"))
    assert len(got) == len(want) and len(got) >= 3
    for a, b in zip(got, want):
        assert [str(s) for s in a["labels"]] == [str(s) for s in b["labels"]]
        assert np.abs(a["probs"] - b["probs"]).max() <= tol
    rb = RingBuffer(5)
    rb.push(np.arange(8, dtype=np.float32))
    assert rb.is_full() and rb.get_buffer().tolist() == [3, 4, 5, 6, 7] and rb.get_slice(1, 3).tolist() == [4, 5]
    rb.clear_from(2)
    assert rb.get_buffer().tolist() == [3, 4, 5] and rb.get_slice(2, 9).size == 0
    tr.engine.close()
