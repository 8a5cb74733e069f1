"""Command line front end - mirror of the reference's transcribe_cli.py:16-114.

Same arguments (``--audio --out --save_clips --save_results``), same console table and ``<stem>_transcription.txt``
format.  Differences: there is no tkinter file dialog (``--audio`` is required), the boolean flags are real
switches, and checkpoints / device can be chosen (``--mlp_ckpt --cnn_ckpt --mlp_root --cnn_root --device``) because
the shipped tree only holds the MLP checkpoint.

    python -m guitar_audio_transcriber_ai_b200.transcribe_cli --audio take.wav --save_results
"""
from __future__ import annotations

import argparse
import tempfile
from pathlib import Path
from pprint import pformat

from .config import CLIP_DURATION, INFERENCE_OUTPUT_ROOT, TARGET_SR
from .transcribe import Transcriber


def format_table(result: dict) -> list[str]:
    """transcribe_cli.py:100-103."""
    lines = ["Idx |  Label |  Confidence | (YIN Note Estimate)"]
    for i, (lab, conf, y_info) in enumerate(zip(result["labels"], result["confidences"], result["dsp_info"])):
        lines.append(f"{i:03d}  {lab:>4}  (conf={conf:.2f})  {y_info[1]['note_name']}")
    return lines


def write_results(out_file: Path, result: dict) -> None:
    """transcribe_cli.py:105-110: ``idx,label,conf`` rows, a blank line, then the pretty-printed result dict."""
    with Path(out_file).open("w", encoding="utf-8") as f:
        for i, (lab, conf) in enumerate(zip(result["labels"], result["confidences"])):
            f.write(f"{i},{lab},{conf:.4f}\n")
        f.write("\nFull result dict:\n")
        f.write(pformat(result))


def main(argv=None) -> dict:
    parser = argparse.ArgumentParser(description="Guitar Audio Transcriber - B200 build of Prototype V1")
    parser.add_argument("--audio", type=str, required=True, help="Path to input .wav file")
    parser.add_argument("--out", type=str, default=None, help="Directory to save output file (default: ./output)")
    parser.add_argument("--save_clips", action="store_true", help="Keep the sliced clips on disk")
    parser.add_argument("--save_results", action="store_true", help="Write <stem>_transcription.txt")
    parser.add_argument("--mlp_ckpt", default=None)
    parser.add_argument("--cnn_ckpt", default=None)
    parser.add_argument("--mlp_root", default=None)
    parser.add_argument("--cnn_root", default=None)
    parser.add_argument("--device", default="cuda")
    args = parser.parse_args(argv)

    audio_path = Path(args.audio)
    if not audio_path.is_file():
        raise FileNotFoundError(f"Audio file not found: {audio_path}")
    if audio_path.suffix.lower() != ".wav":
        raise ValueError(f"Input file must be a .wav file: {audio_path}")
    out_dir = INFERENCE_OUTPUT_ROOT if args.out is None else Path(args.out)
    out_dir.mkdir(parents=True, exist_ok=True)
    out_file = out_dir / f"{audio_path.stem}_transcription.txt"

    transcriber = Transcriber(args.mlp_ckpt, args.cnn_ckpt, args.mlp_root, args.cnn_root, device=args.device)
    if args.save_clips:
        result = transcriber.transcribe(audio_path, out_root=out_dir, audio_name=audio_path.stem, target_sr=TARGET_SR,
                                        clip_duration=CLIP_DURATION)
    else:      # the reference slices into a temporary directory; nothing needs to touch the disk here
        with tempfile.TemporaryDirectory() as tmpdir:
            result = transcriber.transcribe(audio_path, out_root=Path(tmpdir), audio_name=audio_path.stem, target_sr=TARGET_SR,
                                            clip_duration=CLIP_DURATION, save_clips=False)
    print("\nTranscription Results:")
    print("\n".join(format_table(result)))
    if args.save_results:
        write_results(out_file, result)
        print(f"\nSaved transcription to {out_file}")
    return result


if __name__ == "__main__":
    print("\t- TRANSCRIBE CLI - B200 build of Base Version 1.0 -\n")
    main()
