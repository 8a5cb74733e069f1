"""TEST INFRASTRUCTURE ONLY - load the host-emulation build of csrc/gat.cu (see cpu_emu.h).

``install()`` builds tests/emu/libgat_emu.so with g++ if needed and makes ``_lib.load()`` return it, marked
``_host_emulation`` so ``Engine`` accepts device="cpu".  Only tests/test_emu_*.py call this; the product
package contains no path to this library.
"""
from __future__ import annotations

import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
SO = HERE / "libgat_emu.so"


def build() -> pathlib.Path:
    srcs = list((ROOT / "guitar_audio_transcriber_ai_b200" / "csrc").glob("*.cu*")) + [HERE / "cpu_emu.h", ROOT / "include" / "gat.h"]
    if SO.exists() and all(SO.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return SO
    proc = subprocess.run([str(HERE / "build_emu.sh")], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(proc.stdout + proc.stderr)
    return SO


def install():
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    from guitar_audio_transcriber_ai_b200 import _lib
    lib = _lib.GatLib(build())
    lib._host_emulation = True
    _lib._LIB = lib
    _lib.load = lambda: lib
    return lib
