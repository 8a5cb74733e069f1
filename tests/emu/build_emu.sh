#!/bin/sh
# TEST INFRASTRUCTURE ONLY: compile csrc/gat.cu for the HOST against tests/emu/cpu_emu.h (see that file).
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
g++ -std=c++20 -O2 -g -fPIC -shared -pthread -ffp-contract=off -DGAT_CPU_EMU=1 -x c++ \
    -I"$here" -I"$root/guitar_audio_transcriber_ai_b200/csrc" \
    "$root/guitar_audio_transcriber_ai_b200/csrc/gat.cu" -o "$here/libgat_emu.so"
echo "built $here/libgat_emu.so"
