#!/bin/sh
# SASS evidence that the shipped libgat.so runs conv2 / conv3 / FC1 on tcgen05 with TMA-fed operands:
# per kernel, the count of UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
# SYNCS (mbarrier), the packed-FP32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2) and the register / shared-memory footprint.   sh tools/sass_evidence.sh > profiles/r02_sass_tc.txt
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
so="$root/guitar_audio_transcriber_ai_b200/csrc/libgat.so"
echo "# $(basename "$so")  sha256 $(sha256sum "$so" | cut -c1-16)  git $(git -C "$root" rev-parse --short HEAD 2>/dev/null)"
echo "# cuobjdump -sass, instructions per kernel"
cuobjdump -sass "$so" | awk '
  /Function :/ { name=$3; order[++n]=name }
  /UTCHMMA/ { mma[name]++ } /LDTM/ { ldtm[name]++ } /UBLKCP/ { blk[name]++ } /UTCBAR/ { bar[name]++ } /SYNCS/ { syn[name]++ }
  /FFMA2|FADD2|FMUL2/ { pk[name]++ }
  /^[ \t]*\/\*[0-9a-f]+\*\// { ins[name]++ }
  END { printf "%-110s %8s %8s %6s %7s %7s %6s %9s\n", "kernel", "instr", "UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "F32x2 ops";
        for (i=1;i<=n;i++) { k=order[i]; printf "%-110s %8d %8d %6d %7d %7d %6d %9d\n", substr(k,1,110), ins[k], mma[k], ldtm[k], blk[k], bar[k], syn[k], pk[k] } }'
echo
echo "# cuobjdump -res-usage"
cuobjdump -res-usage "$so" 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/Fatbin elf code://' | awk '{ $1=""; print }' | cut -c1-220
