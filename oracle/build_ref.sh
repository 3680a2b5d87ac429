#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's own hot-path translation units, from where
# they lie under /root/reference (nothing is copied), headless for sm_100, and links them with
# oracle/ref_harness.cu into oracle/_ref/libyhref.so (default flags = the shipped arithmetic)
# and oracle/_ref/libyhref_nofma.so (--fmad=false, for the bitwise tier).  The reference's
# Makefile (sm_61, -lglut -lGL -lGLEW -lSOIL) is NOT used; main.cu / openGL_functions.cu need
# GL and are not built.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/reactionDiffusion.cu" ]; then
  echo "build_ref: $REF not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TUS="reactionDiffusion advFDBFECC integralTrapz symmetryReduction tipTracker spaceAPD singleCell linearSolver helper_functions printFunctions saveFiles"
SRCS=""
for t in $TUS; do SRCS="$SRCS $REF/$t.cu"; done
COMMON="-gencode arch=compute_100,code=sm_100 -rdc=true -std=c++11 -O3 -w -Xcompiler -fPIC -I$REF -I$HERE/.. --shared"
nvcc $COMMON -o "$OUT/libyhref.so" $SRCS "$HERE/ref_harness.cu" &
nvcc $COMMON --fmad=false -o "$OUT/libyhref_nofma.so" $SRCS "$HERE/ref_harness.cu" &
wait
ls -la "$OUT"
