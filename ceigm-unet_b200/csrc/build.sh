#!/usr/bin/env bash
# Builds libss2d_b200.so in-tree for sm_100a (the only target). Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libss2d_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --use_fast_math
       -Xcompiler -fPIC -shared -I"$ROOT/include" -I"$HERE" -Xptxas -v "$@")
"$NVCC" "${FLAGS[@]}" -o "$OUT" "$HERE"/api.cu "$HERE"/scan_fwd.cu "$HERE"/scan_bwd.cu "$HERE"/scan_bwd2.cu "$HERE"/scan_par.cu "$HERE"/cross.cu "$HERE"/epilogue.cu "$HERE"/wgrad.cu "$HERE"/layernorm.cu "$HERE"/dwconv.cu \
  2> "$HERE/../build_ptxas.log" || { cat "$HERE/../build_ptxas.log" >&2; exit 1; }
echo "built $OUT"
