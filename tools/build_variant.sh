#!/usr/bin/env bash
# Development A/B builds: tools/build_variant.sh NAME file.cu [-D...]  ->  variants/NAME.so
# (the named source compiled with the extra flags, every other object taken from csrc/build/).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
NAME="$1"; SRC="$2"; shift 2
CS="$ROOT/ceigm-unet_b200/csrc"
mkdir -p "$ROOT/variants/obj"
OBJ="$ROOT/variants/obj/$NAME.o"
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --use_fast_math -Xcompiler -fPIC \
  -I"$ROOT/include" -I"$CS" -Xptxas -v "$@" -c -o "$OBJ" "$CS/$SRC" 2>&1 | grep -E "Used|spill" | grep -v " 0 bytes spill" || true
objs=()
for f in "$CS"/build/*.o; do [[ "$(basename "$f" .o)" == "$(basename "$SRC" .cu)" ]] || objs+=("$f"); done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$ROOT/variants/$NAME.so" "$OBJ" "${objs[@]}"
echo "built variants/$NAME.so"
