#!/bin/bash
# tools/build_variant.sh <name> <translation unit> <extra nvcc flags...>
# Link ab/libsetk_b200_<name>.so from the in-tree objects with ONE translation unit
# recompiled with extra flags (measurement builds for tools/ab_fused.py).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME=$1; TU=$2; shift 2
OBJ="$ROOT/setk_b200/csrc/obj"
mkdir -p "$ROOT/ab"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
  -I"$ROOT/include" -Xptxas -v "$@" -c "$ROOT/setk_b200/csrc/$TU.cu" -o "$ROOT/ab/$TU.$NAME.o" 2> "$ROOT/ab/$TU.$NAME.log"
objs=""
for f in api generic weights weights_coop weights_post stft_cov_fused stft_cov_ws apply_istft_fused stft_spill cov_mma cgmm wpe spatial cm_mask; do
  if [ "$f" = "$TU" ]; then objs="$objs $ROOT/ab/$TU.$NAME.o"; else objs="$objs $OBJ/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$ROOT/ab/libsetk_b200_$NAME.so" $objs
grep -E -A2 "ILi4ELi4" "$ROOT/ab/$TU.$NAME.log" | grep -E "registers|spill" | head -4
echo "built ab/libsetk_b200_$NAME.so"
