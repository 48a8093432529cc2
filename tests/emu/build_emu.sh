#!/bin/bash
# Build the CPU-emulated twin of libsetk_b200.so for the CPU test tier.
# TEST INFRASTRUCTURE ONLY: the product never loads this library.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
SRC="$ROOT/setk_b200/csrc"
OUT="$HERE/libsetk_b200_emu.so"
OBJ="$HERE/obj"
mkdir -p "$OBJ"
pids=()
for f in api generic weights weights_coop weights_post stft_cov_fused stft_cov_ws apply_istft_fused stft_spill cov_mma cgmm wpe spatial cm_mask; do
  g++ -O2 -std=c++17 -DSETK_EMU -DSETK_TABLE_CHUNK=8 -fPIC -pthread -I"$HERE" -I"$ROOT/include" -x c++ -c "$SRC/$f.cu" -o "$OBJ/$f.o" &
  pids+=($!)
done
g++ -O2 -std=c++17 -DSETK_EMU -fPIC -pthread -I"$HERE" -c "$HERE/cuda_emu.cc" -o "$OBJ/cuda_emu.o" &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
g++ -shared -pthread -o "$OUT" "$OBJ"/*.o
echo "built $OUT"
