#!/bin/bash
# Build libsetk_b200.so for sm_100a (B200) in-tree.  No GPU needed (nvcc cross-compiles).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/setk_b200/libsetk_b200.so"
OBJ="$HERE/obj"
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I$ROOT/include ${SETK_NVCC_EXTRA}"
pids=()
for f in api generic weights stft_cov_fused apply_istft_fused; do
  $NVCC $FLAGS -Xptxas -v -c "$HERE/$f.cu" -o "$OBJ/$f.o" > "$OBJ/$f.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
if [ $rc -ne 0 ]; then cat "$OBJ"/*.log | grep -v "^ptxas info" | head -100; exit 1; fi
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/*.o
echo "built $OUT"
