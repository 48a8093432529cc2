#!/bin/bash
# third hardware pass: ws kernel with the tile descriptor table and early mask issue
mkdir -p gpurun_out/c
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "stft_cov or fullsize or bookkeeping or pipeline" > gpurun_out/c/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c/pytest.log
for rep in 1 2; do
  SETK_SC_IMPL=classic timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so classic >> gpurun_out/c/ab.jsonl 2>> gpurun_out/c/ab.err
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ws_v3 >> gpurun_out/c/ab.jsonl 2>> gpurun_out/c/ab.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/c/ws_prof python tools/ab_fused.py > gpurun_out/c/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/c/ncu.log
tail -3 gpurun_out/c/pytest.log; cat gpurun_out/c/ab.jsonl; tail -2 gpurun_out/c/ncu.log
