#!/bin/bash
# second hardware pass: TMA masks + role order / wait-hint variants of the ws kernel
mkdir -p gpurun_out/b
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "stft_cov or fullsize or bookkeeping or pipeline" > gpurun_out/b/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b/pytest.log
for rep in 1 2; do
  SETK_SC_IMPL=classic timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so classic >> gpurun_out/b/ab.jsonl 2>> gpurun_out/b/ab.err
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ws_default >> gpurun_out/b/ab.jsonl 2>> gpurun_out/b/ab.err
  for v in nohint covfirst hint1k; do
    timeout 300 python tools/ab_fused.py ab/libsetk_b200_ws_$v.so ws_$v >> gpurun_out/b/ab.jsonl 2>> gpurun_out/b/ab.err
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/b/ws_prof python tools/ab_fused.py > gpurun_out/b/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/b/ncu.log
tail -3 gpurun_out/b/pytest.log; cat gpurun_out/b/ab.jsonl; tail -2 gpurun_out/b/ncu.log
