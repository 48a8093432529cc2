#!/bin/bash
# pair-window A/B on the metric kernel, WPE corr A/B, new tests, bench N=1
mkdir -p gpurun_out/g
timeout 900 python -m pytest tests -x -q -m gpu -k "opt_in or tensor_core or stft_cov or fullsize or pipeline or golden" > gpurun_out/g/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g/pytest.log
for rep in 1 2 3; do
  SETK_WS_PAIRWIN=0 timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ws_tablewin >> gpurun_out/g/ab.jsonl 2>> gpurun_out/g/ab.err
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ws_pairwin >> gpurun_out/g/ab.jsonl 2>> gpurun_out/g/ab.err
done
for rep in 1 2; do
  for impl in dfma dmma; do
    SETK_WPE_CORR=$impl timeout 600 python tools/bench_configs.py "cfg4" 3 wpe 2>> gpurun_out/g/configs.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$impl', d['stages']['wpe'])" >> gpurun_out/g/wpe.txt
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/g/ws_prof python tools/ab_fused.py > gpurun_out/g/ncu.log 2>&1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/g/bench.json 2> gpurun_out/g/bench.err
tail -4 gpurun_out/g/pytest.log; cat gpurun_out/g/ab.jsonl | cut -c1-200; cat gpurun_out/g/wpe.txt; cut -c1-900 gpurun_out/g/bench.json
