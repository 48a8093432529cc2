#!/bin/bash
# apply_istft ws (dedicated IFFT warps) A/B, weights A/B, WPE DMMA A/B, parity subset
mkdir -p gpurun_out/f
timeout 900 python -m pytest tests -x -q -m gpu -k "apply_istft or fullsize or pipeline or cli or wpe or wpd or weights or non_power or golden" > gpurun_out/f/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f/pytest.log
for rep in 1 2; do
  SETK_AI_IMPL=classic timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ai_classic >> gpurun_out/f/ab.jsonl 2>> gpurun_out/f/ab.err
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ai_ws16 >> gpurun_out/f/ab.jsonl 2>> gpurun_out/f/ab.err
done
for impl in dmma dfma; do
  SETK_WPE_CORR=$impl timeout 600 python tools/bench_configs.py "cfg4" 3 wpe >> gpurun_out/f/wpe_$impl.jsonl 2>> gpurun_out/f/configs.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f/launches_wpe.csv python tools/bench_configs.py "cfg4" 1 wpe > gpurun_out/f/ncu_wpe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_istft_ws -s 2 -c 1 -o gpurun_out/f/aw_prof python tools/ab_fused.py > gpurun_out/f/ncu_aw.log 2>&1
tail -8 gpurun_out/f/pytest.log; cat gpurun_out/f/ab.jsonl; cat gpurun_out/f/wpe_dmma.jsonl gpurun_out/f/wpe_dfma.jsonl | cut -c1-700
