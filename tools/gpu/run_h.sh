#!/bin/bash
# apply_istft ws (2-frame tiles, 4-slot ring) A/B + parity, ncu
mkdir -p gpurun_out/h
timeout 900 python -m pytest tests -x -q -m gpu -k "opt_in or apply_istft or fullsize" > gpurun_out/h/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h/pytest.log
for rep in 1 2 3; do
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ai_classic >> gpurun_out/h/ab.jsonl 2>> gpurun_out/h/ab.err
  SETK_AI_IMPL=ws timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ai_ws_ring4 >> gpurun_out/h/ab.jsonl 2>> gpurun_out/h/ab.err
done
SETK_AI_IMPL=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_istft_ws -s 2 -c 1 -o gpurun_out/h/aw_prof python tools/ab_fused.py > gpurun_out/h/ncu_aw.log 2>&1
tail -4 gpurun_out/h/pytest.log; cut -c1-200 gpurun_out/h/ab.jsonl
