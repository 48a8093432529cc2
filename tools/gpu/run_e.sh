#!/bin/bash
# full GPU test tier (all failures), C > 4 covariance A/B (tensor cores vs CUDA cores), launch list
mkdir -p gpurun_out/e
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/e/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e/pytest.log
for impl in mma cuda; do
  SETK_COV_IMPL=$impl timeout 600 python tools/bench_configs.py "ch" 5 >> gpurun_out/e/configs_$impl.jsonl 2>> gpurun_out/e/configs.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e/launches_8ch.csv python tools/bench_configs.py "cfg5 8ch" 2 > gpurun_out/e/ncu8.log 2>&1
for rep in 1 2; do
  SETK_AI_IMPL=classic timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ai_classic >> gpurun_out/e/ab.jsonl 2>> gpurun_out/e/ab.err
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ai_ws_64_112 >> gpurun_out/e/ab.jsonl 2>> gpurun_out/e/ab.err
  timeout 300 python tools/ab_fused.py ab/libsetk_b200_aw7296.so ai_ws_72_96 >> gpurun_out/e/ab.jsonl 2>> gpurun_out/e/ab.err
done
for impl in dmma dfma; do
  SETK_WPE_CORR=$impl timeout 600 python tools/bench_configs.py "cfg4" 3 wpe >> gpurun_out/e/wpe_$impl.jsonl 2>> gpurun_out/e/configs.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e/launches_wpe.csv python tools/bench_configs.py "cfg4" 1 wpe > gpurun_out/e/ncu_wpe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_istft_ws -s 2 -c 1 -o gpurun_out/e/aw_prof python tools/ab_fused.py > gpurun_out/e/ncu_aw.log 2>&1
tail -25 gpurun_out/e/pytest.log; cat gpurun_out/e/ab.jsonl; cat gpurun_out/e/wpe_dmma.jsonl gpurun_out/e/wpe_dfma.jsonl | cut -c1-600
python - <<'PY'
import json
for impl in ("mma","cuda"):
    for l in open(f"gpurun_out/e/configs_{impl}.jsonl"):
        d=json.loads(l); print(impl, d["config"], round(d["ms_per_batch"],3), {k:round(v,3) for k,v in d["stages"].items() if isinstance(v,float)})
PY
