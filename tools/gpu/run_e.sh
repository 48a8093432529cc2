#!/bin/bash
# full GPU test tier (all failures), C > 4 covariance A/B (tensor cores vs CUDA cores), launch list
mkdir -p gpurun_out/e
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/e/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e/pytest.log
for impl in mma cuda; do
  SETK_COV_IMPL=$impl timeout 600 python tools/bench_configs.py "ch" 5 >> gpurun_out/e/configs_$impl.jsonl 2>> gpurun_out/e/configs.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e/launches_8ch.csv python tools/bench_configs.py "cfg5 8ch" 2 > gpurun_out/e/ncu8.log 2>&1
tail -25 gpurun_out/e/pytest.log
python - <<'PY'
import json
for impl in ("mma","cuda"):
    for l in open(f"gpurun_out/e/configs_{impl}.jsonl"):
        d=json.loads(l); print(impl, d["config"], round(d["ms_per_batch"],3), {k:round(v,3) for k,v in d["stages"].items() if isinstance(v,float)})
PY
