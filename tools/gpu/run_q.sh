#!/bin/bash
# cov_mma with the truncating TF32 split and running pointers: parity + configs
mkdir -p gpurun_out/q
timeout 600 python -m pytest tests -x -q -m gpu -k "many_channel or nfft1024 or config_fixtures" > gpurun_out/q/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q/pytest.log; tail -3 gpurun_out/q/pytest.log
timeout 300 python tools/bench_configs.py "8ch" 10 > gpurun_out/q/configs.jsonl 2> gpurun_out/q/configs.err
timeout 300 python tools/bench_configs.py "16ch" 10 >> gpurun_out/q/configs.jsonl 2>> gpurun_out/q/configs.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:setk:: -c 60 --csv --log-file gpurun_out/q/launches_8ch.csv python tools/bench_configs.py "8ch MVDR" 3 > gpurun_out/q/cfg_8ch.log 2>&1
python - <<'PY'
import json, sys
sys.path.insert(0, "tools")
import launch_summary
for l in open("gpurun_out/q/configs.jsonl"):
    d=json.loads(l); print(d["config"], round(d["utts_per_s"]), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stages"].items()}, d["status_failures"])
launch_summary.main("gpurun_out/q/launches_8ch.csv", only="")
PY
