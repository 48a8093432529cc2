#!/bin/bash
# cov_mma with a 4-step cp.async ring: parity of the C > 4 / 1024 routes, configs, ncu
mkdir -p gpurun_out/p
timeout 900 python -m pytest tests -x -q -m gpu -k "many_channel or nfft1024 or config or cgmm or wpd" > gpurun_out/p/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p/pytest.log; tail -3 gpurun_out/p/pytest.log
timeout 600 python tools/bench_configs.py "ch" 10 > gpurun_out/p/configs.jsonl 2> gpurun_out/p/configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:setk:: -c 100 --csv --log-file gpurun_out/p/launches_8ch.csv python tools/bench_configs.py "8ch MVDR" 3 > gpurun_out/p/cfg_8ch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cov_mma -s 2 -c 1 -o gpurun_out/p/cov_mma8 python tools/bench_configs.py "8ch MVDR" 3 > gpurun_out/p/ncu.log 2>&1
python - <<'PY'
import json, sys
sys.path.insert(0, "tools")
import launch_summary
for l in open("gpurun_out/p/configs.jsonl"):
    d=json.loads(l); print(d["config"], round(d["utts_per_s"]), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stages"].items()}, d["status_failures"])
launch_summary.main("gpurun_out/p/launches_8ch.csv", only="")
PY
