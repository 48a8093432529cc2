#!/bin/bash
# new defaults (TMEM constants, 4-thread weight solve), CM masks: full GPU tests, bench, launch list
mkdir -p gpurun_out/m
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/m/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m/pytest.log
timeout 900 python bench.py > gpurun_out/m/bench.json 2> gpurun_out/m/bench.err
echo "bench rc=$?" >> gpurun_out/m/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:setk:: -c 400 --csv --log-file gpurun_out/m/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs --unique 32 > gpurun_out/m/bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/m/smoke.log
tail -5 gpurun_out/m/pytest.log; tail -3 gpurun_out/m/bench.err; tail -2 gpurun_out/m/smoke.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/m/bench.json").readline())
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}); print(d["roofline"]); print(d["stages"])
e=d["e2e"]; print(e["value"], e["h2d_GBps"], e["f32_mask_variant"], e["f32_host_variant"]["value"])
print({k:(v["value"]) for k,v in d["configs"].items()})
PY
grep -c setk gpurun_out/m/launches.csv
