#!/bin/bash
# final verification of the round: full GPU tests, smoke, bench (both arms), launch list, ncu captures, CLI wall clock
mkdir -p gpurun_out/final
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/final/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final/smoke.log
timeout 900 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err
echo "bench rc=$?" >> gpurun_out/final/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:setk:: -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs --unique 32 > gpurun_out/final/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/final/sc python tools/ab_fused.py > gpurun_out/final/ncu_sc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:apply_istft_kernel -s 2 -c 1 -o gpurun_out/final/ai python tools/ab_fused.py > gpurun_out/final/ncu_ai.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cov_mma -s 2 -c 1 -o gpurun_out/final/cov_mma8 python tools/bench_configs.py "8ch MVDR" 3 > gpurun_out/final/ncu_cm.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:wpe_corr -s 1 -c 1 -o gpurun_out/final/wpe_corr python tools/bench_configs.py "6ch" 1 wpe > gpurun_out/final/ncu_wpe.log 2>&1
timeout 300 python tools/ab_fused.py >> gpurun_out/final/ab.jsonl 2>> gpurun_out/final/ab.err
timeout 600 python tools/cli_wallclock.py 512 256 > gpurun_out/final/cli.json 2> gpurun_out/final/cli.err
tail -4 gpurun_out/final/pytest.log; tail -2 gpurun_out/final/smoke.log; tail -2 gpurun_out/final/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final/bench.json").readline())
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}); print(d["roofline"]); print(d["stages"])
e=d["e2e"]; print(e["value"], e["h2d_GBps"], e["f32_mask_variant"]["value"], e["f32_host_variant"]["value"])
print({k:(v["value"]) for k,v in d["configs"].items()}); print(d["cpu_baseline"]); print(d["clocks"])
PY
cat gpurun_out/final/ab.jsonl; cat gpurun_out/final/cli.json; tail -2 gpurun_out/final/cli.err
