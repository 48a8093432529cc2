#!/bin/bash
# weights pitch, cov_finalize per-matrix threads, apply without DC selects: parity subset + A/B; launch lists of the C > 4 routes
mkdir -p gpurun_out/o
timeout 900 python -m pytest tests -x -q -m gpu -k "stft_cov or weights or apply_istft or end_to_end or fullsize or peer_ring" > gpurun_out/o/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/o/pytest.log; tail -3 gpurun_out/o/pytest.log
for rep in 1 2 3; do
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so new >> gpurun_out/o/ab.jsonl 2>> gpurun_out/o/ab.err
done
for cfg in "8ch MVDR" "16ch MVDR" "8ch GEV"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:setk:: -c 300 --csv --log-file gpurun_out/o/launches_$tag.csv python tools/bench_configs.py "$cfg" 3 > gpurun_out/o/cfg_$tag.log 2>&1
done
python - <<'PY'
import json, sys
sys.path.insert(0, "tools")
import launch_summary
for l in open("gpurun_out/o/ab.jsonl"):
    d=json.loads(l); print(d["label"], d["stft_cov_ms"], d["weights_ms"], d["apply_istft_ms"], d["step_ms"], d["Rs_sum"], d["wave_sum"])
for t in ("8ch_MVDR","16ch_MVDR","8ch_GEV"):
    print("==", t); launch_summary.main(f"gpurun_out/o/launches_{t}.csv", only="")
PY
