#!/bin/bash
# baseline of the current build on a fresh box: GPU tests, bench (ours + reference), launch list, ncu of the side kernels
mkdir -p gpurun_out/j
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/j/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j/pytest.log
timeout 600 python bench.py > gpurun_out/j/bench.json 2> gpurun_out/j/bench.err
echo "bench rc=$?" >> gpurun_out/j/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j/bench_ref.json 2> gpurun_out/j/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/j/bench_under_ncu.log 2>&1
timeout 300 python tools/ab_fused.py >> gpurun_out/j/ab.jsonl 2>> gpurun_out/j/ab.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:weights_kernel -s 2 -c 1 -o gpurun_out/j/weights_prof python tools/ab_fused.py > gpurun_out/j/ncu_w.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:apply_istft_kernel -s 2 -c 1 -o gpurun_out/j/ai_prof python tools/ab_fused.py > gpurun_out/j/ncu_ai.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/j/sc_prof python tools/ab_fused.py > gpurun_out/j/ncu_sc.log 2>&1
timeout 600 python tools/bench_configs.py > gpurun_out/j/configs.jsonl 2> gpurun_out/j/configs.err
tail -5 gpurun_out/j/pytest.log; tail -3 gpurun_out/j/bench.err; cut -c1-600 gpurun_out/j/bench.json; cat gpurun_out/j/ab.jsonl; cut -c1-300 gpurun_out/j/configs.jsonl
