#!/bin/bash
# full GPU test tier + both bench arms at N=1
mkdir -p gpurun_out/d
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/d/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d/pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/d/bench_ref.json 2> gpurun_out/d/bench_ref.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/d/bench.json 2> gpurun_out/d/bench.err
echo "bench rc=$?" >> gpurun_out/d/bench.err
tail -15 gpurun_out/d/pytest.log; tail -5 gpurun_out/d/bench.err; cut -c1-1500 gpurun_out/d/bench.json; cut -c1-600 gpurun_out/d/bench_ref.json
