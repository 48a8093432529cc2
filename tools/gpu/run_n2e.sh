#!/bin/bash
mkdir -p gpurun_out/n2e
for g in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 --no-configs --gather $g > gpurun_out/n2e/bench_$g.json 2> gpurun_out/n2e/bench_$g.err
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/n2e/bench_n1.json 2> gpurun_out/n2e/bench_n1.err
python - <<'PY'
import json
for f in ("peer","nccl","n1"):
    try:
        d=json.loads(open(f"gpurun_out/n2e/bench_{f}.json").readline())
        print(f, d["value"], d["ms_per_step"], d["run"]["per_rank_ms_per_step"], d["gather"] and d["gather"]["gather_GBps"])
    except Exception as e: print(f, "ERR", e)
PY
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
