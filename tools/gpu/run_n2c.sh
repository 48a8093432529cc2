#!/bin/bash
# two-GPU: symmetric-memory peer ring
mkdir -p gpurun_out/n2c
timeout 600 python -m pytest tests/test_gpu_peer_ring.py -x -q -m gpu > gpurun_out/n2c/pytest_ring.log 2>&1; tail -3 gpurun_out/n2c/pytest_ring.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 3 --no-configs > gpurun_out/n2c/bench.json 2> gpurun_out/n2c/bench.err
echo "bench rc=$?" >> gpurun_out/n2c/bench.err
tail -4 gpurun_out/n2c/bench.err | cut -c1-400
python - <<'PY'
import json
lines=open("gpurun_out/n2c/bench.json").read().splitlines()
print("stdout lines:", len(lines))
d=json.loads(lines[0])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}); print(d["gather"]); print(d["e2e"]["value"], d["e2e"]["f32_mask_variant"]["value"])
PY
