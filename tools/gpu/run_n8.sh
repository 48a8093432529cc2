#!/bin/bash
# eight-GPU validation of the driver's scaling run (default bench, reference arm)
mkdir -p gpurun_out/n8
nvidia-smi topo -m > gpurun_out/n8/topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/n8/bench.json 2> gpurun_out/n8/bench.err ) 2> gpurun_out/n8/time.txt
echo "bench rc=$?" >> gpurun_out/n8/bench.err
tail -4 gpurun_out/n8/bench.err | cut -c1-400; cat gpurun_out/n8/time.txt
python - <<'PY'
import json
lines=open("gpurun_out/n8/bench.json").read().splitlines()
print("stdout lines:", len(lines))
d=json.loads(lines[0])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}); print(d["gather"]); print(d["e2e"]["value"], d["e2e"]["h2d_GBps"], d["e2e"]["f32_mask_variant"]["value"], d["run"])
print({k:(v["value"]) for k,v in d["configs"].items()})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/n8/bench_ref.json 2> gpurun_out/n8/bench_ref.err
cut -c1-250 gpurun_out/n8/bench_ref.json
