#!/bin/bash
# two-GPU: the driver's command on the final build
mkdir -p gpurun_out/n2d
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/n2d/bench.json 2> gpurun_out/n2d/bench.err ) 2> gpurun_out/n2d/time.txt
echo "bench rc=$?" >> gpurun_out/n2d/bench.err
tail -2 gpurun_out/n2d/bench.err | cut -c1-300; cat gpurun_out/n2d/time.txt
python - <<'PY'
import json
lines=open("gpurun_out/n2d/bench.json").read().splitlines()
print("stdout lines:", len(lines))
d=json.loads(lines[0])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}); print(d["gather"]); print(d["e2e"]["value"], d["e2e"]["f32_mask_variant"]["value"])
print({k:(v["value"]) for k,v in d["configs"].items()})
PY
