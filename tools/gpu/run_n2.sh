#!/bin/bash
# two-GPU validation of the bench (async gather of every batch, NUMA binding, configs) and of the CLI under torchrun
mkdir -p gpurun_out/n2
nvidia-smi topo -m > gpurun_out/n2/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/n2/bench.json 2> gpurun_out/n2/bench.err
echo "bench rc=$?" >> gpurun_out/n2/bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/n2/bench_ref.json 2> gpurun_out/n2/bench_ref.err
tail -5 gpurun_out/n2/bench.err; cut -c1-300 gpurun_out/n2/bench.json; python - <<'PY'
import json
d=json.loads(open("gpurun_out/n2/bench.json").readline())
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}); print(d["gather"]); print(d["e2e"]["value"], d["run"]); print({k:(v["value"]) for k,v in d["configs"].items()})
PY
cut -c1-200 gpurun_out/n2/bench_ref.json
