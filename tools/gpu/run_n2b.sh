#!/bin/bash
# two-GPU validation: peer-ring gather + NCCL fallback, reference arm under torchrun, peer ring test, CLI under torchrun
mkdir -p gpurun_out/n2
nvidia-smi topo -m > gpurun_out/n2/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_peer_ring.py -x -q -m gpu > gpurun_out/n2/pytest_ring.log 2>&1; tail -3 gpurun_out/n2/pytest_ring.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/n2/bench.json 2> gpurun_out/n2/bench.err
echo "bench rc=$?" >> gpurun_out/n2/bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 40 --warmup 3 --gather nccl --no-configs > gpurun_out/n2/bench_nccl.json 2> gpurun_out/n2/bench_nccl.err
echo "bench nccl rc=$?" >> gpurun_out/n2/bench_nccl.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2/bench_ref.json 2> gpurun_out/n2/bench_ref.err
tail -5 gpurun_out/n2/bench.err; tail -3 gpurun_out/n2/bench_nccl.err
python - <<'PY'
import json
for f in ("bench","bench_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/n2/{f}.json").readline())
        print(f,{k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}); print(d["gather"]); print(d["e2e"]["value"], d["e2e"]["f32_mask_variant"]["value"], d["run"])
        if d.get("configs"): print({k:(v["value"]) for k,v in d["configs"].items()})
    except Exception as e: print(f,"ERR",e)
PY
cut -c1-200 gpurun_out/n2/bench_ref.json
