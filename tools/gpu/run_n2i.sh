#!/bin/bash
mkdir -p gpurun_out/n2i
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 3 --no-configs > gpurun_out/n2i/bench_a.json 2> gpurun_out/n2i/bench_a.err; echo "rc=$?"
SETK_BENCH_PROBE_FORCE_REJECT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 30 --warmup 3 --no-configs > gpurun_out/n2i/bench_b.json 2> gpurun_out/n2i/bench_b.err; echo "rc=$?"
for f in a b; do python -c "
import json
d=json.loads(open('gpurun_out/n2i/bench_$f.json').readline()); print('$f', round(d['value']), round(d['ms_per_step'],4), d['gather']['peer_ring_probe'], d['gather']['how'][:40], d['gather']['gather_GBps'])"; tail -2 gpurun_out/n2i/bench_$f.err | cut -c1-200; done
