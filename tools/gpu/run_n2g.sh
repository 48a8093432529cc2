#!/bin/bash
mkdir -p gpurun_out/n2g
for i in 1 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 100 --warmup 5 --no-configs > gpurun_out/n2g/bench_$i.json 2> gpurun_out/n2g/bench_$i.err
python -c "
import json
d=json.loads(open('gpurun_out/n2g/bench_$i.json').readline()); print($i, round(d['value']), d['ms_per_step'], d['run']['per_rank_ms_per_step'])"
done
