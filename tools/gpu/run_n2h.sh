#!/bin/bash
mkdir -p gpurun_out/n2h
i=0
for mode in inside after inside after inside inside; do
i=$((i+1))
SETK_BENCH_PUSH=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 2 --steps 100 --warmup 5 --no-configs > gpurun_out/n2h/bench_$i.json 2> gpurun_out/n2h/bench_$i.err
python -c "
import json
d=json.loads(open('gpurun_out/n2h/bench_$i.json').readline()); print('$mode', round(d['value']), round(d['ms_per_step'],4), [(r['device'], r['host_enqueue']) for r in d['run']['per_rank_ms_per_step']])"
done
