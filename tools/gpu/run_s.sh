#!/bin/bash
# stft_cov_ws measurement variants: role order, mbarrier wait hint, table chunk
mkdir -p gpurun_out/s
for rep in 1 2 3; do
  for v in base covfirst hint1k hint0 chunk256; do
    lib=ab/libsetk_b200_$v.so; [ $v = base ] && lib=setk_b200/libsetk_b200.so
    timeout 200 python tools/ab_fused.py $lib $v >> gpurun_out/s/ab.jsonl 2>> gpurun_out/s/ab.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/s/ab.jsonl"):
    d=json.loads(l); print(d["label"], d["stft_cov_ms"], d["step_ms"], d["Rs_sum"])
PY
tail -2 gpurun_out/s/ab.err
