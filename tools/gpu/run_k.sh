#!/bin/bash
# stft_cov ws: constants from tensor memory (SETK_WS_CONST=tmem) and direct audio loads, weights C=4 on thread groups: parity + A/B + ncu
mkdir -p gpurun_out/k
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "opt_in" > gpurun_out/k/pytest_optin.log 2>&1; tail -3 gpurun_out/k/pytest_optin.log
SETK_WS_CONST=tmem timeout 900 python -m pytest tests -x -q -m gpu -k "stft_cov or fullsize or end_to_end" > gpurun_out/k/pytest_tmem.log 2>&1
echo "pytest rc=$?" >> gpurun_out/k/pytest_tmem.log; tail -3 gpurun_out/k/pytest_tmem.log
for rep in 1 2 3; do
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so base >> gpurun_out/k/ab.jsonl 2>> gpurun_out/k/ab.err
  SETK_WS_CONST=tmem timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so tmem >> gpurun_out/k/ab.jsonl 2>> gpurun_out/k/ab.err
  SETK_WS_AUDIO=direct timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so direct >> gpurun_out/k/ab.jsonl 2>> gpurun_out/k/ab.err
  SETK_WS_CONST=tmem SETK_WS_AUDIO=direct timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so tmem_direct >> gpurun_out/k/ab.jsonl 2>> gpurun_out/k/ab.err
  SETK_W_IMPL=coop timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so w_coop >> gpurun_out/k/ab.jsonl 2>> gpurun_out/k/ab.err
done
SETK_WS_CONST=tmem timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/k/sc_tmem python tools/ab_fused.py > gpurun_out/k/ncu.log 2>&1
SETK_W_IMPL=coop timeout 600 ncu --set full --clock-control none --import-source on -k regex:weights_coop -s 2 -c 1 -o gpurun_out/k/w_coop python tools/ab_fused.py > gpurun_out/k/ncu_w.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/k/ab.jsonl"):
    d=json.loads(l); print(d["label"], d["stft_cov_ms"], d["weights_ms"], d["apply_istft_ms"], d["step_ms"], d["Rs_sum"], d["wave_sum"])
PY
tail -3 gpurun_out/k/ab.err
