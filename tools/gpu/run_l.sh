#!/bin/bash
# TMEM constants in both fused kernels (stft_cov_ws with uniform warp index + half max|x|, apply_istft), weights on thread groups with M aliased on A
mkdir -p gpurun_out/l
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "opt_in" > gpurun_out/l/pytest_optin.log 2>&1; tail -3 gpurun_out/l/pytest_optin.log
SETK_WS_CONST=tmem SETK_AI_CONST=tmem SETK_W_IMPL=coop timeout 1200 python -m pytest tests -x -q -m gpu -k "stft_cov or fullsize or end_to_end or weights or apply_istft or golden" > gpurun_out/l/pytest_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l/pytest_all.log; tail -3 gpurun_out/l/pytest_all.log
for rep in 1 2 3; do
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so base >> gpurun_out/l/ab.jsonl 2>> gpurun_out/l/ab.err
  SETK_WS_CONST=tmem timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so sc_tmem >> gpurun_out/l/ab.jsonl 2>> gpurun_out/l/ab.err
  SETK_AI_CONST=tmem timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so ai_tmem >> gpurun_out/l/ab.jsonl 2>> gpurun_out/l/ab.err
  SETK_WS_CONST=tmem SETK_AI_CONST=tmem SETK_W_IMPL=coop timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so all >> gpurun_out/l/ab.jsonl 2>> gpurun_out/l/ab.err
done
SETK_AI_CONST=tmem timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_istft_kernel -s 2 -c 1 -o gpurun_out/l/ai_tmem python tools/ab_fused.py > gpurun_out/l/ncu.log 2>&1
SETK_WS_CONST=tmem timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/l/sc_tmem python tools/ab_fused.py > gpurun_out/l/ncu2.log 2>&1
SETK_W_IMPL=coop timeout 600 ncu --set full --clock-control none --import-source on -k regex:weights_coop -s 2 -c 1 -o gpurun_out/l/w_coop python tools/ab_fused.py > gpurun_out/l/ncu_w.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/l/ab.jsonl"):
    d=json.loads(l); print(d["label"], d["stft_cov_ms"], d["weights_ms"], d["apply_istft_ms"], d["step_ms"], d["Rs_sum"], d["wave_sum"])
PY
tail -3 gpurun_out/l/ab.err
