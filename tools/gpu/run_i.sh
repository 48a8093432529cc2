#!/bin/bash
# stft_cov ws: audio by TMA staging (default) vs direct global loads (SETK_WS_AUDIO=direct), parity + A/B + ncu
mkdir -p gpurun_out/i
timeout 400 python -m pytest tests/test_gpu_peer_ring.py -x -q -m gpu > gpurun_out/i/pytest_ring.log 2>&1; tail -15 gpurun_out/i/pytest_ring.log
SETK_WS_AUDIO=direct timeout 900 python -m pytest tests -x -q -m gpu -k "stft_cov or fullsize or end_to_end" > gpurun_out/i/pytest_direct.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i/pytest_direct.log
for rep in 1 2 3; do
  timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so sc_tma >> gpurun_out/i/ab.jsonl 2>> gpurun_out/i/ab.err
  SETK_WS_AUDIO=direct timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so sc_direct >> gpurun_out/i/ab.jsonl 2>> gpurun_out/i/ab.err
  SETK_WS_AUDIO=direct timeout 300 python tools/ab_fused.py ab/libsetk_b200_nopf.so sc_direct_nopf >> gpurun_out/i/ab.jsonl 2>> gpurun_out/i/ab.err
done
SETK_WS_AUDIO=direct timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/i/ws_direct python tools/ab_fused.py > gpurun_out/i/ncu.log 2>&1
tail -4 gpurun_out/i/pytest_direct.log; cut -c1-220 gpurun_out/i/ab.jsonl
