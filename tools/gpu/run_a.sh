#!/bin/bash
# first hardware pass of the warp-specialised stft_cov kernel: parity subset, A/B timing, ncu
mkdir -p gpurun_out/a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "stft_cov or fullsize or bookkeeping" > gpurun_out/a/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a/pytest.log
for impl in classic ws classic ws; do
  SETK_SC_IMPL=$impl timeout 300 python tools/ab_fused.py setk_b200/libsetk_b200.so $impl >> gpurun_out/a/ab.jsonl 2>> gpurun_out/a/ab.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_cov_ws -s 2 -c 1 -o gpurun_out/a/ws_prof python tools/ab_fused.py > gpurun_out/a/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/a/ncu.log
tail -3 gpurun_out/a/pytest.log; cat gpurun_out/a/ab.jsonl; tail -2 gpurun_out/a/ncu.log
