#!/bin/bash
mkdir -p gpurun_out/r
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -m gpu > gpurun_out/r/pytest_cli.log 2>&1; tail -2 gpurun_out/r/pytest_cli.log
timeout 600 python tools/cli_wallclock.py 1024 256 > gpurun_out/r/cli.json 2> gpurun_out/r/cli.err
cat gpurun_out/r/cli.json; tail -2 gpurun_out/r/cli.err
