#!/bin/bash
# round 2, call 23: the CLI with the device reads loader; whole GPU suite; CLI timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2x_pytest.log
python tools/cli_timing.py > gpurun_out/r2x_cli_timing.txt 2>&1; echo "cli timing rc=$?"; tail -32 gpurun_out/r2x_cli_timing.txt
