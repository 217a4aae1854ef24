#!/bin/bash
# CLI parity tests + CLI timing on one GPU box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli_gpu.py tests/test_gpu_ingest.py -x -q > gpurun_out/pytest_cli.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_cli.log
timeout 900 python tools/cli_timing.py 50000000 3000000 > gpurun_out/cli_timing.log 2>&1; echo "cli timing rc=$?"; tail -32 gpurun_out/cli_timing.log
