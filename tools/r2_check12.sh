#!/bin/bash
# round 2, call 12: device-side output formatting (K8): format tests, CLI goldens with both formatters, CLI timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_format.py tests/test_cli_gpu.py tests/test_ref_binding_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2l_pytest.log
python tools/cli_timing.py > gpurun_out/r2l_cli_timing.txt 2>&1; echo "cli timing rc=$?"; tail -30 gpurun_out/r2l_cli_timing.txt
