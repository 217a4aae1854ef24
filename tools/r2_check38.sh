#!/bin/bash
# round 2, call 38: real_gpu_set_text_device_async (the mask by the time of the match call): the whole GPU suite with its new test
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2am_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2am_pytest.log
