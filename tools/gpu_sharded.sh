#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "sharded rc=$?"; tail -15 gpurun_out/pytest_sharded.log
