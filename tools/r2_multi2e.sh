#!/bin/bash
# round 2, 2 GPUs, final tree: the multi-process GPU tests (NCCL rendezvous + CUDA IPC peer windows against one handle)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_prepare.py -m gpu -x -q --timeout 400 --timeout-method thread > gpurun_out/r2ak_pytest_m2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ak_pytest_m2.log
