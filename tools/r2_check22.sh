#!/bin/bash
# round 2, call 22: FASTA reads parsed on the device (real_gpu_set_reads_fasta)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reads_ingest.py tests/test_gpu_ingest.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2w_pytest.log
