#!/bin/bash
# round 2, call 27: staged host->device copies of pageable file bytes, parallel writes of the output; CLI tests; full-size CLI
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_gpu_ingest.py tests/test_gpu_reads_ingest.py tests/test_ref_binding_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2aa_pytest.log
timeout 1500 python tools/cli_fullsize.py > gpurun_out/r2_cli_fullsize.txt 2>&1; echo "fullsize rc=$?"; tail -32 gpurun_out/r2_cli_fullsize.txt
