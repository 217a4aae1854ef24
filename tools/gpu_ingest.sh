#!/bin/bash
# one GPU-box call for K0 (text ingest): its tests, the CLI parity tests, the ingest benchmark, ncu --set full of the three kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py tests/test_cli_gpu.py -x -q --timeout 90 --timeout-method thread > gpurun_out/pytest_ingest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_ingest.log
timeout 900 python tools/ingest_bench.py > gpurun_out/ingest_bench.log 2>&1; rc=$?; echo "ingest bench rc=$rc"; tail -c 1800 gpurun_out/ingest_bench.log
if [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fa_" -s 3 -c 3 -o gpurun_out/prof_ingest_r1 -f python tools/ingest_bench.py --bases 1000000000 --reps 1 --no-cpu --no-e2e > gpurun_out/ncu_ingest.log 2>&1
  echo "ncu rc=$?"
fi
