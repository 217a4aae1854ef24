#!/bin/bash
# round check + smoke + K0 benchmark + CLI timing in one GPU-box call
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
bash tools/gpu_round_check.sh
timeout 600 python tools/ingest_bench.py > gpurun_out/ingest_bench.log 2>&1; echo "ingest bench rc=$?"; tail -c 900 gpurun_out/ingest_bench.log
timeout 900 python tools/cli_timing.py 50000000 3000000 > gpurun_out/cli_timing.log 2>&1; echo "cli timing rc=$?"; tail -22 gpurun_out/cli_timing.log
