#!/bin/bash
# round 2, call 33: explicit real_gpu_prepare_scan keeps a preparation that is already on its way; GPU suite; C3 and C1 lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ag_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2ag_c3.log 2>&1; echo "rc=$?"
timeout 400 $B --workload c1 > gpurun_out/r2ag_c1.log 2>&1; echo "rc=$?"
timeout 400 $B --workload c4 > gpurun_out/r2ag_c4.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2ag_*.log
grep -o '"e2e": {.\{0,900\}' gpurun_out/r2ag_c3.log | head -c 1000; echo
