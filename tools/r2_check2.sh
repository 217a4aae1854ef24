#!/bin/bash
# round 2, call 2: fold tests, list-based partition of a bucket shard, probe occupancy clamp
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest2.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2b_c3.log 2>&1; echo "rc=$?"
for RN in 0/8 3/8 0/4 0/2; do
  TAG=$(echo $RN | sed 's,/,of,')
  timeout 200 $B --as-rank $RN > gpurun_out/r2b_as${TAG}.log 2>&1; echo "rc=$?"
done
REAL_GPU_PROBE_OCC=2 timeout 200 $B > gpurun_out/r2b_c3_occ2.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2b_*.log
