#!/bin/bash
# round 2, call 1: parity after the ReadSrc change, C3 with packed / byte reads, one rank's share of 2/4/8 with both histogram paths
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest1.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest"
timeout 300 $B > gpurun_out/r2_c3_packed.log 2>&1; echo "rc=$?"
timeout 300 $B --no-e2e --reads-format bytes > gpurun_out/r2_c3_bytes.log 2>&1; echo "rc=$?"
for RN in 0/8 0/4 0/2; do
  TAG=$(echo $RN | sed 's,/,of,')
  REAL_GPU_HIST_PICK=0 timeout 200 $B --no-e2e --as-rank $RN > gpurun_out/r2_as${TAG}_pick0.log 2>&1; echo "rc=$?"
  REAL_GPU_HIST_PICK=128 timeout 200 $B --no-e2e --as-rank $RN > gpurun_out/r2_as${TAG}_pick128.log 2>&1; echo "rc=$?"
done
python tools/bline.py gpurun_out/r2_c3_packed.log gpurun_out/r2_c3_bytes.log gpurun_out/r2_as*.log
