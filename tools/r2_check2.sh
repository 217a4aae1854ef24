#!/bin/bash
# round 2, call 2: fold tests, list-based partition of a bucket shard, probe occupancy clamp, parity gate + digest of the bench
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest2.log
timeout 300 python bench.py --workload tiny --steps 3 --warmup 3 --no-ingest --cpu-text 2000000 --cpu-reads 50000 > gpurun_out/r2b_tiny.log 2>&1; echo "tiny rc=$?"
timeout 400 python bench.py --steps 4 --warmup 3 --no-ingest > gpurun_out/r2b_c3.log 2>&1; echo "c3 rc=$?"
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
for RN in 0/8 3/8 0/4 0/2; do
  TAG=$(echo $RN | sed 's,/,of,')
  timeout 200 $B --as-rank $RN > gpurun_out/r2b_as${TAG}.log 2>&1; echo "rc=$?"
done
REAL_GPU_PROBE_OCC=2 timeout 200 $B > gpurun_out/r2b_c3_occ2.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2b_*.log
grep -o '"result_digest": "[0-9a-f]*"' gpurun_out/r2b_tiny.log gpurun_out/r2b_c3.log
grep -o '"parity": {[^}]*}' gpurun_out/r2b_tiny.log gpurun_out/r2b_c3.log
