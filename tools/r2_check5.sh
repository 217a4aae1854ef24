#!/bin/bash
# round 2, call 5: fused index build
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest5.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2e_c3.log 2>&1; echo "rc=$?"
REAL_GPU_NO_FUSED_BUILD=1 timeout 300 $B > gpurun_out/r2e_c3_nofused.log 2>&1; echo "rc=$?"
for RN in 0/8 0/4 0/2; do
  TAG=$(echo $RN | sed 's,/,of,')
  timeout 200 $B --as-rank $RN > gpurun_out/r2e_as${TAG}.log 2>&1; echo "rc=$?"
done
timeout 300 $B --workload c5 > gpurun_out/r2e_c5.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2e_*.log
grep -o '"result_digest": "[0-9a-f]*"' gpurun_out/r2e_c3.log gpurun_out/r2e_c3_nofused.log
N="ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv"
$N --log-file gpurun_out/r2e_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_l5.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2e_launches_c3.csv 2>/dev/null | head -24
