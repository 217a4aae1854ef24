#!/bin/bash
# round 2, call 6: CLI team / stdin / packed reads, k_build_sub3 without claim atomics
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest6.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2f_c3.log 2>&1; echo "rc=$?"
timeout 200 $B --as-rank 0/8 > gpurun_out/r2f_as0of8.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2f_*.log
grep -o '"result_digest": "[0-9a-f]*"' gpurun_out/r2f_c3.log
N="ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv"
$N --log-file gpurun_out/r2f_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_l6.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r2f_launches_c3.csv 2>/dev/null | grep -E "k_build|k_ent|TOTAL"
