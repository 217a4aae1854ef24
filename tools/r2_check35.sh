#!/bin/bash
# round 2, call 35: one rank of 2 with the kept-position list instead of the bucket filter
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline --no-e2e"
REAL_GPU_OWN_LIST_MAX=128 timeout 400 $B --as-rank 0/2 > gpurun_out/r2ai_0of2_list.log 2>&1; echo "rc=$?"
timeout 400 $B --as-rank 0/2 > gpurun_out/r2ai_0of2.log 2>&1; echo "rc=$?"
REAL_GPU_OWN_LIST_MAX=128 REAL_GPU_AUTO_PREPARE=0 timeout 400 $B --as-rank 0/2 > gpurun_out/r2ai_0of2_list_off.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2ai_*.log
