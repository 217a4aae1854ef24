#!/bin/bash
# round 2, call 34: bucket shards of 4+ ranks start their partition beside the index build: one rank of 8 and of 4
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline --no-e2e"
timeout 400 $B --as-rank 0/8 > gpurun_out/r2ah_0of8.log 2>&1; echo "rc=$?"
timeout 400 $B --as-rank 3/8 > gpurun_out/r2ah_3of8.log 2>&1; echo "rc=$?"
timeout 400 $B --as-rank 0/4 > gpurun_out/r2ah_0of4.log 2>&1; echo "rc=$?"
REAL_GPU_AUTO_PREPARE=0 timeout 400 $B --as-rank 0/4 > gpurun_out/r2ah_0of4_off.log 2>&1; echo "rc=$?"
REAL_GPU_AUTO_PREPARE=1 timeout 400 $B --as-rank 0/2 > gpurun_out/r2ah_0of2_on.log 2>&1; echo "rc=$?"
timeout 400 $B --as-rank 0/2 > gpurun_out/r2ah_0of2.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2ah_*.log
