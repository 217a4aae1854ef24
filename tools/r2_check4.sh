#!/bin/bash
# round 2, call 4: bucket-shard partition variants (dense with filter for 2 ranks, list for >= 4), mid-scale goldens
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_midscale.py tests/test_gpu_nccl.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest4.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
for RN in 0/8 0/4 0/2; do
  TAG=$(echo $RN | sed 's,/,of,')
  timeout 200 $B --as-rank $RN > gpurun_out/r2d_as${TAG}.log 2>&1; echo "rc=$?"
done
REAL_GPU_OWN_LIST_MAX=0 timeout 200 $B --as-rank 0/4 > gpurun_out/r2d_as0of4_dense.log 2>&1; echo "rc=$?"
REAL_GPU_OWN_LIST_MAX=0 timeout 200 $B --as-rank 0/8 > gpurun_out/r2d_as0of8_dense.log 2>&1; echo "rc=$?"
REAL_GPU_OWN_LIST_MAX=128 timeout 200 $B --as-rank 0/2 > gpurun_out/r2d_as0of2_list.log 2>&1; echo "rc=$?"
timeout 200 $B > gpurun_out/r2d_c3.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2d_*.log
