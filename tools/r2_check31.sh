#!/bin/bash
# round 2, call 31: bucket shards sized for one chunk (so that real_gpu_prepare_scan serves them): tests, one rank of 2 and of 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prepare.py tests/test_gpu_sharded.py tests/test_gpu_fullsize.py -x -q --timeout 300 --timeout-method thread > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2ae_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-ingest --no-cpu-baseline --no-e2e"
timeout 400 $B --as-rank 0/8 > gpurun_out/r2ae_0of8.log 2>&1; echo "rc=$?"
timeout 400 $B --as-rank 0/2 > gpurun_out/r2ae_0of2.log 2>&1; echo "rc=$?"
timeout 400 $B --as-rank 0/4 > gpurun_out/r2ae_0of4.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2ae_*.log
