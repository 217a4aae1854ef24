#!/bin/bash
# round 2, call 17: sub-bucket build specialised for -l 32
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_midscale.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2r_c3.log 2>&1; echo "rc=$?"
REAL_GPU_BUILD_GENERAL=1 timeout 300 $B > gpurun_out/r2r_c3_general.log 2>&1; echo "rc=$?"
timeout 300 $B --as-rank 0/8 > gpurun_out/r2r_as0of8.log 2>&1; echo "rc=$?"
timeout 300 $B --workload c2 > gpurun_out/r2r_c2.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2r_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2r_c3*.log
