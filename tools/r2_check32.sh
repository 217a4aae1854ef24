#!/bin/bash
# round 2, call 32: the partition of a text starts when the text is set (beside the index build); records re-used by the next scan of the same text
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2af_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-ingest --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2af_c3.log 2>&1; echo "rc=$?"
REAL_GPU_AUTO_PREPARE=0 timeout 400 $B --no-e2e > gpurun_out/r2af_c3_off.log 2>&1; echo "rc=$?"
timeout 400 $B --workload c2 > gpurun_out/r2af_c2.log 2>&1; echo "rc=$?"
timeout 400 $B --workload c4 > gpurun_out/r2af_c4.log 2>&1; echo "rc=$?"
timeout 400 $B --workload c5 > gpurun_out/r2af_c5.log 2>&1; echo "rc=$?"
timeout 400 $B --workload c1 > gpurun_out/r2af_c1.log 2>&1; echo "rc=$?"
timeout 400 $B --no-e2e --as-rank 0/8 > gpurun_out/r2af_0of8.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2af_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2af_*.log | sort | uniq -c
