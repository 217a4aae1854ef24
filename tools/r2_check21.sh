#!/bin/bash
# round 2, call 21: partition kernels with fewer block barriers (warp scan + one barrier, counters cleared under the copy-out barrier)
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2v_c3.log 2>&1; echo "rc=$?"
timeout 300 $B --as-rank 0/8 > gpurun_out/r2v_as0of8.log 2>&1; echo "rc=$?"
timeout 300 $B --as-rank 0/2 > gpurun_out/r2v_as0of2.log 2>&1; echo "rc=$?"
timeout 300 $B --workload c2 > gpurun_out/r2v_c2.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2v_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2v_c3.log
grep -o '"probe_ms": [0-9.]*' gpurun_out/r2v_*.log | sort -u
