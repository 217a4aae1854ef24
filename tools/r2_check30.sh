#!/bin/bash
# round 2, call 30: real_gpu_prepare_scan (text first, reads second): its tests, then the end-to-end leg of C3 in both orders
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prepare.py -x -q --timeout 300 --timeout-method thread > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2ad_pytest.log
timeout 400 python bench.py --steps 4 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2ad_c3.log 2>&1; echo "rc=$?"
timeout 400 python bench.py --steps 3 --warmup 3 --no-ingest --no-cpu-baseline --e2e-order reads-first > gpurun_out/r2ad_c3_rf.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2ad_c*.log
for f in gpurun_out/r2ad_c3.log gpurun_out/r2ad_c3_rf.log; do grep -o '"e2e": {.\{0,900\}' $f | head -c 1000; echo; done
tail -3 gpurun_out/r2ad_c3.log | head -c 600
