#!/bin/bash
# round 2, call 29: asynchronous text upload (partition starts on the words, the mask travels meanwhile), CLI with the direct rewritten path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ac_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-ingest > gpurun_out/r2ac_c3.log 2>&1; echo "rc=$?"
timeout 300 python bench.py --steps 4 --warmup 3 --no-ingest --no-cpu-baseline --workload c2 > gpurun_out/r2ac_c2.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2ac_c*.log
grep -o '"e2e": {.\{0,700\}' gpurun_out/r2ac_c3.log | head -c 900; echo
grep -o '"digest_ok": [a-z]*' gpurun_out/r2ac_c3.log
