#!/bin/bash
# round 2, call 11: scoring / gapped DP with batched loads, gate with the reference's undefined reads masked, memory query only when needed
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-ingest"
for W in c1 c2 c4; do
  timeout 400 $B --workload $W > gpurun_out/r2k_$W.log 2>&1; echo "$W rc=$?"
done
timeout 400 $B --no-cpu-baseline > gpurun_out/r2k_c3.log 2>&1; echo "c3 rc=$?"
python tools/bline.py gpurun_out/r2k_c*.log
grep -o '"parity": {[^}]*}' gpurun_out/r2k_c4.log
grep -o '"gap_[a-z]*_ms": [0-9.]*' gpurun_out/r2k_c4.log
