#!/bin/bash
# parity tests + default bench (no CPU arm) + one rank's share of 2 and of 8
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["phases_ms"], "e2e", d["e2e"]["ms_per_step"])'
for rn in 0/2 0/8; do
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --as-rank $rn > gpurun_out/asrank_${rn/\//of}.log 2>&1
  echo "$rn rc=$? $(tail -1 gpurun_out/asrank_${rn/\//of}.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],2), round(d["phases_ms"]["index_ms"],2), round(d["phases_ms"]["scan_ms"],2), d["counts"]["hits"])' 2>&1)"
done
