#!/bin/bash
# one rank's share of a bucket-sharded C3 step, ranks of 2 and of 8 (bounded)
mkdir -p gpurun_out
for rn in 0/2 0/8; do
  timeout 70 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --as-rank $rn > gpurun_out/asrank_${rn/\//of}.log 2>&1
  echo "$rn rc=$? $(tail -1 gpurun_out/asrank_${rn/\//of}.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],2), round(d["phases_ms"]["pack_ms"],2), round(d["phases_ms"]["index_ms"],2), round(d["phases_ms"]["scan_ms"],2), d["counts"]["hits"])' 2>&1)"
done
