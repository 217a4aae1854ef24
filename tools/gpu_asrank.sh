#!/bin/bash
# launch list of ONE rank's share of an N-rank bucket-sharded job (usage: gpu_asrank.sh R/N)
mkdir -p gpurun_out
RN=${1:-0/8}; TAG=$(echo $RN | sed 's,/,of,')
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --as-rank $RN > gpurun_out/asrank_$TAG.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/asrank_$TAG.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["phases_ms"], d["counts"])' || tail -5 gpurun_out/asrank_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_asrank_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --as-rank $RN > gpurun_out/ncu_asrank.log 2>&1
echo "ncu rc=$?"
