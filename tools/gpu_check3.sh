#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["phases_ms"], d["counts"])' || tail -5 gpurun_out/bench_quick.log
bash tools/gpu_asrank.sh ${1:-0/8}
