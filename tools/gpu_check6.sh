#!/bin/bash
# bounded final check: parity tests (per-test limit), default bench without the CPU arm
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q --timeout 90 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["phases_ms"], "e2e", d["e2e"]["ms_per_step"], d["counts"])'
