#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e1.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_e2e1.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["phases_ms"], "e2e", d["e2e"])' || tail -5 gpurun_out/bench_e2e1.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e$N.log 2>&1; echo "bench$N rc=$?"
tail -1 gpurun_out/bench_e2e$N.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["phases_ms"], "e2e", d["e2e"])' || tail -25 gpurun_out/bench_e2e$N.log
