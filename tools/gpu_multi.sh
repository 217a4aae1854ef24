#!/bin/bash
# N-GPU bench of the multi-GPU forms; usage: gpu_multi.sh N [modes...]   (default modes: buckets tables text)
N=${1:-2}; shift
MODES=${@:-buckets tables text}
mkdir -p gpurun_out
for par in $MODES; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parallel $par > gpurun_out/bench_n${N}_$par.log 2>&1
  echo "$par rc=$?"
  tail -1 gpurun_out/bench_n${N}_$par.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["n_gpus"], d["ms_per_step"], d["phases_ms"], d["counts"])' || tail -20 gpurun_out/bench_n${N}_$par.log
done
