#!/bin/bash
# N-GPU default bench (buckets mode, with e2e); usage: gpu_n.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}.log 2>&1; echo "bench$N rc=$?"
tail -1 gpurun_out/bench_n${N}.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["n_gpus"], d["ms_per_step"], d["phases_ms"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])' || tail -25 gpurun_out/bench_n${N}.log
