#!/bin/bash
# 2-GPU box: the default bench as the driver launches it (torchrun, buckets mode, with e2e), the reference arm under torchrun, the multi-rank tests
mkdir -p gpurun_out
bash tools/gpu_n.sh 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_reference_n2.log 2>&1; echo "reference arm rc=$?"; tail -1 gpurun_out/bench_reference_n2.log | cut -c1-500
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "sharded tests rc=$?"; tail -3 gpurun_out/pytest_sharded.log
