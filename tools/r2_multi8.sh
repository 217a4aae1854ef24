#!/bin/bash
# round 2, 8 GPUs: C3 at N=8 and N=4 (bucket shards, peer-memory fold), C5 at N=8 (BASELINE config 5 is an 8-GPU config)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 500 $T8 bench.py --gpus 8 --steps 10 --warmup 3 --no-ingest > gpurun_out/r2u_n8_c3.log 2>&1; echo "n8 c3 rc=$?"
timeout 500 $T8 bench.py --gpus 8 --steps 5 --warmup 3 --no-ingest --workload c5 > gpurun_out/r2u_n8_c5.log 2>&1; echo "n8 c5 rc=$?"
timeout 500 $T4 bench.py --gpus 4 --steps 10 --warmup 3 --no-ingest > gpurun_out/r2u_n4_c3.log 2>&1; echo "n4 c3 rc=$?"
python tools/bline.py gpurun_out/r2u_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2u_*.log
grep -o '"fold_ms": [0-9.]*' gpurun_out/r2u_*.log | sort -u
