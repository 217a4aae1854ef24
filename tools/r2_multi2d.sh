#!/bin/bash
# round 2, 2 GPUs, final tree: C3 at N=2 (bucket shards, peer fold, text-first end-to-end leg)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus 2 --steps 6 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2w_n2_c3.log 2>&1; echo "n2 rc=$?"
python tools/bline.py gpurun_out/r2w_n2_c3.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2w_n2_c3.log | sort | uniq -c
