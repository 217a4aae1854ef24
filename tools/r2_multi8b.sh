#!/bin/bash
# round 2, 8 GPUs: C3 at N=8 and N=4 with the text-first end-to-end leg, C5 at N=8
mkdir -p gpurun_out
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 400 $T8 bench.py --gpus 8 --steps 6 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2w_n8_c3.log 2>&1; echo "n8 c3 rc=$?"
timeout 400 $T8 bench.py --gpus 8 --steps 5 --warmup 3 --no-ingest --no-cpu-baseline --workload c5 > gpurun_out/r2w_n8_c5.log 2>&1; echo "n8 c5 rc=$?"
timeout 400 $T4 bench.py --gpus 4 --steps 6 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2w_n4_c3.log 2>&1; echo "n4 c3 rc=$?"
python tools/bline.py gpurun_out/r2w_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2w_*.log | sort | uniq -c
for f in gpurun_out/r2w_n8_c3.log gpurun_out/r2w_n4_c3.log; do grep -o '"e2e": {.\{0,800\}' $f | head -c 900; echo; done
