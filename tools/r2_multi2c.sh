#!/bin/bash
# round 2, 2 GPUs: end-to-end leg with the text first (records formed while the reads are uploaded and gathered), both orders
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus 2 --steps 4 --warmup 3 --no-ingest --no-cpu-baseline > gpurun_out/r2n_n2_tf.log 2>&1; echo "n2 text-first rc=$?"
timeout 400 $T bench.py --gpus 2 --steps 3 --warmup 3 --no-ingest --no-cpu-baseline --e2e-order reads-first > gpurun_out/r2n_n2_rf.log 2>&1; echo "n2 reads-first rc=$?"
python tools/bline.py gpurun_out/r2n_*.log
for f in gpurun_out/r2n_n2_tf.log gpurun_out/r2n_n2_rf.log; do grep -o '"e2e": {.\{0,900\}' $f | head -c 1000; echo; done
grep -o '"digest_ok": [a-z]*' gpurun_out/r2n_*.log
tail -5 gpurun_out/r2n_n2_tf.log | cut -c1-300
