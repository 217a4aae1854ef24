#!/bin/bash
# round 2, 2 GPUs: the real multi-process path (NCCL rendezvous + peer-memory fold over CUDA IPC / NVLink), the CLI on two devices, bench at N=2
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_cli_gpu.py -m gpu -x -q --timeout 400 --timeout-method thread > gpurun_out/r2n_pytest_m2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2n_pytest_m2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus 2 --steps 5 --warmup 3 --no-ingest > gpurun_out/r2n_n2_peer.log 2>&1; echo "n2 peer rc=$?"
timeout 400 $T bench.py --gpus 2 --steps 5 --warmup 3 --no-ingest --no-e2e --fold nccl > gpurun_out/r2n_n2_nccl.log 2>&1; echo "n2 nccl rc=$?"
python tools/bline.py gpurun_out/r2n_*.log
grep -o '"result_digest": "[0-9a-f]*"' gpurun_out/r2n_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2n_*.log
grep -o '"fold_ms": [0-9.]*' gpurun_out/r2n_*.log
