#!/bin/bash
# round 2, call 16: partition kernel without the staged bucket byte; why C5's scan got slower (read layout, chunking)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_midscale.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2q_c3.log 2>&1; echo "rc=$?"
timeout 300 $B --as-rank 0/8 > gpurun_out/r2q_as0of8.log 2>&1; echo "rc=$?"
timeout 300 $B --workload c5 > gpurun_out/r2q_c5.log 2>&1; echo "rc=$?"
timeout 300 $B --workload c5 --reads-format bytes > gpurun_out/r2q_c5_bytes.log 2>&1; echo "rc=$?"
REAL_GPU_CHUNK_MPOS=1024 timeout 300 $B --workload c5 > gpurun_out/r2q_c5_chunk.log 2>&1; echo "rc=$?"
REAL_GPU_CHUNK_MPOS=1024 timeout 300 $B --workload c5 --reads-format bytes > gpurun_out/r2q_c5_chunk_bytes.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2q_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2q_c3.log
grep -o '"probe_ms": [0-9.]*' gpurun_out/r2q_c*.log | head
