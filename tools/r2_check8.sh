#!/bin/bash
# round 2, call 8: k_build_sub3 split per table; one chunk for the whole text (32-bit chunk positions) with entry cache policies
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_midscale.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2h_c3.log 2>&1; echo "rc=$?"
REAL_GPU_BUILD_SPLIT=0 timeout 300 $B > gpurun_out/r2h_c3_nosplit.log 2>&1; echo "rc=$?"
for D in 0 2 4; do
  REAL_GPU_CHUNK_MPOS=4000 REAL_GPU_DEBUG=$D timeout 300 $B > gpurun_out/r2h_c3_one_d$D.log 2>&1; echo "rc=$?"
done
REAL_GPU_CHUNK_MPOS=2048 REAL_GPU_DEBUG=4 timeout 300 $B > gpurun_out/r2h_c3_2g_d4.log 2>&1; echo "rc=$?"
REAL_GPU_CHUNK_MPOS=4000 REAL_GPU_DEBUG=4 REAL_GPU_L2_SLICE_MB=24 timeout 300 $B > gpurun_out/r2h_c3_one_d4_s24.log 2>&1; echo "rc=$?"
timeout 200 $B --as-rank 0/8 > gpurun_out/r2h_as0of8.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2h_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2h_*.log
