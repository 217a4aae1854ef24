#!/bin/bash
# round 2, call 19: partition CTAs of 512 threads (second library build); compact matchAll rows
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_format.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest.log
REAL_GPU_LIB=$PWD/real_b200/libreal_gpu_ps512.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2t_pytest512.log 2>&1; echo "pytest512 rc=$?"; tail -2 gpurun_out/r2t_pytest512.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2t_c3.log 2>&1; echo "rc=$?"
REAL_GPU_LIB=$PWD/real_b200/libreal_gpu_ps512.so timeout 300 $B > gpurun_out/r2t_c3_ps512.log 2>&1; echo "rc=$?"
REAL_GPU_LIB=$PWD/real_b200/libreal_gpu_ps512.so timeout 300 $B --as-rank 0/8 > gpurun_out/r2t_as0of8_ps512.log 2>&1; echo "rc=$?"
REAL_GPU_LIB=$PWD/real_b200/libreal_gpu_ps512.so timeout 300 $B --as-rank 0/2 > gpurun_out/r2t_as0of2_ps512.log 2>&1; echo "rc=$?"
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --workload c2 > gpurun_out/r2t_c2.log 2>&1; echo "rc=$?"
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --workload c5 > gpurun_out/r2t_c5.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2t_*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2t_c3*.log
grep -o '"probe_ms": [0-9.]*' gpurun_out/r2t_*.log | sort -u
