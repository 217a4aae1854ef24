#!/bin/bash
# round 2, call 10: C4 gate after the bench fix, C1 scan time traced
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-ingest"
timeout 400 $B --workload c4 > gpurun_out/r2j_c4.log 2>&1; echo "c4 rc=$?"
REAL_GPU_TRACE=1 timeout 200 $B --workload c1 --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/r2j_c1_trace.log 2>&1; echo "c1 rc=$?"
REAL_GPU_CHUNK_MPOS=1024 timeout 200 $B --workload c1 --no-e2e --no-cpu-baseline > gpurun_out/r2j_c1_chunk.log 2>&1; echo "c1 rc=$?"
timeout 200 $B --workload c1 --no-e2e --no-cpu-baseline --reads-format bytes > gpurun_out/r2j_c1_bytes.log 2>&1; echo "c1 rc=$?"
python tools/bline.py gpurun_out/r2j_c*.log
grep -o '"parity": {[^}]*}' gpurun_out/r2j_c4.log
grep trace gpurun_out/r2j_c1_trace.log | tail -12
