#!/bin/bash
# development experiment: C3 step (N=1, and one rank of 8 / of 2) with the default build and with real_b200/variants/*.so
mkdir -p gpurun_out
run() { # name, extra args, env...
  name=$1; extra=$2; shift; shift
  env "$@" python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-ingest $extra > gpurun_out/var_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/var_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); p=d["phases_ms"]; print(round(d["ms_per_step"],3), "index", round(p["index_ms"],2), "scan", round(p["scan_ms"],2), "part", round(p["part_ms"],2), "probe", round(p["probe_ms"],2), d["counts"]["hits"], d.get("digest_ok"), d["result_digest"])' 2>&1)"
}
run default "" A=1
run default_0of8 "--as-rank 0/8" REAL_GPU_AUTO_PREPARE=0
run default_0of2 "--as-rank 0/2" REAL_GPU_AUTO_PREPARE=0
for v in real_b200/variants/*.so; do
  n=$(basename $v .so)
  run $n "" REAL_GPU_LIB=$PWD/$v
  run ${n}_0of8 "--as-rank 0/8" REAL_GPU_LIB=$PWD/$v REAL_GPU_AUTO_PREPARE=0
  run ${n}_0of2 "--as-rank 0/2" REAL_GPU_LIB=$PWD/$v REAL_GPU_AUTO_PREPARE=0
done
