#!/bin/bash
# development experiment: the C3 step with differently tuned builds (real_b200/variants/*.so, built with -D flags)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/var_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/var_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["ms_per_step"], d["phases_ms"]["pack_ms"], d["phases_ms"]["index_ms"], d["phases_ms"]["scan_ms"], d["counts"]["hits"])' 2>&1)"
}
run default A=1
for v in real_b200/variants/*.so; do run $(basename $v .so) REAL_GPU_LIB=$PWD/$v; done
