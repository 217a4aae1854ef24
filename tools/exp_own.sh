#!/bin/bash
# development experiment: one rank's share of a bucket-sharded C3 step (ranks of 2 and of 8) with differently tuned own-partition kernels
# (variants = real_b200/variants/v_<name>.so, built from csrc/real_gpu.cu with -DREAL_PF_THREADS=... -DREAL_PF_BATCH=...; the directory is scratch)
mkdir -p gpurun_out
run() { # name, rank spec, env...
  name=$1; rn=$2; shift; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --as-rank $rn > gpurun_out/own_${name}_${rn/\//of}.log 2>&1
  echo "$name $rn rc=$? $(tail -1 gpurun_out/own_${name}_${rn/\//of}.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],2), round(d["phases_ms"]["index_ms"],2), round(d["phases_ms"]["scan_ms"],2), d["counts"]["hits"])' 2>&1)"
}
for rn in 0/2 0/8; do
  run default $rn A=1
  for v in real_b200/variants/*.so; do run $(basename $v .so) $rn REAL_GPU_LIB=$PWD/$v; done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_asrank_0of2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --as-rank 0/2 > gpurun_out/ncu_asrank.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_part_scatter_own" -s 3 -c 1 -o gpurun_out/prof_own_n2 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --as-rank 0/2 > gpurun_out/ncu_own_n2.log 2>&1
echo "ncu rc=$?"
