#!/bin/bash
# one GPU-box call: gpu tests, default bench, ncu launch list, ncu --set full of the top kernels (each only after the plain run exited 0)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.log 2>&1; rc=$?; echo "bench rc=$rc"; tail -1 gpurun_out/bench_default.log | cut -c1-400
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_part_hist|k_build_sub|k_ent_scatter|k_ent2_scatter|k_pack_both" -s 10 -c 14 -o gpurun_out/prof_c3_r1e -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
