#!/bin/bash
# round 2, call 14: CLI test of the kept rewritten file, launch lists + ncu --set full of the current tree (N=1 and one rank of 8)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread -k "rewritten or stdin" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
B2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/r2o_c3.log 2>&1; echo "c3 rc=$?"
python tools/bline.py gpurun_out/r2o_c3.log
N="ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv"
$N --log-file gpurun_out/r2o_launches_c3.csv $B2 > gpurun_out/ncu_l14a.log 2>&1; echo "ncu rc=$?"
$N --log-file gpurun_out/r2o_launches_0of8.csv $B2 --as-rank 0/8 > gpurun_out/ncu_l14b.log 2>&1; echo "ncu rc=$?"
for f in c3 0of8; do echo "== $f"; python tools/launch_summary.py gpurun_out/r2o_launches_$f.csv > gpurun_out/r2o_launches_${f}_summary.txt; grep -E "k_|TOTAL" gpurun_out/r2o_launches_${f}_summary.txt | head -20; done
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_part_hist|k_build_sub|k_ent|k_seeds_packed" -s 8 -c 12 -o gpurun_out/r2o_prof_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_f14a.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_own_list|k_build_sub|k_ent" -s 9 -c 12 -o gpurun_out/r2o_prof_0of8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e --as-rank 0/8 > gpurun_out/ncu_f14b.log 2>&1; echo "ncu full 0of8 rc=$?"
