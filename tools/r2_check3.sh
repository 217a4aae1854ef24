#!/bin/bash
# round 2, call 3: parity of the long-segment sort / overflow / packed gaps, C3 after the verify fix, launch lists of N=1, 0/8, 0/2
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest3.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e"
timeout 300 $B > gpurun_out/r2c_c3.log 2>&1; echo "rc=$?"
timeout 300 $B --reads-format bytes > gpurun_out/r2c_c3_bytes.log 2>&1; echo "rc=$?"
python tools/bline.py gpurun_out/r2c_*.log
N="ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv"
B2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
$N --log-file gpurun_out/r2_launches_c3.csv $B2 > gpurun_out/ncu_l1.log 2>&1; echo "ncu rc=$?"
$N --log-file gpurun_out/r2_launches_0of8.csv $B2 --as-rank 0/8 > gpurun_out/ncu_l2.log 2>&1; echo "ncu rc=$?"
$N --log-file gpurun_out/r2_launches_0of2.csv $B2 --as-rank 0/2 > gpurun_out/ncu_l3.log 2>&1; echo "ncu rc=$?"
for f in c3 0of8 0of2; do echo "== $f"; python tools/launch_summary.py gpurun_out/r2_launches_$f.csv | head -16; done
ncu --set full --clock-control none --import-source on -k regex:"k_part_scatter|k_own_list" -s 3 -c 4 -o gpurun_out/r2_prof_own8 -f $B2 --as-rank 0/8 > gpurun_out/ncu_f1.log 2>&1; echo "ncu full rc=$?"
