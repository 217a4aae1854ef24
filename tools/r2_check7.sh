#!/bin/bash
# round 2, call 7 (re-entry): the whole GPU suite on the current tree, the default bench line, launch list and
# ncu --set full of the top kernels (each ncu pass only after the plain command exited 0)
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_c3.log 2>&1; rc=$?; echo "bench rc=$rc"
python tools/bline.py gpurun_out/r2g_c3.log
B2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2g_launches_c3.csv $B2 > gpurun_out/ncu_l7.log 2>&1; echo "ncu launches rc=$?"
  python tools/launch_summary.py gpurun_out/r2g_launches_c3.csv | head -30
  ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_part_hist|k_build_sub|k_ent|k_read_seeds|k_pack" -s 12 -c 16 -o gpurun_out/r2g_prof_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_f7.log 2>&1; echo "ncu full rc=$?"
fi
for RN in 0/8 0/2; do
  TAG=$(echo $RN | sed 's,/,of,')
  timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ingest --no-e2e --as-rank $RN > gpurun_out/r2g_as${TAG}.log 2>&1; echo "rc=$?"
done
python tools/bline.py gpurun_out/r2g_as*.log
