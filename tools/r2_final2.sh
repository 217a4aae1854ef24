#!/bin/bash
# round 2, final tree: the bench lines of every workload on one GPU (C3 = the default run, with the CPU arm and K0), launch lists
# (C3, C4, rank 0 of 8) and ncu --set full of one rank of 8 (its kernels now run over one chunk).  The kernels themselves are the
# ones of the r2f captures (tools/r2_final_profile.sh): only the host orchestration changed since.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python bench.py > gpurun_out/r2g_c3_default.log 2>gpurun_out/r2g_c3_default.err; echo "default rc=$?"
: > gpurun_out/r2g_workloads.jsonl
for w in c1 c2 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-ingest 2>> gpurun_out/r2g_workloads.err | tail -1 >> gpurun_out/r2g_workloads.jsonl; echo "$w rc=$?"
done
python tools/bline.py gpurun_out/r2g_c3_default.log gpurun_out/r2g_workloads.jsonl
B2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches_c3.csv $B2 > gpurun_out/ncu_g1.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/r2g_launches_c3.csv > gpurun_out/r2g_launches_c3_summary.txt; grep -E "k_|TOTAL" gpurun_out/r2g_launches_c3_summary.txt | head -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_g2.log 2>&1; echo "ncu launches c4 rc=$?"
python tools/launch_summary.py gpurun_out/r2g_launches_c4.csv > gpurun_out/r2g_launches_c4_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches_0of8.csv $B2 --as-rank 0/8 > gpurun_out/ncu_g3.log 2>&1; echo "ncu launches 0of8 rc=$?"
python tools/launch_summary.py gpurun_out/r2g_launches_0of8.csv > gpurun_out/r2g_launches_0of8_summary.txt; grep -E "k_|TOTAL" gpurun_out/r2g_launches_0of8_summary.txt | head -12
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_own_list|k_build_sub|k_ent" -s 9 -c 12 -o /tmp/r2g_prof_0of8 -f $B1 --as-rank 0/8 > gpurun_out/ncu_g4.log 2>&1; echo "ncu full 0of8 rc=$?"
ncu -i /tmp/r2g_prof_0of8.ncu-rep --page raw --csv > gpurun_out/r2g_raw_0of8.csv 2>/dev/null
gzip -f gpurun_out/r2g_launches_c3.csv gpurun_out/r2g_launches_c4.csv gpurun_out/r2g_launches_0of8.csv
du -sh gpurun_out
