#!/bin/bash
# round 2, final evidence: launch list + ncu --set full of the final tree (C3 on one B200), the gapped workload under ncu,
# one rank's share of 2 and of 8.  The .ncu-rep files stay on the box: their CSV pages travel.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
B2="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_c3.csv $B2 > gpurun_out/ncu_lf.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/r2f_launches_c3.csv > gpurun_out/r2f_launches_c3_summary.txt; grep -E "k_|TOTAL" gpurun_out/r2f_launches_c3_summary.txt | head -14
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_part_hist|k_build_sub|k_ent|k_seeds_packed" -s 10 -c 10 -o /tmp/r2f_prof_c3 -f $B1 > gpurun_out/ncu_ff.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2f_prof_c3.ncu-rep --page raw --csv > gpurun_out/r2f_raw_c3.csv 2>/dev/null
for k in k_bucket_probe k_part_scatter k_build_sub3; do ncu -i /tmp/r2f_prof_c3.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:$k --launch-count 1 2>/dev/null | gzip > gpurun_out/r2f_src_$k.csv.gz; done
ncu --set full --clock-control none --import-source on -k regex:"k_gap_dp|k_gap_replay|k_unique_replay|k_score_hits|k_fmt" -c 8 -o /tmp/r2f_prof_c4 -f python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_fc4.log 2>&1; echo "ncu c4 rc=$?"
ncu -i /tmp/r2f_prof_c4.ncu-rep --page raw --csv > gpurun_out/r2f_raw_c4.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_lc4.log 2>&1; echo "ncu launches c4 rc=$?"
python tools/launch_summary.py gpurun_out/r2f_launches_c4.csv > gpurun_out/r2f_launches_c4_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_0of8.csv $B2 --as-rank 0/8 > gpurun_out/ncu_l8.log 2>&1; echo "ncu launches 0of8 rc=$?"
python tools/launch_summary.py gpurun_out/r2f_launches_0of8.csv > gpurun_out/r2f_launches_0of8_summary.txt
gzip -f gpurun_out/r2f_launches_c3.csv gpurun_out/r2f_launches_c4.csv gpurun_out/r2f_launches_0of8.csv
du -sh gpurun_out; ls -la gpurun_out | tail -16
