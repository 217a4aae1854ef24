#!/bin/bash
# round 2, call 15: ncu --set full of the current tree (N=1 and one rank of 8); the reports are turned into CSV pages on the box
# (the .ncu-rep files are too large to travel) 
mkdir -p gpurun_out
cd gpurun_out && rm -f *.ncu-rep && cd ..
B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_part_hist|k_build_sub|k_ent|k_seeds_packed" -s 8 -c 12 -o /tmp/r2p_prof_c3 -f $B1 > gpurun_out/ncu_f15a.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2p_prof_c3.ncu-rep --page raw --csv > gpurun_out/r2p_raw_c3.csv 2>/dev/null
ncu -i /tmp/r2p_prof_c3.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_bucket_probe --launch-count 1 2>/dev/null | gzip > gpurun_out/r2p_src_probe.csv.gz
ncu -i /tmp/r2p_prof_c3.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_part_scatter --launch-count 1 2>/dev/null | gzip > gpurun_out/r2p_src_scatter.csv.gz
ncu -i /tmp/r2p_prof_c3.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_build_sub3 --launch-count 1 2>/dev/null | gzip > gpurun_out/r2p_src_build.csv.gz
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_probe|k_part_scatter|k_own_list|k_build_sub|k_ent" -s 9 -c 12 -o /tmp/r2p_prof_0of8 -f $B1 --as-rank 0/8 > gpurun_out/ncu_f15b.log 2>&1; echo "ncu full 0of8 rc=$?"
ncu -i /tmp/r2p_prof_0of8.ncu-rep --page raw --csv > gpurun_out/r2p_raw_0of8.csv 2>/dev/null
ncu -i /tmp/r2p_prof_0of8.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_own_list --launch-count 1 2>/dev/null | gzip > gpurun_out/r2p_src_ownlist.csv.gz
ncu -i /tmp/r2p_prof_0of8.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_part_scatter --launch-count 1 2>/dev/null | gzip > gpurun_out/r2p_src_scatter_list.csv.gz
ls -la gpurun_out | tail -12
du -sh gpurun_out
