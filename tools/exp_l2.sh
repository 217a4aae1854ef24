#!/bin/bash
# development experiment: probe kernel time / DRAM bytes / L2 hit rate for different bucket counts,
# with (debug=0) and without (debug=1) following the set slot bits
for w in tiny c2; do for pb in 0 4 8; do for dbg in 0 1; do
REAL_GPU_PASS_BITS=$pb REAL_GPU_DEBUG=$dbg ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_bucket" -s 1 -c 1 --csv --log-file gpurun_out/exp_${w}_${pb}_${dbg}.csv python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
done; done; done
