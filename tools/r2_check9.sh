#!/bin/bash
# round 2, call 9: automatic chunk size (one chunk for C3), all workloads, gapped pass under ncu
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-ingest"
timeout 400 $B > gpurun_out/r2i_c3.log 2>&1; echo "rc=$?"
for W in c1 c2 c4 c5; do
  timeout 400 $B --workload $W > gpurun_out/r2i_$W.log 2>&1; echo "$W rc=$?"
done
python tools/bline.py gpurun_out/r2i_c*.log
grep -o '"digest_ok": [a-z]*' gpurun_out/r2i_c*.log
ncu --set full --clock-control none --import-source on -k regex:"k_gap_dp|k_gap_replay|k_unique_replay|k_score_hits" -c 6 -o gpurun_out/r2i_prof_c4 -f python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_f9.log 2>&1; echo "ncu c4 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2i_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-ingest --no-e2e > gpurun_out/ncu_l9.log 2>&1; echo "ncu launches c4 rc=$?"
python tools/launch_summary.py gpurun_out/r2i_launches_c4.csv > gpurun_out/r2i_launches_c4_summary.txt; head -16 gpurun_out/r2i_launches_c4_summary.txt
