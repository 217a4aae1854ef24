#!/bin/bash
# the bench line of every BASELINE.json configuration that fits one GPU (C3 is the default run; C4 is covered by tests/test_gpu_fullsize.py)
mkdir -p gpurun_out
: > gpurun_out/bench_workloads.log
for w in c1 c2 c5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench_workloads.err | tail -1 >> gpurun_out/bench_workloads.log; echo "$w rc=$?"
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_workloads.log"):
    d = json.loads(l)
    print(d["config"]["workload"][:40], "| ms/step", round(d["ms_per_step"], 2), "| reads/s %.3g" % d["value"], "| scan Gbp/s", round(d["scan_only"]["text_gbp_per_s_per_gpu"], 1),
          "| e2e reads/s %.3g" % d["e2e"]["value"], "| hits", d["counts"]["matchall_hits"] or d["counts"]["hits"], "| frac", round(d["roofline"]["frac"], 2), round(d["roofline"]["design_frac"], 2))
PY
