#!/usr/bin/env python
"""profiles/rNN_kernels_<workload>.json from an `ncu --set full` capture of one bench step.

usage: ncu -i x.ncu-rep --page raw --csv > raw.csv; python tools/kernel_table.py raw.csv c3 "<source note>" [scan_steps] > profiles/r02_kernels_c3.json
scan_steps = how many steps' worth of scan launches the capture holds (the capture window may span the warm-up step), default 1.

Per kernel of the step (summed over its launches): launches, time under ncu, DRAM bytes read + written, and -- from the
launch with the longest duration -- registers, achieved occupancy, issue-slot use, the busiest unit (DRAM / L1TEX / LTS /
SM) and the two largest stall reasons.  `scan_dram_bytes_per_step` = DRAM bytes of the scan kernels (k_part_hist,
k_own_list, k_part_scatter, k_bucket_probe): bench.py reports it as roofline.traffic."""
import collections
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
SCAN = ("k_part_hist", "k_own_list", "k_part_scatter", "k_bucket_probe")
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "mio_throttle", "lg_throttle", "math_pipe_throttle", "wait", "branch_resolving", "not_selected", "no_instruction"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=0.0):
    if name not in col or r[col[name]] in ("", "n/a"):
        return default
    return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)


agg = collections.OrderedDict()
for r in data:
    full = r[col["Kernel Name"]]
    name = full.split("(")[0].replace("void ", "").replace("realgpu::", "").strip()
    if not name.startswith("k_"):
        continue
    a = agg.setdefault(name, {"kernel": name, "launches": 0, "ms_under_ncu": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "_best": -1.0})
    ms = val(r, "gpu__time_duration.sum")
    a["launches"] += 1
    a["ms_under_ncu"] += ms
    a["dram_read_bytes"] += val(r, "dram__bytes_read.sum")
    a["dram_write_bytes"] += val(r, "dram__bytes_write.sum")
    if ms > a["_best"]:
        a["_best"] = ms
        units_pct = {"dram": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "l1tex": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                     "lts": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), "sm": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")}
        stalls = sorted(((val(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s), s) for s in STALLS), reverse=True)
        a.update(longest_launch_ms=ms, registers=int(val(r, "launch__registers_per_thread")), grid=int(val(r, "launch__grid_size")), block=int(val(r, "launch__block_size")),
                 warps_active_pct=round(val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
                 issue_active_pct=round(val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                 unit_pct_of_peak={k: round(v, 1) for k, v in units_pct.items()}, busiest_unit=max(units_pct, key=units_pct.get),
                 warp_instructions=val(r, "smsp__inst_executed.sum"), l2_hit_rate_pct=round(val(r, "lts__t_sector_hit_rate.pct"), 1),
                 top_stalls=[{"reason": s, "cycles_per_issue": round(v, 2)} for v, s in stalls[:2]])
kernels = []
for a in agg.values():
    a.pop("_best")
    a["dram_GBps_under_ncu"] = round((a["dram_read_bytes"] + a["dram_write_bytes"]) / max(a["ms_under_ncu"], 1e-9) / 1e6, 1)
    kernels.append(a)
kernels.sort(key=lambda a: -a["ms_under_ncu"])
scan_steps = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
scan = sum(a["dram_read_bytes"] + a["dram_write_bytes"] for a in kernels if any(a["kernel"].startswith(s) for s in SCAN)) / scan_steps
print(json.dumps({"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "workload": sys.argv[2], "n_gpus": 1,
                  "scan_dram_bytes_per_step": scan, "scan_steps_in_capture": scan_steps,
                  "note": "one step under `ncu --set full --clock-control none` (cold caches, serialised launches: shares, not absolute times); "
                          "scan_dram_bytes_per_step = dram__bytes_read.sum + dram__bytes_write.sum of the scan kernels",
                  "kernels": kernels}, indent=1))
