#!/usr/bin/env python
"""profiles/r01_traffic_<workload>.json from an `ncu --set full` capture of one bench step.

usage: ncu -i x.ncu-rep --page raw --csv > raw.csv; python tools/traffic_json.py raw.csv c3 "<source note>" > profiles/r01_traffic_c3.json
Sums dram__bytes_read.sum + dram__bytes_write.sum over the scan launches (k_part_hist, k_part_scatter, k_bucket_probe) of the capture;
bench.py reads `scan_dram_bytes_per_step` as roofline.traffic.
"""
import csv
import json
import sys

SCAN = ("k_part_hist", "k_part_scatter", "k_bucket_probe")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    return float(r[col[name]]) * UNIT[units[col[name]]]


launches = []
for r in data:
    name = r[col["Kernel Name"]]
    short = next((k for k in SCAN if k in name), None)
    if short is None:
        continue
    launches.append({"kernel": short, "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                     "ms_under_ncu": val(r, "gpu__time_duration.sum")})
total = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches)
print(json.dumps({"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "workload": sys.argv[2], "n_gpus": 1,
                  "scan_dram_bytes_per_step": total,
                  "note": "dram__bytes_read.sum + dram__bytes_write.sum of the %d scan launches of one step (k_part_hist, k_part_scatter, k_bucket_probe per chunk)" % len(launches),
                  "launches": launches}, indent=1))
