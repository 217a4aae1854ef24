#!/usr/bin/env python
"""Prints the essentials of the bench lines found in the given log files (one row per file)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        lines = [l for l in open(path) if l.startswith("{")]
        d = json.loads(lines[-1])
        ph = d.get("phases_ms", {})
        e2e = d.get("e2e") or {}
        print("%-34s ms/step %8.3f | pack %.2f index %.2f scan %.2f (part %.2f probe %.2f) post %.2f exch %.2f | e2e %s | parity %s" % (
            path.split("/")[-1], d["ms_per_step"], ph.get("pack_ms", 0), ph.get("index_ms", 0), ph.get("scan_ms", 0), ph.get("part_ms", 0), ph.get("probe_ms", 0), ph.get("post_ms", 0),
            ph.get("exchange_ms", 0), ("%.2f" % e2e["ms_per_step"]) if e2e.get("ms_per_step") else "-", d.get("parity_checked")))
    except Exception as e:
        print("%-34s unreadable: %s" % (path.split("/")[-1], e))
