#!/usr/bin/env python
"""Top source lines by stall samples.
   ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> > src.csv ; ncu_hot_lines.py src.csv [n]
Source rows (first column = line number) carry the per-line aggregates; SASS rows (empty first column) are skipped."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = ""
hdr = None
out = []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        si, ie, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        sl = hdr.index("stall_long_sb"); sb = hdr.index("stall_barrier"); ss = hdr.index("stall_short_sb")
        continue
    if hdr and len(r) > ti and r[0].strip().isdigit():
        try:
            out.append((int(r[si]), int(r[ie]), int(r[ti]), fname, int(r[0]), r[1].strip(), int(r[sl] or 0), int(r[sb] or 0), int(r[ss] or 0)))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print("samples %d, warp instructions %d" % (tot, toti))
print("%7s %7s %5s %7s %6s %6s  %s" % ("samp%", "inst%", "lanes", "longsb%", "barr%", "shsb%", "source"))
for o in sorted(out, key=lambda x: -x[0])[:n]:
    print("%6.2f%% %6.2f%% %5.1f %6.1f%% %5.1f%% %5.1f%%  %s:%d %s" % (100.0 * o[0] / tot, 100.0 * o[1] / toti, o[2] / max(1, o[1]), 100.0 * o[6] / max(1, o[0]),
                                                        100.0 * o[7] / max(1, o[0]), 100.0 * o[8] / max(1, o[0]), o[3], o[4], o[5][:110]))
