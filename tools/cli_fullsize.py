#!/usr/bin/env python
"""The `real` command line (real_b200/bin/real) on inputs of BASELINE.json's C3 size: a 3.1 Gbp FASTA text (24 records, 0.1 % N)
and 50 M x 100 bp FASTA reads, written to files first.  Phase times of the driver (REAL_TIMING=1), wall time, output size.
The stock binary is not run at this size (its text loader alone needs ~43 s and its matching ~9 minutes on 16 cores,
DESIGN.md 7); the byte-identity of the outputs is what tests/test_cli_gpu.py and tools/cli_timing.py check at sizes it finishes.

usage: cli_fullsize.py [text_bases] [reads] [workdir]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch

from real_b200 import build as rbuild, devsynth
import ingest_bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
work = sys.argv[3] if len(sys.argv) > 3 else ("/dev/shm/cli_fullsize" if os.path.isdir("/dev/shm") else "/tmp/cli_fullsize")
L = 100
os.makedirs(work, exist_ok=True)
rbuild.build(); rbuild.build_host()

t0 = time.time()
data, sym, starts = ingest_bench.make_fasta_on_device(torch, n, 24, 7)
with open(os.path.join(work, "t.fa"), "wb") as f:
    step = 1 << 28
    for o in range(0, data.numel(), step):
        f.write(data[o:o + step].cpu().numpy().tobytes())
text_bytes = data.numel()
del data
# reads cut from the text on the device (the generator kernels of the bench), written as 112-byte FASTA records
words, mask = ingest_bench.expected_words(torch, sym)
words = torch.cat([words, torch.zeros(4, dtype=torch.int64, device="cuda")])          # the generator reads a word past the last base
mask = torch.cat([mask, torch.zeros(4, dtype=torch.int64, device="cuda")])
del sym
torch.cuda.empty_cache()
lut = torch.tensor([65, 67, 71, 84, 78], dtype=torch.uint8, device="cuda")
digits = torch.tensor([10 ** k for k in range(7, -1, -1)], dtype=torch.int64, device="cuda")
reads_bytes = 0
with open(os.path.join(work, "r.fa"), "wb") as f:
    per = 5_000_000
    for first in range(0, R, per):
        c = min(per, R - first)
        mapped, _, _ = devsynth.reads_device(8, words, mask, n, R, L, 0.01, first=first, count=c)
        rec = torch.empty((c, 1 + 1 + 8 + 1 + L + 1), dtype=torch.uint8, device="cuda")
        rec[:, 0] = 62; rec[:, 1] = 114                                                    # ">r"
        idx = torch.arange(first, first + c, dtype=torch.int64, device="cuda")
        rec[:, 2:10] = ((idx[:, None] // digits[None, :]) % 10 + 48).to(torch.uint8)
        rec[:, 10] = 10
        rec[:, 11:11 + L] = lut[mapped.view(c, L).clamp(max=4).long()]
        rec[:, 11 + L] = 10
        b = rec.reshape(-1).cpu().numpy().tobytes()
        f.write(b); reads_bytes += len(b)
        del mapped, rec
del words, mask
torch.cuda.empty_cache()
print("inputs written in %.1f s: text %.2f GB, reads %.2f GB" % (time.time() - t0, text_bytes / 1e9, reads_bytes / 1e9), flush=True)

flags = ["-u", "1", "-s", "2", "-e", "4", "-l", "32", "-q", "0"]
for name, out, env in (("to a file", os.path.join(work, "out.txt"), {}), ("to /dev/null", "/dev/null", {}), ("to /dev/null, 2nd run", "/dev/null", {})):
    t1 = time.time()
    p = subprocess.run([rbuild.HOST_BIN, "-t", os.path.join(work, "t.fa"), "-p", os.path.join(work, "r.fa"), "-o", out] + flags,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, REAL_TIMING="1", **env))
    dt = time.time() - t1
    size = os.path.getsize(out) if out != "/dev/null" and os.path.exists(out) else 0
    print("%-24s rc=%d wall %.2f s%s" % (name, p.returncode, dt, (", output %.2f GB" % (size / 1e9)) if size else ""), flush=True)
    for l in p.stderr.splitlines():
        if l.startswith("[timing]") or l.startswith("unique:") or l.startswith("Number of patterns"):
            print("    " + l)
    if p.returncode != 0:
        print(p.stderr[-1500:])
for fn in ("t.fa", "r.fa", "out.txt"):
    try:
        os.remove(os.path.join(work, fn))
    except OSError:
        pass
