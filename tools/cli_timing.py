#!/usr/bin/env python
"""End-to-end wall time of the `real` command line (this repo's GPU path vs the stock CPU binary oracle/_ref/real) on one
mid-size synthetic case, with the phase times of the GPU driver (REAL_TIMING=1) and a byte comparison of the outputs.
usage: cli_timing.py [text_bases] [reads]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from real_b200 import build as rbuild, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
work = tempfile.mkdtemp(prefix="cli_timing_")
t0 = time.time()
text = synth.make_text(11, n, nrecords=4, n_per_million=500)
reads = synth.make_reads(text, 12, R, 100, 0.01, fastq=False)
synth.write_fasta(os.path.join(work, "t.fa"), text)
synth.write_reads(os.path.join(work, "r.fa"), reads, False)
print("inputs: %d bp, %d reads, generated in %.1f s" % (n, R, time.time() - t0), flush=True)
rbuild.build(); rbuild.build_host()
flags = ["-u", "1", "-s", "2", "-e", "4", "-l", "32", "-q", "0", "-R", "0"]
res = {}
# one discarded invocation first: the first process on a fresh box pays for paging in the CUDA libraries
subprocess.run([rbuild.HOST_BIN, "-t", os.path.join(work, "t.fa"), "-p", os.path.join(work, "r.fa"), "-o", os.path.join(work, "warm.txt")] + flags,
               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
for name, exe, extra, env in (("gpu_T1", rbuild.HOST_BIN, ["-T", "1"], {"REAL_TIMING": "1"}),
                              ("gpu_Tall", rbuild.HOST_BIN, [], {"REAL_TIMING": "1"}),
                              ("gpu_Tall_hostreads", rbuild.HOST_BIN, [], {"REAL_TIMING": "1", "REAL_READS_LOADER": "host"}),
                              ("gpu_Tall_hostfmt", rbuild.HOST_BIN, [], {"REAL_TIMING": "1", "REAL_FORMAT": "host"}),
                              ("gpu_Tall_hosttext", rbuild.HOST_BIN, [], {"REAL_TIMING": "1", "REAL_TEXT_LOADER": "host"}),
                              ("stock_cpu", os.path.join(ROOT, "oracle", "_ref", "real"), ["-T", str(os.cpu_count() or 1)], {})):
    if not os.path.exists(exe):
        print(name, "binary missing"); continue
    out = os.path.join(work, name + ".txt")
    t0 = time.time()
    p = subprocess.run([exe, "-t", os.path.join(work, "t.fa"), "-p", os.path.join(work, "r.fa"), "-o", out] + flags + extra,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, **env))
    dt = time.time() - t0
    res[name] = out
    print("%-10s rc=%d wall %.2f s, %d output lines" % (name, p.returncode, dt, sum(1 for _ in open(out)) if os.path.exists(out) else -1), flush=True)
    for l in p.stderr.splitlines():
        if l.startswith("[timing]"):
            print("    " + l)
if "gpu_Tall" in res and "stock_cpu" in res:
    same = open(res["gpu_Tall"], "rb").read() == open(res["stock_cpu"], "rb").read()
    print("outputs identical to the stock binary:", same)
if "gpu_Tall_hosttext" in res and "gpu_Tall" in res:
    print("outputs identical between the device and the host text loader:", open(res["gpu_Tall_hosttext"], "rb").read() == open(res["gpu_Tall"], "rb").read())
if "gpu_Tall_hostreads" in res and "gpu_Tall" in res:
    print("outputs identical between the device and the host reads loader:", open(res["gpu_Tall_hostreads"], "rb").read() == open(res["gpu_Tall"], "rb").read())
if "gpu_Tall_hostfmt" in res and "gpu_Tall" in res:
    print("outputs identical between the device and the host formatter:", open(res["gpu_Tall_hostfmt"], "rb").read() == open(res["gpu_Tall"], "rb").read())
if "gpu_T1" in res and "gpu_Tall" in res:
    print("outputs identical between -T 1 and all threads:", open(res["gpu_T1"], "rb").read() == open(res["gpu_Tall"], "rb").read())
