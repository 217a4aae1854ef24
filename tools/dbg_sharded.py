"""development: two ranks of one process on one GPU, verbose"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("REAL_GPU_COMM_TIMEOUT_MS", "3000")
import numpy as np
from real_b200 import matcher, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NTEXT = int(sys.argv[2]) if len(sys.argv) > 2 else 600_000
NREADS = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
ROUND = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 18
text = synth.make_text(5, NTEXT, nrecords=3, n_per_million=1000)
reads = synth.make_reads(text, 6, NREADS, 100, 0.01, fastq=False)
words, nmask = text.packed()
kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
ms = [matcher.AllMatcher(matcher.RealOptions(**kw)) for _ in range(N)]
for r, m in enumerate(ms):
    m.handle.comm_init(r, N, ROUND)
for m in ms:
    m.handle.comm_connect_local([x.handle for x in ms])
for m in ms:
    m.set_reads(reads.mapped, reads.offsets, None)
    m.set_text(words, nmask, text.n, text.record_starts)
t0 = time.time()
res = [None] * N
def run(i):
    print("rank", i, "start", round(time.time() - t0, 3), flush=True)
    try:
        res[i] = len(ms[i].match())
    except Exception as e:
        res[i] = repr(e)
    print("rank", i, "end", round(time.time() - t0, 3), res[i], flush=True)
ts = [threading.Thread(target=run, args=(i,)) for i in range(N)]
for t in ts: t.start()
for t in ts: t.join()
print("result", res, "stats", [m.stats()["n_hits"] for m in ms])
