#!/usr/bin/env python
"""Development: the C4 gate sample in detail -- which reads differ between the GPU path and the reference harness."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from real_b200 import matcher
from oracle import oracle_py as O

wl = bench.WORKLOADS["c4"]
nt, nr = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 250_000
cpu = bench.cpu_reference(wl, nt, nr, 1)
text, reads = cpu["_sample"]
ref = cpu["_result"]
opts = matcher.RealOptions(totalkmax=wl["e"], scores=True)
opts.gaps = True
words, nmask = text.packed()
for stage in ("match", "gaps"):
    m = matcher.UniqueMatcher(opts, table_bits=bench.table_bits_of(wl))
    m.set_reads(reads.mapped, reads.offsets, reads.quality)
    m.handle.set_block_windows(0)
    m.set_text(words, nmask, text.n, text.record_starts)
    m.match()
    if stage == "gaps":
        m.matchGaps(0)
    info, sc = m.info()
    grows = m.gaps()
    m.close()
    if stage == "match":
        # the oracle port for the state before the gapped pass
        iref, sref = O.unique_init(reads.nreads, True)
        O.match_unique(text, reads, iref, sref, totalkmax=wl["e"], scores=True, ll=O.build_ll())
        bad = np.nonzero((info != iref) | (sc.view(np.uint32) != sref.view(np.uint32)))[0]
        print("before the gapped pass: %d reads differ from the oracle port" % len(bad))
        for r in bad[:10]:
            print("  read %d gpu %016x %r port %016x %r" % (r, info[r], sc[r], iref[r], sref[r]))
        continue
    rd, rs = ref["unique"]["data"], ref["unique"]["score"]
    bad = np.nonzero((info != rd) | (sc.view(np.uint32) != rs.view(np.uint32)))[0]
    print("after the gapped pass: %d reads differ from the harness" % len(bad))
    st_g, st_r = matcher.umi_state(info), matcher.umi_state(rd)
    for r in bad[:20]:
        print("  read %d gpu %016x st %d %r | ref %016x st %d %r | L %d" % (r, info[r], st_g[r], sc[r], rd[r], st_r[r], rs[r], reads.offsets[r+1] - reads.offsets[r]))
    import collections
    print("  state pairs (gpu, ref):", collections.Counter(zip(st_g[bad].tolist(), st_r[bad].tolist())))
    g = grows[grows["present"] == 1]
    r = ref["gaps"]
    print("gap rows gpu %d ref %d" % (len(g), len(r)))
    sg, sr = set(g["patid"].tolist()), set(r["patid"].tolist())
    print("  only gpu %d only ref %d" % (len(sg - sr), len(sr - sg)), sorted(sg - sr)[:5], sorted(sr - sg)[:5])
    # the oracle port on the same sample
    iref, sref = O.unique_init(reads.nreads, True)
    ll = O.build_ll()
    O.match_unique(text, reads, iref, sref, totalkmax=wl["e"], scores=True, ll=ll)
    gref = np.zeros(reads.nreads, dtype=O.GAP_DTYPE)
    O.match_gaps(text, reads, iref, sref, gref, totalkmax=wl["e"], scores=True, ll=ll)
    print("port vs harness: %d words differ; port vs gpu: %d" % (int((iref != rd).sum()), int((iref != info).sum())))
