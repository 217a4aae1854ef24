"""CPU: fresh seeded inputs through the REAL reference (oracle/_ref/ref_harness, built from the
reference's own sources by oracle/Makefile) against the oracle restatement.  Skipped where the
harness binary does not exist."""
import os
import tempfile

import numpy as np
import pytest

from oracle import oracle_py as O
from real_b200 import synth
from util import canon_hits

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/ref_harness not built (needs /root/reference)")


def _run_ref(mode, text, reads, fastq, args, n_list=0):
    with tempfile.TemporaryDirectory() as work:
        synth.write_fasta(os.path.join(work, "t.fa"), text)
        rf = os.path.join(work, "r.fq" if fastq else "r.fa")
        synth.write_reads(rf, reads, fastq)
        env = {"REAL_HARNESS_NLIST": str(n_list)} if n_list else {}
        a = ["-t", os.path.join(work, "t.fa"), "-p", rf, "-o", "x", "-u", "1" if mode == "unique" else "0", "-R", "0"] + args
        _, dump, _ = O.run_ref(mode, work, a, env=env)
        return np.fromfile(dump, dtype=O.HIT_DTYPE if mode == "all" else O.UNIQUE_DTYPE)


@pytest.mark.parametrize("seed,L,e,scores,nrec,npm,n_list", [
    (901, 36, 2, False, 1, 0, 0),
    (902, 100, 4, True, 3, 2000, 0),
    (903, 75, 5, True, 5, 4000, 17000),
])
def test_all_fresh(seed, L, e, scores, nrec, npm, n_list):
    text = synth.make_text(seed, 40000, nrecords=nrec, n_per_million=npm)
    reads = synth.make_reads(text, seed + 1, 400, L, 0.02, fastq=True)
    ref = _run_ref("all", text, reads, True, ["-e", str(e), "-q", "1" if scores else "0", "-Q", "33"], n_list)
    got = O.match_all(text, reads, totalkmax=e, scores=scores, n_list=n_list)
    a, b = canon_hits(got, with_block=True), canon_hits(ref, with_block=True)
    assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("seed,scores,n_list", [(911, False, 0), (912, True, 15000)])
def test_unique_fresh(seed, scores, n_list):
    text = synth.make_text(seed, 40000, nrecords=2, n_per_million=1000)
    sym = text.symbols.copy()
    sym[20000:24000] = sym[5000:9000]          # a repeat: NonUnique reads
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, seed + 1, 500, 64, 0.015, fastq=True)
    ref = _run_ref("unique", text, reads, True, ["-e", "4", "-q", "1" if scores else "0", "-Q", "33"], n_list)
    info, sc = O.unique_init(reads.nreads, scores)
    O.match_unique(text, reads, info, sc, totalkmax=4, scores=scores, n_list=n_list)
    assert np.array_equal(info, ref["data"])
    if scores:
        assert np.array_equal(sc.view(np.uint32), ref["score"].view(np.uint32))
