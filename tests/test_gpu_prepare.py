"""GPU: real_gpu_prepare_scan (include/real_gpu.h) -- text first, reads second.  The records of the text scan are formed
ahead of the match call, on a stream of their own, while the reads travel; the match call uses them when it would form the
very same ones and forms its own otherwise.  Whatever the order of the calls, the result is the one of the plain order
(set_reads, set_text, match) and of the oracle: match.hpp:335-416 through AllMatcher::match / UniqueMatcher::match."""
import numpy as np
import pytest
import torch

from oracle import oracle_py as O
from real_b200 import matcher, synth
from util import canon_hits

pytestmark = pytest.mark.gpu


def _fresh(seed, n=900_000, nreads=20_000, L=100, nrec=5, npm=1500, sub=0.012):
    text = synth.make_text(seed, n, nrecords=nrec, n_per_million=npm)
    sym = text.symbols.copy()
    sym[n // 2:n // 2 + 20000] = sym[1000:21000]          # a repeat => multi-hit reads
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, seed + 1, nreads, L, sub, fastq=False)
    return text, reads


def _pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


@pytest.mark.parametrize("async_text", [False, True])
@pytest.mark.parametrize("table_bits", [0, 32])
def test_prepared_match_all_equals_oracle(async_text, table_bits):
    text, reads = _fresh(311)
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    ref = O.match_all(text, reads, **kw)
    words, nmask = text.packed()
    keep_w, words = _pinned(words.view(np.uint64))
    keep_m, nmask = _pinned(nmask.view(np.uint64))
    m = matcher.AllMatcher(matcher.RealOptions(**kw), table_bits=table_bits)
    try:
        for rep in range(2):            # the second round finds every buffer in place (no allocation between the calls)
            m.handle.set_text(words, nmask, text.n, text.record_starts, async_copy=async_text)
            m.handle.prepare_scan(100)
            m.set_reads(reads.mapped, reads.offsets, None)
            got = m.match()
            st = m.stats()
            assert st["prepared_scans"] == rep + 1
            a, b = canon_hits(got), canon_hits(ref)
            assert len(b) > reads.nreads // 2
            assert a.shape == b.shape and np.array_equal(a, b)
        # a scan without preparation after one with: it forms its own records again
        got = m.match()
        assert m.stats()["prepared_scans"] == 2
        assert np.array_equal(canon_hits(got), canon_hits(ref))
    finally:
        m.close()


def test_prepared_match_unique_two_files():
    """Two text files against one read set, the second one prepared while nothing else happens; states equal the oracle's."""
    t0, reads0 = _fresh(31, n=500_000, nreads=8000)
    t1, reads1 = _fresh(33, n=300_000, nreads=4000, nrec=2)
    sym = t1.symbols.copy()
    sym[5000:45000] = t0.symbols[30000:70000]            # cross-file repeats
    t1 = synth.Text(sym, t1.records)
    reads = synth.concat_reads([reads0, reads1])
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    info_ref, _ = O.unique_init(reads.nreads, False)
    for fi, t in enumerate((t0, t1)):
        O.match_unique(t, reads, info_ref, None, fileid=fi, **kw)
    m = matcher.UniqueMatcher(matcher.RealOptions(**kw))
    try:
        keep = []
        for fi, t in enumerate((t0, t1)):
            words, nmask = t.packed()
            kw_, w = _pinned(words.view(np.uint64))
            km_, nm = _pinned(nmask.view(np.uint64))
            keep += [kw_, km_]
            m.handle.set_text(w, nm, t.n, t.record_starts, fileid=fi, async_copy=True)
            m.handle.prepare_scan(100)
            if fi == 0:
                m.set_reads(reads.mapped, reads.offsets, None)
            m.match()
        assert m.stats()["prepared_scans"] == 2
        info = m.info()[0]
    finally:
        m.close()
    assert np.array_equal(matcher.canonical_unique(info), matcher.canonical_unique(info_ref))


@pytest.mark.parametrize("nranks", [2, 8])
def test_prepared_bucket_shards(nranks):
    """Bucket shards (dense kernels with a filter at two ranks, the kept-position list at eight): every rank prepares its
    own share; the union of the hits is the oracle's and the kept positions add up."""
    text, reads = _fresh(200 + nranks)
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    ref = O.match_all(text, reads, **kw)
    words, nmask = text.packed()
    ms = [matcher.AllMatcher(matcher.RealOptions(**kw)) for _ in range(nranks)]
    try:
        parts, kept = [], []
        for r, m in enumerate(ms):
            m.handle.set_bucket_shard(r, nranks)
            m.set_text(words, nmask, text.n, text.record_starts)
            m.handle.prepare_scan(100)
            m.set_reads(reads.mapped, reads.offsets, None)
            parts.append(m.match())
            st = m.stats()
            assert st["prepared_scans"] == 1
            kept.append(st["n_windows"])
    finally:
        for m in ms:
            m.close()
    a, b = canon_hits(np.concatenate(parts)), canon_hits(ref)
    assert a.shape == b.shape and np.array_equal(a, b)
    assert sum(kept) == text.n - 32 + 1 + 2 * 8 and min(kept) > 0


def test_text_set_after_reads_prepares_itself():
    """The usual order (reads, then text after text): real_gpu_set_text starts the partition of the new text itself; the
    match call finds the records, a second match call on the same text re-uses them (nothing is formed twice), a scan of the
    same reads after another text forms new ones.  Results equal the oracle's every time."""
    text, reads = _fresh(511, n=700_000, nreads=9000)
    other, _ = _fresh(512, n=500_000, nreads=10)
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    ref = canon_hits(O.match_all(text, reads, **kw))
    ref_other = canon_hits(O.match_all(other, reads, **kw))
    m = matcher.AllMatcher(matcher.RealOptions(**kw))
    try:
        m.set_reads(reads.mapped, reads.offsets, None)
        m.set_text(*text.packed(), text.n, text.record_starts)
        assert np.array_equal(canon_hits(m.match()), ref)
        st = m.stats()
        assert st["prepared_scans"] == 1 and st["part_ms"] > 0
        assert np.array_equal(canon_hits(m.match()), ref)
        st = m.stats()
        assert st["prepared_scans"] == 1 and st["part_ms"] == 0          # the records of the last scan
        m.set_text(*other.packed(), other.n, other.record_starts)
        assert np.array_equal(canon_hits(m.match()), ref_other)
        assert m.stats()["prepared_scans"] == 2
        m.set_text(*text.packed(), text.n, text.record_starts)
        m.handle.prepare_scan(100)                                       # already on its way: nothing happens
        assert np.array_equal(canon_hits(m.match()), ref)
        assert m.stats()["prepared_scans"] == 3
    finally:
        m.close()


def test_preparation_that_does_not_fit_is_ignored(monkeypatch):
    """A text shard prepared for reads of 60 bases and matched with reads of 100 (other last window), a text replaced after
    the preparation, a bucket shard set after it: the scan forms its own records; results equal the plain order's.
    (REAL_GPU_AUTO_PREPARE=0: only the explicit calls prepare.)"""
    monkeypatch.setenv("REAL_GPU_AUTO_PREPARE", "0")
    text, reads = _fresh(411, n=600_000, nreads=8000)
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    ref = O.match_all(text, reads, **kw)
    words, nmask = text.packed()
    shards = matcher.shard_ranges(text.n, 2, 100)
    parts = []
    m = matcher.AllMatcher(matcher.RealOptions(**kw))
    try:
        for sh in shards:
            m.set_text(words, nmask, text.n, text.record_starts, shard=sh)
            m.handle.prepare_scan(60)
            m.set_reads(reads.mapped, reads.offsets, None)
            parts.append(m.match())
        n_used = m.stats()["prepared_scans"]
        assert n_used <= 1                      # the last shard ends with the text: its window range does not depend on the read length
        a, b = canon_hits(np.concatenate(parts)), canon_hits(ref)
        assert a.shape == b.shape and np.array_equal(a, b)
        # replaced text
        other, _ = _fresh(412, n=600_000, nreads=10)
        ow, om = other.packed()
        m.set_text(ow, om, other.n, other.record_starts)
        m.handle.prepare_scan(100)
        m.set_text(words, nmask, text.n, text.record_starts)
        got = m.match()
        assert m.stats()["prepared_scans"] == n_used
        assert np.array_equal(canon_hits(got), canon_hits(ref))
        # prepared for the whole signature space, then made a bucket shard
        m.handle.prepare_scan(100)
        m.handle.set_bucket_shard(0, 2)
        m.set_reads(reads.mapped, reads.offsets, None)
        half = m.match()
        assert m.stats()["prepared_scans"] == n_used
        assert 0 < len(half) < len(ref)
    finally:
        m.close()


def test_device_text_with_late_mask():
    """real_gpu_set_text_device_async: the words are taken at once, the wildcard mask when the match call starts -- the caller
    completes it in between (ranks that all-gather their inputs gather the mask while the index builds).  Reads cut across the
    text's N runs with A in place of N: they match only if the mask is ignored, so a stale mask would show."""
    text, reads = _fresh(611, n=600_000, nreads=6000, npm=3000)
    sym = text.symbols
    npos = np.flatnonzero(sym == 4)
    assert npos.size > 500
    rng = np.random.default_rng(5)
    picks = rng.choice(npos[(npos > 200) & (npos < text.n - 200)], size=300, replace=False)
    seqs = []
    for p in picks:
        s0 = int(p) - int(rng.integers(0, 100))
        q = sym[s0:s0 + 100].copy()
        q[q == 4] = 0
        seqs.append(q)
    across = synth.reads_from_list(seqs)
    allreads = synth.concat_reads([reads, across])
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    ref = canon_hits(O.match_all(text, allreads, **kw))
    words, nmask = text.packed()
    dev = torch.device("cuda", 0)
    d_words = torch.from_numpy(np.ascontiguousarray(words).view(np.int64)).to(dev)
    d_mask = torch.zeros(np.ascontiguousarray(nmask).size, dtype=torch.int64, device=dev)          # not there yet
    torch.cuda.synchronize()
    m = matcher.AllMatcher(matcher.RealOptions(**kw))
    try:
        for rep in range(2):
            d_mask.zero_()
            torch.cuda.synchronize()
            m.handle.set_text_device(d_words.data_ptr(), d_mask.data_ptr(), text.n, text.record_starts, async_copy=True)
            m.handle.prepare_scan(100)
            m.set_reads(allreads.mapped, allreads.offsets, None)
            d_mask.copy_(torch.from_numpy(np.ascontiguousarray(nmask).view(np.int64)))                 # the mask arrives after the reads
            torch.cuda.synchronize()
            got = canon_hits(m.match())
            assert got.shape == ref.shape and np.array_equal(got, ref)
        assert m.stats()["prepared_scans"] == 2
        # the same with the mask left empty: the reads across the N runs match -- the test does look at the mask
        d_mask.zero_()
        torch.cuda.synchronize()
        m.handle.set_text_device(d_words.data_ptr(), d_mask.data_ptr(), text.n, text.record_starts, async_copy=True)
        wrong = canon_hits(m.match())
        assert len(wrong) > len(ref)
    finally:
        m.close()
