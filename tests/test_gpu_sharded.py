"""GPU: the sharded-table form of the scan (include/real_gpu.h, real_gpu_comm_*), the multi-GPU path of SURVEY.md 8(e).

One GPU is enough to exercise it: the ranks are handles of this process on the same device, connected with
real_gpu_comm_connect_local and driven from one host thread each -- exactly what one process per GPU does, except
that the peer windows are reached through plain device pointers instead of CUDA IPC mappings.  Every rank builds the
tables of its own buckets only, partitions its slice of every round's text positions and writes the window records
into the owners' windows; the results (union of the hits / merged unique states) must equal the oracle's."""
import threading

import numpy as np
import pytest

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


def _ranks(cls, opts, nranks, round_positions, table_bits=0):
    ms = [cls(opts, table_bits=table_bits) for _ in range(nranks)]
    for r, m in enumerate(ms):
        m.handle.comm_init(r, nranks, round_positions)
    for m in ms:
        m.handle.comm_connect_local([x.handle for x in ms])
    return ms


def _run_threads(fns):
    """One host thread per rank (the ranks wait for each other on the device)."""
    errs, outs = [], [None] * len(fns)

    def run(i):
        try:
            outs[i] = fns[i]()
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=run, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts), "a rank is stuck"
    if errs:
        raise errs[0]
    return outs


@pytest.mark.parametrize("nranks,round_positions,e,table_bits", [(2, 1 << 18, 4, 0), (4, 1 << 19, 3, 0), (3, 1 << 30, 4, 24), (8, 1 << 17, 4, 0), (4, 1 << 18, 4, 32)])
def test_sharded_match_all_vs_oracle(nranks, round_positions, e, table_bits):
    text, reads = _fresh(100 + nranks)
    kw = dict(seedl=32, seedkmax=2, totalkmax=e, scores=False)
    ref = O.match_all(text, reads, **kw)
    ms = _ranks(matcher.AllMatcher, matcher.RealOptions(**kw), nranks, round_positions, table_bits)
    try:
        words, nmask = text.packed()
        for m in ms:
            m.set_reads(reads.mapped, reads.offsets, None)
            m.set_text(words, nmask, text.n, text.record_starts)
        parts = _run_threads([m.match for m in ms])
        entries = [m.stats()["n_candidates"] for m in ms]
    finally:
        for m in ms:
            m.close()
    got = np.concatenate(parts)
    a, b = canon_hits(got), canon_hits(ref)
    assert len(b) > reads.nreads // 2
    assert a.shape == b.shape and np.array_equal(a, b)       # every hit found by exactly one rank
    assert sum(1 for p in parts if len(p)) == nranks and min(entries) > 0


@pytest.mark.parametrize("nranks,round_positions", [(2, 1 << 19), (4, 1 << 18)])
def test_sharded_match_unique_two_files_vs_oracle(nranks, round_positions):
    import torch
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

    ms = _ranks(matcher.UniqueMatcher, matcher.RealOptions(**kw), nranks, round_positions)
    try:
        dev = torch.device("cuda", 0)
        for m in ms:
            m.set_reads(reads.mapped, reads.offsets, None)
        for fi, t in enumerate((t0, t1)):
            words, nmask = t.packed()
            for m in ms:
                m.set_text(words, nmask, t.n, t.record_starts, fileid=fi)
            _run_threads([m.match for m in ms])
            # the exchange of real_b200.dist.unique_exchange (MIN of the keys, SUM of the ties), emulated on one device
            keys = [torch.empty(reads.nreads, dtype=torch.int64, device=dev) for _ in ms]
            ties = [torch.empty(reads.nreads, dtype=torch.uint8, device=dev) for _ in ms]
            for g, m in enumerate(ms):
                m.handle.unique_export_keys(keys[g].data_ptr())
            kmin = torch.stack(keys).min(dim=0).values.contiguous()
            for g, m in enumerate(ms):
                m.handle.unique_export_ties(kmin.data_ptr(), ties[g].data_ptr())
            tsum = torch.stack(ties).sum(dim=0).to(torch.uint8).contiguous()
            torch.cuda.synchronize()
            for m in ms:
                m.handle.unique_import(kmin.data_ptr(), tsum.data_ptr())
        infos = [m.info()[0] for m in ms]
    finally:
        for m in ms:
            m.close()
    want = matcher.canonical_unique(info_ref)
    states = matcher.umi_state(info_ref)
    assert (states == 4).sum() > 50 and (states == 1).sum() > 1000 and (states == 2).sum() > 1000
    for info in infos:
        assert np.array_equal(matcher.canonical_unique(info), want)


def _bucket_ranks(cls, opts, nranks, table_bits=0):
    ms = [cls(opts, table_bits=table_bits) for _ in range(nranks)]
    for r, m in enumerate(ms):
        m.handle.set_bucket_shard(r, nranks)
    return ms


@pytest.mark.parametrize("nranks,e,table_bits,n", [(2, 4, 0, 900_000), (8, 4, 0, 900_000), (3, 3, 24, 700_000), (5, 4, 0, 2_300_000),
                                                   (4, 4, 32, 900_000), (8, 4, 32, 700_000), (2, 3, 32, 700_000)])      # 32: the fused index build
def test_bucket_shards_match_all_vs_oracle(nranks, e, table_bits, n):
    """Bucket shards (real_gpu_set_bucket_shard): every rank reads the whole text but keeps the positions of its own
    buckets; no exchange.  The union of the ranks' hits must be the oracle's hit set, every hit found exactly once."""
    text, reads = _fresh(200 + nranks, n=n)
    kw = dict(seedl=32, seedkmax=2, totalkmax=e, scores=False)
    ref = O.match_all(text, reads, **kw)
    ms = _bucket_ranks(matcher.AllMatcher, matcher.RealOptions(**kw), nranks, table_bits)
    try:
        words, nmask = text.packed()
        parts, kept = [], []
        for m in ms:
            m.set_reads(reads.mapped, reads.offsets, None)
            m.set_text(words, nmask, text.n, text.record_starts)
            parts.append(m.match())
            kept.append(m.stats()["n_windows"])
    finally:
        for m in ms:
            m.close()
    a, b = canon_hits(np.concatenate(parts)), canon_hits(ref)
    assert len(b) > reads.nreads // 2
    assert a.shape == b.shape and np.array_equal(a, b)
    assert sum(kept) == text.n - 32 + 1 + 2 * 8          # every probed position (seed windows + two fragments) is kept by exactly one rank
    assert min(kept) > 0


def test_bucket_shards_low_complexity_text():
    """A text whose positions all fall into few buckets (long single-base and dinucleotide runs): one rank keeps nearly
    everything, the list of kept positions is flushed every sub-tile, others keep next to nothing."""
    text, reads = _fresh(77, n=400_000, nreads=6000)
    sym = text.symbols.copy()
    sym[50_000:150_000] = 0                       # poly-A
    sym[200_000:260_000:2] = 1                    # CGCG...
    sym[200_001:260_000:2] = 2
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, 78, 6000, 100, 0.01, fastq=False)
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    info_ref, _ = O.unique_init(reads.nreads, False)
    O.match_unique(text, reads, info_ref, None, **kw)
    import torch
    nranks = 4
    ms = _bucket_ranks(matcher.UniqueMatcher, matcher.RealOptions(**kw), nranks)
    try:
        dev = torch.device("cuda", 0)
        words, nmask = text.packed()
        for m in ms:
            m.set_reads(reads.mapped, reads.offsets, None)
            m.set_text(words, nmask, text.n, text.record_starts)
            m.match()
        keys = [torch.empty(reads.nreads, dtype=torch.int64, device=dev) for _ in ms]
        ties = [torch.empty(reads.nreads, dtype=torch.uint8, device=dev) for _ in ms]
        for g, m in enumerate(ms):
            m.handle.unique_export_keys(keys[g].data_ptr())
        kmin = torch.stack(keys).min(dim=0).values.contiguous()
        for g, m in enumerate(ms):
            m.handle.unique_export_ties(kmin.data_ptr(), ties[g].data_ptr())
        tsum = torch.stack(ties).sum(dim=0).to(torch.uint8).contiguous()
        torch.cuda.synchronize()
        for m in ms:
            m.handle.unique_import(kmin.data_ptr(), tsum.data_ptr())
        infos = [m.info()[0] for m in ms]
    finally:
        for m in ms:
            m.close()
    want = matcher.canonical_unique(info_ref)
    for info in infos:
        assert np.array_equal(matcher.canonical_unique(info), want)


def test_sharded_mode_refusals():
    """Order dependent folds need the whole table set in one handle; a rank that was never connected must not scan."""
    from real_b200 import lib as rlib
    text, reads = _fresh(77, n=200_000, nreads=1000)
    words, nmask = text.packed()
    m = matcher.UniqueMatcher(matcher.RealOptions(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    try:
        m.handle.comm_init(0, 2, 1 << 18)
        m.set_reads(reads.mapped, reads.offsets, None)
        m.set_text(words, nmask, text.n, text.record_starts)
        with pytest.raises(rlib.RealGpuError):
            m.match()
    finally:
        m.close()


def _ipc_rank(rank, nranks, conn, seed):
    """One rank in its own process (what one process per GPU does); the windows are mapped with CUDA IPC."""
    import os
    os.environ["REAL_GPU_COMM_TIMEOUT_MS"] = "20000"
    from real_b200 import matcher as M
    text, reads = _fresh(seed, n=400_000, nreads=6000)
    m = M.AllMatcher(M.RealOptions(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    try:
        conn.send(m.handle.comm_init(rank, nranks, 1 << 18))
        m.handle.comm_connect(conn.recv())
        m.set_reads(reads.mapped, reads.offsets, None)
        words, nmask = text.packed()
        m.set_text(words, nmask, text.n, text.record_starts)
        conn.send("ready")
        conn.recv()
        conn.send(m.match())
        conn.recv()                 # keep the window mapped until every rank is done
    finally:
        m.close()


def test_sharded_two_processes_ipc():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    nranks, seed = 2, 55
    pipes = [ctx.Pipe() for _ in range(nranks)]
    procs = [ctx.Process(target=_ipc_rank, args=(r, nranks, pipes[r][1], seed)) for r in range(nranks)]
    for p in procs:
        p.start()
    try:
        def get(c):
            assert c.poll(180), "a rank process did not answer"
            return c.recv()
        handles = b"".join(get(pipes[r][0]) for r in range(nranks))
        for r in range(nranks):
            pipes[r][0].send(handles)
        for r in range(nranks):
            assert get(pipes[r][0]) == "ready"
        for r in range(nranks):
            pipes[r][0].send("go")
        parts = [get(pipes[r][0]) for r in range(nranks)]
        for r in range(nranks):
            pipes[r][0].send("done")
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    text, reads = _fresh(seed, n=400_000, nreads=6000)
    ref = O.match_all(text, reads, seedl=32, seedkmax=2, totalkmax=4, scores=False)
    a, b = canon_hits(np.concatenate(parts)), canon_hits(ref)
    assert len(b) > 3000 and a.shape == b.shape and np.array_equal(a, b)
    assert all(len(p) for p in parts)


# ---- peer-memory fold of the unique state (real_gpu_fold_*) ----------------------------------------------------------

def _two_file_job(seed=41):
    t0, reads0 = _fresh(seed, n=500_000, nreads=8000)
    t1, reads1 = _fresh(seed + 2, n=300_000, nreads=4000, nrec=2)
    sym = t1.symbols.copy()
    sym[5000:45000] = t0.symbols[30000:70000]            # cross-file repeats
    t1 = synth.Text(sym, t1.records)
    reads = synth.concat_reads([reads0, reads1])
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    info_ref, _ = O.unique_init(reads.nreads, False)
    for fi, t in enumerate((t0, t1)):
        O.match_unique(t, reads, info_ref, None, fileid=fi, **kw)
    return (t0, t1), reads, kw, info_ref


@pytest.mark.parametrize("nranks,packed,table_bits", [(2, False, 0), (3, True, 0), (8, False, 0), (4, True, 32)])
def test_bucket_shards_fold_group_vs_oracle(nranks, packed, table_bits):
    """Bucket shards of ONE process folded with real_gpu_fold_unique_group (k_fold_push / k_fold_merge over peer pointers):
    after every file rank r holds the merged words of its own reads; they must equal the oracle's words and the numpy
    restatement of the fold (real_b200.dist.fold_reduce_scatter_reference) applied to the ranks' pre-fold states."""
    from real_b200 import dist as rdist
    from real_b200 import lib as rlib
    texts, reads, kw, info_ref = _two_file_job()
    ms = _bucket_ranks(matcher.UniqueMatcher, matcher.RealOptions(**kw), nranks, table_bits)
    R = reads.nreads
    try:
        for r, m in enumerate(ms):
            m.handle.fold_init(r, nranks, R + 5)
        for m in ms:
            m.handle.fold_connect_local([x.handle for x in ms])
        for m in ms:
            if packed:
                pk, boffs, lens, flags = synth.pack_reads_2bit(reads)
                m.handle.set_reads_packed(pk, R, byte_offsets=boffs, lengths=lens, wildcard_flags=flags)
            else:
                m.set_reads(reads.mapped, reads.offsets, None)
        for fi, t in enumerate(texts):
            words, nmask = t.packed()
            for m in ms:
                m.set_text(words, nmask, t.n, t.record_starts, fileid=fi)
                m.match()
            before = [m.handle.get_unique()[0] for m in ms]
            rlib.Handle.fold_unique_group([m.handle for m in ms])
            for r, m in enumerate(ms):
                lo, hi = rdist.own_read_range(R, r, nranks)
                got, _ = m.handle.get_unique(first=lo, count=hi - lo)
                assert np.array_equal(got, rdist.fold_reduce_scatter_reference(before, r)), (fi, r)
        merged = np.concatenate([m.handle.get_unique(first=rdist.own_read_range(R, r, nranks)[0],
                                                     count=rdist.own_read_range(R, r, nranks)[1] - rdist.own_read_range(R, r, nranks)[0])[0]
                                 for r, m in enumerate(ms)])
        fold_ms = [m.stats()["fold_ms"] for m in ms]
    finally:
        for m in ms:
            m.close()
    st = matcher.umi_state(info_ref)
    assert (st == 4).sum() > 50 and (st == 1).sum() > 1000 and (st == 2).sum() > 1000
    assert np.array_equal(matcher.canonical_unique(merged), matcher.canonical_unique(info_ref))
    assert all(t > 0 for t in fold_ms)


def test_fold_refusals():
    from real_b200 import lib as rlib
    text, reads = _fresh(79, n=200_000, nreads=1000)
    words, nmask = text.packed()
    m = matcher.UniqueMatcher(matcher.RealOptions(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    try:
        m.set_reads(reads.mapped, reads.offsets, None)
        m.set_text(words, nmask, text.n, text.record_starts)
        m.match()
        with pytest.raises(rlib.RealGpuError):           # never initialised
            m.handle.fold_unique()
        m.handle.fold_init(0, 2, 500)                    # sized for fewer reads than are set, and never connected
        with pytest.raises(rlib.RealGpuError):
            m.handle.fold_unique()
        with pytest.raises(rlib.RealGpuError):
            m.handle.fold_init(0, 2, 5000)               # twice
    finally:
        m.close()


def _fold_ipc_rank(rank, nranks, conn, seed):
    """One rank in its own process: bucket shard + fold window mapped with CUDA IPC, hand-over with device flags."""
    import os
    os.environ["REAL_GPU_COMM_TIMEOUT_MS"] = "60000"
    from real_b200 import matcher as M
    from real_b200 import dist as D
    text, reads = _fresh(seed, n=400_000, nreads=6000)
    m = M.UniqueMatcher(M.RealOptions(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    try:
        m.handle.set_bucket_shard(rank, nranks)
        conn.send(m.handle.fold_init(rank, nranks, reads.nreads))
        m.handle.fold_connect(conn.recv())
        m.set_reads(reads.mapped, reads.offsets, None)
        words, nmask = text.packed()
        m.set_text(words, nmask, text.n, text.record_starts)
        conn.send("ready")
        conn.recv()
        outs = []
        for _ in range(3):                               # three exchanges in a row: both staging areas are reused
            m.match()
            m.handle.fold_unique()
            lo, hi = D.own_read_range(reads.nreads, rank, nranks)
            outs.append(m.handle.get_unique(first=lo, count=hi - lo)[0])
        conn.send(outs)
        conn.recv()                 # keep the window mapped until every rank is done
    finally:
        m.close()


def test_fold_two_processes_ipc():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    nranks, seed = 2, 57
    pipes = [ctx.Pipe() for _ in range(nranks)]
    procs = [ctx.Process(target=_fold_ipc_rank, args=(r, nranks, pipes[r][1], seed)) for r in range(nranks)]
    for p in procs:
        p.start()
    try:
        def get(c):
            assert c.poll(240), "a rank process did not answer"
            return c.recv()
        handles = b"".join(get(pipes[r][0]) for r in range(nranks))
        for r in range(nranks):
            pipes[r][0].send(handles)
        for r in range(nranks):
            assert get(pipes[r][0]) == "ready"
        for r in range(nranks):
            pipes[r][0].send("go")
        parts = [get(pipes[r][0]) for r in range(nranks)]
        for r in range(nranks):
            pipes[r][0].send("done")
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    text, reads = _fresh(seed, n=400_000, nreads=6000)
    want, _ = O.unique_init(reads.nreads, False)
    O.match_unique(text, reads, want, None, seedl=32, seedkmax=2, totalkmax=4, scores=False)
    for it in range(3):
        got = np.concatenate([parts[r][it] for r in range(nranks)])
        assert np.array_equal(matcher.canonical_unique(got), matcher.canonical_unique(want)), it
    assert (matcher.umi_state(want) != 0).sum() > 3000


def test_sharded_tables_hit_buffer_overflow_is_collective():
    """Sharded tables, one rank finds far more hits than its buffer holds (poly-A reads on a poly-A stretch: every hit lies
    in bucket AAAA = rank 0).  That rank must NOT rescan on its own -- its extra rounds would pair with the peers' next
    call -- but fail with REAL_GPU_E_LIMIT after enlarging its buffer; the repeated call on every rank then succeeds and
    the rounds of the ranks are still in step (the union of the rows is the oracle's)."""
    from real_b200 import lib as rlib
    text, reads = _fresh(91, n=400_000, nreads=3000)
    sym = text.symbols.copy()
    sym[150_000:151_200] = 0
    text = synth.Text(sym, text.records)
    rng = np.random.RandomState(5)
    polya = []
    for _ in range(300):
        s = np.zeros(100, np.uint8)
        s[rng.randint(40, 100, size=2)] = rng.randint(1, 4, size=2)       # <= 2 substitutions behind the seed
        polya.append(s)
    reads = synth.concat_reads([reads, synth.reads_from_list(polya)])
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    ref = O.match_all(text, reads, **kw)
    assert len(ref) > 200_000
    nranks = 2
    ms = _ranks(matcher.AllMatcher, matcher.RealOptions(**kw), nranks, 1 << 18)

    def attempt(m):
        def run():
            try:
                return m.match()
            except rlib.RealGpuError as e:
                return e.code
        return run
    try:
        words, nmask = text.packed()
        for m in ms:
            m.set_reads(reads.mapped, reads.offsets, None)
            m.set_text(words, nmask, text.n, text.record_starts)
        first = _run_threads([attempt(m) for m in ms])
        assert first[0] == rlib.REAL_GPU_E_LIMIT and not isinstance(first[1], int)
        second = _run_threads([attempt(m) for m in ms])
        assert not any(isinstance(p, int) for p in second)
        third = _run_threads([attempt(m) for m in ms])            # and again: the rounds are still in step
    finally:
        for m in ms:
            m.close()
    b = canon_hits(ref)
    for parts in (second, third):
        a = canon_hits(np.concatenate(parts))
        assert a.shape == b.shape and np.array_equal(a, b)
    assert len(second[0]) > 200_000
