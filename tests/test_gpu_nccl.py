"""GPU, two or more devices: the REAL multi-process path -- one process per GPU, torch.distributed over NCCL, bucket shards,
the peer-memory fold (real_gpu_fold_unique over CUDA IPC windows and NVLink) and the NCCL form of the exchange
(real_b200.dist.unique_exchange), and the matchAll gather -- compared with what ONE handle holding everything produces.
Skipped on a single-GPU box (tests/test_gpu_sharded.py covers the same kernels there with several handles / processes on
one device)."""
import os
import socket

import numpy as np
import pytest

from real_b200 import matcher, synth

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _job(seed=71):
    text = synth.make_text(seed, 1_200_000, nrecords=4, n_per_million=1200)
    sym = text.symbols.copy()
    sym[600_000:630_000] = sym[2000:32_000]            # a repeat => NonUnique reads, ties between the ranks
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, seed + 1, 30_000, 100, 0.012, fastq=False)
    return text, reads


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    from real_b200 import dist as rdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        text, reads = _job()
        words, nmask = text.packed()
        kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
        R = reads.nreads
        out = {}
        # matchUnique, both forms of the exchange
        for form in ("peer", "nccl"):
            m = matcher.UniqueMatcher(matcher.RealOptions(**kw), device=rank)
            try:
                m.handle.set_bucket_shard(rank, world)
                if form == "peer":
                    rdist.connect_fold(m.handle, dev, R)
                m.set_reads(reads.mapped, reads.offsets, None)
                m.set_text(words, nmask, text.n, text.record_starts)
                for it in range(2):                         # twice: the state persists, the fold is idempotent
                    m.match()
                    if form == "peer":
                        m.handle.fold_unique()
                    else:
                        rdist.unique_exchange(rdist.HandleShard(m.handle, R))
                lo, hi = rdist.own_read_range(R, rank, world)
                out[form + "_own"] = m.handle.get_unique(first=lo, count=hi - lo)[0]
                out[form + "_digest"] = np.asarray([m.handle.unique_checksum(lo, hi - lo)], dtype=np.uint64)
                if form == "nccl":
                    out["nccl_all"] = m.handle.get_unique()[0]
            finally:
                m.close()
        # matchAll: the rows of the ranks gathered and merged on rank 0
        a = matcher.AllMatcher(matcher.RealOptions(**kw), device=rank)
        try:
            a.handle.set_bucket_shard(rank, world)
            a.set_reads(reads.mapped, reads.offsets, None)
            a.set_text(words, nmask, text.n, text.record_starts)
            merged = rdist.gather_match_all(a.match(), dst=0)
            if rank == 0:
                out["all_merged"] = merged
        finally:
            a.close()
        np.savez(out_path % rank, **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_ranks_equal_single_handle(tmp_path, world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from real_b200 import dist as rdist
    text, reads = _job()
    words, nmask = text.packed()
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    # the single handle
    m = matcher.UniqueMatcher(matcher.RealOptions(**kw))
    try:
        m.set_reads(reads.mapped, reads.offsets, None)
        m.set_text(words, nmask, text.n, text.record_starts)
        m.match()
        want = m.info()[0]
        want_digest = m.handle.unique_checksum()
    finally:
        m.close()
    a = matcher.AllMatcher(matcher.RealOptions(**kw))
    try:
        a.set_reads(reads.mapped, reads.offsets, None)
        a.set_text(words, nmask, text.n, text.record_starts)
        want_all = a.match()
    finally:
        a.close()
    st = matcher.umi_state(want)
    assert (st == 4).sum() > 100 and (st == 1).sum() > 5000 and (st == 2).sum() > 5000

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = [np.load(out % r) for r in range(world)]
    for form in ("peer", "nccl"):
        own = np.concatenate([res[r][form + "_own"] for r in range(world)])
        assert np.array_equal(matcher.canonical_unique(own), matcher.canonical_unique(want)), form
        digest = sum(int(res[r][form + "_digest"][0]) for r in range(world)) & 0xFFFFFFFFFFFFFFFF
        assert digest == want_digest == matcher.unique_checksum(want), form
    for r in range(world):
        assert np.array_equal(matcher.canonical_unique(res[r]["nccl_all"]), matcher.canonical_unique(want))
    got_all = res[0]["all_merged"]
    assert len(got_all) == len(want_all) and all(np.array_equal(got_all[f], want_all[f]) for f in got_all.dtype.names)
