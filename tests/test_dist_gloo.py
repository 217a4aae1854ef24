"""CPU, world_size 2, gloo: the cross-shard fold of the per-read unique state (real_b200.dist.unique_exchange)
-- the one exchange step of the path.  Each rank holds the state a text shard would produce (built here from
the oracle's complete hit set, cut by hit start position), the exchange runs over torch.distributed, and the
merged state must equal the oracle's matchUnique over the whole text.  The three key transforms that the
device kernels implement (k_unique_export / k_unique_ties / k_unique_import, csrc/post.cuh) are restated in
numpy for the CPU participants."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_py as O
from real_b200 import dist as rdist
from real_b200 import matcher, synth

NONE = rdist.UNIQUE_KEY_NONE


def fold_hits(nreads, hits):
    """UpdateUniqueInfo<false>::update over a hit list (order independent form)."""
    info = np.zeros(nreads, dtype=np.uint64)
    best = {}
    for h in hits:
        r = int(h["patid"])
        cur = best.get(r)
        key = (int(h["k"]), int(h["file"]), int(h["pos"]), int(h["inverted"]), int(h["frag"]))
        if cur is None or key[0] < cur[0][0]:
            best[r] = [key, False]
        elif key[0] == cur[0][0]:
            if key[1:3] != cur[0][1:3]:
                cur[1] = True
                if key[1:3] < cur[0][1:3]:
                    cur[0] = key
            elif key[3] < cur[0][3]:
                cur[0] = key
    for r, (key, non) in best.items():
        k, f, p, inv, frag = key
        st = 4 if non else (2 if inv else 1)
        info[r] = p | (f << 35) | (k << 41) | (frag << 45) | (st << 61)
    return info


np_key = matcher.unique_key


class NumpyShard:
    def __init__(self, info):
        self.info = info
        self.nreads = info.size
        self.device = torch.device("cpu")

    def export_keys(self, keys):
        keys.copy_(torch.from_numpy(np_key(self.info)))

    def export_ties(self, min_keys, ties):
        mine, win = np_key(self.info), min_keys.numpy()
        samepos = (((mine ^ win) >> 17) & ((1 << 41) - 1)) == 0
        ties.copy_(torch.from_numpy(((mine != NONE) & ((mine >> 59) == (win >> 59)) & ~samepos).astype(np.uint8)))

    def import_merged(self, min_keys, tie_sums):
        key, ts = min_keys.numpy(), tie_sums.numpy()
        has = key != NONE
        uniq = (((key >> 58) & 1) == 1) & (ts == 0)
        strand = (key >> 16) & 1
        st = np.where(uniq, np.where(strand == 1, 2, 1), 4)
        merged = ((key >> 17) & ((1 << 35) - 1)) | (((key >> 52) & 63) << 35) | ((key >> 59) << 41) | ((key & 0xFFFF) << 45) | (st << 61)
        self.info = np.where(has, merged.astype(np.uint64), self.info)

    def sync(self):
        pass


class FakeHandle:
    """Stands in for real_b200.lib.Handle in the window set-up (no GPU here): records what it is given."""

    def __init__(self):
        self.got = None

    def comm_init(self, rank, nranks, round_positions):
        self.args = (rank, nranks, round_positions)
        return bytes([rank + 1]) * 64

    def comm_connect(self, all_handles):
        self.got = all_handles

    def fold_init(self, rank, nranks, max_reads):
        self.fargs = (rank, nranks, max_reads)
        return bytes([rank + 101]) * 64

    def fold_connect(self, all_handles):
        self.fgot = all_handles


def _worker(rank, world, port, shard_infos, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fh = FakeHandle()
        rdist.connect_sharded_tables(fh, torch.device("cpu"), round_positions=1 << 20)
        assert fh.args == (rank, world, 1 << 20)
        assert fh.got == b"".join(bytes([r + 1]) * 64 for r in range(world))       # the handles of all ranks, in rank order
        rdist.connect_fold(fh, torch.device("cpu"), max_reads=1234)
        assert fh.fargs == (rank, world, 1234)
        assert fh.fgot == b"".join(bytes([r + 101]) * 64 for r in range(world))
        sh = NumpyShard(shard_infos[rank].copy())
        # the reduce-scatter form of the fold (real_gpu_fold_unique): every rank sees the words of all ranks for ITS reads --
        # here they travel by an all-gather over gloo -- and folds them in one pass
        mine = torch.from_numpy(shard_infos[rank].view(np.int64).copy())
        allw = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allw, mine)
        own = rdist.fold_reduce_scatter_reference([w.numpy().view(np.uint64) for w in allw], rank)
        np.save((out_path % rank) + ".own.npy", own)
        rdist.unique_exchange(sh)
        np.save(out_path % rank, sh.info)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("split", ["text", "tables"])
def test_unique_exchange_world2_gloo(tmp_path, split):
    text = synth.make_text(77, 120_000, nrecords=3, n_per_million=1500)
    sym = text.symbols.copy()
    sym[70_000:76_000] = sym[10_000:16_000]          # repeats across the shard boundary => cross-shard ties
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, 78, 1500, 64, 0.015, fastq=False)
    kw = dict(seedl=32, seedkmax=2, totalkmax=4, scores=False)
    hits = O.match_all(text, reads, **kw)
    want, _ = O.unique_init(reads.nreads, False)
    O.match_unique(text, reads, want, None, **kw)
    world = 2
    shards = matcher.shard_ranges(text.n, world, 64)
    if split == "text":
        # text sharded: a rank holds the hits that start in its range
        infos = [fold_hits(reads.nreads, hits[(hits["pos"] >= ob) & (hits["pos"] < oe)]) for ob, oe, _, _ in shards]
    else:
        # tables sharded: a hit is found by the rank that owns the bucket of the seed window it is reported through, which
        # scatters the hits of a read -- even the two strands at one position -- over the ranks
        owner = (hits["pos"].astype(np.int64) * 2654435761 + hits["inverted"].astype(np.int64) * 40503 + hits["patid"].astype(np.int64)) >> 7
        infos = [fold_hits(reads.nreads, hits[(owner % world) == r]) for r in range(world)]
    assert np.array_equal(matcher.canonical_unique(fold_hits(reads.nreads, hits)), matcher.canonical_unique(want))   # the fold itself is right
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank%d.npy")
    mp.spawn(_worker, args=(world, port, infos, out), nprocs=world, join=True)
    st = matcher.umi_state(want)
    assert (st == 4).sum() > 20 and (st == 1).sum() > 300 and (st == 2).sum() > 300
    for r in range(world):
        got = np.load(out % r)
        assert np.array_equal(matcher.canonical_unique(got), matcher.canonical_unique(want))
    # the reduce-scatter form: the ranks' own ranges, put together, are the same state -- bit for bit what the all-reduce form gives
    own = np.concatenate([np.load((out % r) + ".own.npy") for r in range(world)])
    assert np.array_equal(own, np.load(out % 0))


def _upload_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)           # every rank holds the same host inputs
        a = torch.randint(0, 256, (100_003,), dtype=torch.uint8, generator=g)
        b = torch.randint(-2**62, 2**62, (777,), dtype=torch.int64, generator=g)
        c = torch.randint(0, 256, (5,), dtype=torch.uint8, generator=g)
        up = rdist.ShardedUpload([("a", a), ("b", b.view(torch.uint8)), ("c", c)], torch.device("cpu"))
        assert up.chunk * world >= a.numel() + b.numel() * 8 + c.numel() and up.chunk % rdist.ShardedUpload.ALIGN == 0
        # a rank only ever reads its own slice of the blob: wipe the rest to prove it
        lo = rank * up.chunk
        keep = up.blob[lo:lo + up.chunk].clone()
        up.blob.zero_()
        up.blob[lo:lo + up.chunk] = keep
        d = up.run()
        ok = torch.equal(d["a"], a) and torch.equal(d["b"].view(torch.int64), b) and torch.equal(d["c"], c)
        np.save(out_path % rank, np.asarray([int(ok)]))
    finally:
        dist.destroy_process_group()


def test_sharded_upload_world2_gloo(tmp_path):
    """real_b200.dist.ShardedUpload: every rank contributes 1/N of the input bytes, the all-gather completes them everywhere."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "up%d.npy")
    mp.spawn(_upload_worker, args=(2, port, out), nprocs=2, join=True)
    assert all(int(np.load(out % r)[0]) == 1 for r in range(2))


def _lib_hits(h):
    """oracle rows (one text block) in the row type the library returns"""
    from real_b200 import lib as rlib
    out = np.zeros(len(h), dtype=rlib.HIT_DTYPE)
    for f in ("patid", "pos", "file", "frag", "k", "inverted", "score"):
        out[f] = h[f]
    return out


def _all_case():
    text = synth.make_text(91, 90_000, nrecords=2, n_per_million=1000)
    sym = text.symbols.copy()
    sym[50_000:54_000] = sym[5_000:9_000]            # repeats => reads with several rows
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, 92, 1200, 64, 0.015, fastq=True)
    ll = O.build_ll()
    hits = _lib_hits(O.match_all(text, reads, seedl=32, seedkmax=2, totalkmax=4, scores=True, ll=ll))
    owner = (hits["pos"].astype(np.int64) * 2654435761 + hits["inverted"].astype(np.int64) * 40503 + hits["patid"].astype(np.int64)) >> 7
    return hits, owner


def test_merge_match_all_restores_single_handle_order():
    """matcher.merge_match_all: rows scattered over shards come back in the order one handle returns them"""
    hits, owner = _all_case()
    assert len(hits) > 1200 and (np.bincount(hits["patid"].astype(np.int64)) > 1).sum() > 30
    parts = [hits[(owner % 3) == r] for r in (2, 0, 1)]
    parts.append(hits[5:9])                           # rows reported twice are dropped like unifyMatches' std::unique does
    got = matcher.merge_match_all(parts)
    assert got.dtype == hits.dtype and got.tobytes() == hits.tobytes()
    assert len(matcher.merge_match_all([hits[:0], hits[:0]])) == 0


def _gather_worker(rank, world, port, parts, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = rdist.gather_match_all(parts[rank], dst=0)
        if rank == 0:
            np.save(out_path, got)
        else:
            assert got is None
        empty = rdist.gather_match_all(parts[rank][:0], dst=1)         # nobody found anything
        assert (empty is None) if rank != 1 else (len(empty) == 0)
    finally:
        dist.destroy_process_group()


def test_gather_match_all_world2_gloo(tmp_path):
    """real_b200.dist.gather_match_all: the rows of two ranks (one of them may hold none) merged on rank 0"""
    hits, owner = _all_case()
    for parts in ([hits[(owner % 2) == 0], hits[(owner % 2) == 1]], [hits[:0], hits]):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        out = str(tmp_path / "merged.npy")
        mp.spawn(_gather_worker, args=(2, port, parts, out), nprocs=2, join=True)
        assert np.load(out).tobytes() == hits.tobytes()
