"""Multi-GPU layer of the matching path: one process per GPU.

Default form (connect_sharded_tables): the signature tables are sharded -- every rank indexes 1/N of the scan
buckets, partitions 1/N of the text positions of every round, and its partition kernel stores the window records
straight into the windows of the bucket owners (peer memory over NVLink, include/real_gpu.h "sharded tables").
torch.distributed only carries the 64-byte window handles at set-up and the matchUnique fold below.  The older
form (text sharded with a read-length halo, read index replicated on every rank) needs no set-up at all.

matchAll needs no exchange (every shard reports the hits that START in its own range).  matchUnique
has exactly one exchange step per text file, the cross-shard fold of the per-read UniqueMatchInfo
(the reference folds hits into `uniqueinfo[patID]` inside one process,
matchUniqueImplementation.cpp:1097,1282): a MIN all-reduce of order-preserving 63-bit keys followed
by a SUM all-reduce of tie flags (include/real_gpu.h, "cross-shard exchange").  torch.distributed is
the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests); the key transforms are device
kernels behind the C ABI.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

UNIQUE_KEY_NONE = 0x7FFFFFFFFFFFFFFF


class HandleShard:
    """Adapter: one real_gpu handle as a participant of the exchange."""

    def __init__(self, handle, nreads: int):
        self.handle = handle
        self.nreads = nreads
        self.device = torch.device("cuda", handle.device)

    def export_keys(self, keys: torch.Tensor) -> None:
        self.handle.unique_export_keys(keys.data_ptr())

    def export_ties(self, min_keys: torch.Tensor, ties: torch.Tensor) -> None:
        self.handle.unique_export_ties(min_keys.data_ptr(), ties.data_ptr())

    def import_merged(self, min_keys: torch.Tensor, tie_sums: torch.Tensor) -> None:
        self.handle.unique_import(min_keys.data_ptr(), tie_sums.data_ptr())

    def sync(self) -> None:
        torch.cuda.synchronize(self.device)


def unique_exchange(shard, group: Optional[dist.ProcessGroup] = None, keys: Optional[torch.Tensor] = None,
                    ties: Optional[torch.Tensor] = None) -> None:
    """Folds the per-read unique state of all ranks into the state every rank would hold had it
    scanned the whole text.  `shard` provides export_keys / export_ties / import_merged / sync and
    the attributes nreads, device (HandleShard on GPUs)."""
    n = shard.nreads
    if keys is None:
        keys = torch.empty(n, dtype=torch.int64, device=shard.device)
    if ties is None:
        ties = torch.empty(n, dtype=torch.uint8, device=shard.device)
    shard.export_keys(keys)                                   # returns after the kernel has finished
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
        shard.sync()
    shard.export_ties(keys, ties)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(ties, op=dist.ReduceOp.SUM, group=group)
        shard.sync()
    shard.import_merged(keys, ties)


def connect_fold(handle, device: torch.device, max_reads: int, group: Optional[dist.ProcessGroup] = None) -> None:
    """Sets up the peer-memory fold of the unique state (include/real_gpu.h, real_gpu_fold_*) between the ranks of `group`:
    every rank allocates its window, the 64-byte CUDA IPC handles travel over torch.distributed, the peers' windows are
    mapped.  Afterwards fold_unique(handle) replaces unique_exchange: ONE exchange over NVLink peer stores instead of two
    NCCL all-reduces, in reduce-scatter form -- rank r ends up with the merged words of the reads [R r / N, R (r+1) / N)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = handle.fold_init(rank, world, max_reads)
    t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    handle.fold_connect(b"".join(bytes(p.cpu().tolist()) for p in parts))
    dist.barrier(group=group)


def own_read_range(nreads: int, rank: int, world: int):
    """The reads whose merged state rank `rank` holds after fold_unique (real_gpu_fold_unique)."""
    return (nreads * rank) // world, (nreads * (rank + 1)) // world


def fold_reduce_scatter_reference(words_by_rank, rank: int):
    """numpy restatement of real_gpu_fold_unique for the CPU tests: the merged UniqueMatchInfo words of rank `rank`'s own
    reads from the state words of all ranks (k_fold_merge, csrc/post.cuh; the rule of unique_exchange in one pass)."""
    import numpy as np
    from . import matcher as _m
    world = len(words_by_rank)
    n = len(words_by_rank[0])
    lo, hi = own_read_range(n, rank, world)
    keys = np.stack([_m.unique_key(np.asarray(w[lo:hi], dtype=np.uint64)) for w in words_by_rank])
    win = keys.min(axis=0)
    same = (((keys ^ win[None, :]) >> 17) & ((1 << 41) - 1)) == 0
    ties = ((keys != UNIQUE_KEY_NONE) & ((keys >> 59) == (win >> 59)[None, :]) & ~same).sum(axis=0)
    out = np.asarray(words_by_rank[rank][lo:hi], dtype=np.uint64).copy()
    hit = win != UNIQUE_KEY_NONE
    out[hit] = _m.unique_word_from_key(win, ties)[hit]
    return out


def connect_sharded_tables(handle, device: torch.device, round_positions: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """Puts `handle` (this rank's real_gpu handle) into sharded-table mode with all ranks of `group`: allocates the
    rank's window, exchanges the CUDA IPC handles (64 bytes per rank) and maps the peers' windows.  Call before
    set_reads; afterwards every rank is given the whole text and the whole read set."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = handle.comm_init(rank, world, round_positions)
    t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    handle.comm_connect(b"".join(bytes(p.cpu().tolist()) for p in parts))
    dist.barrier(group=group)


class ShardedUpload:
    """Host -> all GPUs for inputs that every rank needs in full (bucket shards: the whole read set and the whole text).
    The byte sections are laid out back to back in one pinned host blob; every step each rank copies ITS 1/N slice of the
    blob to its GPU and one NCCL all-gather over NVLink completes the blob on every GPU -- N times less PCIe traffic per
    GPU than N full uploads.  run() returns device views of the sections (valid until the next run())."""

    ALIGN = 256

    def __init__(self, sections, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.offsets = {}
        total = 0
        for name, t in sections:
            self.offsets[name] = (total, t.numel())
            total += (t.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        unit = self.world * self.ALIGN
        total = (total + unit - 1) // unit * unit
        self.chunk = total // self.world
        self.blob = torch.zeros(total, dtype=torch.uint8)
        if device.type == "cuda":
            self.blob = self.blob.pin_memory()
        for name, t in sections:
            o, n = self.offsets[name]
            self.blob[o:o + n].copy_(t.reshape(-1).view(torch.uint8))
        self.d_in = torch.empty(self.chunk, dtype=torch.uint8, device=device)
        self.d_all = torch.empty(total, dtype=torch.uint8, device=device)

    def run(self):
        lo = self.rank * self.chunk
        self.d_in.copy_(self.blob[lo:lo + self.chunk], non_blocking=True)
        dist.all_gather_into_tensor(self.d_all, self.d_in, group=self.group)
        if self.d_all.device.type == "cuda":
            torch.cuda.current_stream(self.d_all.device).synchronize()
        return {name: self.d_all[o:o + n] for name, (o, n) in self.offsets.items()}


def gather_match_all(hits, dst: int = 0, group: Optional[dist.ProcessGroup] = None):
    """matchAll of a job sharded over the ranks of `group`: every rank passes the rows its handle returned, rank `dst`
    gets them merged in single-handle order (matcher.merge_match_all), the others None.  The rows travel as bytes over
    the group's backend (NCCL: through device memory; gloo: host) -- an all-gather of variable-length pieces, no reduction."""
    import numpy as np
    from . import lib as _lib
    from . import matcher as _matcher
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    raw = np.ascontiguousarray(np.asarray(hits, dtype=_lib.HIT_DTYPE)).view(np.uint8).reshape(-1)
    n = torch.tensor([raw.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if raw.size:
        buf[:raw.size] = torch.from_numpy(raw.copy()).to(dev)
    pieces = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(pieces, buf, group=group)
    if rank != dst:
        return None
    parts = [p[:sz].cpu().numpy().view(_lib.HIT_DTYPE) for p, sz in zip(pieces, sizes) if sz]
    return _matcher.merge_match_all(parts)
