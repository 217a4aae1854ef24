"""Benchmark-scale synthetic inputs generated directly in device memory (same counter-based
formulas as real_b200/synth.py, run by the generator kernels behind real_gpu_synth_text /
real_gpu_synth_reads).  torch only owns the device buffers."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import lib as _lib


def text_device(seed: int, n: int, device: int = 0, n_per_million: int = 0, first_word: int = 0,
                nwords: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Packed text words [first_word, first_word+nwords) and the matching N-mask words (int64 tensors
    holding the u64 bit patterns).  Default: the whole text of n bases."""
    L = _lib.load()
    if nwords is None:
        nwords = (n + 31) // 32 - first_word
    assert first_word % 2 == 0
    dev = torch.device("cuda", device)
    words = torch.zeros(nwords + 2, dtype=torch.int64, device=dev)
    nmask = torch.zeros((nwords + 1) // 2 + 2, dtype=torch.int64, device=dev)
    rc = L.real_gpu_synth_text(device, seed, first_word, nwords, n_per_million, words.data_ptr(), nmask.data_ptr())
    if rc != 0:
        raise _lib.RealGpuError(rc, "real_gpu_synth_text")
    # bases beyond n in the last word are zero in the host generator (pack_text pads with A)
    last = n - 32 * first_word
    if 0 < last <= nwords * 32 and last % 32:
        w = last // 32
        keep = -1 << (64 - 2 * (last % 32))
        words[w] = words[w] & keep
        words[w + 1:] = 0
    if 0 < last <= nwords * 32 and last % 64:
        m = last // 64
        nmask[m] = nmask[m] & (-1 << (64 - last % 64))
        nmask[m + 1:] = 0
    return words, nmask


def reads_device(seed: int, words: torch.Tensor, nmask: Optional[torch.Tensor], text_n: int, total: int, length: int,
                 sub_rate: float, device: int = 0, first: int = 0, count: Optional[int] = None,
                 quality: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """Reads [first, first+count) of a set of `total`: (mapped u8, quality u8 or None, offsets i64)."""
    L = _lib.load()
    count = total - first if count is None else count
    dev = torch.device("cuda", device)
    mapped = torch.empty(count * length, dtype=torch.uint8, device=dev)
    qual = torch.empty(count * length, dtype=torch.uint8, device=dev) if quality else None
    thr = int(round(sub_rate * 16384))
    rc = L.real_gpu_synth_reads(device, seed, words.data_ptr(), nmask.data_ptr() if nmask is not None else None, text_n,
                                total, first, count, length, thr, mapped.data_ptr(), qual.data_ptr() if quality else None)
    if rc != 0:
        raise _lib.RealGpuError(rc, "real_gpu_synth_reads")
    offsets = torch.arange(count + 1, dtype=torch.int64, device=dev) * length
    return mapped, qual, offsets


def unpack_symbols(words: torch.Tensor, nmask: torch.Tensor, n: int) -> torch.Tensor:
    """2-bit text + N mask (int64 tensors with the u64 bit patterns) -> uint8 symbols 0..3, 4 = N (checkers, planted indels)."""
    dev = words.device
    sym = torch.empty(n, dtype=torch.uint8, device=dev)
    sh2 = torch.arange(62, -2, -2, device=dev, dtype=torch.int64)
    sh1 = torch.arange(63, -1, -1, device=dev, dtype=torch.int64)
    step = 1 << 23                                            # words per slice
    nw = (n + 31) // 32
    for w0 in range(0, nw, step):
        w1 = min(nw, w0 + step)
        s = ((words[w0:w1, None] >> sh2[None, :]) & 3).to(torch.uint8).reshape(-1)
        m0, m1 = w0 // 2, (w1 + 1) // 2
        mk = ((nmask[m0:m1, None] >> sh1[None, :]) & 1).to(torch.uint8).reshape(-1)[: s.numel()]
        s = torch.where(mk != 0, torch.full_like(s, 4), s)
        lo, hi = w0 * 32, min(n, w1 * 32)
        sym[lo:hi] = s[: hi - lo]
    return sym


def read_plan_device(seed: int, nreads: int, span: int, dev: torch.device):
    """Planted position and strand of every read (synth.read_plan_total), as device tensors."""
    import numpy as np
    from . import synth
    pos, strand = synth.read_plan_total(seed, nreads, span, 0, nreads)
    return torch.as_tensor(pos.astype(np.int64), device=dev), torch.as_tensor(strand.astype(bool), device=dev)


def plant_deletions(mapped: torch.Tensor, sym: torch.Tensor, ppos: torch.Tensor, pstrand: torch.Tensor, n: int, L: int) -> torch.Tensor:
    """C4: 30 % of the '+' strand reads get one deletion of 1..3 bases behind the seed -- read[off+g:] moves up
    and the tail is refilled from the text behind the window.  In place on `mapped` (R*L bytes); returns the planted mask."""
    R = mapped.numel() // L
    dev = mapped.device
    rid = torch.arange(R, device=dev)
    g = 1 + rid % 3
    off = 60 + rid % 50
    planted = (rid % 10 < 3) & ~pstrand & (ppos + L + 3 < n)
    m2 = mapped.view(R, L)
    col = torch.arange(L, device=dev)[None, :]
    for r0 in range(0, R, 1 << 20):
        r1 = min(R, r0 + (1 << 20))
        sel = planted[r0:r1]
        src_col = torch.where(col >= off[r0:r1, None], col + g[r0:r1, None], col)
        from_read = src_col < L
        a = torch.gather(m2[r0:r1], 1, src_col.clamp(max=L - 1))
        b = sym[ppos[r0:r1, None] + src_col]
        new = torch.where(from_read, a, b)
        m2[r0:r1] = torch.where(sel[:, None], new, m2[r0:r1])
    torch.cuda.synchronize()      # the library reads the device buffers on its own stream (include/real_gpu.h: the caller synchronizes)
    return planted
