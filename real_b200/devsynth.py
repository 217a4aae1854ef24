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
