"""GPU: K0, the FASTA text loader on the device (real_gpu_set_text_fasta, csrc/ingest.cuh), through the C ABI against
 (1) the reference's own getText on awkward files (tests/golden/text_quirks.npz),
 (2) the restated loader (oracle_py.fasta_text) on fuzzed byte strings of tile-boundary sizes,
 (3) the packed-text path: matching after set_text_fasta == matching after set_text on the same file."""
import io

import numpy as np
import pytest

from oracle import oracle_py as O
from real_b200 import lib as rlib
from real_b200 import matcher, synth
from util import canon_hits, text_quirk_cases

pytestmark = pytest.mark.gpu


def _check(m, data, symbols, starts, names):
    n, ranges = m.set_text_fasta(data)
    assert n == symbols.size
    assert m.handle._text_nrec == len(names)
    if n and len(names):
        assert [r[1] for r in ranges] == [int(x) for x in starts]
        assert [r[0] for r in ranges[:-1]] == list(names)
        words, nmask = m.handle.get_text_packed(n)
        ew, em = synth.pack_text(symbols)
        assert np.array_equal(words, ew)
        assert np.array_equal(nmask, em)
    else:
        with pytest.raises(rlib.RealGpuError):      # no text was set
            m.handle.get_text_packed(1)


@pytest.fixture(scope="module")
def m():
    mm = matcher.AllMatcher(matcher.RealOptions(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    yield mm
    mm.close()


def test_reference_quirk_fixture(m):
    for name, data, symbols, starts, names in text_quirk_cases():
        _check(m, data, symbols, starts, names)


@pytest.mark.parametrize("nbytes", [0, 1, 15, 16, 17, 4095, 4096, 4097, 8192, 70001, 1 << 20])
@pytest.mark.parametrize("alphabet", [b"ACGT", b"ACGTN" * 8 + b"acgtn>\n\r x", b"ACGT" * 30 + b">\n\n", b">\n"])
def test_fuzz_against_restatement(m, nbytes, alphabet):
    rng = np.random.RandomState(nbytes * 7 + len(alphabet))
    a = np.frombuffer(alphabet, dtype=np.uint8)
    data = a[rng.randint(0, a.size, nbytes)].tobytes()
    symbols, names, starts = O.fasta_text(data)
    _check(m, data, symbols, starts, names)


def test_long_headers_across_tiles(m):
    rng = np.random.RandomState(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    parts = []
    for i in range(12):
        parts.append(b">" + bytes(rng.randint(32, 127, rng.randint(1, 9000)).astype(np.uint8)) + b"\n")
        parts.append(acgt[rng.randint(0, 4, rng.randint(0, 9000))].tobytes() + (b"\n" if i % 3 else b""))
    data = b"".join(parts)
    symbols, names, starts = O.fasta_text(data)
    _check(m, data, symbols, starts, names)


def test_device_buffer(m):
    import torch
    t = synth.make_text(31, 300000, nrecords=5, n_per_million=2000)
    f = io.BytesIO()
    synth.write_fasta_file(f, t)
    data = f.getvalue()
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    n, nrec = m.handle.set_text_fasta(None, device_ptr=d.data_ptr(), nbytes=len(data))
    assert (n, nrec) == (t.n, 5)
    words, nmask = m.handle.get_text_packed(n)
    ew, em = t.packed()
    assert np.array_equal(words, ew) and np.array_equal(nmask, em)
    starts, _ = m.handle.get_text_records()
    assert np.array_equal(starts, t.record_starts)


def test_matching_after_fasta_ingest_equals_packed_path():
    t = synth.make_text(77, 400000, nrecords=3, n_per_million=1500)
    reads = synth.make_reads(t, 78, 20000, 100, 0.01, False)
    f = io.BytesIO()
    synth.write_fasta_file(f, t)
    mm = matcher.AllMatcher(matcher.RealOptions(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    try:
        mm.set_reads(reads.mapped, reads.offsets, reads.quality)
        words, nmask = t.packed()
        mm.set_text(words, nmask, t.n, t.record_starts)
        a = mm.match()
        n, ranges = mm.set_text_fasta(f.getvalue())
        assert n == t.n and [r[1] for r in ranges] == [int(x) for x in t.record_starts]
        assert [r[0].decode() for r in ranges[:-1]] == [name for name, _ in t.records]
        b = mm.match()
    finally:
        mm.close()
    assert len(a) > 15000
    assert np.array_equal(canon_hits(a), canon_hits(b))


def test_argument_errors(m):
    import ctypes as C
    import torch
    L = m.handle.L
    n, r = C.c_uint64(), C.c_uint64()
    data = np.frombuffer(b">a\nACGT\n", dtype=np.uint8)
    assert L.real_gpu_set_text_fasta(m.handle.h, 0, None, 8, C.byref(n), C.byref(r)) == -1                 # REAL_GPU_E_ARG: null pointer
    assert L.real_gpu_set_text_fasta(m.handle.h, 0, data.ctypes.data, 8, None, C.byref(r)) == -1
    assert L.real_gpu_set_text_fasta(m.handle.h, 64, data.ctypes.data, 8, C.byref(n), C.byref(r)) == -4    # REAL_GPU_E_LIMIT: fileid over UniqueMatchInfo's 6 bits
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    assert L.real_gpu_set_text_fasta_device(m.handle.h, 0, d.data_ptr() + 1, 8, C.byref(n), C.byref(r)) == -1   # unaligned device buffer
    assert m.set_text_fasta(b">a\nACGT\n")[0] == 4
    starts, ends = m.handle.get_text_records()
    assert list(starts) == [0, 4] and list(ends) == [2]
    words, nmask = m.handle.get_text_packed(4)
    assert int(words[0]) == 0x1B << 56 and int(nmask[0]) == 0
    with pytest.raises(rlib.RealGpuError):          # buffers sized for another length are refused, not overrun
        m.handle.get_text_packed(5)
    # the record table belongs to the fasta loader: a packed text replaces it
    w, k = synth.pack_text(np.asarray([0, 1, 2, 3], dtype=np.uint8))
    m.set_text(w, k, 4, np.asarray([0, 4], dtype=np.uint64))
    with pytest.raises(rlib.RealGpuError):
        m.handle.get_text_records()
