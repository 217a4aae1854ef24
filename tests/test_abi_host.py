"""CPU: the C-ABI library loads and exports every symbol include/real_gpu.h declares (no compute
calls without a GPU), it refuses to work without a device, and the host-side mirror of the
reference's option handling behaves like RealOptions.cpp."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from real_b200 import lib as rlib
from real_b200 import matcher


def test_header_symbols_exported():
    L = rlib.load()
    names = rlib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", rlib.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert L.real_gpu_abi_version() == 1


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", rlib.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    archs = {tok for l in out.splitlines() for tok in l.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_create_rejects_bad_options():
    L = rlib.load()
    h = C.c_void_p()
    for kw in (dict(seedl=30), dict(seedl=0), dict(seedl=68), dict(seedkmax=3), dict(totalkmax=16)):
        a = dict(seedl=32, seedkmax=2, totalkmax=5)
        a.update(kw)
        P = rlib.Params(C.sizeof(rlib.Params), 0, a["seedl"], a["seedkmax"], a["totalkmax"], 0, 0.0, None, 0, 0)
        assert L.real_gpu_create(C.byref(P), C.byref(h)) == rlib.REAL_GPU_E_ARG
        assert not h.value
    P = rlib.Params(4, 0, 32, 2, 5, 0, 0.0, None, 0, 0)     # wrong struct_size
    assert L.real_gpu_create(C.byref(P), C.byref(h)) == rlib.REAL_GPU_E_ARG
    assert L.real_gpu_create(None, C.byref(h)) == rlib.REAL_GPU_E_ARG


def test_no_cpu_fallback():
    """Without a CUDA device the library must fail loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rlib.RealGpuError) as e:
        rlib.Handle()
    assert e.value.code == rlib.REAL_GPU_E_CUDA


def test_null_handle_calls():
    L = rlib.load()
    assert L.real_gpu_match_unique(None) == rlib.REAL_GPU_E_ARG
    assert L.real_gpu_destroy(None) == 0
    assert L.real_gpu_last_error(None) == b"null handle"


def test_real_options_parse():
    o = matcher.RealOptions.parse(["-t", "t.fa", "-p", "r.fq", "-o", "out", "-e", "99", "-s", "7", "-l", "70", "-u", "0", "--bogus"])
    assert (o.totalkmax, o.seedkmax, o.seedl, o.match_unique) == (15, 2, 64, False)
    assert any("Ignoring argument --bogus" in w for w in o.warnings)
    o = matcher.RealOptions.parse(["-t", "t", "-p", "p", "-o", "o", "-l", "30", "-filter_level", "9"])
    assert o.seedl == 28 and o.filter_level == 4
    assert o.filter_mult == pytest.approx(3.0 * 5 / 70.0)
    with pytest.raises(RuntimeError):
        matcher.RealOptions.parse(["-p", "p", "-o", "o"])
    with pytest.raises(RuntimeError):
        matcher.RealOptions.parse(["-t", "t", "-p", "p", "-o"])
    with pytest.raises(RuntimeError):
        matcher.RealOptions.parse(["-t", "t", "-p", "p", "-o", "o", "-l", "3"])


def test_shard_ranges_cover_and_halo():
    for n, g, L in ((1000003, 8, 100), (64, 3, 36), (250_000_000, 4, 250), (500, 8, 100)):
        sh = matcher.shard_ranges(n, g, L)
        assert len(sh) == g
        assert sh[0][0] == 0 and sh[-1][1] == n
        for (ob, oe, sb, sl), nxt in zip(sh, sh[1:] + [None]):
            assert ob % 64 == 0 and sb % 64 == 0 and sb <= ob <= oe <= n
            assert sb + sl >= min(n, oe + L)
            assert sb + sl <= n
            if nxt is not None:
                assert nxt[0] == oe


def test_stats_struct_mirrors_header():
    """real_b200.lib.Stats (ctypes) lists the fields of real_gpu_stats in the header's order and types: the struct is copied
    whole by real_gpu_get_stats, a drift would shift every later field."""
    import re
    src = open(rlib.HEADER_PATH).read()
    body = re.search(r"typedef struct\s*\{([^}]*)\}\s*real_gpu_stats;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(float|uint32_t|uint64_t)\s+([a-z_0-9]+)\s*;", body)
    ctype = {"float": C.c_float, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64}
    assert [(n, ctype[t]) for t, n in fields] == list(rlib.Stats._fields_)
    assert C.sizeof(rlib.Stats) % 8 == 0
