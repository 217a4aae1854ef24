"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/liboracle.so and oracle/_ref/ref_harness.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (real_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")
REF_BOUND = os.path.join(HERE, "_ref", "real_bound")

HIT_DTYPE = np.dtype([("patid", "<u8"), ("pos", "<u8"), ("file", "<u4"), ("frag", "<u4"),
                      ("k", "<u4"), ("inverted", "<u4"), ("score", "<f4"), ("block", "<u4")])
UNIQUE_DTYPE = np.dtype([("data", "<u8"), ("score", "<f4"), ("pad", "<u4")])
GAP_DTYPE = np.dtype([("patid", "<u4"), ("mingap", "<u4"), ("where", "<u4"), ("start", "<u4"),
                      ("gap_pos", "<u4"), ("present", "<u4")])


class Params(C.Structure):
    _fields_ = [("seedl", C.c_uint32), ("seedkmax", C.c_uint32), ("totalkmax", C.c_uint32), ("scores", C.c_uint32),
                ("filter_mult", C.c_double), ("ll", C.c_void_p), ("n_list", C.c_uint64)]


class TextS(C.Structure):
    _fields_ = [("words", C.c_void_p), ("nmask", C.c_void_p), ("n", C.c_uint64),
                ("record_starts", C.c_void_p), ("nrecords", C.c_uint32), ("fileid", C.c_uint32)]


class ReadsS(C.Structure):
    _fields_ = [("mapped", C.c_void_p), ("quality", C.c_void_p), ("offsets", C.c_void_p), ("nreads", C.c_uint64)]


def build(force: bool = False) -> None:
    """Compile liboracle.so (and, where /root/reference exists, oracle/_ref/ref_harness)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "real_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])
        # the reference binary with its matching loops bound to libreal_gpu.so; needs the library
        if os.path.exists(os.path.join(HERE, "..", "real_b200", "libreal_gpu.so")):
            subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref_bound"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_match_all.restype = C.c_int64
        L.oracle_match_all.argtypes = [C.POINTER(Params), C.POINTER(TextS), C.POINTER(ReadsS), C.c_void_p, C.c_uint64]
        L.oracle_match_unique.restype = C.c_int
        L.oracle_match_unique.argtypes = [C.POINTER(Params), C.POINTER(TextS), C.POINTER(ReadsS), C.c_void_p, C.c_void_p]
        L.oracle_match_gaps.restype = C.c_int
        L.oracle_match_gaps.argtypes = [C.POINTER(Params), C.POINTER(TextS), C.POINTER(ReadsS), C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_match_gaps_flagged.restype = C.c_int
        L.oracle_match_gaps_flagged.argtypes = [C.POINTER(Params), C.POINTER(TextS), C.POINTER(ReadsS), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_unique_init.restype = None
        L.oracle_unique_init.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p]
        L.oracle_build_ll.restype = None
        L.oracle_build_ll.argtypes = [C.c_double] * 5 + [C.c_void_p]
        L.oracle_fragments.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_reverse_fragments.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_pair_signatures.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_rest_words.restype = C.c_uint32
        L.oracle_rest_words.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.oracle_diffcount64.restype = C.c_uint32
        L.oracle_diffcount64.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_text_word.restype = C.c_uint64
        L.oracle_text_word.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.oracle_dontcare_free.restype = C.c_int
        L.oracle_dontcare_free.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.oracle_position_to_range.restype = C.c_uint32
        L.oracle_position_to_range.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.oracle_position_valid.restype = C.c_int
        L.oracle_position_valid.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32]
        L.oracle_rest_distance.restype = C.c_uint32
        L.oracle_rest_distance.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.oracle_compute_score.restype = C.c_float
        L.oracle_compute_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int]
        L.oracle_count_windows.restype = C.c_uint64
        L.oracle_count_windows.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


DEFAULT_SCORING = (0.995, 0.41, 0.71, 0.00, 2.0)  # similarity, gc, trans, err, gcmut_bias (Scoring.cpp:204-208)


def build_ll(similarity=0.995, gc=0.41, trans=0.71, err=0.0, gcmut_bias=2.0) -> np.ndarray:
    ll = np.zeros(1024, dtype=np.float64)
    lib().oracle_build_ll(similarity, gc, trans, err, gcmut_bias, ll.ctypes.data)
    return ll


def filter_mult(totalkmax: int, filter_level: int = 2) -> float:
    """RealOptions.cpp:455-463"""
    m = {1: 0.5, 2: 1.0, 3: 2.0, 4: 3.0}.get(filter_level, 0.0) * totalkmax
    return m / 70.0


class _Keep:
    """Holds numpy arrays alive next to the ctypes structs that point into them."""

    def __init__(self):
        self.refs = []

    def ptr(self, a: Optional[np.ndarray], dtype) -> Optional[int]:
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        self.refs.append(a)
        return a.ctypes.data


def _mk(keep: _Keep, seedl, seedkmax, totalkmax, scores, fmult, ll, n_list, text, reads, fileid):
    P = Params(seedl, seedkmax, totalkmax, 1 if scores else 0, fmult, keep.ptr(ll, np.float64), n_list)
    words, nmask = text.packed()
    # one spare word so unaligned extracts at the very end stay in bounds
    words = np.concatenate([words, np.zeros(2, np.uint64)])
    nmask = np.concatenate([nmask, np.zeros(2, np.uint64)])
    rs = text.record_starts
    T = TextS(keep.ptr(words, np.uint64), keep.ptr(nmask, np.uint64), text.n, keep.ptr(rs, np.uint64), len(text.records), fileid)
    R = ReadsS(keep.ptr(reads.mapped, np.uint8), keep.ptr(reads.quality, np.uint8) if reads.quality is not None else None,
               keep.ptr(reads.offsets, np.uint64), reads.nreads)
    return P, T, R


def match_all(text, reads, seedl=32, seedkmax=2, totalkmax=5, scores=False, ll=None, n_list=0, fileid=0,
              filter_level=2) -> np.ndarray:
    keep = _Keep()
    if scores and ll is None:
        ll = build_ll()
    P, T, R = _mk(keep, seedl, seedkmax, totalkmax, scores, filter_mult(totalkmax, filter_level), ll, n_list, text, reads, fileid)
    cap = max(1024, 8 * reads.nreads)
    while True:
        out = np.zeros(cap, dtype=HIT_DTYPE)
        n = lib().oracle_match_all(C.byref(P), C.byref(T), C.byref(R), out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("oracle_match_all failed")
        if n <= cap:
            return out[:n]
        cap = int(n)


def unique_init(nreads: int, scores: bool) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    info = np.zeros(nreads, dtype=np.uint64)
    sc = np.zeros(nreads, dtype=np.float32) if scores else None
    lib().oracle_unique_init(nreads, info.ctypes.data, sc.ctypes.data if sc is not None else None)
    return info, sc


def match_unique(text, reads, info, score, seedl=32, seedkmax=2, totalkmax=5, scores=False, ll=None, n_list=0,
                 fileid=0, filter_level=2) -> None:
    keep = _Keep()
    if scores and ll is None:
        ll = build_ll()
    P, T, R = _mk(keep, seedl, seedkmax, totalkmax, scores, filter_mult(totalkmax, filter_level), ll, n_list, text, reads, fileid)
    r = lib().oracle_match_unique(C.byref(P), C.byref(T), C.byref(R), info.ctypes.data, score.ctypes.data if score is not None else None)
    if r != 0:
        raise RuntimeError("oracle_match_unique failed")


def match_gaps(text, reads, info, score, gaps, seedl=32, seedkmax=2, totalkmax=5, scores=True, ll=None, n_list=0,
               fileid=0, filter_level=2, undefined=None) -> None:
    """undefined (uint8 per read, optional): set where the reference's own result is not defined -- a candidate whose band
    never reaches MINscore leaves ::matchGaps' uninitialised locals unassigned (match.hpp:518-522)."""
    keep = _Keep()
    if ll is None:
        ll = build_ll()
    P, T, R = _mk(keep, seedl, seedkmax, totalkmax, scores, filter_mult(totalkmax, filter_level), ll, n_list, text, reads, fileid)
    r = lib().oracle_match_gaps_flagged(C.byref(P), C.byref(T), C.byref(R), info.ctypes.data,
                                        score.ctypes.data if score is not None else None, gaps.ctypes.data,
                                        undefined.ctypes.data if undefined is not None else None)
    if r != 0:
        raise RuntimeError("oracle_match_gaps failed")


# ------------------------------------------------------------------ the real reference, through the harness

def have_ref() -> bool:
    return os.path.exists(REF_HARNESS)


def run_ref(mode: str, workdir: str, real_args, gap_dump: bool = False, env=None, threads: Optional[int] = None):
    """Runs oracle/_ref/ref_harness; returns (timing dict, dump path, gap dump path or None)."""
    dump = os.path.join(workdir, "ref_%s.bin" % mode)
    cmd = [REF_HARNESS, mode, dump]
    gpath = None
    if gap_dump:
        gpath = os.path.join(workdir, "ref_gaps.bin")
        cmd.append(gpath)
    cmd.append("--")
    cmd += [str(a) for a in real_args]
    e = dict(os.environ)
    if env:
        e.update(env)
    if threads is not None:
        e["OMP_NUM_THREADS"] = str(threads)
    p = subprocess.run(cmd, cwd=workdir, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        raise RuntimeError("ref_harness failed: %s\n%s" % (" ".join(cmd), p.stderr[-2000:]))
    timing = json.loads(p.stdout.strip().splitlines()[-1])
    # text files in the order the reference visited them (getFileList.cpp uses raw readdir order)
    order = []
    for tok in p.stderr.split("Computing length of file ")[1:]:
        name = tok.split("...")[0]
        if name not in order:
            order.append(name)
    timing["file_order"] = order
    return timing, dump, gpath


# ------------------------------------------------------------------ text loader (restatement)

def fasta_text(data: bytes):
    """The reference's text loader restated: countLength + readFile (countReads.cpp:28-125).
    Returns (symbols uint8 0..4, record names (bytes), record starts incl. the terminal entry).
    '>' opens a header wherever it stands and clears the name; '\\n' closes a header (filing the record at the base count of
    the '>') ; inside a header every other byte joins the name; outside, A C G T N are kept and everything else is dropped.
    Pinned against the reference's own getText by tests/golden/text_quirks.npz (tests/test_oracle_golden.py)."""
    code = {65: 0, 67: 1, 71: 2, 84: 3, 78: 4}
    syms = bytearray()
    names, starts = [], []
    header = False
    name = bytearray()
    idcnt = 0
    for c in data:
        if c == 62:            # '>'
            header = True
            idcnt = len(syms)
            name = bytearray()
        elif c == 10:          # '\n'
            if header:
                names.append(bytes(name))
                starts.append(idcnt)
            header = False
        elif header:
            name.append(c)
        elif c in code:
            syms.append(code[c])
    starts.append(len(syms))
    return np.frombuffer(bytes(syms), dtype=np.uint8), names, np.asarray(starts, dtype=np.uint64)
