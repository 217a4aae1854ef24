"""Host-side mirror of the reference interface for the matching path, on top of the C ABI.

Names and argument meaning follow the reference: `RealOptions` (RealOptions.hpp:29-77,
RealOptions.cpp:122-463), `Scoring` (Scoring.cpp:61-171), `AllMatcher` / `UniqueMatcher`
(matchAllImplementation.cpp:244-355, matchUniqueImplementation.cpp:348-500).  The per-read
`match(pattern, ...)` calls of the reference are batched: one call matches the whole read set
against the current text file or shard on the GPU.  Nothing here computes matches on the CPU.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import lib as _lib

# ---------------------------------------------------------------------------------------------
# options
# ---------------------------------------------------------------------------------------------

_Q_PRB = [
    1.0000000, 0.7943282, 0.6309573, 0.5011872, 0.3981072, 0.3162278, 0.2511886, 0.1995262, 0.1584893, 0.1258925,
    0.1000000, 0.0794328, 0.0630957, 0.0501187, 0.0398107, 0.0316228, 0.0251189, 0.0199526, 0.0158489, 0.0125893,
    0.0100000, 0.0079433, 0.0063096, 0.0050119, 0.0039811, 0.0031623, 0.0025119, 0.0019953, 0.0015849, 0.0012589,
    0.0010000, 0.0007943, 0.0006310, 0.0005012, 0.0003981, 0.0003162, 0.0002512, 0.0001995, 0.0001585, 0.0001259,
    0.0001000, 0.0000794, 0.0000631, 0.0000501, 0.0000398, 0.0000316, 0.0000251, 0.0000200, 0.0000158, 0.0000126,
    0.0000100, 0.0000079, 0.0000063, 0.0000050, 0.0000040, 0.0000032, 0.0000025, 0.0000020, 0.0000016, 0.0000013,
    0.0000010, 0.0000008, 0.0000006, 0.0000005, 0.0000004,
]


@dataclasses.dataclass
class RealOptions:
    """Same fields, defaults and clamps as the reference (RealOptions.hpp:29-77)."""
    textfilename: str = ""
    patternfilename: str = ""
    outputfilename: str = ""
    seedkmax: int = 2
    totalkmax: int = 5
    seedl: int = 32
    match_unique: bool = True
    fracmem: float = 0.75
    scores: bool = True
    qualityOffset: int = 0
    rewritepatterns: bool = True
    filter_level: int = 2
    similarity: float = 0.995
    err: float = 0.00
    trans: float = 0.71
    gc: float = 0.41
    gcmut_bias: float = 2.0
    gaps: bool = False
    threads: int = 0
    warnings: List[str] = dataclasses.field(default_factory=list)

    @property
    def filter_mult(self) -> float:
        """RealOptions.cpp:455-463"""
        m = {1: 0.5, 2: 1.0, 3: 2.0, 4: 3.0}.get(self.filter_level, 0.0) * self.totalkmax
        return m / 70.0

    def getFilterValue(self, patl: int) -> float:
        return self.filter_mult * patl

    @classmethod
    def parse(cls, argv: Sequence[str]) -> "RealOptions":
        """The hand-rolled argv loop of RealOptions.cpp:140-396: unknown arguments are ignored,
        a missing parameter raises like the reference's runtime_error."""
        o = cls()
        i = 0
        argv = list(argv)

        def need(flag):
            if i + 1 >= len(argv):
                raise RuntimeError("Parameter for argument %s is missing." % flag)
            return argv[i + 1]

        while i < len(argv):
            a = argv[i]
            if a == "-t": o.textfilename = need(a); i += 2
            elif a == "-p": o.patternfilename = need(a); i += 2
            elif a == "-o": o.outputfilename = need(a); i += 2
            elif a == "-s": o.seedkmax = _atoi(need(a)); i += 2
            elif a == "-e":
                o.totalkmax = _atoi(need(a)); i += 2
                if o.totalkmax > 15:
                    o.totalkmax = 15
                    o.warnings.append("Warning: reducing maximum amount of errors to 15")
            elif a == "-l": o.seedl = _atoi(need(a)); i += 2
            elif a == "-u": o.match_unique = bool(_atoi(need(a))); i += 2
            elif a == "-g": o.gaps = bool(_atoi(need(a))); i += 2
            elif a == "-R": o.rewritepatterns = bool(_atoi(need(a))); i += 2
            elif a in ("-m", "-f"): o.fracmem = float(need(a)); i += 2
            elif a == "-q": o.scores = bool(_atoi(need(a))); i += 2
            elif a == "-Q": o.qualityOffset = _atoi(need(a)); i += 2
            elif a == "-T":
                o.threads = _atoi(need(a)); i += 2
                if o.threads < 1:
                    raise RuntimeError("Argument for -T parameter is invalid (<1)")
            elif a == "-similarity": o.similarity = float(need(a)); i += 2
            elif a == "-err": o.err = float(need(a)); i += 2
            elif a == "-trans": o.trans = float(need(a)); i += 2
            elif a == "-gc": o.gc = float(need(a)); i += 2
            elif a == "-gcmut_bias": o.gcmut_bias = float(need(a)); i += 2
            elif a == "-filter_level":
                o.filter_level = min(4, max(0, _atoi(need(a)))); i += 2
            elif a == "-h":
                raise RuntimeError("Help requested.")
            else:
                o.warnings.append("Ignoring argument %s" % a)
                i += 1
        if not o.textfilename:
            raise RuntimeError("Mandatory argument -t (text file name) is not given.")
        if not o.patternfilename:
            raise RuntimeError("Mandatory argument -p (pattern file name) is not given.")
        if not o.outputfilename:
            raise RuntimeError("Mandatory argument -o (output file name) is not given.")
        o.fracmem = min(1.0, o.fracmem)
        if o.seedl > 64:
            o.seedl = 64
        if o.seedl % 4:
            o.seedl -= o.seedl % 4
        if o.seedl < 4:
            raise RuntimeError("cannot handle seed length < 4")
        if o.seedkmax > 2:
            o.seedkmax = 2
        return o


def _atoi(s: str) -> int:
    """C atoi: leading integer, 0 when there is none."""
    s = s.strip()
    j = 0
    if j < len(s) and s[j] in "+-":
        j += 1
    while j < len(s) and s[j].isdigit():
        j += 1
    try:
        return int(s[:j])
    except ValueError:
        return 0


def scoring_table(similarity: float = 0.995, gc: float = 0.41, trans: float = 0.71, err: float = 0.0,
                  gcmut_bias: float = 2.0) -> np.ndarray:
    """Scoring::init + Scoring::getScore(ref, read, q) (Scoring.cpp:61-133, 155-171): the 4x4x64 table
    LL[(ref<<8)|(read<<6)|q] = log2(odds[ref][read]) * (1 - 10^(-q/10)), same operation order in IEEE
    doubles."""
    odds = [[0.0] * 4 for _ in range(4)]
    transit = trans * (1 - similarity)
    transver = (1 - trans) * (1 - similarity)
    bg = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    bias = gcmut_bias * (1 - gc) / gc
    odds[0][2] = transit / (bias + 1) / (1 - gc)
    odds[3][1] = transit / (bias + 1) / (1 - gc)
    odds[2][0] = transit / (bias + 1) / gc * bias
    odds[1][3] = transit / (bias + 1) / gc * bias
    odds[0][1] = transver / 2 / (bias + 1) / (1 - gc)
    odds[3][2] = transver / 2 / (bias + 1) / (1 - gc)
    odds[0][3] = transver / 2 / (bias + 1) / (1 - gc)
    odds[3][0] = transver / 2 / (bias + 1) / (1 - gc)
    odds[1][0] = transver / 2 / (bias + 1) / gc * bias
    odds[2][3] = transver / 2 / (bias + 1) / gc * bias
    odds[1][2] = transver / 2 / (bias + 1) / gc * bias
    odds[2][1] = transver / 2 / (bias + 1) / gc * bias
    odds[0][0] = 1 - odds[0][1] - odds[0][2] - odds[0][3]
    odds[3][3] = 1 - odds[3][0] - odds[3][1] - odds[3][2]
    odds[2][2] = 1 - odds[2][0] - odds[2][1] - odds[2][3]
    odds[1][1] = 1 - odds[1][0] - odds[1][2] - odds[1][3]
    for x in range(4):
        for y in range(4):
            odds[x][y] *= 1 - err
            odds[x][y] /= bg[y]
    ll = np.zeros(1024, dtype=np.float64)
    log2 = math.log(2.0)
    for c0 in range(4):
        for c1 in range(4):
            for q in range(64):
                ll[(c0 << 8) | (c1 << 6) | q] = math.log(odds[c0][c1]) / log2 * (1 - _Q_PRB[q])
    return ll


# ---------------------------------------------------------------------------------------------
# sharding
# ---------------------------------------------------------------------------------------------

def shard_ranges(n_total: int, nshards: int, maxlen: int) -> List[Tuple[int, int, int, int]]:
    """Cuts [0,n_total) into nshards contiguous chunks on 64-base boundaries.
    Returns (own_begin, own_end, shard_begin, shard_len): a shard reports hits STARTING in
    [own_begin, own_end) and loads the text up to own_end + maxlen (read-length halo)."""
    out = []
    per = -(-n_total // nshards)
    per = -(-per // 64) * 64
    for s in range(nshards):
        ob = min(n_total, s * per)
        oe = min(n_total, (s + 1) * per)
        sb = ob
        se = min(n_total, oe + maxlen + 64)
        out.append((ob, oe, sb, se - sb))
    return out


# ---------------------------------------------------------------------------------------------
# matchers
# ---------------------------------------------------------------------------------------------

class MatcherBase:
    """Counterpart of MatcherBase (MatcherBase.hpp:14-35): owns the device handle with the read-side
    index and borrows the text of the current file."""

    def __init__(self, opts: RealOptions, device: int = 0, table_bits: int = 0, ll_table: Optional[np.ndarray] = None):
        self.opts = opts
        need_ll = opts.scores or opts.gaps
        if ll_table is None and need_ll:
            ll_table = scoring_table(opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias)
        self.ll = ll_table
        self.handle = _lib.Handle(seedl=opts.seedl, seedkmax=opts.seedkmax, totalkmax=opts.totalkmax, scores=opts.scores,
                                  filter_mult=opts.filter_mult, ll_table=ll_table, device=device, table_bits=table_bits)
        self.maxlen = 0

    def close(self):
        self.handle.close()

    def set_reads(self, mapped: np.ndarray, offsets: np.ndarray, quality: Optional[np.ndarray] = None):
        """Packs the reads and builds the signature index (K1+K2)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.maxlen = int(np.max(np.diff(offsets.astype(np.int64)))) if offsets.size > 1 else 0
        self.handle.set_reads(mapped, offsets, quality)

    def set_text(self, words: np.ndarray, nmask: np.ndarray, n_total: int, record_starts: np.ndarray, fileid: int = 0,
                 shard: Optional[Tuple[int, int, int, int]] = None):
        """Text of file `fileid`; with `shard` = (own_begin, own_end, shard_begin, shard_len) only that part."""
        if shard is None:
            self.handle.set_text(words, nmask, n_total, record_starts, fileid)
        else:
            ob, oe, sb, sl = shard
            w0, w1 = sb // 32, (sb + sl + 31) // 32
            m0, m1 = sb // 64, (sb + sl + 63) // 64
            self.handle.set_text(np.ascontiguousarray(words[w0:w1]), np.ascontiguousarray(nmask[m0:m1]), n_total, record_starts,
                                 fileid, shard_begin=sb, shard_len=sl, own_begin=ob, own_end=oe)

    def set_text_fasta(self, data, fileid: int = 0):
        """Text of file `fileid` from the bytes of its FASTA file, parsed and packed on the device (getText, getText.hpp:31-58).
        Returns (n_bases, [(record name, start)] + [("terminal", n)]) like countLength's `ranges` (countReads.cpp:28-85)."""
        n, nrec = self.handle.set_text_fasta(data, fileid)
        if not n or not nrec:
            return n, [(b"terminal", n)]
        starts, ends = self.handle.get_text_records()
        buf = bytes(data)
        ranges = [(buf[buf.rfind(b">", 0, int(e)) + 1:int(e)], int(s)) for s, e in zip(starts[:-1], ends)]
        return n, ranges + [(b"terminal", n)]

    def stats(self) -> dict:
        return self.handle.stats()


class AllMatcher(MatcherBase):
    """AllMatcher::match + unifyMatches for every read at once (matchAllImplementation.cpp:261-355,150-161)."""

    def match(self) -> np.ndarray:
        return self.handle.match_all()


class UniqueMatcher(MatcherBase):
    """UniqueMatcher::match for every read at once (matchUniqueImplementation.cpp:369-500); the per-read
    UniqueMatchInfo array lives on the device across files like `uniqueinfo` (:1097)."""

    def match(self):
        self.handle.match_unique()

    def info(self):
        return self.handle.get_unique()

    def reset(self):
        self.handle.reset_unique()

    def matchGaps(self, n_list: int = 0):
        """UniqueMatcher::matchGaps for every read still NoMatch/Gapped (matchUniqueImplementation.cpp:501-572)."""
        self.handle.match_gaps(n_list)

    def gaps(self):
        return self.handle.get_gaps()


def merge_match_all(parts) -> np.ndarray:
    """matchAll rows of several shards of one job (text shards or bucket shards: every hit is found by exactly one of them)
    -> one array in the order a single handle returns them: by read, and inside a read by (k, pos, file, frag, score,
    inverted), the order of MatchPosAndError::operator< that unifyMatches sorts with (matchAllImplementation.cpp:122-161);
    rows equal in all of these are dropped like its std::unique does.  No collective is needed for matchAll: each
    rank's rows go to the host and are merged here."""
    parts = [np.asarray(p, dtype=_lib.HIT_DTYPE) for p in parts if len(p)]
    if not parts:
        return np.zeros(0, dtype=_lib.HIT_DTYPE)
    h = np.concatenate(parts)
    order = np.lexsort((h["inverted"], h["score"], h["frag"], h["file"], h["pos"], h["k"], h["patid"]))
    h = h[order]
    if len(h) > 1:
        same = np.ones(len(h) - 1, dtype=bool)
        for f in ("patid", "k", "pos", "file", "frag", "inverted"):
            same &= h[f][1:] == h[f][:-1]
        same &= h["score"][1:] == h["score"][:-1]
        h = h[np.concatenate(([True], ~same))]
    return h


# UniqueMatchInfo field access (UniqueMatchInfo.hpp:26-39) on numpy arrays
def umi_state(d): return (np.asarray(d, dtype=np.uint64) >> np.uint64(61)).astype(np.int64)
def umi_pos(d): return (np.asarray(d, dtype=np.uint64) & np.uint64((1 << 35) - 1)).astype(np.int64)
def umi_file(d): return ((np.asarray(d, dtype=np.uint64) >> np.uint64(35)) & np.uint64(63)).astype(np.int64)
def umi_err(d): return ((np.asarray(d, dtype=np.uint64) >> np.uint64(41)) & np.uint64(15)).astype(np.int64)
def umi_frag(d): return ((np.asarray(d, dtype=np.uint64) >> np.uint64(45)) & np.uint64(0xFFFF)).astype(np.int64)


UNIQUE_KEY_NONE = 0x7FFFFFFFFFFFFFFF


def unique_key(d) -> np.ndarray:
    """The order-preserving key of the cross-shard fold (unique_key, csrc/post.cuh), as int64: err(4) | unique flag |
    file(6) | pos(35) | strand | frag(16); NoMatch/Gapped = UNIQUE_KEY_NONE."""
    st = umi_state(d)
    key = (umi_err(d) << 59) | ((st != 4).astype(np.int64) << 58) | (umi_file(d) << 52) | (umi_pos(d) << 17) | \
          ((st == 2).astype(np.int64) << 16) | umi_frag(d)
    key[(st == 0) | (st == 3)] = UNIQUE_KEY_NONE
    return key.astype(np.int64)


def unique_word_from_key(key, tie_sums) -> np.ndarray:
    """UniqueMatchInfo words of the winning keys (k_unique_import / k_fold_merge, csrc/post.cuh); entries whose key is
    UNIQUE_KEY_NONE are meaningless."""
    key = np.asarray(key, dtype=np.int64)
    uniq = (((key >> 58) & 1) == 1) & (np.asarray(tie_sums) == 0)
    strand = (key >> 16) & 1
    st = np.where(uniq, np.where(strand == 1, 2, 1), 4)
    return (((key >> 17) & ((1 << 35) - 1)) | (((key >> 52) & 63) << 35) | ((key >> 59) << 41) | ((key & 0xFFFF) << 45) | (st << 61)).astype(np.uint64)


def canonical_unique(d: np.ndarray) -> np.ndarray:
    """What is defined about a UniqueMatchInfo word independently of visiting order: everything for
    Straight/Reverse, (state, errors) for NonUnique (the reference leaves the position of whichever
    hit it saw first there), only the state for NoMatch."""
    d = np.asarray(d, dtype=np.uint64).copy()
    st = umi_state(d)
    non = st == 4
    d[non] &= np.uint64((7 << 61) | (15 << 41))
    d[st == 0] = 0
    return d


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def unique_checksum(info, first: int = 0) -> int:
    """numpy form of real_gpu_unique_checksum: digest of the canonical state of reads first .. first+len(info)-1."""
    d = canonical_unique(info)
    idx = np.arange(first, first + d.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum(_splitmix64(d ^ _splitmix64(idx)), dtype=np.uint64))


def hits_checksum(hits) -> int:
    """Order independent digest of matchAll rows (patid, pos, file, frag, k, inverted, score bits)."""
    h = np.asarray(hits)
    if h.size == 0:
        return 0
    with np.errstate(over="ignore"):
        a = _splitmix64(h["patid"].astype(np.uint64)) ^ _splitmix64(h["pos"].astype(np.uint64) + np.uint64(0x1234567))
        b = (h["file"].astype(np.uint64) << np.uint64(40)) | (h["frag"].astype(np.uint64) << np.uint64(8)) | (h["k"].astype(np.uint64) << np.uint64(1)) | h["inverted"].astype(np.uint64)
        c = np.ascontiguousarray(h["score"]).view(np.uint32).astype(np.uint64)
        return int(np.sum(_splitmix64(a ^ _splitmix64(b ^ (c << np.uint64(32)))), dtype=np.uint64))
