"""Seeded synthetic inputs in the reference's file formats.

The reference's own generators (`randstr.cpp:35-52`, `genpat.cpp:92-158`) seed with
`time(0)` and are not reproducible, so this module restates their *output formats*
(FASTA header `> random_<len>` + 60-column lines; read ids `p<pos>[_inv][_<j><from><to>]...`;
FASTQ qualities `D` for an unchanged base and `*` for a changed one) on top of a
counter-based generator: every text word and every read is a pure function of
(seed, index), so the same data can be produced independently by numpy here and by the
CUDA generator kernels (`real_b200/csrc/synth.cu`) at benchmark scale.

Layouts produced here are the ones the drop-in boundary takes (include/real_gpu.h):
  * text: u64 words, 32 bases per word, base i at bits 63-2*(i%32)..62-2*(i%32)
    (`AutoTextArray.hpp:27-43`), codes A=0 C=1 G=2 T=3, N stored as 0;
  * N mask: u64 words, bit i at bit 63-(i%64) (`AutoTextArray.hpp:45-61`);
  * reads: one byte per base, 0..3 = ACGT, 4 = anything else (`Pattern.hpp:105-128`).
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
BASES = np.frombuffer(b"ACGTN", dtype=np.uint8)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream(seed: int, tag: int) -> np.uint64:
    """Independent 64-bit stream key for (seed, tag); consumers hash `key ^ counter`."""
    with np.errstate(over="ignore"):
        a = splitmix64(np.asarray([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))
        return splitmix64(a + np.uint64(tag))[0]


# ------------------------------------------------------------------ packing helpers

def pack_text(symbols: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """symbols (uint8, 0..4) -> (2-bit words, N-mask words), reference layouts."""
    symbols = np.asarray(symbols, dtype=np.uint8)
    n = symbols.size
    nw = (n + 31) // 32
    pad = np.zeros(nw * 32, dtype=np.uint64)
    pad[:n] = symbols & 3  # N is stored as A (AutoTextArray.hpp:39)
    pad[:n][symbols > 3] = 0
    shifts = (np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64))
    words = np.bitwise_or.reduce(pad.reshape(nw, 32) << shifts, axis=1).astype(np.uint64)
    nmw = (n + 63) // 64
    nb = np.zeros(nmw * 64, dtype=np.uint64)
    nb[:n] = (symbols > 3)
    mshifts = (np.uint64(63) - np.arange(64, dtype=np.uint64))
    nmask = np.bitwise_or.reduce(nb.reshape(nmw, 64) << mshifts, axis=1).astype(np.uint64)
    return words, nmask


def unpack_text(words: np.ndarray, n: int, nmask: Optional[np.ndarray] = None) -> np.ndarray:
    shifts = (np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64))
    sym = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n]
    if nmask is not None:
        mshifts = (np.uint64(63) - np.arange(64, dtype=np.uint64))
        isn = ((nmask[:, None] >> mshifts) & np.uint64(1)).astype(bool).reshape(-1)[:n]
        sym = sym.copy()
        sym[isn] = 4
    return sym


def revcomp_mapped(m: np.ndarray) -> np.ndarray:
    """Reverse complement of mapped bytes; 4 (N) stays 4 (`acgtnMap.hpp` invertN)."""
    r = m[::-1].copy()
    ok = r < 4
    r[ok] = 3 - r[ok]
    return r


# ------------------------------------------------------------------ text

@dataclasses.dataclass
class Text:
    """One text file: concatenated symbols plus the record table of `countReads.cpp:28-84`."""
    symbols: np.ndarray                       # uint8, 0..4
    records: List[Tuple[str, int]]            # (header text after '>', start offset)

    @property
    def n(self) -> int:
        return int(self.symbols.size)

    @property
    def record_starts(self) -> np.ndarray:
        """Start offsets plus the terminal sentinel n (`countReads.cpp:81`)."""
        return np.asarray([s for _, s in self.records] + [self.n], dtype=np.uint64)

    def packed(self) -> Tuple[np.ndarray, np.ndarray]:
        return pack_text(self.symbols)


def text_words(seed: int, nwords: int, first_word: int = 0) -> np.ndarray:
    """Packed text words [first_word, first_word+nwords): word w = splitmix64(stream(seed,1) ^ w)."""
    idx = np.arange(first_word, first_word + nwords, dtype=np.uint64) ^ stream(seed, 1)
    return splitmix64(idx)


def nrun_mask_words(seed: int, nwords64: int, per_million: int, first_word: int = 0) -> np.ndarray:
    """N mask: each 64-base word is all-N with probability per_million / 1e6 (runs of 64)."""
    if per_million <= 0:
        return np.zeros(nwords64, dtype=np.uint64)
    idx = np.arange(first_word, first_word + nwords64, dtype=np.uint64) ^ stream(seed, 2)
    h = splitmix64(idx)
    hit = (h % np.uint64(1000000)) < np.uint64(per_million)
    return np.where(hit, _M64, np.uint64(0)).astype(np.uint64)


def make_text(seed: int, n: int, nrecords: int = 1, n_per_million: int = 0) -> Text:
    """Uniform random ACGT text of n bases, optionally cut into unequal records and
    sprinkled with 64-base N runs.  Header of a single record is `> random_<n>` exactly
    like `randstr.cpp:38`."""
    words = text_words(seed, (n + 31) // 32)
    nm = nrun_mask_words(seed, (n + 63) // 64, n_per_million)
    sym = unpack_text(words, n, nm)
    if nrecords <= 1:
        recs = [(" random_%d" % n, 0)]
    else:
        # unequal record lengths: cut points from the same counter-based stream
        cuts = sorted(set(int(x % np.uint64(n)) for x in splitmix64(np.arange(nrecords - 1, dtype=np.uint64) ^ stream(seed, 3))))
        cuts = [c for c in cuts if c > 0]
        starts = [0] + cuts
        recs = [(" random_%d_part%d" % (n, i), s) for i, s in enumerate(starts)]
    return Text(symbols=sym, records=recs)


def write_fasta_file(f, text: Text, line: int = 60) -> None:
    """FASTA writer in `randstr.cpp` style (60 columns) onto an open binary file."""
    starts = [s for _, s in text.records] + [text.n]
    for (name, s), e in zip(text.records, starts[1:]):
        f.write(b">" + name.encode() + b"\n")
        chunk = BASES[text.symbols[s:e]]
        full = (chunk.size // line) * line
        if full:
            rows = np.empty((full // line, line + 1), dtype=np.uint8)
            rows[:, :line] = chunk[:full].reshape(-1, line)
            rows[:, line] = 10
            f.write(rows.tobytes())
        if chunk.size > full:
            f.write(chunk[full:].tobytes() + b"\n")


def write_fasta(path: str, text: Text, line: int = 60) -> None:
    with open(path, "wb") as f:
        write_fasta_file(f, text, line)


# ------------------------------------------------------------------ reads

@dataclasses.dataclass
class Reads:
    """A read set in boundary layout plus what is needed to write FASTA/FASTQ."""
    mapped: np.ndarray                 # uint8 concatenated bases 0..4
    offsets: np.ndarray                # uint64, nreads+1
    quality: Optional[np.ndarray]      # uint8 concatenated PHRED values (already minus offset) or None
    ids: List[str]

    @property
    def nreads(self) -> int:
        return int(self.offsets.size - 1)

    def read(self, i: int) -> np.ndarray:
        return self.mapped[int(self.offsets[i]):int(self.offsets[i + 1])]

    def qual(self, i: int) -> Optional[np.ndarray]:
        if self.quality is None:
            return None
        return self.quality[int(self.offsets[i]):int(self.offsets[i + 1])]


def read_plan_total(seed: int, total: int, span: int, first: int, count: int):
    """Start position (stratified uniform, hence ascending like genpat's sorted positions)
    and strand of reads [first, first+count) out of a set of `total` reads."""
    r = np.arange(first, first + count, dtype=np.uint64)
    lo = (r * np.uint64(span)) // np.uint64(total)
    hi = ((r + np.uint64(1)) * np.uint64(span)) // np.uint64(total)
    width = np.maximum(hi - lo, np.uint64(1))
    h0 = splitmix64(r ^ stream(seed, 4))
    h1 = splitmix64(r ^ stream(seed, 5))
    pos = lo + h0 % width
    pos = np.minimum(pos, np.uint64(span - 1))
    strand = (h1 >> np.uint64(63)).astype(np.uint8)
    return pos, strand


def substitution_plan(seed: int, first: int, count: int, length: int, sub_per_16384: int):
    """(count, length) arrays: substitute flag and delta in 1..3 for every read base."""
    r = np.arange(first, first + count, dtype=np.uint64)[:, None]
    j = np.arange(length, dtype=np.uint64)[None, :]
    hc = splitmix64((r * np.uint64(64) + (j >> np.uint64(2))) ^ stream(seed, 6))
    u16 = (hc >> (np.uint64(16) * (j & np.uint64(3)))) & np.uint64(0xFFFF)
    sub = (u16 >> np.uint64(2)) < np.uint64(sub_per_16384)
    delta = (np.uint64(1) + (u16 & np.uint64(3)) % np.uint64(3)).astype(np.uint8)
    return sub, delta


def make_reads(text: Text, seed: int, nreads: int, length: int, sub_rate: float,
               fastq: bool, total: Optional[int] = None, first: int = 0) -> Reads:
    """genpat-model reads (`genpat.cpp:92-158`): window of the text, reverse-complemented with
    p=0.5, per-base substitutions, ids that encode the truth.  Uniform length."""
    total = nreads if total is None else total
    span = text.n - length + 1
    assert span > 0
    pos, strand = read_plan_total(seed, total, span, first, nreads)
    thr = int(round(sub_rate * 16384))
    sub, delta = substitution_plan(seed, first, nreads, length, thr)
    idx = pos[:, None].astype(np.int64) + np.arange(length, dtype=np.int64)[None, :]
    win = text.symbols[idx]                                   # (nreads, L) forward text
    rc = win[:, ::-1].copy()
    okrc = rc < 4
    rc[okrc] = 3 - rc[okrc]
    base = np.where(strand[:, None] == 1, rc, win)
    changed = sub & (base < 4)
    out = base.copy()
    out[changed] = (base[changed] + delta[changed]) & 3
    ids: List[str] = []
    for i in range(nreads):
        s = "p%d" % int(pos[i])
        if strand[i]:
            s += "_inv"
        for jj in np.nonzero(changed[i])[0]:
            s += "_%d%s%s" % (jj, "ACGTN"[base[i, jj]], "ACGTN"[out[i, jj]])
        if fastq:
            s += " length=%d" % length
        ids.append(s)
    quality = None
    if fastq:
        # 'D' (68) unchanged, '*' (42) changed, Sanger offset 33 (genpat.cpp:153-157)
        quality = np.where(changed, 42 - 33, 68 - 33).astype(np.uint8).reshape(-1)
    offsets = (np.arange(nreads + 1, dtype=np.uint64) * np.uint64(length))
    return Reads(mapped=out.astype(np.uint8).reshape(-1), offsets=offsets, quality=quality, ids=ids)


def plant_deletions(text: Text, reads: Reads, seed: int, length: int) -> np.ndarray:
    """C4: 30 % of the '+' strand reads get one deletion of 1..3 bases behind the seed -- read[off+g:] moves up
    and the tail is refilled from the text behind the window (the formula of devsynth.plant_deletions).  In place on
    reads.mapped (uniform `length`, made by make_reads with the same seed); returns the planted mask."""
    R = reads.nreads
    pos, strand = read_plan_total(seed, R, text.n - length + 1, 0, R)
    pos = pos.astype(np.int64)
    rid = np.arange(R, dtype=np.int64)
    g = 1 + rid % 3
    off = 60 + rid % 50
    planted = (rid % 10 < 3) & (strand == 0) & (pos + length + 3 < text.n)
    m2 = reads.mapped.reshape(R, length)
    col = np.arange(length, dtype=np.int64)[None, :]
    src_col = np.where(col >= off[:, None], col + g[:, None], col)
    a = np.take_along_axis(m2, np.minimum(src_col, length - 1), axis=1)
    b = text.symbols[np.minimum(pos[:, None] + src_col, text.n - 1)]
    new = np.where(src_col < length, a, b)
    m2[planted] = new[planted]
    return planted


def reads_from_list(seqs: Sequence[np.ndarray], quals: Optional[Sequence[np.ndarray]] = None,
                    ids: Optional[Sequence[str]] = None) -> Reads:
    """Ragged read set from explicit arrays (tests)."""
    lens = np.asarray([len(s) for s in seqs], dtype=np.uint64)
    offsets = np.concatenate([np.zeros(1, dtype=np.uint64), np.cumsum(lens, dtype=np.uint64)])
    mapped = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    quality = None
    if quals is not None:
        quality = np.concatenate([np.asarray(q, dtype=np.uint8) for q in quals]) if len(quals) else np.zeros(0, np.uint8)
    if ids is None:
        ids = ["r%d" % i for i in range(len(seqs))]
    return Reads(mapped=mapped, offsets=offsets, quality=quality, ids=list(ids))


def concat_reads(parts: Sequence[Reads]) -> Reads:
    seqs, quals, ids = [], [], []
    has_q = all(p.quality is not None for p in parts)
    for p in parts:
        for i in range(p.nreads):
            seqs.append(p.read(i))
            if has_q:
                quals.append(p.qual(i))
        ids.extend(p.ids)
    return reads_from_list(seqs, quals if has_q else None, ids)


def write_reads(path: str, reads: Reads, fastq: bool, qoffset: int = 33) -> None:
    """FASTA (`>id\\nSEQ\\n`) or FASTQ (`@id\\nSEQ\\n+\\nQUAL\\n`) exactly as genpat prints them."""
    with open(path, "wb") as f:
        for i in range(reads.nreads):
            seq = BASES[np.minimum(reads.read(i), 4)].tobytes()
            if fastq:
                q = reads.qual(i)
                if q is None:
                    q = np.full(len(seq), 30, dtype=np.uint8)
                f.write(b"@" + reads.ids[i].encode() + b"\n" + seq + b"\n+\n" + (q + qoffset).astype(np.uint8).tobytes() + b"\n")
            else:
                f.write(b">" + reads.ids[i].encode() + b"\n" + seq + b"\n")


def tmp_path(dirname: str, name: str) -> str:
    os.makedirs(dirname, exist_ok=True)
    return os.path.join(dirname, name)


def pack_reads_2bit(reads: Reads):
    """Reads in the layout of the reference's rewritten pattern file (TemporaryFile.hpp:231-268): 4 bases per byte,
    first base in bits 7..6, every read on a byte boundary.  Returns (packed, byte_offsets, lengths, wildcard_flags);
    reads with a wildcard are flagged and their bases stored as A."""
    n = reads.nreads
    lens = np.diff(reads.offsets.astype(np.int64)).astype(np.int64)
    nbytes = (lens + 3) // 4
    boffs = np.concatenate([np.zeros(1, np.int64), np.cumsum(nbytes)]).astype(np.uint64)
    packed = np.zeros(int(boffs[-1]), dtype=np.uint8)
    flags = np.zeros(n, dtype=np.uint8)
    for r in range(n):
        s = reads.read(r)
        if (s > 3).any():
            flags[r] = 1
            s = np.where(s > 3, 0, s)
        pad = np.zeros(int(nbytes[r]) * 4, dtype=np.uint8)
        pad[:len(s)] = s
        q = pad.reshape(-1, 4)
        packed[int(boffs[r]):int(boffs[r + 1])] = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
    return packed, boffs, lens.astype(np.uint32), flags
