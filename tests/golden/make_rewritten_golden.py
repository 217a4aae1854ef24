#!/usr/bin/env python
"""Fixtures of the reference's rewritten pattern file (next-4): tests/golden/rewritten_{fa,fq}.npz.

Runs the reference's OWN writer + decoder (src/reorderPat.cpp, compiled in place by `make -C oracle ref_reorder`) on small
pattern files built to exercise the format: ragged lengths (also not multiples of 4 and of 2), reads with wildcards (the
4 bit/base sections), empty sections, multi-line records, ids with blanks.  Stored: the bytes of the input file, the bytes
reorderPat wrote, and what its decoder printed back (ordinal, bases, id) -- the order and content the host driver's reader
must reproduce.  Only runs where /root/reference exists; the fixtures are committed."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REORDER = os.path.join(ROOT, "oracle", "_ref", "reorderPat")


def make_inputs():
    rng = np.random.RandomState(77)
    fa, fq = [], []
    for i in range(240):
        L = int(rng.choice([5, 17, 32, 33, 36, 50, 63, 64, 100]))
        s = "".join("ACGT"[x] for x in rng.randint(0, 4, L))
        if i % 9 == 4:
            k = int(rng.randint(0, L))
            s = s[:k] + "N" + s[k + 1:]
        if i % 31 == 7:
            s = s.lower()
        ident = "read_%d" % i + (" with a comment" if i % 3 == 0 else "")
        wrapped = s if i % 5 else "\n".join(s[a:a + 13] for a in range(0, L, 13))        # multi-line records
        fa.append(">%s\n%s\n" % (ident, wrapped))
        q = "".join(chr(33 + int(x)) for x in rng.randint(0, 41, L))
        fq.append("@%s\n%s\n+%s\n%s\n" % (ident, s, ident if i % 2 else "", q))
    return "".join(fa).encode(), "".join(fq).encode()


def run(data: bytes, suffix: str):
    work = tempfile.mkdtemp(prefix="rewritten_")
    src, dst = os.path.join(work, "r" + suffix), os.path.join(work, "r.bin")
    with open(src, "wb") as f:
        f.write(data)
    p = subprocess.run([REORDER, src, dst], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    err = p.stderr.decode("latin1")
    # the decoder prints "<ordinal>\t<bases>" per read, then "<ordinal>\t<id>" per read
    rows = re.findall(r"^(\d+)\t(.*)$", err, flags=re.M)
    n = len(rows) // 2
    bases = [r[1] for r in rows[:n]]
    ids = [r[1] for r in rows[n:]]
    assert [int(r[0]) for r in rows[:n]] == list(range(n)) and [int(r[0]) for r in rows[n:]] == list(range(n))
    return open(dst, "rb").read(), bases, ids


if __name__ == "__main__":
    if not os.path.exists(REORDER):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref_reorder"])
    fa, fq = make_inputs()
    for name, data, suffix in (("fa", fa, ".fa"), ("fq", fq, ".fq")):
        out, bases, ids = run(data, suffix)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rewritten_%s.npz" % name), input=np.frombuffer(data, dtype=np.uint8),
                            rewritten=np.frombuffer(out, dtype=np.uint8), bases=np.asarray(bases), ids=np.asarray(ids))
        print(name, len(data), "->", len(out), "bytes,", len(ids), "reads")
