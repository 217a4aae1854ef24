"""Golden fixture of the reference's TEXT LOADER on awkward FASTA files, from the reference itself.

Run in the build container only (needs /root/reference and oracle/_ref/ref_harness):
    python tests/golden/make_text_golden.py
Each case is a FASTA byte string; the expected symbols and record table are what the reference's own
getText (countReads.cpp countLength/readFile + AutoTextArray) produced for it, as dumped by the
harness's `kat` mode (ranges, symbols).  Stored in tests/golden/text_quirks.npz.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle_py as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases():
    rng = np.random.RandomState(20261018)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def seq(n, alphabet=acgt):
        return alphabet[rng.randint(0, alphabet.size, n)].tobytes()

    def lines(s, w=60, eol=b"\n"):
        return b"".join(s[o:o + w] + eol for o in range(0, len(s), w))

    out = {}
    # plain multi-record file, 60 columns, N runs
    s = bytearray(seq(9000)); s[1000:1100] = b"N" * 100; s[5000:5003] = b"NNN"
    out["plain"] = b">chr1 first\n" + lines(bytes(s[:4000])) + b">chr2\n" + lines(bytes(s[4000:])) + b">empty\n>tail\n" + lines(seq(700))
    # soft-masked (lower case is dropped), CR LF line ends, IUPAC codes, blanks and tabs
    mixed = np.frombuffer(b"ACGTNacgtnRYKM \t", dtype=np.uint8)
    out["lowercase_crlf"] = b">r1 x\r\n" + lines(seq(6000, mixed), 70, b"\r\n") + b">r2\r\n" + lines(seq(3000, mixed), 70, b"\r\n")
    # '>' in the middle of a sequence line and inside headers, bases in front of the first header, empty lines,
    # empty name, no newline at the end of the file, header at the end of the file that is never closed
    out["stray_markers"] = (seq(500) + b"\n" + b">a>b>c d\n" + seq(2500) + b">mid line header\n\n\n" + seq(1200) + b"\n>\n" + seq(800)
                            + b"\n>x\n>y\n" + seq(900) + b">unclosed header ACGT")
    # header longer than a 4096-byte tile, sequence lines of 5000 columns, three consecutive markers at 4095..4097
    body = seq(4095 - 3) + b"\n" + b">" + b">" + b">z\n"
    out["tiles"] = b">h\n" + body[3:] + seq(5000) + b"\n>" + seq(4500, np.frombuffer(b"ACGTNxyz >", dtype=np.uint8)).replace(b"\n", b"") + b"\n" + lines(seq(12000), 5000)
    # no header at all
    out["no_header"] = lines(seq(3000))
    return out


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref/ref_harness missing: run `make -C oracle ref` where /root/reference exists")
    payload = {}
    names = []
    for name, data in cases().items():
        work = tempfile.mkdtemp(prefix="golden_text_")
        try:
            with open(os.path.join(work, "t.fa"), "wb") as f:
                f.write(data)
            with open(os.path.join(work, "r.fq"), "wb") as f:
                f.write(b"@r0\n" + b"ACGT" * 10 + b"\n+\n" + b"I" * 40 + b"\n")
            _, dump, _ = O.run_ref("kat", work, ["-t", os.path.join(work, "t.fa"), "-p", os.path.join(work, "r.fq"), "-o", "x", "-Q", "33"])
            with open(dump, newline="") as f:          # names may hold a '\r'
                doc = json.load(f, strict=False)
        finally:
            shutil.rmtree(work, ignore_errors=True)
        rn = [r[0].encode("latin-1") for r in doc["ranges"]]
        assert rn[-1] == b"terminal" and all(b"\n" not in x for x in rn), rn
        payload[name + "_fasta"] = np.frombuffer(data, dtype=np.uint8)
        payload[name + "_symbols"] = np.asarray(doc["symbols"], dtype=np.uint8)
        payload[name + "_starts"] = np.asarray([r[1] for r in doc["ranges"]], dtype=np.uint64)
        payload[name + "_names"] = np.frombuffer(b"\n".join(rn[:-1]), dtype=np.uint8)
        payload[name + "_nrecords"] = np.asarray(len(rn) - 1)
        names.append(name)
        print("%-16s bytes=%d bases=%d records=%d" % (name, len(data), len(doc["symbols"]), len(rn) - 1))
    payload["cases"] = np.frombuffer("\n".join(names).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "text_quirks.npz"), **payload)


if __name__ == "__main__":
    main()
