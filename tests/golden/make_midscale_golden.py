"""Mid-scale golden digests from the REFERENCE ITSELF (oracle/_ref/ref_harness): inputs too large to commit, so the
fixture stores the generator parameters (real_b200/synth.py is counter based: same seeds, same bytes) and the sha256 of the
reference's canonical result.  Run in the build container only:   python tests/golden/make_midscale_golden.py

  midscale_unique : 32 Mbp text (4 records, 0.1 % N, one planted 30 kb repeat), 500 k x 100 bp reads, matchUnique -e 4 -q 0
  midscale_all    : the same text, 300 k x 100 bp FASTQ reads, matchAll -e 4 with scores
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from real_b200 import matcher, synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "midscale.json")

CASES = {
    "midscale_unique": dict(seed=7001, n=32_000_000, nrec=4, npm=1000, reads=500_000, L=100, sub=0.01, e=4, mode="unique", scores=False),
    "midscale_all": dict(seed=7003, n=32_000_000, nrec=4, npm=1000, reads=300_000, L=100, sub=0.012, e=4, mode="all", scores=True),
}


def make_inputs(c):
    text = synth.make_text(c["seed"], c["n"], nrecords=c["nrec"], n_per_million=c["npm"])
    sym = text.symbols.copy()
    sym[c["n"] // 2:c["n"] // 2 + 30_000] = sym[5000:35_000]       # a repeat: NonUnique reads / multi-hit reads
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, c["seed"] + 1, c["reads"], c["L"], c["sub"], fastq=c["scores"])
    return text, reads


def canonical_digest(c, result) -> str:
    """sha256 of the canonical result: unique = canonical words (matcher.canonical_unique) as little-endian u64;
    all = rows (patid, k, pos, frag, inverted, score bits) as int64, sorted lexicographically."""
    if c["mode"] == "unique":
        a = matcher.canonical_unique(np.asarray(result, dtype=np.uint64))
        return hashlib.sha256(np.ascontiguousarray(a).astype("<u8").tobytes()).hexdigest()
    h = result
    a = np.stack([h["patid"].astype(np.int64), h["k"].astype(np.int64), h["pos"].astype(np.int64), h["frag"].astype(np.int64),
                  h["inverted"].astype(np.int64), np.ascontiguousarray(h["score"]).view(np.uint32).astype(np.int64)], 1)
    a = a[np.lexsort(a.T[::-1])]
    return hashlib.sha256(np.ascontiguousarray(a).astype("<i8").tobytes()).hexdigest()


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref/ref_harness missing: run `make -C oracle ref` where /root/reference exists")
    doc = {}
    for name, c in CASES.items():
        text, reads = make_inputs(c)
        work = tempfile.mkdtemp(prefix="midscale_")
        try:
            synth.write_fasta(os.path.join(work, "t.fa"), text)
            rf = os.path.join(work, "r.fq" if c["scores"] else "r.fa")
            synth.write_reads(rf, reads, c["scores"])
            args = ["-t", os.path.join(work, "t.fa"), "-p", rf, "-o", "x", "-u", "1" if c["mode"] == "unique" else "0", "-R", "0",
                    "-s", "2", "-e", str(c["e"]), "-l", "32", "-q", "1" if c["scores"] else "0"]
            if c["scores"]:
                args += ["-Q", "33"]
            timing, dump, _ = O.run_ref(c["mode"], work, args, env={"REAL_HARNESS_NLIST": str(c["n"])})
            if c["mode"] == "unique":
                res = np.fromfile(dump, dtype=O.UNIQUE_DTYPE)["data"]
                st = matcher.umi_state(res)
                counts = {"straight": int((st == 1).sum()), "reverse": int((st == 2).sum()), "nonunique": int((st == 4).sum()), "nomatch": int((st == 0).sum())}
            else:
                res = np.fromfile(dump, dtype=O.HIT_DTYPE)
                counts = {"rows": int(len(res)), "reads_with_hits": int(np.unique(res["patid"]).size)}
            doc[name] = dict(params=c, sha256=canonical_digest(c, res), counts=counts,
                             reference_seconds={"index": timing["index_s"], "match": timing["match_s"]})
            print(name, doc[name]["sha256"], counts, doc[name]["reference_seconds"])
        finally:
            shutil.rmtree(work, ignore_errors=True)
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
