"""Generates the committed golden fixtures from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference and oracle/_ref/ref_harness, built by
`make -C oracle ref`):   python tests/golden/make_golden.py

Every expected output in tests/golden/*.npz / kat_*.json.gz comes from the reference's own
objects driven by oracle/ref_harness (AllMatcher::match, UniqueMatcher::match, matchGaps,
SignatureConstruction, RestWordBuffer, ComputeScore, Scoring ...).  The inputs are produced by
real_b200/synth.py with fixed seeds and stored alongside, so the fixtures are self-contained:
the tests never need the reference at run time.
"""
from __future__ import annotations

import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from real_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def indel_reads(text, seed, nreads, L, sub_rate, frac_indel=0.5):
    """genpat-model reads with one planted indel of 1..3 bases behind the seed ('+' strand only,
    the reference's gapped pass ignores the '-' strand, match.hpp:499)."""
    base = synth.make_reads(text, seed, nreads, L + 3, sub_rate, fastq=True)
    rng = np.random.RandomState(seed)
    seqs, quals, ids = [], [], []
    for i in range(nreads):
        s = base.read(i).copy()
        q = base.qual(i).copy()
        idn = base.ids[i].replace(" length=%d" % (L + 3), "")
        if rng.rand() < frac_indel and "_inv" not in idn:
            g = rng.randint(1, 4)
            p = rng.randint(40, L - 10)
            if rng.rand() < 0.5:
                s = np.concatenate([s[:p], s[p + g:]])
                q = np.concatenate([q[:p], q[p + g:]])
                idn += "_del%d@%d" % (g, p)
            else:
                ins = rng.randint(0, 4, size=g).astype(np.uint8)
                s = np.concatenate([s[:p], ins, s[p:]])
                q = np.concatenate([q[:p], np.full(g, 35, np.uint8), q[p:]])
                idn += "_ins%d@%d" % (g, p)
        seqs.append(s[:L])
        quals.append(q[:L])
        ids.append(idn)
    return synth.reads_from_list(seqs, quals, ids)


def save_case(name, texts, reads, fastq, mode, real_args, params, n_list=0, gaps=False):
    work = tempfile.mkdtemp(prefix="golden_")
    try:
        tdir = os.path.join(work, "txt")
        os.makedirs(tdir)
        for i, t in enumerate(texts):
            synth.write_fasta(os.path.join(tdir, "t%02d.fa" % i), t)
        rf = os.path.join(work, "r.fq" if fastq else "r.fa")
        synth.write_reads(rf, reads, fastq)
        env = {"REAL_HARNESS_NLIST": str(n_list)} if n_list else {}
        targ = tdir if len(texts) > 1 else os.path.join(tdir, "t00.fa")
        args = ["-t", targ, "-p", rf, "-o", "x", "-u", "1" if mode == "unique" else "0", "-R", "0"] + list(real_args)
        if gaps:
            args += ["-g", "1"]
        timing, dump, gdump = O.run_ref(mode, work, args, gap_dump=gaps, env=env, threads=1 if gaps else None)
        order = [int(os.path.basename(f)[1:3]) for f in timing["file_order"]]
        payload = dict(
            params=json.dumps(dict(params, n_list=n_list, fastq=fastq, mode=mode, gaps=gaps, real_args=list(real_args))),
            file_order=np.asarray(order, dtype=np.int32),
            ntexts=np.int32(len(texts)),
            mapped=reads.mapped, offsets=reads.offsets,
            quality=reads.quality if reads.quality is not None else np.zeros(0, np.uint8),
            has_quality=np.int32(reads.quality is not None),
        )
        for i, t in enumerate(texts):
            payload["text%d_symbols" % i] = t.symbols
            payload["text%d_starts" % i] = t.record_starts
        if mode == "all":
            payload["ref_hits"] = np.fromfile(dump, dtype=O.HIT_DTYPE)
        else:
            payload["ref_unique"] = np.fromfile(dump, dtype=O.UNIQUE_DTYPE)
            if gaps:
                payload["ref_gaps"] = np.fromfile(gdump, dtype=O.GAP_DTYPE)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
        n = len(payload.get("ref_hits", payload.get("ref_unique")))
        print("%-22s %s rows=%d" % (name, mode, n))
    finally:
        shutil.rmtree(work, ignore_errors=True)


def save_kat(name, text, reads, real_args):
    work = tempfile.mkdtemp(prefix="golden_")
    try:
        synth.write_fasta(os.path.join(work, "t.fa"), text)
        synth.write_reads(os.path.join(work, "r.fq"), reads, True)
        _, dump, _ = O.run_ref("kat", work, ["-t", os.path.join(work, "t.fa"), "-p", os.path.join(work, "r.fq"), "-o", "x", "-Q", "33"] + list(real_args))
        with open(dump) as f:
            doc = json.load(f)
        with gzip.open(os.path.join(OUT, name + ".json.gz"), "wt") as f:
            json.dump(doc, f, separators=(",", ":"))
        print("%-22s kat reads=%d" % (name, len(doc["reads"])))
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref/ref_harness missing: run `make -C oracle ref` where /root/reference exists")

    t = synth.make_text(101, 60000)
    # C1-like: FASTA 36 bp, -e 2, no scores
    save_case("all_c1", [t], synth.make_reads(t, 1, 700, 36, 0.02, False), False, "all",
              ["-s", "2", "-e", "2", "-l", "32", "-q", "0"], dict(seedl=32, seedkmax=2, totalkmax=2, scores=False))
    # C2-like: FASTQ 100 bp, -e 4, scores
    save_case("all_c2", [t], synth.make_reads(t, 2, 600, 100, 0.01, True), True, "all",
              ["-e", "4", "-q", "1", "-Q", "33"], dict(seedl=32, seedkmax=2, totalkmax=4, scores=True))
    # records + wildcards, two index blocks
    t3 = synth.make_text(103, 70000, nrecords=6, n_per_million=3000)
    save_case("all_records_n", [t3], synth.make_reads(t3, 3, 700, 48, 0.03, False), False, "all",
              ["-e", "4", "-q", "0"], dict(seedl=32, seedkmax=2, totalkmax=4, scores=False), n_list=30000)
    # C5-like: long reads, -e 8
    save_case("all_c5", [t], synth.make_reads(t, 4, 300, 250, 0.01, False), False, "all",
              ["-e", "8", "-q", "0"], dict(seedl=32, seedkmax=2, totalkmax=8, scores=False))
    # repeats: many hits per read, palindromic text so both strands hit one position
    sym = t.symbols.copy()
    sym[30000:36000] = sym[10000:16000]
    sym[40000:43000] = sym[10000:13000]
    pal = sym[50000:50060].copy()
    sym[50060:50120] = synth.revcomp_mapped(pal)
    trep = synth.Text(sym, t.records)
    rrep = synth.make_reads(trep, 5, 700, 60, 0.01, True)
    pal_reads = synth.reads_from_list([trep.symbols[50030:50090].copy(), trep.symbols[50010:50110].copy()],
                                      [np.full(60, 35, np.uint8), np.full(100, 35, np.uint8)], ["pal60", "pal100"])
    rrep = synth.concat_reads([rrep, pal_reads])
    save_case("all_repeats", [trep], rrep, True, "all", ["-e", "4", "-q", "1", "-Q", "33"],
              dict(seedl=32, seedkmax=2, totalkmax=4, scores=True))
    save_case("all_repeats_noscore", [trep], rrep, True, "all", ["-e", "4", "-q", "0", "-Q", "33"],
              dict(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    # other seed geometries
    r6 = synth.make_reads(t, 6, 500, 80, 0.02, True)
    save_case("all_l20", [t], r6, True, "all", ["-e", "5", "-q", "1", "-Q", "33", "-l", "20"], dict(seedl=20, seedkmax=2, totalkmax=5, scores=True))
    save_case("all_l28_s1", [t], r6, True, "all", ["-e", "5", "-q", "0", "-Q", "33", "-l", "28", "-s", "1"], dict(seedl=28, seedkmax=1, totalkmax=5, scores=False))
    save_case("all_l32_s0", [t], r6, True, "all", ["-e", "6", "-q", "0", "-Q", "33", "-s", "0"], dict(seedl=32, seedkmax=0, totalkmax=6, scores=False))
    save_case("all_l64", [t], r6, True, "all", ["-e", "5", "-q", "1", "-Q", "33", "-l", "64"], dict(seedl=64, seedkmax=2, totalkmax=5, scores=True))
    # ragged lengths, short reads, reads with N
    rag = []
    rng = np.random.RandomState(7)
    for i in range(400):
        L = int(rng.choice([20, 31, 32, 33, 36, 63, 64, 65, 96, 100, 129, 150]))
        p = int(rng.randint(0, t.n - L))
        s = t.symbols[p:p + L].copy()
        if rng.rand() < 0.5:
            s = synth.revcomp_mapped(s)
        for _ in range(int(rng.randint(0, 3))):
            j = int(rng.randint(0, L)); s[j] = (s[j] + 1 + rng.randint(0, 3)) % 4
        if i % 37 == 0:
            s[int(rng.randint(0, L))] = 4
        rag.append(s)
    rag_reads = synth.reads_from_list(rag, [np.full(len(s), 20 + (i % 30), np.uint8) for i, s in enumerate(rag)], ["rag%d" % i for i in range(len(rag))])
    save_case("all_ragged", [t], rag_reads, True, "all", ["-e", "3", "-q", "1", "-Q", "33"], dict(seedl=32, seedkmax=2, totalkmax=3, scores=True))

    # unique
    save_case("unique_plain", [trep], rrep, True, "unique", ["-e", "4", "-q", "0", "-Q", "33"], dict(seedl=32, seedkmax=2, totalkmax=4, scores=False))
    save_case("unique_scores_blocks", [trep], rrep, True, "unique", ["-e", "4", "-q", "1", "-Q", "33"],
              dict(seedl=32, seedkmax=2, totalkmax=4, scores=True), n_list=25000)
    t4 = synth.make_text(104, 50000, nrecords=4, n_per_million=2000)
    both = synth.Text(trep.symbols[:30000].copy(), [(" part", 0)])
    rmix = synth.concat_reads([synth.make_reads(t4, 8, 300, 60, 0.02, True), synth.make_reads(both, 9, 300, 60, 0.02, True)])
    save_case("unique_multifile", [t4, both, trep], rmix, True, "unique", ["-e", "4", "-q", "0", "-Q", "33"],
              dict(seedl=32, seedkmax=2, totalkmax=4, scores=False), n_list=20000)
    save_case("unique_multifile_scores", [t4, both, trep], rmix, True, "unique", ["-e", "4", "-q", "1", "-Q", "33"],
              dict(seedl=32, seedkmax=2, totalkmax=4, scores=True), n_list=20000)
    save_case("unique_ragged", [t], rag_reads, True, "unique", ["-e", "3", "-q", "0", "-Q", "33"], dict(seedl=32, seedkmax=2, totalkmax=3, scores=False))

    # gapped
    tg = synth.make_text(105, 80000)
    rg = indel_reads(tg, 10, 500, 150, 0.01)
    save_case("gaps_c4", [tg], rg, True, "unique", ["-e", "3", "-q", "1", "-Q", "33"], dict(seedl=32, seedkmax=2, totalkmax=3, scores=True), gaps=True)
    tg2 = synth.make_text(106, 60000, nrecords=5, n_per_million=1500)
    rg2 = indel_reads(tg2, 11, 400, 120, 0.02)
    save_case("gaps_records_blocks", [tg2], rg2, True, "unique", ["-e", "3", "-q", "1", "-Q", "33"],
              dict(seedl=32, seedkmax=2, totalkmax=3, scores=True), gaps=True, n_list=25000)

    # unit-level known answers
    tk = synth.make_text(107, 3000, nrecords=3, n_per_million=30000)
    rk = synth.concat_reads([synth.make_reads(tk, 12, 12, L, 0.03, True) for L in (32, 33, 36, 64, 96, 100, 150, 250)])
    save_kat("kat_l32", tk, rk, [])
    save_kat("kat_l20", tk, rk, ["-l", "20"])
    save_kat("kat_l64", tk, synth.concat_reads([synth.make_reads(tk, 13, 12, L, 0.03, True) for L in (64, 65, 100, 150)]), ["-l", "64"])


if __name__ == "__main__":
    main()
