"""Inputs of the command-line parity cases (shared by the golden generator and the GPU test)."""
import os

import numpy as np

from real_b200 import synth

CASES = ["unique_fa_R1", "unique_fq_R0", "unique_dir_ragged", "unique_fq_scores_default", "unique_fq_l64_scores"]


def make_case(name, work):
    if name == "unique_fa_R1":
        text = synth.make_text(301, 80000, nrecords=3, n_per_million=2000)
        reads = synth.concat_reads([synth.make_reads(text, 302, 300, 50, 0.02, False), synth.make_reads(text, 303, 300, 36, 0.02, False),
                                    synth.make_reads(text, 304, 300, 75, 0.02, False)])
        synth.write_fasta(os.path.join(work, "t.fa"), text)
        rf = os.path.join(work, "r.fa")
        synth.write_reads(rf, reads, False)
        return os.path.join(work, "t.fa"), rf, ["-e", "3", "-q", "0"]                      # default -u 1 -R 1
    if name == "unique_fq_R0":
        text = synth.make_text(311, 60000)
        sym = text.symbols.copy()
        sym[30000:33000] = sym[2000:5000]
        text = synth.Text(sym, text.records)
        reads = synth.make_reads(text, 312, 800, 100, 0.01, True)
        synth.write_fasta(os.path.join(work, "t.fa"), text)
        rf = os.path.join(work, "r.fq")
        synth.write_reads(rf, reads, True)
        return os.path.join(work, "t.fa"), rf, ["-e", "4", "-q", "0", "-R", "0", "-Q", "33"]
    if name == "unique_dir_ragged":
        tdir = os.path.join(work, "txt")
        os.makedirs(tdir)
        t0 = synth.make_text(321, 40000, nrecords=2)
        synth.write_fasta(os.path.join(tdir, "only.fa"), t0)
        with open(os.path.join(tdir, "ignored.txt"), "w") as f:
            f.write(">x\nACGT\n")
        rng = np.random.RandomState(5)
        seqs = []
        for i in range(500):
            L = int(rng.choice([20, 32, 33, 40, 64, 90]))
            p = int(rng.randint(0, t0.n - L))
            s = t0.symbols[p:p + L].copy()
            if rng.rand() < 0.5:
                s = synth.revcomp_mapped(s)
            if i % 5 == 0:
                s[int(rng.randint(0, L))] = (s[0] + 1) % 4
            if i % 41 == 0:
                s[3] = 4
            seqs.append(s)
        reads = synth.reads_from_list(seqs, None, ["q%d extra words" % i for i in range(len(seqs))])
        rf = os.path.join(work, "r.fa")
        synth.write_reads(rf, reads, False)
        return tdir, rf, ["-e", "2", "-q", "0"]
    if name == "unique_fq_scores_default":
        # the stock default mode: -u 1 -q 1 -R 1 (quality-aware scores, epsilon filter level 2)
        text = synth.make_text(331, 90000, nrecords=2, n_per_million=1000)
        sym = text.symbols.copy()
        seg = sym[3000:9000].copy()
        seg[::97] = (seg[::97] + 1) % 4
        sym[50000:56000] = seg
        text = synth.Text(sym, text.records)
        reads = synth.concat_reads([synth.make_reads(text, 332, 700, 100, 0.015, True), synth.make_reads(text, 333, 300, 64, 0.02, True)])
        synth.write_fasta(os.path.join(work, "t.fa"), text)
        rf = os.path.join(work, "r.fq")
        synth.write_reads(rf, reads, True)
        return os.path.join(work, "t.fa"), rf, ["-e", "4", "-Q", "33"]
    if name == "unique_fq_l64_scores":
        # 64-base seeds (the reference's u_int64_t signatures), scores, repeats within epsilon
        text = synth.make_text(341, 90000, nrecords=3, n_per_million=1000)
        sym = text.symbols.copy()
        seg = sym[4000:9000].copy()
        seg[::83] = (seg[::83] + 2) % 4
        sym[60000:65000] = seg
        text = synth.Text(sym, text.records)
        reads = synth.concat_reads([synth.make_reads(text, 342, 700, 100, 0.015, True), synth.make_reads(text, 343, 300, 70, 0.02, True),
                                    synth.make_reads(text, 344, 100, 50, 0.01, True)])      # the 50-base reads are shorter than the seed
        synth.write_fasta(os.path.join(work, "t.fa"), text)
        rf = os.path.join(work, "r.fq")
        synth.write_reads(rf, reads, True)
        return os.path.join(work, "t.fa"), rf, ["-e", "4", "-Q", "33", "-l", "64"]
    raise KeyError(name)
