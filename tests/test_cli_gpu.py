"""GPU: the `real` command line on the GPU path (real_b200/bin/real) against output files written by the
STOCK reference binary (tests/golden/cli_*.txt, made by tests/golden/make_cli_golden.py)."""
import os
import subprocess

import pytest

from real_b200 import build as rbuild
from cli_cases import CASES, make_case

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("formatter", ["device", "host", "device+hostreads"])
@pytest.mark.parametrize("name", CASES)
def test_cli_output_identical_to_stock_real(name, formatter, tmp_path):
    """The lines are formatted on the device (K8, real_gpu_format_*: the default) or by the host team (REAL_FORMAT=host), FASTA
    pattern files are parsed on the device (real_gpu_set_reads_fasta: the default with the device formatter) or by the host team
    (REAL_READS_LOADER=host): every combination writes the stock binary's bytes."""
    rbuild.build()
    rbuild.build_host()
    targ, rf, flags = make_case(name, str(tmp_path))
    out = tmp_path / "out.txt"
    p = subprocess.run([rbuild.HOST_BIN, "-t", targ, "-p", rf, "-o", str(out)] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, REAL_STRICT_EXIT="1", REAL_FORMAT=formatter.split("+")[0], REAL_READS_LOADER="host" if "hostreads" in formatter else "device"))
    assert p.returncode == 0, p.stderr[-2000:]
    want = open(os.path.join(GOLDEN, "cli_%s.txt" % name)).read()
    got = out.read_text()
    assert len(want.splitlines()) > 100
    assert got == want


@pytest.mark.parametrize("name", CASES[:2])
def test_cli_output_identical_with_threaded_formatter(name, tmp_path):
    """The output is formatted by several host threads, wave by wave, and written in order: tiny waves and 5 threads
    must give the stock binary's bytes too."""
    rbuild.build()
    rbuild.build_host()
    targ, rf, flags = make_case(name, str(tmp_path))
    out = tmp_path / "out.txt"
    p = subprocess.run([rbuild.HOST_BIN, "-t", targ, "-p", rf, "-o", str(out)] + flags + ["-T", "5"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, REAL_STRICT_EXIT="1", REAL_FORMAT_CHUNK="7", REAL_FORMAT="host"))
    assert p.returncode == 0, p.stderr[-2000:]
    assert out.read_text() == open(os.path.join(GOLDEN, "cli_%s.txt" % name)).read()


def test_cli_match_all_complete_output(tmp_path):
    """matchAll through the CLI: every oracle hit is printed once (the stock CLI truncates)."""
    import numpy as np
    from oracle import oracle_py as O
    from real_b200 import synth
    rbuild.build()
    rbuild.build_host()
    text = synth.make_text(401, 50000, nrecords=2)
    reads = synth.make_reads(text, 402, 500, 60, 0.02, True)
    synth.write_fasta(str(tmp_path / "t.fa"), text)
    synth.write_reads(str(tmp_path / "r.fq"), reads, True)
    out = tmp_path / "o.txt"
    p = subprocess.run([rbuild.HOST_BIN, "-t", str(tmp_path / "t.fa"), "-p", str(tmp_path / "r.fq"), "-o", str(out), "-u", "0", "-e", "4", "-q", "1", "-Q", "33"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, REAL_STRICT_EXIT="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    ref = O.match_all(text, reads, totalkmax=4, scores=True)
    rows = [l.split("\t") for l in out.read_text().splitlines()]
    assert len(rows) == len(ref)
    starts = text.record_starts
    want = sorted((reads.ids[int(h["patid"])], "-" if h["inverted"] else "+", text.records[int(h["frag"])][0], int(h["pos"]) - int(starts[int(h["frag"])]) + 1,
                   int(h["k"]), "%g" % np.float32(h["score"])) for h in ref)
    got = sorted((r[0], r[6], r[7], int(r[8]), int(r[10]), r[2]) for r in rows)
    assert got == want


@pytest.mark.parametrize("name,ngpus", [("unique_fa_R1", 2), ("unique_fq_R0", 3), ("unique_dir_ragged", 2), ("unique_fq_scores_default", 2)])
def test_cli_several_handles_identical_to_stock_real(name, ngpus, tmp_path):
    """REAL_GPUS=N: one process, N handles (bucket shards; here all on the one device of the test box), matchUnique folded over
    peer memory (real_gpu_fold_unique_group).  The bytes written are the stock binary's.  With scores the fold depends on the
    visiting order: the request is ignored (one handle) and said so."""
    rbuild.build()
    rbuild.build_host()
    targ, rf, flags = make_case(name, str(tmp_path))
    out = tmp_path / "out.txt"
    p = subprocess.run([rbuild.HOST_BIN, "-t", targ, "-p", rf, "-o", str(out)] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, REAL_STRICT_EXIT="1", REAL_GPUS=str(ngpus)))
    assert p.returncode == 0, p.stderr[-2000:]
    assert out.read_text() == open(os.path.join(GOLDEN, "cli_%s.txt" % name)).read()
    assert ("ignored" in p.stderr) == (name == "unique_fq_scores_default")


@pytest.mark.parametrize("name", ["unique_fq_R0", "unique_fa_R1"])
def test_cli_patterns_from_stdin(name, tmp_path):
    """-p - : the pattern file comes from standard input (type from its first byte, rewriting forced on, RealOptions.cpp:418-426)."""
    rbuild.build()
    rbuild.build_host()
    targ, rf, flags = make_case(name, str(tmp_path))
    out = tmp_path / "out.txt"
    with open(rf, "rb") as fin:
        p = subprocess.run([rbuild.HOST_BIN, "-t", targ, "-p", "-", "-o", str(out)] + flags, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           env=dict(os.environ, REAL_STRICT_EXIT="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    want = open(os.path.join(GOLDEN, "cli_stdin_%s.txt" % name)).read()
    assert len(want.splitlines()) > 100 and out.read_text() == want
    if name == "unique_fq_R0":
        assert b"switching on pattern rewriting" in p.stderr


def test_cli_match_all_several_handles(tmp_path):
    """matchAll with three handles: the rows of the handles merged on the host are the rows of one handle, byte for byte."""
    from real_b200 import synth
    rbuild.build()
    rbuild.build_host()
    text = synth.make_text(411, 70000, nrecords=3, n_per_million=1500)
    sym = text.symbols.copy()
    sym[40000:43000] = sym[1000:4000]
    text = synth.Text(sym, text.records)
    reads = synth.make_reads(text, 412, 900, 70, 0.02, True)
    synth.write_fasta(str(tmp_path / "t.fa"), text)
    synth.write_reads(str(tmp_path / "r.fq"), reads, True)
    outs = []
    for n in (1, 3):
        out = tmp_path / ("o%d.txt" % n)
        p = subprocess.run([rbuild.HOST_BIN, "-t", str(tmp_path / "t.fa"), "-p", str(tmp_path / "r.fq"), "-o", str(out), "-u", "0", "-e", "4", "-q", "1", "-Q", "33"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, REAL_STRICT_EXIT="1", REAL_GPUS=str(n)))
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(out.read_text())
    assert len(outs[0].splitlines()) > 900 and outs[0] == outs[1]


@pytest.mark.parametrize("name", ["unique_fa_R1", "unique_fq_scores_default"])
def test_cli_kept_rewritten_pattern_file(name, tmp_path):
    """next-4: REAL_KEEP_REWRITTEN keeps the reference's rewritten pattern file (which the stock binary deletes); given back as
    -p it is read instead of parsing the FASTA/FASTQ file again, and the output is the stock binary's."""
    rbuild.build()
    rbuild.build_host()
    targ, rf, flags = make_case(name, str(tmp_path))
    want = open(os.path.join(GOLDEN, "cli_%s.txt" % name)).read()
    kept = tmp_path / "kept.bin"
    out = tmp_path / "out.txt"
    p = subprocess.run([rbuild.HOST_BIN, "-t", targ, "-p", rf, "-o", str(out)] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, REAL_STRICT_EXIT="1", REAL_KEEP_REWRITTEN=str(kept)))
    assert p.returncode == 0, p.stderr[-2000:]
    assert out.read_text() == want and kept.stat().st_size > 1000 and kept.read_bytes()[0] == 0
    out2 = tmp_path / "out2.txt"
    flags2 = [f for f in flags]
    p = subprocess.run([rbuild.HOST_BIN, "-t", targ, "-p", str(kept), "-o", str(out2)] + flags2, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, REAL_STRICT_EXIT="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    assert "rewritten pattern file" in p.stderr
    assert out2.read_text() == want
