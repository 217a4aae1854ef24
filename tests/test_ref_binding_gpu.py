"""GPU: the reference-side binding of INTEGRATION.md section 3, compiled and run.

oracle/_ref/real_bound (oracle/Makefile, target ref_bound) is the REFERENCE's own real.cpp, option parser, pattern
rewriting, readers, text loader and output code, linked with oracle/ref_binding/bound_{unique,all}.cpp -- specialisations of
EnumerateUniqueMatches / EnumerateAllMatches ::doMatching whose text-block loop is the C ABI of include/real_gpu.h -- and
with libreal_gpu.so.  Its output files must be the stock binary's (tests/golden/cli_*.txt).  Skipped where the binary was
not built (it needs /root/reference at build time)."""
import os
import subprocess

import pytest

from oracle import oracle_py as O
from cli_cases import CASES, make_case
from real_b200 import build as rbuild

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(O.REF_BOUND), reason="oracle/_ref/real_bound not built (needs /root/reference)")]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", CASES)
def test_bound_reference_binary_writes_the_stock_output(name, tmp_path):
    targ, rf, flags = make_case(name, str(tmp_path))
    out = tmp_path / "out.txt"
    p = subprocess.run([O.REF_BOUND, "-t", targ, "-p", rf, "-o", str(out)] + flags, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0 and "libreal_gpu.so" in p.stderr, p.stderr[-2000:]
    want = open(os.path.join(GOLDEN, "cli_%s.txt" % name)).read()
    assert len(want.splitlines()) > 100
    assert out.read_text() == want


def test_bound_reference_binary_match_all(tmp_path):
    """-u 0 through the bound reference binary: the same lines as this repo's own command line (which is checked against the
    oracle in tests/test_cli_gpu.py); FASTA reads, because the stock -u 0 dispatch parses FASTQ files with the FASTA reader
    (real.cpp:325-328)."""
    from real_b200 import synth
    rbuild.build()
    rbuild.build_host()
    text = synth.make_text(421, 60000, nrecords=2, n_per_million=1000)
    reads = synth.make_reads(text, 422, 600, 64, 0.02, False)
    synth.write_fasta(str(tmp_path / "t.fa"), text)
    synth.write_reads(str(tmp_path / "r.fa"), reads, False)
    outs = []
    for exe in (O.REF_BOUND, rbuild.HOST_BIN):
        out = tmp_path / ("o_%s.txt" % os.path.basename(exe))
        p = subprocess.run([exe, "-t", str(tmp_path / "t.fa"), "-p", str(tmp_path / "r.fa"), "-o", str(out), "-u", "0", "-e", "4", "-q", "1"],
                           cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(out.read_text())
    assert len(outs[0].splitlines()) > 500 and outs[0] == outs[1]
