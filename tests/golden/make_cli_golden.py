"""Generates tests/golden/cli_*.txt with the STOCK reference command line (oracle/_ref/real, built by
`make -C oracle ref_real` from the reference's own sources).  Only matchUnique runs are used: the stock
matchAll output is truncated and thread dependent.  Inputs are re-created by the
tests from the same seeds (tests/test_cli_gpu.py:cli_case)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cli_cases import CASES, make_case  # noqa: E402

REAL = os.path.join(ROOT, "oracle", "_ref", "real")
OUT = os.path.dirname(os.path.abspath(__file__))

for name in CASES:
    with tempfile.TemporaryDirectory() as work:
        targ, rf, flags = make_case(name, work)
        out = os.path.join(work, "out.txt")
        p = subprocess.run([REAL, "-t", targ, "-p", rf, "-o", out] + flags, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        lines = open(out).read()
        open(os.path.join(OUT, "cli_%s.txt" % name), "w").write(lines)
        print(name, len(lines.splitlines()), "lines")

# patterns from standard input (-p -, RealOptions.cpp:418-426: the type is taken from the first byte, rewriting is forced on)
STDIN_CASES = ["unique_fq_R0", "unique_fa_R1"]
for name in STDIN_CASES:
    with tempfile.TemporaryDirectory() as work:
        targ, rf, flags = make_case(name, work)
        out = os.path.join(work, "out.txt")
        with open(rf, "rb") as fin:
            p = subprocess.run([REAL, "-t", targ, "-p", "-", "-o", out] + flags, cwd=work, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        lines = open(out).read()
        open(os.path.join(OUT, "cli_stdin_%s.txt" % name), "w").write(lines)
        print("stdin", name, len(lines.splitlines()), "lines")
