"""CPU: the C++ host layer (real_b200/host) -- FASTA/FASTQ parsing, text packing, record table, scoring table
and option handling -- against the reference's own outputs in the KAT fixtures and against the layouts of
real_b200/synth.py.  Uses real_b200/bin/real_host_dump (no GPU calls)."""
import json
import os
import subprocess

import numpy as np
import pytest

from real_b200 import build as rbuild
from real_b200 import synth
from util import load_kat


@pytest.fixture(scope="module")
def dump():
    rbuild.build()
    rbuild.build_host()

    def run(*args, ok=True, env=None):
        p = subprocess.run([rbuild.HOST_DUMP] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                           env=dict(os.environ, **env) if env else None)
        if ok:
            assert p.returncode == 0, p.stderr
            return json.loads(p.stdout)
        return p
    return run


def test_text_loader_matches_reference_layout(dump, tmp_path):
    kat = load_kat("kat_l32")
    tk = synth.make_text(107, 3000, nrecords=3, n_per_million=30000)
    f = tmp_path / "t.fa"
    synth.write_fasta(str(f), tk)
    d = dump("text", f)
    assert d["n"] == kat["n"]
    assert [[a, b] for a, b in d["ranges"]] == kat["ranges"]
    ref = kat["textwords"]
    assert d["words"][:len(ref)] == ref
    w, m = tk.packed()
    assert d["nmask"][:m.size] == [int(x) for x in m]


def test_text_loader_quirks(dump, tmp_path):
    # lower case and IUPAC codes are dropped, '>' inside a line starts a header (countReads.cpp:44-76)
    f = tmp_path / "q.fa"
    f.write_bytes(b"> one two\nACGTacgtNNRYAC\nGT>mid line\nTTTT\n\nGG\n")
    d = dump("text", f)
    assert d["ranges"] == [[" one two", 0], ["mid line", 10], ["terminal", 16]]
    sym = synth.unpack_text(np.asarray(d["words"], dtype=np.uint64), 16, np.asarray(d["nmask"], dtype=np.uint64))
    assert "".join("ACGTN"[c] for c in sym) == "ACGTNNACGTTTTTGG"


def test_scoring_table_bits(dump):
    kat = load_kat("kat_l32")
    p = kat["scoring_params"]
    assert dump("ll", *["%.17g" % x for x in p]) == kat["ll_bits"]
    assert dump("ll", 0.9, 0.5, 0.6, 0.01, 1.5) == [int(x) for x in __import__("oracle.oracle_py", fromlist=["x"]).build_ll(0.9, 0.5, 0.6, 0.01, 1.5).view(np.uint64)]


@pytest.mark.parametrize("fastq", [False, True])
def test_read_parsers(dump, tmp_path, fastq):
    kat = load_kat("kat_l32")
    tk = synth.make_text(107, 3000, nrecords=3, n_per_million=30000)
    rk = synth.concat_reads([synth.make_reads(tk, 12, 12, L, 0.03, True) for L in (32, 33, 36, 64, 96, 100, 150, 250)])
    f = tmp_path / ("r.fq" if fastq else "r.fa")
    synth.write_reads(str(f), rk, fastq)
    d = dump("reads", f, int(fastq), 0, 0)
    assert d["ids"] == rk.ids
    assert d["offsets"] == [int(x) for x in rk.offsets]
    assert d["mapped"] == [int(x) for x in rk.mapped]
    for r, ref in enumerate(kat["reads"]):
        b, e = d["offsets"][r], d["offsets"][r + 1]
        assert d["mapped"][b:e] == ref["mapped"]
        if fastq:
            assert d["quality"][b:e] == ref["quality"]
    if fastq:
        assert d["qoff"] == 33


def test_fastq_multiline_and_offset_detection(dump, tmp_path):
    f = tmp_path / "m.fq"
    f.write_bytes(b"@r1 x\nACGT\nAC\n+r1\nhhhh\nhh\n@r2\nNNAC\n+\n~~~~\n")
    d = dump("reads", f, 1, 0, 0)
    assert d["qoff"] == 64 and d["ids"] == ["r1 x", "r2"]
    assert d["mapped"] == [0, 1, 2, 3, 0, 1, 4, 4, 0, 1]
    assert d["quality"] == [40] * 6 + [62] * 4


def test_rewrite_order(dump, tmp_path):
    # -R 1: grouped by length, wildcard-free reads first, file order inside a group (ReorderFastA.hpp, TemporaryFile.hpp)
    f = tmp_path / "o.fa"
    f.write_bytes(b">a\nACGTAC\n>b\nACG\n>c\nACNTAC\n>d\nAAA\n>e\nGGGGGG\n>f\nNNN\n")
    d = dump("reads", f, 0, 0, 1)
    assert d["ids"] == ["b", "d", "f", "a", "e", "c"]


def test_options(dump, tmp_path):
    r = tmp_path / "r.fa"
    r.write_bytes(b">x\nACGT\n")
    d = dump("opts", "-t", "t.fa", "-p", r, "-o", "o", "-e", "77", "-s", "9", "-l", "70", "-u", "0", "-bogus", "-filter_level", "3")
    assert (d["totalkmax"], d["seedkmax"], d["seedl"], d["match_unique"], d["filter_level"], d["fastq"]) == (15, 2, 64, 0, 3, 0)
    assert np.asarray([d["filter_mult_bits"]], dtype=np.uint64).view(np.float64)[0] == 2 * 15 / 70.0
    p = dump("opts", "-p", r, "-o", "o", ok=False)
    assert p.returncode == 1 and "Mandatory argument -t" in p.stderr
    p = dump("opts", "-t", "t", "-p", r, "-o", ok=False)
    assert p.returncode == 1 and "Parameter for argument -o is missing." in p.stderr
    q = tmp_path / "r.fq"
    q.write_bytes(b"@x\nACGT\n+\nIIII\n")
    assert dump("opts", "-t", "t", "-p", q, "-o", "o")["fastq"] == 1


def test_file_list(dump, tmp_path):
    (tmp_path / "d").mkdir()
    (tmp_path / "d" / "a.fa").write_bytes(b">a\nA\n")
    (tmp_path / "d" / "b.txt").write_bytes(b">a\nA\n")
    (tmp_path / "d" / "sub").mkdir()
    (tmp_path / "d" / "sub" / "c.fa").write_bytes(b">a\nA\n")
    got = sorted(os.path.basename(x) for x in dump("files", tmp_path / "d"))
    assert got == ["a.fa", "c.fa"]
    assert dump("files", tmp_path / "d" / "b.txt") == []


def test_text_loader_reference_quirk_fixture(dump, tmp_path):
    """host getText against the reference's own getText on the awkward files of tests/golden/text_quirks.npz"""
    from util import text_quirk_cases
    for name, data, symbols, starts, names in text_quirk_cases():
        f = tmp_path / (name + ".fa")
        f.write_bytes(data)
        d = dump("text", f)
        assert d["n"] == symbols.size, name
        assert [r[1] for r in d["ranges"]] == [int(x) for x in starts], name
        assert [r[0].encode("latin-1") for r in d["ranges"][:-1]] == names, name
        sym = synth.unpack_text(np.asarray(d["words"], dtype=np.uint64), symbols.size, np.asarray(d["nmask"], dtype=np.uint64))
        assert np.array_equal(sym, symbols), name


def _awkward_read_files():
    """(name, fastq, bytes): files on which a guessed record start is wrong somewhere"""
    rng = np.random.RandomState(99)
    acgt = np.frombuffer(b"ACGTN", dtype=np.uint8)

    def seq(n):
        return acgt[rng.randint(0, 5 if rng.rand() < 0.2 else 4, n)].tobytes()

    fa = b"junk in front\n"
    for i in range(120):
        s = seq(rng.randint(1, 90))
        cut = rng.randint(0, len(s) + 1)
        fa += b">id%d >not a marker\n" % i + s[:cut] + (b"\n" if i % 3 else b"\r\n") + s[cut:] + (b"" if i % 7 == 0 else b"\n")
    fa += b">last without newline"
    fq = b""
    for i in range(150):
        n = rng.randint(1, 80)
        s = seq(n)
        q = bytes(rng.randint(33, 74, n).astype(np.uint8))
        if i % 5 == 0:
            q = b"@" + q[1:]                      # a quality line that starts like a record
        if i % 11 == 0 and n > 3:
            q = q[:1] + b"+" + q[2:]
        if i % 4 == 0 and n > 10:                  # sequence and qualities over two lines
            fq += b"@r%d\n" % i + s[:5] + b"\n" + s[5:] + b"\n+r%d\n" % i + q[:7] + b"\n" + q[7:] + b"\n"
        else:
            fq += b"@r%d\n" % i + s + b"\n+\n" + q + b"\n"
    fq_trunc = fq + b"@cut\nACGTACGT\n+\nIII"      # the reader rejects the last record
    return [("fa", 0, fa), ("fq", 1, fq), ("fq_trunc", 1, fq_trunc)]


@pytest.mark.parametrize("chunk", [1, 37, 256, 1000])
def test_parallel_read_parser_equals_serial(dump, tmp_path, chunk):
    """the team parser (ranges of REAL_PARSE_CHUNK bytes, guessed record starts, checked hand-over) against the one-range parse"""
    for name, fastq, data in _awkward_read_files():
        f = tmp_path / (name + (".fq" if fastq else ".fa"))
        f.write_bytes(data)
        serial = dump("reads", f, fastq, 33 if fastq else 0, 0, env={"REAL_PARSE_CHUNK": str(1 << 40)})
        team = dump("reads", f, fastq, 33 if fastq else 0, 0, env={"REAL_PARSE_CHUNK": str(chunk)})
        assert len(serial["ids"]) > 100, name
        assert team == serial, name


def test_host_text_loader_fuzz_against_restatement(dump, tmp_path):
    """host getText (REAL_TEXT_LOADER=host path of the command line) against the restated loader on random byte strings"""
    from oracle import oracle_py as O
    rng = np.random.RandomState(4242)
    for k, alphabet in enumerate([b"ACGTN" * 8 + b"acgtn>\n\r x", b"ACGT" * 30 + b">\n\n", b">\nA"]):
        a = np.frombuffer(alphabet, dtype=np.uint8)
        data = a[rng.randint(0, a.size, 3000 + 517 * k)].tobytes()
        f = tmp_path / ("fz%d.fa" % k)
        f.write_bytes(data)
        symbols, names, starts = O.fasta_text(data)
        d = dump("text", f)
        assert d["n"] == symbols.size
        assert [r[1] for r in d["ranges"]] == [int(x) for x in starts]
        assert [r[0].encode("latin-1") for r in d["ranges"][:-1]] == names
        sym = synth.unpack_text(np.asarray(d["words"], dtype=np.uint64), symbols.size, np.asarray(d["nmask"], dtype=np.uint64))
        assert np.array_equal(sym, symbols)


def test_fasta_writer_roundtrip():
    """synth.write_fasta_file (the generator of the GPU ingest tests) read back by the restated loader"""
    import io
    from oracle import oracle_py as O
    t = synth.make_text(31, 30007, nrecords=5, n_per_million=2000)
    f = io.BytesIO()
    synth.write_fasta_file(f, t)
    s, names, st = O.fasta_text(f.getvalue())
    assert np.array_equal(s, t.symbols) and np.array_equal(st, t.record_starts)
    assert [x.decode() for x in names] == [a for a, _ in t.records]


@pytest.mark.parametrize("n,nreads,frac,scores", [(3_000_000, 2000, 0.0062, 1), (3_000_000, 2000, 0.0063, 0), (1_234_567, 777, 0.0068, 1), (1_234_567, 777, 0.75, 0)])
def test_memory_planner_equals_stock_binary(dump, tmp_path, n, nreads, frac, scores):
    """The text-block size (n_list) decides the visiting order of the order-dependent folds: the host driver's planner against
    the 'Using n_list' line of the stock binary on texts that need several blocks (small -f).  Needs oracle/_ref/real (built
    where /root/reference exists); the figure depends on the machine's MemTotal, so it is compared live."""
    stock = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "real")
    if not os.path.exists(stock):
        pytest.skip("stock binary not built")
    text = synth.make_text(5, n, nrecords=3, n_per_million=500)
    reads = synth.make_reads(text, 6, nreads, 50, 0.01, fastq=True)
    synth.write_fasta(str(tmp_path / "t.fa"), text)
    synth.write_reads(str(tmp_path / "r.fq"), reads, True)
    args = ["-t", str(tmp_path / "t.fa"), "-p", str(tmp_path / "r.fq"), "-o", str(tmp_path / "o.txt"), "-u", "1", "-q", str(scores), "-Q", "33", "-R", "0", "-f", str(frac)]
    p = subprocess.run([stock] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    want = [l for l in p.stderr.splitlines() if l.startswith("Using n_list")]
    assert want, p.stderr[-500:]
    q = dump("plan", text.n, nreads, *args, ok=False)
    assert q.returncode == 0, q.stderr[-500:]
    got = [l for l in q.stderr.splitlines() if l.startswith("Using n_list")]
    assert got == want[:1]
    assert json.loads(q.stdout)["n_list"] == int(want[0].split("=")[1])
