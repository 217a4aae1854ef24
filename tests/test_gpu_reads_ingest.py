"""K0 in pattern-file mode: FASTA reads parsed, packed and ordered on the device (real_gpu_set_reads_fasta) against the host
driver's reader (real_b200/bin/real_host_dump reads ..., the restatement of FastAReader / Pattern::computeMapped /
reorderFastA that the KAT fixtures and the command-line goldens pin to the reference)."""
import json
import subprocess

import numpy as np
import pytest

from real_b200 import build as rbuild, lib as rlib, matcher, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dump():
    rbuild.build()
    rbuild.build_host()
    return rbuild.HOST_DUMP


def host_reads(dump, path, rewrite):
    return json.loads(subprocess.run([dump, "reads", str(path), "0", "0", "1" if rewrite else "0"], check=True, stdout=subprocess.PIPE).stdout)


def check_against_host(dump, tmp_path, data: bytes, rewrite: bool, name: str):
    f = tmp_path / (name + ".fa")
    f.write_bytes(data)
    if not data.lstrip(b"\x00 \t\r\n\x0b\x0c").startswith(b">") and b">" not in data:
        want = {"ids": [], "offsets": [0], "mapped": []}
    else:
        want = host_reads(dump, f, rewrite)
    h = rlib.Handle(totalkmax=3)
    try:
        n = h.set_reads_fasta(data, rewrite_order=rewrite)
        assert n == len(want["ids"]), (name, n, len(want["ids"]))
        if n == 0:
            return
        ids = h.get_read_ids()
        assert [x.decode("latin1") for x in ids] == [x for x in want["ids"]], name
        ln, fl = h.get_read_table()
        offs = want["offsets"]
        assert [int(x) for x in ln] == [offs[i + 1] - offs[i] for i in range(n)], name
        got = h.get_reads_mapped()
        m = want["mapped"]
        for i in range(n):
            w = np.asarray(m[offs[i]:offs[i + 1]], dtype=np.int64)
            assert bool(fl[i]) == bool((w > 3).any()), (name, i)
            assert np.array_equal(got[i], np.where(w > 3, 0, w)), (name, i)
    finally:
        h.close()


QUIRKS = [
    b">r1 first\nACGTACGTAC\n>r2\nACGNNACGTA\n>r3 x\nACGTA\n>r4\nTTTTTGGGGG\n",
    b"garbage ACGT before\n>a\nAC GT\tAC\r\nGT\n>b>c\nacgtACGTnN\n>empty\n>d\nA>e\nCCCC",                   # blanks, CR, lower case, '>' in an id, '>' behind a base
    b">only one, no newline at the end\nACGTTGCA",
    b">a\nACGT\n>unclosed id at the end of the file",
    b"no marker at all\nACGT\n",
    b"",
    b">\n\n>\nA\n",
    b">x\n" + b"ACGT" * 20000 + b"\n>y\n" + b"T" * 33 + b"\n",
]


@pytest.mark.parametrize("rewrite", [False, True])
def test_quirks(dump, tmp_path, rewrite):
    for k, data in enumerate(QUIRKS):
        if rewrite and data.count(b"ACGT" * 20000):
            continue                      # (reads above 65535 bases are refused)
        if data.count(b"ACGT" * 20000):
            h = rlib.Handle()
            try:
                with pytest.raises(rlib.RealGpuError):
                    h.set_reads_fasta(data)
            finally:
                h.close()
            continue
        check_against_host(dump, tmp_path, data, rewrite, "quirk%d" % k)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_against_host_reader(dump, tmp_path, seed):
    rng = np.random.RandomState(seed)
    parts = []
    if seed == 2:
        parts.append(b"ACGT junk in front of the first marker\n")
    nreads = [3000, 700, 12000][seed - 1]
    for i in range(nreads):
        L = int(rng.choice([0, 1, 3, 4, 5, 31, 32, 33, 36, 50, 63, 64, 65, 100, 127, 128, 250]))
        alphabet = b"ACGT" if rng.rand() < 0.8 else b"ACGTNacgtRYKM-*"
        s = bytes(alphabet[x] for x in rng.randint(0, len(alphabet), L))
        ident = b"read_%d" % i
        if i % 7 == 0:
            ident += b" some description > with a marker"
        if i % 501 == 0:
            ident += b"x" * int(rng.randint(1, 9000))                 # ids longer than a tile of the parser's neighbours
        width = int(rng.choice([7, 60, 61, 1000]))
        body = b"\n".join(s[a:a + width] for a in range(0, max(L, 1), width))
        if i % 11 == 0:
            body = body.replace(b"A", b"A ", 2).replace(b"\n", b"\r\n", 1)
        parts.append(b">" + ident + b"\n" + body + (b"\n" if (i % 13 or i == nreads - 1) else b""))
    data = b"".join(parts)
    for rewrite in (False, True):
        check_against_host(dump, tmp_path, data, rewrite, "fuzz%d_%d" % (seed, int(rewrite)))


def test_matching_after_the_device_reader_equals_the_host_path(dump, tmp_path):
    text = synth.make_text(31, 400_000, nrecords=3, n_per_million=400)
    reads = synth.concat_reads([synth.make_reads(text, 32, 5000, 100, 0.01, False), synth.make_reads(text, 33, 2000, 64, 0.02, False)])
    f = tmp_path / "r.fa"
    synth.write_reads(str(f), reads, False)
    data = f.read_bytes()
    want = host_reads(dump, f, True)
    words, nmask = text.packed()
    names = [("rec%d" % i).encode() for i in range(len(text.record_starts) - 1)]
    out = []
    for device_reader in (False, True):
        h = rlib.Handle(totalkmax=4)
        try:
            if device_reader:
                assert h.set_reads_fasta(data, rewrite_order=True) == reads.nreads
            else:
                h.set_reads(np.asarray(want["mapped"], dtype=np.uint8), np.asarray(want["offsets"], dtype=np.uint64))
                h.set_read_ids([x.encode("latin1") for x in want["ids"]])
            h.set_text(words, nmask, text.n, text.record_starts)
            h.set_record_names(names, text.record_starts[:-1])
            h.match_unique()
            lines, nl = h.format_unique()
            assert nl > 6000
            out.append(lines)
        finally:
            h.close()
    assert out[0] == out[1]
