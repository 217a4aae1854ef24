"""CPU: the reference's rewritten pattern file (next-4; TemporaryFile.hpp:194-403, FastDecoder.hpp:66-130) as the host driver
writes and reads it, against fixtures made by the reference's own reorderPat (tests/golden/make_rewritten_golden.py)."""
import json
import os
import subprocess

import numpy as np
import pytest

from real_b200 import build as rbuild

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dump():
    if not os.path.exists(rbuild.LIB):
        rbuild.build()
    rbuild.build_host()
    return rbuild.HOST_DUMP


@pytest.mark.parametrize("kind", ["fa", "fq"])
def test_writer_and_reader_against_reference_fixture(dump, kind, tmp_path):
    z = np.load(os.path.join(GOLDEN, "rewritten_%s.npz" % kind))
    src = tmp_path / ("r." + kind)
    src.write_bytes(z["input"].tobytes())
    # writer: the bytes the reference's reorderPat wrote
    out = tmp_path / "ours.bin"
    subprocess.run([dump, "rewrite", str(src), "1" if kind == "fq" else "0", "0", str(out)], check=True, stdout=subprocess.PIPE)
    assert out.read_bytes() == z["rewritten"].tobytes()
    # reader, on the reference's file: order, bases and ids as the reference's decoder printed them
    ref = tmp_path / "ref.bin"
    ref.write_bytes(z["rewritten"].tobytes())
    r = json.loads(subprocess.run([dump, "unrewrite", str(ref)], check=True, stdout=subprocess.PIPE).stdout)
    assert r["fastq"] == (1 if kind == "fq" else 0)
    assert r["ids"] == [str(x) for x in z["ids"]]
    offs, mapped = r["offsets"], r["mapped"]
    got = ["".join("ACGTN"[c] for c in mapped[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]
    assert got == [str(x) for x in z["bases"]] and len(got) == 240
    # and it equals parsing the pattern file + the rewritten order
    p = json.loads(subprocess.run([dump, "reads", str(src), "1" if kind == "fq" else "0", "0", "1"], check=True, stdout=subprocess.PIPE).stdout)
    for k in ("ids", "offsets", "mapped", "quality"):
        assert p[k] == r[k], k


@pytest.mark.parametrize("kind", ["fa", "fq"])
def test_rewritten_file_straight_into_the_device_layout(dump, kind, tmp_path):
    """readRewrittenPacked (the ACGT sections copied as they stand) gives what packing the parsed reads gives."""
    z = np.load(os.path.join(GOLDEN, "rewritten_%s.npz" % kind))
    ref = tmp_path / "ref.bin"
    ref.write_bytes(z["rewritten"].tobytes())
    src = tmp_path / ("r." + kind)
    src.write_bytes(z["input"].tobytes())
    r = json.loads(subprocess.run([dump, "unrewrite_packed", str(ref)], check=True, stdout=subprocess.PIPE).stdout)
    p = json.loads(subprocess.run([dump, "reads", str(src), "1" if kind == "fq" else "0", "0", "1"], check=True, stdout=subprocess.PIPE).stdout)
    offs, mapped = p["offsets"], np.asarray(p["mapped"], dtype=np.int64)
    assert r["ids"] == p["ids"] and r["quality"] == p["quality"] and r["fastq"] == (1 if kind == "fq" else 0)
    assert r["lengths"] == [offs[i + 1] - offs[i] for i in range(len(offs) - 1)]
    packed = np.asarray(r["packed"], dtype=np.int64)
    for i in range(len(offs) - 1):
        m = mapped[offs[i]:offs[i + 1]]
        wild = bool((m > 3).any())
        assert r["wildcard"][i] == int(wild)
        got = packed[r["byte_offsets"][i]:r["byte_offsets"][i + 1]]
        want = np.zeros((len(m) + 3) // 4, dtype=np.int64)
        if not wild:
            for j, c in enumerate(m):
                want[j // 4] |= int(c) << (6 - 2 * (j % 4))
        assert np.array_equal(got, want), i


def test_reader_rejects_broken_files(dump, tmp_path):
    z = np.load(os.path.join(GOLDEN, "rewritten_fa.npz"))
    data = z["rewritten"].tobytes()
    for name, bad in (("truncated", data[:len(data) // 2]), ("magic", data[:8] + b"\x00\x00\x00\x07" + data[12:])):
        f = tmp_path / (name + ".bin")
        f.write_bytes(bad)
        p = subprocess.run([dump, "unrewrite", str(f)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode != 0 and b"rewritten pattern file" in p.stderr
