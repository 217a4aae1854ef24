"""CPU: the oracle's unit functions against known answers dumped from the reference's own objects
(tests/golden/kat_*.json.gz, produced by oracle/ref_harness/harness_kat.cpp)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as O
from real_b200 import matcher, synth
from util import load_kat

KATS = ["kat_l20", "kat_l32", "kat_l64"]


@pytest.fixture(scope="module", params=KATS)
def kat(request):
    d = load_kat(request.param)
    sym = np.asarray(d["symbols"], dtype=np.uint8)
    words, nmask = synth.pack_text(sym)
    d["_words"] = np.concatenate([words, np.zeros(2, np.uint64)])
    d["_nmask"] = np.concatenate([nmask, np.zeros(2, np.uint64)])
    d["_starts"] = np.asarray([s for _, s in d["ranges"]], dtype=np.uint64)
    return d


def test_text_packing_matches_reference_words(kat):
    ref = np.asarray(kat["textwords"], dtype=np.uint64)
    assert np.array_equal(kat["_words"][:ref.size], ref)


def test_text_word_wildcards_records(kat):
    L = O.lib()
    w, m, st = kat["_words"], kat["_nmask"], kat["_starts"]
    nrec = st.size - 1
    for q in kat["textqueries"]:
        i, l = q["i"], q["l"]
        assert L.oracle_text_word(w.ctypes.data, i, l) == q["word"]
        assert L.oracle_dontcare_free(m.ctypes.data, i, l) == q["dcf"]
        assert L.oracle_position_to_range(st.ctypes.data, nrec, i) == q["range"]
        if q["inrange"]:
            assert L.oracle_dontcare_free(m.ctypes.data, i, q["patl"]) == q["dcf_patl"]
        assert L.oracle_position_valid(st.ctypes.data, nrec, i, q["patl"]) == q["valid"]


def test_diffcountpair(kat):
    L = O.lib()
    for a, b, c in kat["diffcountpair64"]:
        assert L.oracle_diffcount64(a, b) == c
    for a, b, c in kat["diffcountpair32"]:
        assert L.oracle_diffcount64(a, b) == c


def test_signatures_and_rest_words(kat):
    L = O.lib()
    seedl = kat["seedl"]
    n_usable = 0
    for r in kat["reads"]:
        mapped = np.asarray(r["mapped"], dtype=np.uint8)
        assert np.array_equal(synth.revcomp_mapped(mapped), np.asarray(r["transposed"], dtype=np.uint8))
        if not r["usable"]:
            continue
        n_usable += 1
        m = (C.c_uint32 * 4)()
        s = (C.c_uint64 * 6)()
        L.oracle_fragments(mapped.ctypes.data, seedl, m)
        assert list(m) == r["m"]
        L.oracle_pair_signatures(m, seedl, s)
        assert list(s) == r["fw"]
        L.oracle_reverse_fragments(mapped.ctypes.data, seedl, m)
        assert list(m) == r["im"]
        L.oracle_pair_signatures(m, seedl, s)
        assert list(s) == r["rv"]
        out = np.zeros(16, dtype=np.uint64)
        nw = L.oracle_rest_words(mapped.ctypes.data, r["len"], seedl, 0, out.ctypes.data)
        assert list(out[:nw]) == r["rest_straight"]
        nw = L.oracle_rest_words(mapped.ctypes.data, r["len"], seedl, 1, out.ctypes.data)
        assert list(out[:nw]) == r["rest_reverse"]
    assert n_usable > 10


def test_rest_distance_and_scores(kat):
    L = O.lib()
    seedl = kat["seedl"]
    ll = np.asarray(kat["ll_bits"], dtype=np.uint64).view(np.float64)
    w = kat["_words"]
    checked = 0
    for r in kat["reads"]:
        if not r["usable"]:
            continue
        mapped = np.asarray(r["mapped"], dtype=np.uint8)
        q = np.asarray(r["quality"], dtype=np.uint8)
        rs = np.zeros(16, dtype=np.uint64)
        rr = np.zeros(16, dtype=np.uint64)
        L.oracle_rest_words(mapped.ctypes.data, r["len"], seedl, 0, rs.ctypes.data)
        L.oracle_rest_words(mapped.ctypes.data, r["len"], seedl, 1, rr.ctypes.data)
        for at in r["at"]:
            pos = at["pos"]
            assert L.oracle_rest_distance(rs.ctypes.data, r["len"], seedl, w.ctypes.data, pos + r["straighttextrestoffset"]) == at["rest_fw"]
            assert L.oracle_rest_distance(rr.ctypes.data, r["len"], seedl, w.ctypes.data, pos) == at["rest_rv"]
            sf = np.float32(L.oracle_compute_score(ll.ctypes.data, w.ctypes.data, mapped.ctypes.data, q.ctypes.data, pos, r["len"], 0))
            sr = np.float32(L.oracle_compute_score(ll.ctypes.data, w.ctypes.data, mapped.ctypes.data, q.ctypes.data, pos, r["len"], 1))
            assert int(sf.view(np.uint32)) == at["score_fw"]
            assert int(sr.view(np.uint32)) == at["score_rv"]
            checked += 1
    assert checked > 20


def test_scoring_table_bit_exact(kat):
    """Scoring::init/getScore restated twice (oracle C, product host layer) == the reference's table."""
    ref = np.asarray(kat["ll_bits"], dtype=np.uint64)
    p = kat["scoring_params"]
    got_c = O.build_ll(*[p[0], p[1], p[2], p[3], p[4]]).view(np.uint64)
    assert np.array_equal(got_c, ref)
    got_py = matcher.scoring_table(similarity=p[0], gc=p[1], trans=p[2], err=p[3], gcmut_bias=p[4]).view(np.uint64)
    assert np.array_equal(got_py, ref)


def test_filter_mult(kat):
    ref = np.asarray([kat["filter_mult_bits"]], dtype=np.uint64).view(np.float64)[0]
    assert O.filter_mult(5, 2) == ref
    o = matcher.RealOptions()
    assert o.filter_mult == ref
