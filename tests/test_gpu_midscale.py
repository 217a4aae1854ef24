"""GPU: mid-scale inputs (32 Mbp x 300-500 k reads) against digests of the REFERENCE's own result (tests/golden/midscale.json,
made by tests/golden/make_midscale_golden.py through oracle/_ref/ref_harness).  The inputs are regenerated from the
committed seeds (real_b200/synth.py is counter based); the comparison is the sha256 of the canonical result."""
import importlib.util
import json
import os

import numpy as np
import pytest

from real_b200 import matcher

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _maker():
    spec = importlib.util.spec_from_file_location("make_midscale_golden", os.path.join(GOLDEN, "make_midscale_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,packed,nranks", [("midscale_unique", False, 1), ("midscale_unique", True, 4), ("midscale_all", False, 1)])
def test_midscale_digest_of_the_reference(name, packed, nranks):
    mk = _maker()
    with open(os.path.join(GOLDEN, "midscale.json")) as f:
        doc = json.load(f)[name]
    c = doc["params"]
    text, reads = mk.make_inputs(c)
    words, nmask = text.packed()
    opts = matcher.RealOptions(totalkmax=c["e"], scores=c["scores"])
    if c["mode"] == "all":
        m = matcher.AllMatcher(opts)
        try:
            m.set_reads(reads.mapped, reads.offsets, reads.quality)
            m.set_text(words, nmask, text.n, text.record_starts)
            got = m.match()
        finally:
            m.close()
        assert len(got) == doc["counts"]["rows"]
        assert mk.canonical_digest(c, got) == doc["sha256"]
        return
    from real_b200 import dist as rdist
    from real_b200 import lib as rlib
    from real_b200 import synth
    ms = [matcher.UniqueMatcher(opts) for _ in range(nranks)]
    try:
        if nranks > 1:
            for r, m in enumerate(ms):
                m.handle.set_bucket_shard(r, nranks)
                m.handle.fold_init(r, nranks, reads.nreads)
            for m in ms:
                m.handle.fold_connect_local([x.handle for x in ms])
        if packed:
            # uniform 100-base reads, 2 bit/base: vectorised form of synth.pack_reads_2bit
            m2 = reads.mapped.reshape(reads.nreads, c["L"])
            flags = (m2 > 3).any(axis=1).astype(np.uint8)
            q = np.where(m2 > 3, 0, m2).reshape(reads.nreads, c["L"] // 4, 4)
            pk = ((q[:, :, 0] << 6) | (q[:, :, 1] << 4) | (q[:, :, 2] << 2) | q[:, :, 3]).astype(np.uint8).reshape(-1)
        for m in ms:
            if packed:
                m.handle.set_reads_packed(pk, reads.nreads, uniform_length=c["L"], wildcard_flags=flags)
            else:
                m.set_reads(reads.mapped, reads.offsets, None)
            m.set_text(words, nmask, text.n, text.record_starts)
            m.match()
        if nranks > 1:
            rlib.Handle.fold_unique_group([m.handle for m in ms])
            info = np.concatenate([m.handle.get_unique(first=rdist.own_read_range(reads.nreads, r, nranks)[0],
                                                       count=rdist.own_read_range(reads.nreads, r, nranks)[1] - rdist.own_read_range(reads.nreads, r, nranks)[0])[0]
                                   for r, m in enumerate(ms)])
            digest = sum(m.handle.unique_checksum(*(lambda lo, hi: (lo, hi - lo))(*rdist.own_read_range(reads.nreads, r, nranks)))
                         for r, m in enumerate(ms)) & 0xFFFFFFFFFFFFFFFF
        else:
            info = ms[0].info()[0]
            digest = ms[0].handle.unique_checksum()
    finally:
        for m in ms:
            m.close()
    st = matcher.umi_state(info)
    assert int((st == 4).sum()) == doc["counts"]["nonunique"] and int((st == 0).sum()) == doc["counts"]["nomatch"]
    assert mk.canonical_digest(c, info) == doc["sha256"]
    assert digest == matcher.unique_checksum(info)          # the device-side digest is the numpy one
