#!/usr/bin/env python
"""K0 (FASTA text ingest on the device, csrc/ingest.cuh) on a synthetic genome file: kernel-side throughput against the HBM
roofline, end to end from pinned host bytes, and the reference's own loader (countLength + readFile, timed by the stock
binary itself) on a bounded sample of the same file on the host.

usage: ingest_bench.py [--bases N] [--records R] [--reps K] [--cpu-bases M] [--no-cpu]
The file is built on the device (24 records, 60-column lines, 0.1 % N in runs of 64 -- the layout randstr.cpp writes), then
 (1) real_gpu_set_text_fasta_device: summary + scan + clear + pack, CUDA events on the library's stream (stats.h2d_text_ms),
 (2) real_gpu_set_text_fasta from pinned host memory (adds the H2D copy of the file bytes),
 (3) result check: the packed words equal the symbols the file was made from (a checksum over all words), record table exact.
Prints one JSON line."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_fasta_on_device(torch, n, nrecords, seed):
    """(uint8 device tensor with the file bytes, packed words (uint64 device tensor), N-mask words, record starts)"""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84, 78], dtype=torch.uint8, device="cuda")
    cuts = sorted(set(int(x) for x in np.random.RandomState(seed).randint(1, n, nrecords - 1))) if nrecords > 1 else []
    starts = [0] + cuts + [n]
    parts, syms = [], []
    for r in range(len(starts) - 1):
        m = starts[r + 1] - starts[r]
        parts.append(torch.frombuffer(bytearray(b"> random_%d_part%d\n" % (n, r)), dtype=torch.uint8).cuda())
        done = 0
        while done < m:
            c = min(m - done, 60 * (1 << 22))
            s = torch.randint(0, 4, (c,), device="cuda", dtype=torch.uint8, generator=g)
            # N runs: 64-base stretches, about 0.1 % of the bases
            nruns = max(1, c // 64000)
            at = torch.randint(0, max(1, c - 64), (nruns,), device="cuda", generator=g)
            idx = (at[:, None] + torch.arange(64, device="cuda")[None, :]).reshape(-1)
            s[idx[idx < c]] = 4
            syms.append(s)
            ch = lut[s.long()]
            full = (c // 60) * 60
            rows = torch.empty((full // 60, 61), dtype=torch.uint8, device="cuda")
            rows[:, :60] = ch[:full].reshape(-1, 60)
            rows[:, 60] = 10
            parts.append(rows.reshape(-1))
            if c > full:
                parts.append(torch.cat([ch[full:], torch.tensor([10], dtype=torch.uint8, device="cuda")]))
            done += c
            del ch, rows
    data = torch.cat(parts)
    del parts
    sym = torch.cat(syms)
    del syms
    return data, sym, np.asarray(starts, dtype=np.uint64)



def expected_words(torch, sym):
    """packed 2-bit words and N-mask words of the symbols, on the device, in pieces"""
    n = sym.numel()
    nw, nm = (n + 31) // 32, (n + 63) // 64
    words = torch.zeros(nw, dtype=torch.int64, device="cuda")
    mask = torch.zeros(nm, dtype=torch.int64, device="cuda")
    step = 1 << 26
    sh = (62 - 2 * torch.arange(32, device="cuda", dtype=torch.int64))
    msh = (63 - torch.arange(64, device="cuda", dtype=torch.int64))
    for o in range(0, n, step):
        e = min(n, o + step)
        s = sym[o:e].long()
        pad = (-(e - o)) % 64
        if pad:
            s = torch.cat([s, torch.zeros(pad, dtype=torch.int64, device="cuda")])
        code = torch.where(s > 3, torch.zeros_like(s), s)
        words[o // 32:o // 32 + code.numel() // 32] = (code.reshape(-1, 32) << sh).sum(1)[: (nw - o // 32)]
        isn = (s > 3).long()
        mask[o // 64:o // 64 + isn.numel() // 64] = (isn.reshape(-1, 64) << msh).sum(1)
    return words, mask


def reference_loader_seconds(path):
    """countLength and readFile of the stock binary, by its own clocks (getText.hpp:34-44); the process is stopped there."""
    exe = os.path.join(ROOT, "oracle", "_ref", "real")
    if not os.path.exists(exe):
        return None
    work = tempfile.mkdtemp(prefix="ingest_ref_")
    rp = os.path.join(work, "r.fa")
    with open(rp, "w") as f:
        f.write(">r0\n" + "ACGT" * 10 + "\n")
    p = subprocess.Popen([exe, "-t", path, "-p", rp, "-o", os.path.join(work, "o.txt"), "-u", "0", "-q", "0", "-T", "1"],
                         stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    clocks, t0, buf = [], time.time(), ""
    try:
        while len(clocks) < 2 and time.time() - t0 < 600:
            ch = p.stderr.read(1)
            if not ch:
                break
            buf += ch
            if ch == "\n":
                clocks += [float(x) for x in re.findall(r"done\. clocks ([0-9.eE+-]+)", buf)]
                buf = ""
    finally:
        p.kill()
        p.wait()
    return {"count_length_s": clocks[0], "read_file_s": clocks[1]} if len(clocks) == 2 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=int, default=3_100_000_000)
    ap.add_argument("--records", type=int, default=24)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-bases", type=int, default=256_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    import torch
    from real_b200 import lib as rlib
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this benchmark has no CPU path")
    data, sym, starts = make_fasta_on_device(torch, args.bases, args.records, 7)
    nbytes = data.numel()
    h = rlib.Handle(seedl=32, seedkmax=2, totalkmax=4)
    ms = []
    for _ in range(args.reps + 2):
        n, nrec = h.set_text_fasta(None, device_ptr=data.data_ptr(), nbytes=nbytes)
        ms.append(h.stats()["h2d_text_ms"])
    assert n == args.bases and nrec == args.records, (n, nrec)
    dev_ms = float(np.median(ms[2:]))
    # result check on the device: the library's words against the words of the symbols the file was made from
    got_starts, _ = h.get_text_records()
    ok_records = bool(np.array_equal(got_starts, starts))
    ew, em = expected_words(torch, sym)
    hw = np.zeros((n + 31) // 32, dtype=np.uint64)
    hm = np.zeros((n + 63) // 64, dtype=np.uint64)
    assert h.L.real_gpu_get_text_packed(h.h, n, hw.ctypes.data, hm.ctypes.data) == 0
    ok_words = bool(np.array_equal(hw.view(np.int64), ew.cpu().numpy())) and bool(np.array_equal(hm.view(np.int64), em.cpu().numpy()))
    del ew, em, hw, hm
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6538.0)) if isinstance(peaks, dict) else 6538.0
    alg = nbytes + n * 3 / 8.0                          # file bytes read once + 2-bit text and 1-bit mask written
    design = 2 * nbytes + 2 * n * 3 / 8.0                # two passes over the file, arrays cleared then written (tile summaries: 36 B per 16 KB)
    out = {"kernel": "K0 k_fa_summary + k_fa_scan + k_fa_pack (+ clearing the arrays)", "file_bytes": nbytes, "bases": n, "records": nrec,
           "device_ms": dev_ms, "device_ms_all": [round(x, 3) for x in ms], "file_GBps": nbytes / dev_ms / 1e6, "gbp_per_s": n / dev_ms / 1e6,
           "roofline": {"bound": "hbm", "achieved": alg / dev_ms / 1e6, "peak": peak, "unit": "GB/s", "frac": alg / dev_ms / 1e6 / peak,
                        "design_bytes": design, "design_frac": design / dev_ms / 1e6 / peak},
           "parity": {"words_and_mask_equal": ok_words, "record_table_equal": ok_records}}
    if not args.no_e2e:
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.copy_(data)
        arr = host.numpy()
        e = []
        for _ in range(3):
            t0 = time.perf_counter()
            h.set_text_fasta(arr)
            e.append((time.perf_counter() - t0) * 1e3)
        out["e2e_pinned_host"] = {"ms": float(np.median(e)), "file_GBps": nbytes / float(np.median(e)) / 1e6, "h2d_bytes": nbytes}
    if not args.no_cpu:
        m = min(args.cpu_bases, args.bases)
        # the first m bases of the same file (header lines included)
        cut = int(m + m // 60 + 64 * args.records)
        sample = data[:min(cut, nbytes)].cpu().numpy().tobytes()
        sample = sample[:sample.rfind(b"\n") + 1]
        path = os.path.join(tempfile.mkdtemp(prefix="ingest_"), "t.fa")
        with open(path, "wb") as f:
            f.write(sample)
        ref = reference_loader_seconds(path)
        if ref:
            s = ref["count_length_s"] + ref["read_file_s"]
            out["cpu_baseline"] = {"kind": "reference", "cores": 1, "sample": "first %d bytes of the file; countLength %.2f s + readFile %.2f s by the stock binary's own clocks" % (len(sample), ref["count_length_s"], ref["read_file_s"]),
                                   "file_GBps": len(sample) / s / 1e9, "extrapolated_full_s": s * nbytes / len(sample)}
        os.remove(path)
    h.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
