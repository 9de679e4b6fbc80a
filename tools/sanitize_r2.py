#!/usr/bin/env python3
"""Tiny runs of the round-2 kernels for compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize_r2.py

dvs_count_select (counting launches on a second stream, the selection kernel trailing on a high-priority stream),
the sliced rounds of the cooperative selection kernel at k=7, the sparse counting passes at k=12 (bitmap pass 2,
counter form, CTA form for a homopolymer's giant bucket), the Euclidean Gram kernels and the mash sketch."""
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from diverseseq_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
flat, off = _lib.synth_host(11, 400, 6, 20_000)
ss = _lib.SeqSet.upload(ctx, flat, off)
order = np.random.default_rng(2).permutation(400).astype(np.uint32)
kf, idx, delta, stats = _lib.KFreqs.count_select(ctx, ss, 6, order, _lib.MODE_NMOST, 20, 20)
print("count_select", idx[:6].tolist(), int(ctx._lib.dvs_select_last_trail_accepts(ctx.handle)), flush=True)
ref = _lib.KFreqs.count(ctx, ss, 6).select(order, _lib.MODE_NMOST, 20, 20)
assert idx.tolist() == ref[0].tolist()

kf7 = _lib.KFreqs.count(ctx, ss, 7)
idx7, _, _ = kf7.select(order, _lib.MODE_NMOST, 12, 12)
print("k=7 nmost (sliced rounds)", idx7[:6].tolist(), flush=True)
eu = kf7.euclidean()
print("euclid", float(eu[0, 1]), flush=True)

seqs = [np.random.default_rng(3).integers(0, 4, size=300_000, dtype=np.uint8), np.zeros(80_000, dtype=np.uint8),
        np.random.default_rng(4).integers(0, 5, size=5_000, dtype=np.uint8)]
f2, o2 = _lib.concat(seqs)
sp = _lib.KSparse.count(ctx, _lib.SeqSet.upload(ctx, f2, o2), 12, want_entropy=True)
nnz, tot, ent, valid = sp.stats()
print("sparse k=12", nnz.tolist(), tot.tolist(), flush=True)
i0, c0 = sp.record(1)
assert i0.tolist() == [0] and int(c0[0]) == 80_000 - 11

sk = _lib.Sketches.sketch(ctx, ss, 16, 200, 4, True)
print("sketch", sk.download()[1][:4].tolist(), flush=True)
