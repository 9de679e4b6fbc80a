#!/usr/bin/env python3
"""Tiny selection runs for compute-sanitizer (memcheck / racecheck) over the persistent kernels:

    compute-sanitizer --tool racecheck python tools/sanitize_select.py
"""
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from diverseseq_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
flat, off = _lib.synth_host(7, 120, 5, 3000)
kf = _lib.KFreqs.count(ctx, _lib.SeqSet.upload(ctx, flat, off), 6)
order = np.random.default_rng(1).permutation(120).astype(np.uint32)
for n in (4, 30):
    idx, delta, stats = kf.select(order, _lib.MODE_NMOST, n)
    print("nmost", n, idx[:6].tolist(), int(ctx._lib.dvs_select_last_accepts(ctx.handle)), flush=True)
idx, delta, stats = kf.select(order, _lib.MODE_MAX_STDEV, 4, 12)
print("max", idx.tolist(), flush=True)
