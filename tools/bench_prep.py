#!/usr/bin/env python3
"""Throughput of the `dvs prep` encode kernels (csrc/prep.cu) on device-resident FASTA text.

    python tools/bench_prep.py [--files 1024] [--mean-len 4000000] [--width 80]

Synthetic genomes (the bench generator) are rendered as wrapped FASTA on the host, copied to the
device once, and dvs_prep_fasta(text_on_device=1) is timed with the ctx's CUDA events.  Prints one
JSON line: Gbp/s, GB/s against the algorithmic traffic (text read once + record bytes written once).
"""
from __future__ import annotations

import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def render(seq: np.ndarray, width: int, label: bytes) -> np.ndarray:
    letters = np.frombuffer(b"TCAGN", dtype=np.uint8)[np.minimum(seq, 4)]
    n = letters.size
    rows = (n + width - 1) // width
    pad = np.full(rows * width, ord("\n"), dtype=np.uint8)
    pad[:n] = letters
    body = np.concatenate([pad.reshape(rows, width), np.full((rows, 1), ord("\n"), np.uint8)], axis=1).reshape(-1)
    body = body[: n + (n + width - 1) // width]  # drop the padding of the last row, keep its newline
    if body.size:
        body[-1] = ord("\n")
    return np.concatenate([np.frombuffer(b">" + label + b"\n", dtype=np.uint8), body])


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=1024)
    ap.add_argument("--mean-len", type=int, default=4_000_000)
    ap.add_argument("--width", type=int, default=80)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch

    from diverseseq_b200 import _lib

    ctx = _lib.Context(0)
    ctx.enable_timing(True)
    ss = _lib.SeqSet.synth(ctx, 20261017, a.files, 64, a.mean_len)
    off = ss.offsets()
    flat = ss.download()
    files = [render(flat[int(off[i]):int(off[i + 1])], a.width, b"genome%d some description" % i) for i in range(a.files)]
    text, toff = _lib.concat(files)
    d = torch.empty(text.size + 64, dtype=torch.uint8, device="cuda:0")
    d[:text.size] = torch.from_numpy(text).cuda()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        out = _lib.SeqSet.prep_fasta(ctx, None, toff, device_ptr=d.data_ptr())
        wall = time.perf_counter() - t0
        ms.append((ctx.phase_ms(_lib.PHASE_PREP), wall * 1e3))
        same = out.total_bases == ss.total_bases
        del out
    out = _lib.SeqSet.prep_fasta(ctx, None, toff, device_ptr=d.data_ptr())
    ok = bool(same and np.array_equal(out.download(), np.where(flat > 3, 5, flat)))
    dev = float(np.median([m[0] for m in ms]))
    print(json.dumps({"files": a.files, "text_bytes": int(text.size), "bases": int(ss.total_bases), "device_ms": dev,
                      "wall_ms": float(np.median([m[1] for m in ms])), "gbp_per_s": ss.total_bases / dev / 1e6,
                      "algorithmic_gb_per_s": (text.size + ss.total_bases) / dev / 1e6, "matches_generator": ok}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
