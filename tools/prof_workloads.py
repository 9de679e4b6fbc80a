#!/usr/bin/env python3
"""Small invocations of the side-path kernels for ncu captures (python tools/prof_workloads.py euclid|mash|sparse|k12|select8)"""
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from diverseseq_b200 import _lib  # noqa: E402

SEED = 20261017


def main():
    what = sys.argv[1]
    ctx = _lib.Context(0)
    ctx.enable_timing(True)
    if what == "euclid":  # 4,096 genomes x 4^8: 8.4 M pairs, the shape of configs[4] at 0.15x the pairs
        ss = _lib.SeqSet.synth(ctx, SEED, 4096, 64, 400_000)
        kf = _lib.KFreqs.count(ctx, ss, 8)
        d = _lib.DeviceBuffer(ctx, 4096 * 4096 * 8)
        for _ in range(2):
            kf.euclidean_into(d.ptr)
        print("euclid ms", ctx.phase_ms(_lib.PHASE_EUCLID), "fallback pairs", ctx._lib.dvs_euclid_last_fallback_pairs(ctx.handle))
    elif what == "mash":  # 200 genomes, k=16, s=3000
        ss = _lib.SeqSet.synth(ctx, SEED, 200, 64, 4_000_000)
        for _ in range(2):
            sk = _lib.Sketches.sketch(ctx, ss, 16, 3000, 4, True)
        print("sketch ms", ctx.phase_ms(_lib.PHASE_SKETCH))
        sk1000 = _lib.Sketches.from_host(ctx, np.sort(np.random.default_rng(1).integers(0, 2**32, (1000, 3000), dtype=np.uint64).astype(np.uint32), axis=1),
                                         np.full(1000, 3000, dtype=np.uint32))
        d = _lib.DeviceBuffer(ctx, 1000 * 1000 * 8)
        for _ in range(2):
            sk1000.distances_into(d.ptr, 16, 3000)
        print("pairs ms", ctx.phase_ms(_lib.PHASE_MASH_PAIRS))
    elif what == "sparse":
        ss = _lib.SeqSet.synth(ctx, SEED, 128, 64, 4_000_000)
        for _ in range(2):
            sp = _lib.KSparse.count(ctx, ss, 12)
        print("sparse ms", ctx.phase_ms(_lib.PHASE_SPARSE), "Gbp/s", ss.total_bases / ctx.phase_ms(_lib.PHASE_SPARSE) / 1e6)
    elif what == "k12":
        ss = _lib.SeqSet.synth(ctx, SEED, 16, 64, 4_000_000)
        for _ in range(2):
            kf = _lib.KFreqs.count(ctx, ss, 12)
        print("k12 dense ms", ctx.phase_ms(_lib.PHASE_COUNT_KERNEL))
    elif what == "select8":
        ss = _lib.SeqSet.synth(ctx, SEED, 10500, 64, 400_000)
        kf = _lib.KFreqs.count(ctx, ss, 8)
        order = np.random.default_rng(SEED).permutation(10500).astype(np.uint32)
        for _ in range(2):  # nmost: the cooperative rounds (k_sel_persist); `max` below its max_size runs the grow kernels
            kf.select(order, _lib.MODE_NMOST, 100, 100)
        print("nmost ms", ctx.phase_ms(_lib.PHASE_SELECT))
        for _ in range(2):
            kf.select(order, _lib.MODE_MAX_STDEV, 10, 100)
        print("max ms", ctx.phase_ms(_lib.PHASE_SELECT))


if __name__ == "__main__":
    main()
