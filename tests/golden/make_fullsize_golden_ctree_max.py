#!/usr/bin/env python3
"""Full-size goldens for BASELINE.json configs[2..4] from the CPU ORACLE alone (no GPU involved):

    python tests/golden/make_fullsize_golden_ctree_max.py [mash] [k8]

  mash  configs[3]: 1,000 synthetic genomes (~4 Mbp), canonical mash sketches k=16 s=3000 and the whole
        distance matrix -> fullsize_mash_k16_s3000.npz: sketch lengths, an order-sensitive 64-bit checksum of
        every sketch, row checksums of the intersection / union matrices, 4,000 sampled pair distances.
  k8    configs[2] and [4]: the 10,500-genome set counted at k=8 -> fullsize_k8.npz: per-record totals and
        entropy bit patterns, `max` selections (stdev 5..10, stdev 10..100, cov 10..100) and nmost n=100 (ids,
        delta_jsd bit patterns, total_jsd), and 4,000 sampled Euclidean pair distances.
Takes ~10 min (mash) and ~15 min (k8) on 8 cores; the k8 part needs ~7 GB of RAM."""
import pathlib
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from diverseseq_b200 import _lib  # noqa: E402  (host generator only; no CUDA call)
from oracle import oracle as orc  # noqa: E402

SEED, NFAM, MEAN_LEN = 20261017, 64, 4_000_000
MIX = np.uint64(0x9E3779B97F4A7C15)
OUT = ROOT / "tests" / "golden"


def checksum_rows(a: np.ndarray) -> np.ndarray:
    w = np.arange(a.shape[1], dtype=np.uint64) * MIX + np.uint64(1)
    with np.errstate(over="ignore"):
        return (a.astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)


def mash():
    nrec, k, s, chunk = 1000, 16, 3000, 8
    sk = np.zeros((nrec, s), dtype=np.uint32)
    lens = np.zeros(nrec, dtype=np.uint32)

    def work(first):
        flat, off = _lib.synth_host(SEED, nrec, NFAM, MEAN_LEN, first, min(chunk, nrec - first))
        a, ln = orc.mash_sketch_batch(flat, off, k, s, canonical=True, threads=1)
        return first, a, ln

    with ThreadPoolExecutor(8) as pool:
        for first, a, ln in pool.map(work, range(0, nrec, chunk)):
            sk[first:first + len(ln), : a.shape[1]] = a
            lens[first:first + len(ln)] = ln
            if first % 200 == 0:
                print("sketched", first, flush=True)
    dist, inter, uni = orc.mash_matrix(sk, lens, k, s, threads=8)
    rng = np.random.default_rng(SEED)
    pi, pj = rng.integers(0, nrec, 4000), rng.integers(0, nrec, 4000)
    np.savez_compressed(OUT / "fullsize_mash_k16_s3000.npz", lens=lens, sketch_checksums=checksum_rows(sk),
                        inter_row_checksums=checksum_rows(inter), union_row_checksums=checksum_rows(uni),
                        pair_i=pi.astype(np.uint32), pair_j=pj.astype(np.uint32), pair_dist=dist[pi, pj],
                        mean_dist=np.float64(dist.mean()))
    print("mash: mean distance", dist.mean(), "sketch lens", lens.min(), lens.max())


def k8():
    nrec, k, chunk = 10500, 8, 10
    dim = 4 ** k
    freqs = np.zeros((nrec, dim))
    ent = np.zeros(nrec)
    totals = np.zeros(nrec, dtype=np.uint64)
    valid = np.zeros(nrec, dtype=np.uint8)

    def work(first):
        flat, off = _lib.synth_host(SEED, nrec, NFAM, MEAN_LEN, first, min(chunk, nrec - first))
        counts, f, e, v = orc.count_batch(flat, off, k, threads=1)
        return first, counts.sum(axis=1).astype(np.uint64), f, e, v

    with ThreadPoolExecutor(8) as pool:
        for first, t, f, e, v in pool.map(work, range(0, nrec, chunk)):
            n = len(t)
            totals[first:first + n], freqs[first:first + n], ent[first:first + n], valid[first:first + n] = t, f, e, v
            if first % 1000 == 0:
                print("counted", first, flush=True)
    order = np.random.default_rng(SEED).permutation(nrec).astype(np.uint32)
    out = {"totals": totals, "entropy_bits": ent.view(np.uint64), "valid": valid}
    for name, mode, lo, hi in (("stdev_5_10", "stdev", 5, 10), ("stdev_10_100", "stdev", 10, 100),
                               ("cov_10_100", "cov", 10, 100), ("nmost_100", "nmost", 100, 100)):
        r = orc.select_rows(freqs, ent, order, mode, lo, hi, valid=valid)
        out[f"{name}_ids"] = np.asarray(r.ids, dtype=np.uint32)
        out[f"{name}_delta_bits"] = np.asarray(r.delta_jsd).view(np.uint64)
        out[f"{name}_total_jsd"] = np.float64(r.total_jsd)
        print(name, "size", len(r.ids), "total_jsd", r.total_jsd, flush=True)
    rng = np.random.default_rng(SEED + 1)
    pi, pj = rng.integers(0, nrec, 4000), rng.integers(0, nrec, 4000)
    out["pair_i"], out["pair_j"] = pi.astype(np.uint32), pj.astype(np.uint32)
    out["pair_euclid"] = np.sqrt(((freqs[pi] - freqs[pj]) ** 2).sum(axis=1))  # np.linalg.norm's definition
    np.savez_compressed(OUT / "fullsize_k8.npz", **out)
    print("k8 done")


if __name__ == "__main__":
    which = set(sys.argv[1:]) or {"mash", "k8"}
    if "mash" in which:
        mash()
    if "k8" in which:
        k8()
