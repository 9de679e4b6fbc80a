#!/usr/bin/env python3
"""Full-size golden for BASELINE.json configs[1] from the CPU ORACLE alone (no GPU involved):

    python tests/golden/make_fullsize_golden.py        # ~2 min on 8 cores, writes tests/golden/fullsize_k6_n100.npz

The 10,500 synthetic genomes (~4 Mbp each, 42 Gbp; the host generator is the twin of the device one) are
generated chunk by chunk, counted at k=6 by the oracle (counts -> frequencies -> reference-order entropy),
and the oracle's nmost n=100 is run over numpy.random.default_rng(20261017).permutation(10500) — the
order bench.py uses.  Stored: per-record k-mer totals and entropies (bit patterns), a 64-bit checksum of
every record's count row, the selected ids in Vec order, their delta_jsd, total_jsd and summed entropies."""
import pathlib
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from diverseseq_b200 import _lib  # noqa: E402  (host generator only; no CUDA call)
from oracle import oracle as orc  # noqa: E402

SEED, NREC, NFAM, MEAN_LEN, K, N = 20261017, 10500, 64, 4_000_000, 6, 100
CHUNK = 25
MIX = np.uint64(0x9E3779B97F4A7C15)


def row_checksum(counts: np.ndarray) -> np.ndarray:
    """order-sensitive 64-bit checksum per row: sum_i count_i * (i * MIX + 1)  (mod 2^64)"""
    w = (np.arange(counts.shape[1], dtype=np.uint64) * MIX + np.uint64(1))
    with np.errstate(over="ignore"):
        return (counts.astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)


def work(first):
    cnt = min(CHUNK, NREC - first)
    flat, off = _lib.synth_host(SEED, NREC, NFAM, MEAN_LEN, first, cnt)
    counts, freqs, ent, valid = orc.count_batch(flat, off, K, threads=1)
    return first, counts.sum(axis=1).astype(np.uint64), row_checksum(counts), freqs, ent, valid


def main():
    dim = 4 ** K
    freqs = np.zeros((NREC, dim))
    ent = np.zeros(NREC)
    totals = np.zeros(NREC, dtype=np.uint64)
    sums = np.zeros(NREC, dtype=np.uint64)
    valid = np.zeros(NREC, dtype=np.uint8)
    with ThreadPoolExecutor(8) as pool:
        for first, t, s, f, e, v in pool.map(work, range(0, NREC, CHUNK)):
            n = len(t)
            totals[first:first + n], sums[first:first + n] = t, s
            freqs[first:first + n], ent[first:first + n], valid[first:first + n] = f, e, v
            if first % 1000 == 0:
                print("counted", first, flush=True)
    order = np.random.default_rng(SEED).permutation(NREC).astype(np.uint32)
    res = orc.select_rows(freqs, ent, order, "nmost", N, valid=valid)
    out = ROOT / "tests" / "golden" / "fullsize_k6_n100.npz"
    np.savez_compressed(out, totals=totals, count_checksums=sums, entropy_bits=ent.view(np.uint64), valid=valid,
                        ids=np.asarray(res.ids, dtype=np.uint32), delta_jsd_bits=np.asarray(res.delta_jsd).view(np.uint64),
                        total_jsd=np.float64(res.total_jsd), summed_entropies=np.float64(res.summed_entropies),
                        n_accepts=np.int64(len(res.trace)) if hasattr(res, "trace") else np.int64(-1))
    print("wrote", out, "ids head", res.ids[:8], "total_jsd", res.total_jsd)


if __name__ == "__main__":
    main()
