"""Full-size parity for BASELINE.json configs[2..4] against oracle-only goldens
(tests/golden/make_fullsize_golden_ctree_max.py): configs[3] = mash k=16 s=3000 on 1,000 genomes,
configs[2] = `max` sweeps at k=8 over the 10,500-genome set, configs[4] = Euclidean k=8 on the same set."""
import pathlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED, NFAM, MEAN_LEN = 20261017, 64, 4_000_000
GOLD = pathlib.Path(__file__).resolve().parent / "golden"
MIX = np.uint64(0x9E3779B97F4A7C15)


def checksum_rows(a: np.ndarray) -> np.ndarray:
    w = np.arange(a.shape[1], dtype=np.uint64) * MIX + np.uint64(1)
    with np.errstate(over="ignore"):
        return (a.astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)


def test_fullsize_mash_equals_the_oracle():
    from diverseseq_b200 import _lib

    g = np.load(GOLD / "fullsize_mash_k16_s3000.npz")
    ctx = _lib.Context(0)
    ss = _lib.SeqSet.synth(ctx, SEED, 1000, NFAM, MEAN_LEN)
    sk = _lib.Sketches.sketch(ctx, ss, 16, 3000, 4, True)
    data, lens = sk.download()
    assert np.array_equal(lens, g["lens"])
    assert np.array_equal(checksum_rows(data[:, :3000]), g["sketch_checksums"])  # every sketch, order-sensitive
    dist = sk.distances(16, 3000)
    np.testing.assert_allclose(dist[g["pair_i"], g["pair_j"]], g["pair_dist"], rtol=1e-9, atol=0)
    np.testing.assert_allclose(dist.mean(), float(g["mean_dist"]), rtol=1e-12)


def test_fullsize_k8_selections_and_euclid_equal_the_oracle():
    from diverseseq_b200 import _lib

    g = np.load(GOLD / "fullsize_k8.npz")
    nrec = 10500
    ctx = _lib.Context(0)
    ss = _lib.SeqSet.synth(ctx, SEED, nrec, NFAM, MEAN_LEN)
    kf = _lib.KFreqs.count(ctx, ss, 8)
    ent = np.zeros(nrec)
    totals = np.zeros(nrec, dtype=np.uint64)
    for first in range(0, nrec, 250):  # 250 x 65536 x u64 = 131 MB at a time
        cnt = min(250, nrec - first)
        c, _, e, _v = kf.download(first, cnt, freqs=False)
        ent[first:first + cnt] = e
        totals[first:first + cnt] = c.sum(axis=1, dtype=np.uint64)
    assert np.array_equal(totals, g["totals"])
    assert np.array_equal(ent.view(np.uint64), g["entropy_bits"])
    order = np.random.default_rng(SEED).permutation(nrec).astype(np.uint32)
    for name, mode, lo, hi in (("stdev_5_10", _lib.MODE_MAX_STDEV, 5, 10), ("stdev_10_100", _lib.MODE_MAX_STDEV, 10, 100),
                               ("cov_10_100", _lib.MODE_MAX_COV, 10, 100), ("nmost_100", _lib.MODE_NMOST, 100, 100)):
        idx, delta, stats = kf.select(order, mode, lo, hi)
        assert idx.tolist() == g[f"{name}_ids"].tolist(), name
        assert np.array_equal(delta.view(np.uint64), g[f"{name}_delta_bits"]), name
        assert stats[0] == float(g[f"{name}_total_jsd"]), name
    pi, pj = g["pair_i"].astype(np.int64), g["pair_j"].astype(np.int64)
    rows = np.unique(pi)[:64]  # a band of the matrix: 64 rows x all columns
    for r in rows:
        d = kf.euclidean(int(r), int(r) + 1)[0]
        sel = pi == r
        np.testing.assert_allclose(d[pj[sel]], g["pair_euclid"][sel], rtol=1e-9, atol=1e-18)
