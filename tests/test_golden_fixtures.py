"""Committed golden vectors (tests/golden/brca1_expected.json, made by make_expected_outputs.py from
the pinned oracle on the reference's brca1 fixture = BASELINE.json configs[0]).
CPU: the oracle still reproduces them.  GPU (-m gpu): the CUDA path hits them bit-for-bit."""
import json
import pathlib

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
EXP = json.loads((ROOT / "tests" / "golden" / "brca1_expected.json").read_text())
FIVE = ["Human", "Chimpanzee", "Manatee", "Dugong", "Rhesus"]


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def _flat(brca1):
    names = EXP["names"]
    assert names == list(brca1)
    arrs = [brca1[n] for n in names]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([a.size for a in arrs])
    return names, np.concatenate(arrs), off


def _checksum(c):
    return [int(x) for x in (c.astype(np.uint64) * (np.arange(c.shape[1], dtype=np.uint64) + 1)).sum(axis=1)]


def test_oracle_reproduces_golden(brca1):
    from oracle import oracle as orc
    names, flat, off = _flat(brca1)
    for k, e in EXP["count"].items():
        c, f, ent, v = orc.count_batch(flat, off, int(k))
        assert np.array_equal(ent, unhex(e["entropy"])) and _checksum(c) == e["counts_checksum"]
    for g in EXP["nmost"]:
        order = np.random.default_rng(g["seed"]).permutation(len(names))
        s = orc.select_seqs(flat, off, order, g["k"], "nmost", g["n"])
        assert [names[i] for i in s.ids] == g["names"] and np.array_equal(s.delta_jsd, unhex(g["delta_jsd"]))
        assert s.total_jsd == float.fromhex(g["total_jsd"])
    for g in EXP["max"]:
        order = np.random.default_rng(g["seed"]).permutation(len(names))
        s = orc.select_seqs(flat, off, order, g["k"], g["stat"], g["min"], g["max"])
        assert [names[i] for i in s.ids] == g["names"] and np.array_equal(s.delta_jsd, unhex(g["delta_jsd"]))


@pytest.mark.gpu
def test_cuda_hits_golden(brca1):
    from diverseseq_b200 import _lib
    ctx = _lib.Context(0)
    names, flat, off = _flat(brca1)
    ss = _lib.SeqSet.upload(ctx, flat, off)
    for k, e in EXP["count"].items():
        kf = _lib.KFreqs.count(ctx, ss, int(k))
        c, f, ent, v = kf.download()
        assert np.array_equal(ent, unhex(e["entropy"])) and _checksum(c) == e["counts_checksum"]
        assert v.tolist() == e["valid"]
        if e["human_counts"] is not None:
            assert c[names.index("Human")].tolist() == e["human_counts"]
    for g in EXP["nmost"]:
        kf = _lib.KFreqs.count(ctx, ss, g["k"])
        order = np.random.default_rng(g["seed"]).permutation(len(names)).astype(np.uint32)
        idx, delta, stats = kf.select(order, _lib.MODE_NMOST, g["n"])
        assert [names[i] for i in idx] == g["names"] and np.array_equal(delta, unhex(g["delta_jsd"]))
        assert [stats[0], stats[1], stats[2], stats[3]] == [float.fromhex(g[x]) for x in ("total_jsd", "mean", "std", "cov")]
    for g in EXP["max"]:
        kf = _lib.KFreqs.count(ctx, ss, g["k"])
        order = np.random.default_rng(g["seed"]).permutation(len(names)).astype(np.uint32)
        mode = _lib.MODE_MAX_STDEV if g["stat"] == "stdev" else _lib.MODE_MAX_COV
        idx, delta, stats = kf.select(order, mode, g["min"], g["max"])
        assert [names[i] for i in idx] == g["names"] and np.array_equal(delta, unhex(g["delta_jsd"]))
    five = _lib.SeqSet.from_seqs(ctx, [brca1[n] for n in FIVE])
    for key, e in EXP["sketch"].items():
        k, s, canon = int(key.split("_")[0][1:]), int(key.split("_")[1][1:]), key.endswith("c1")
        sk = _lib.Sketches.sketch(ctx, five, k, s, 4, canon)
        data, lens = sk.download()
        assert lens.tolist() == e["lens"]
        assert [data[i, :8].tolist() for i in range(5)] == e["head"]
        assert [int(np.bitwise_xor.reduce(data[i, :lens[i]])) for i in range(5)] == e["xor"]
        dist, inter, uni = sk.distances(k, s, want_counts=True)
        m = EXP["mash"][key]
        assert inter.ravel().tolist() == m["inter"] and uni.ravel().tolist() == m["union"]
        np.testing.assert_allclose(dist.ravel(), unhex(m["dist"]), rtol=1e-9, atol=0)
    for k, e in EXP["euclid"].items():
        kf = _lib.KFreqs.count(ctx, five, int(k))
        np.testing.assert_allclose(kf.euclidean().ravel(), unhex(e["dist"]), rtol=1e-9, atol=0)
