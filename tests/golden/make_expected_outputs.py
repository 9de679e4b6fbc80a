#!/usr/bin/env python3
"""Freezes oracle outputs on the reference's brca1 fixture into tests/golden/brca1_expected.json.

The oracle (oracle/dvs_oracle.cpp) is pinned to the reference's own known-answer values by
tests/test_oracle_golden.py; this script records what that pinned oracle produces on the
55-sequence brca1 set (BASELINE.json configs[0]) so that (i) a later change to the oracle cannot
silently move the target and (ii) the GPU tests have committed vectors to hit.  f64 values are
stored as hex strings (bit-exact).  Run from the repo root: python tests/golden/make_expected_outputs.py
"""
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    z = np.load(ROOT / "tests" / "golden" / "brca1.npz")
    names = [str(n) for n in z["names"]]
    off = z["offsets"].astype(np.uint64)
    flat = np.ascontiguousarray(z["data"])
    out = {"names": names, "count": {}, "nmost": [], "max": [], "sketch": {}, "mash": {}, "euclid": {}}
    for k in (1, 3, 6):
        c, f, e, v = orc.count_batch(flat, off, k)
        out["count"][str(k)] = {"entropy": hx(e), "valid": v.tolist(),
                                "counts_checksum": [int(x) for x in (c * (np.arange(c.shape[1], dtype=np.uint64) + 1)).sum(axis=1)],
                                "human_counts": c[names.index("Human")].tolist() if k <= 3 else None}
    for k, n, seed in ((6, 10, 1), (6, 10, 2), (3, 5, 3), (1, 3, 4)):
        order = np.random.default_rng(seed).permutation(len(names))
        s = orc.select_seqs(flat, off, order, k, "nmost", n)
        out["nmost"].append({"k": k, "n": n, "seed": seed, "names": [names[i] for i in s.ids], "delta_jsd": hx(s.delta_jsd),
                             "total_jsd": float(s.total_jsd).hex(), "mean": float(s.mean_delta_jsd).hex(),
                             "std": float(s.std_delta_jsd).hex(), "cov": float(s.cov_delta_jsd).hex()})
    for k, lo, hi, stat, seed in ((2, 3, 12, "stdev", 1), (5, 5, 20, "cov", 2), (4, 2, 55, "stdev", 3)):
        order = np.random.default_rng(seed).permutation(len(names))
        s = orc.select_seqs(flat, off, order, k, stat, lo, hi)
        out["max"].append({"k": k, "min": lo, "max": hi, "stat": stat, "seed": seed, "names": [names[i] for i in s.ids],
                           "delta_jsd": hx(s.delta_jsd), "total_jsd": float(s.total_jsd).hex()})
    five = ["Human", "Chimpanzee", "Manatee", "Dugong", "Rhesus"]
    rows = [names.index(n) for n in five]
    seqs = [flat[int(off[r]):int(off[r + 1])] for r in rows]
    for k, s_, canon in ((16, 400, True), (12, 3000, False)):
        key = f"k{k}_s{s_}_c{int(canon)}"
        sk = [orc.mash_sketch(q, k, s_, 4, canon) for q in seqs]
        out["sketch"][key] = {"names": five, "lens": [len(x) for x in sk], "head": [x[:8].tolist() for x in sk],
                              "xor": [int(np.bitwise_xor.reduce(x)) for x in sk]}
        stride = max(len(x) for x in sk)
        mat = np.zeros((5, stride), dtype=np.uint32)
        for i, x in enumerate(sk):
            mat[i, :len(x)] = x
        d, inter, uni = orc.mash_matrix(mat, np.array([len(x) for x in sk], dtype=np.uint32), k, s_)
        out["mash"][key] = {"dist": hx(d), "inter": inter.ravel().tolist(), "union": uni.ravel().tolist()}
    for k in (5,):
        f = np.stack([orc.kfreqs_unchecked(q, k) for q in seqs])
        out["euclid"][str(k)] = {"names": five, "dist": hx(orc.euclid_matrix(f))}
    (ROOT / "tests" / "golden" / "brca1_expected.json").write_text(json.dumps(out, indent=1))
    print("wrote", ROOT / "tests" / "golden" / "brca1_expected.json")


if __name__ == "__main__":
    main()
