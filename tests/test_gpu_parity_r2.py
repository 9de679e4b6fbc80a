"""Round-2 GPU parity tests: the gaps VERDICT r1 listed.

* `max` (stdev / cov) and nmost at k=8 against the oracle on a 1,600-record synthetic set - k=8 rows do not
  fit shared memory, so these go through the global-state persistent kernel and the batched grow attempts,
  i.e. the path BASELINE.json configs[2] takes (/root/reference/src/records.rs:390-454);
* Euclidean distances on near-duplicate rows (0.1 % substitution family members and exact copies), the case
  that separates the Gram form from the difference form (/root/reference/diverse_seq/distance.py:335-336);
* `max` with max_size < min_size / max_size == 0 (only the reference's CLI rejects it, cli.py:311).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def lib():
    from diverseseq_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def ctx(lib):
    return lib.Context(0)


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def k8_set(lib, ctx, orc):
    """1,600 synthetic genomes of ~60 kbp in 16 families, counted at k=8 on the device and by the oracle"""
    nrec = 1600
    flat, off = lib.synth_host(8088, nrec, 16, 60_000)
    kf = lib.KFreqs.count(ctx, lib.SeqSet.upload(ctx, flat, off), 8)
    _, of, oe, ov = orc.count_batch(flat, off, 8, threads=max(1, orc.hardware_threads()), want_counts=False)
    _, f, e, v = kf.download(counts=False)
    assert np.array_equal(v, ov) and np.array_equal(e, oe) and np.array_equal(f, of)  # bit-identical rows
    order = np.random.default_rng(8).permutation(nrec).astype(np.uint32)
    return kf, of, oe, ov, order


@pytest.mark.parametrize("mode,lo,hi", [("stdev", 5, 10), ("stdev", 10, 100), ("cov", 10, 40), ("nmost", 30, 30)])
def test_select_k8_matches_oracle(lib, ctx, orc, k8_set, mode, lo, hi):
    kf, of, oe, ov, order = k8_set
    m = {"stdev": lib.MODE_MAX_STDEV, "cov": lib.MODE_MAX_COV, "nmost": lib.MODE_NMOST}[mode]
    exp = orc.select_rows(of, oe, order, mode, lo, hi, valid=ov)
    idx, delta, stats = kf.select(order, m, lo, hi)
    assert idx.tolist() == exp.ids.tolist()
    assert np.array_equal(delta, exp.delta_jsd)  # bitwise
    assert stats[0] == exp.total_jsd and stats[1] == exp.mean_delta_jsd and stats[2] == exp.std_delta_jsd
    assert lo <= idx.size <= hi
    assert len(exp.trace) >= 2  # the pass really changed the set


@pytest.mark.parametrize("k", [5, 8])
def test_euclidean_near_duplicate_rows(lib, ctx, orc, k):
    """rows that differ by ~1e-3 relative (and some that do not differ at all): d^2 is ~1e-6 of ||a||^2, where
    a Gram-form evaluation loses its digits to cancellation; 1e-9 relative must still hold and copies give 0"""
    rng = np.random.default_rng(2026 + k)
    base = rng.integers(0, 4, size=300_000, dtype=np.uint8)
    seqs = []
    for i in range(24):
        s = base.copy()
        hit = rng.random(s.size) < 0.001
        s[hit] = rng.integers(0, 4, size=int(hit.sum()), dtype=np.uint8)
        seqs.append(s)
    seqs += [seqs[3].copy(), seqs[7].copy(), base.copy(), base.copy()]              # exact copies
    seqs += [rng.integers(0, 4, size=250_000, dtype=np.uint8) for _ in range(4)]    # unrelated rows
    flat, off = lib.concat(seqs)
    kf = lib.KFreqs.count(ctx, lib.SeqSet.upload(ctx, flat, off), k)
    _, of, _, _ = orc.count_batch(flat, off, k, want_counts=False)
    exp = orc.euclid_matrix(of)
    got = kf.euclidean()
    assert exp[3, 24] == 0.0 and exp[26, 27] == 0.0 and got[3, 24] == 0.0 and got[26, 27] == 0.0 and got[7, 25] == 0.0
    near = exp[:24, :24][np.triu_indices(24, 1)]
    far = exp[:24, 28:]
    assert near.max() < far.min() / 5  # the family really is near-duplicate
    np.testing.assert_allclose(got, exp, rtol=RTOL, atol=0)
    assert np.array_equal(got, got.T) and (np.diag(got) == 0).all()
    assert np.array_equal(kf.euclidean(5, 29), got[5:29])


def test_max_with_max_size_below_min_size_grows_like_the_reference(lib, ctx, orc):
    """records.rs:427-451: `size == max_size` never holds, so the set only ever grows (memory-safely)"""
    flat, off = lib.synth_host(31, 400, 8, 3000)
    kf = lib.KFreqs.count(ctx, lib.SeqSet.upload(ctx, flat, off), 3)
    _, of, oe, ov = orc.count_batch(flat, off, 3)
    order = np.random.default_rng(4).permutation(400).astype(np.uint32)
    for mode, omode, lo, hi in ((lib.MODE_MAX_STDEV, "stdev", 5, 3), (lib.MODE_MAX_COV, "cov", 6, 0),
                                (lib.MODE_MAX_STDEV, "stdev", 4, 0)):
        exp = orc.select_rows(of, oe, order, omode, lo, hi, valid=ov)
        idx, delta, stats = kf.select(order, mode, lo, hi)
        assert idx.tolist() == exp.ids.tolist() and np.array_equal(delta, exp.delta_jsd)
        assert idx.size >= lo and stats[0] == exp.total_jsd
    from diverseseq_b200 import _dvs as dvs
    st = dvs.make_zarr_store()
    names = [f"s{i}" for i in range(60)]
    for i, nm in enumerate(names):
        st.write(nm, flat[int(off[i]):int(off[i + 1])].tobytes())
    r = dvs.max_divergent(st, 5, 3, 3, seqids=names)  # must not overrun anything
    o = orc.select_rows(of[:60], oe[:60], np.arange(60), "stdev", 5, 3, valid=ov[:60])
    assert r.record_names == [names[i] for i in o.ids] and r.size == len(o.ids) >= 5


def test_count_refuses_nothing_below_4g_bases(lib, ctx):
    """the 2^32-base guard must not trigger on ordinary records (the limit itself cannot be tested at size)"""
    kf = lib.KFreqs.count(ctx, lib.SeqSet.from_seqs(ctx, [np.zeros(70_000, dtype=np.uint8)]), 2)
    c = kf.download()[0]
    assert int(c[0, 0]) == 69_999 and int(c.sum()) == 69_999


def test_include_rerun_matches_oracle(lib, orc, brca1):
    """`dvs nmost --include` (cli.py:445-474): shuffle with the seed, select, then nmost over selected + included
    names with n = their number; a name already selected is listed twice and contributes two rows, as upstream"""
    from diverseseq_b200 import _dvs as dvs, records
    store = dvs.make_zarr_store()
    names = list(brca1)
    for nm in names:
        store.write(nm, brca1[nm].tobytes())
    k, n, seed = 4, 8, 11
    order = records.shuffled_seqids(store.get_seqids(), seed)
    rng = np.random.default_rng(seed=seed)
    expect_order = list(names)
    rng.shuffle(expect_order)
    assert order == expect_order
    flat, off = lib.concat([brca1[nm] for nm in order])
    first = orc.select_seqs(flat, off, np.arange(len(order)), k, "nmost", n)
    sel_names = [order[i] for i in first.ids]
    include = ["Human", sel_names[2]]  # one new record and one that is already selected
    res = records.select_nmost(store, n, k, seed=seed, include=include)
    final_names = sel_names + include
    flat2, off2 = lib.concat([brca1[nm] for nm in final_names])
    second = orc.select_seqs(flat2, off2, np.arange(len(final_names)), k, "nmost", len(final_names))
    assert res.record_names == [final_names[i] for i in second.ids] and res.size == len(final_names)
    assert [r[2] for r in res.records] == second.delta_jsd.tolist() and res.total_jsd == second.total_jsd
    with pytest.raises(ValueError):
        records.select_nmost(store, n, k, seed=seed, include=["not-there"])
    m = records.select_max(store, 5, 9, k, stat="cov", seed=seed)
    mexp = orc.select_seqs(flat, off, np.arange(len(order)), k, "cov", 5, 9)
    assert m.record_names == [order[i] for i in mexp.ids]


@pytest.mark.parametrize("slices", ["0", "1", "2"])
def test_k8_round_slicing_modes_give_the_same_selection(lib, ctx, orc, k8_set, slices, monkeypatch):
    """DVS_SELECT_SLICES: candidates (1) and member slots (2) cut into slices over the idle CTAs of the cooperative
    kernel - partial sums combined in slice order, the extra additions covered by the error bounds, so every form
    must give the oracle's selection bit for bit"""
    kf, of, oe, ov, order = k8_set
    monkeypatch.setenv("DVS_SELECT_SLICES", slices)
    exp = orc.select_rows(of, oe, order, "nmost", 24, 24, valid=ov)
    idx, delta, stats = kf.select(order, lib.MODE_NMOST, 24, 24)
    assert idx.tolist() == exp.ids.tolist()
    assert np.array_equal(delta, exp.delta_jsd)
    assert stats[0] == exp.total_jsd and stats[1] == exp.mean_delta_jsd and stats[2] == exp.std_delta_jsd
    exp2 = orc.select_rows(of, oe, order, "cov", 6, 12, valid=ov)
    idx2, delta2, _ = kf.select(order, lib.MODE_MAX_COV, 6, 12)
    assert idx2.tolist() == exp2.ids.tolist() and np.array_equal(delta2, exp2.delta_jsd)
