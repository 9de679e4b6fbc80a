"""dvs_count_select: counting with the nmost rounds trailing it on the same GPU.

The call must give exactly what dvs_count_kmers followed by dvs_select give - rows, entropies, selection, deltas and
statistics - whatever the chunking, because the rounds only ever examine positions whose rows have been published
(reference semantics: /root/reference/src/records.rs:311-342 nmost over the records /root/reference/src/record.rs
KmerSeq::new builds; the oracle pins the sequential form).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


def _same(lib, a, b):
    (fa, ia, da, sa), (fb, ib, db, sb) = a, b
    assert ia.tolist() == ib.tolist()
    assert np.array_equal(da, db) and np.array_equal(sa, sb)
    ca, ra, ea, va = fa.download()
    cb, rb, eb, vb = fb.download()
    assert np.array_equal(ca, cb) and np.array_equal(ra, rb) and np.array_equal(ea, eb) and np.array_equal(va, vb)


@pytest.fixture(scope="module")
def families(lib, ctx):
    """3,000 synthetic genomes of ~120 kbp in 24 families, a few invalid runs"""
    nrec = 3000
    flat, off = lib.synth_host(4242, nrec, 24, 120_000)
    return lib.SeqSet.upload(ctx, flat, off), flat, off, nrec


@pytest.mark.parametrize("k", [4, 5, 6])
@pytest.mark.parametrize("n,chunks", [(3, 2), (20, 8), (100, 0), (100, 16), (256, 5), (40, 64)])
def test_count_select_equals_the_two_calls(lib, ctx, families, k, n, chunks):
    ss, _, _, nrec = families
    order = np.random.default_rng(100 * k + n).permutation(nrec).astype(np.uint32)
    kf = lib.KFreqs.count(ctx, ss, k)
    idx, delta, stats = kf.select(order, lib.MODE_NMOST, n, n)
    accepts = int(ctx._lib.dvs_select_last_accepts(ctx.handle))
    got = lib.KFreqs.count_select(ctx, ss, k, order, lib.MODE_NMOST, n, n, chunks=chunks)
    assert int(ctx._lib.dvs_select_last_accepts(ctx.handle)) == accepts
    _same(lib, got, (kf, idx, delta, stats))
    # a plain count after the sequenced one rebuilds its work list
    kf2 = lib.KFreqs.count(ctx, ss, k)
    assert np.array_equal(kf2.download()[0], kf.download()[0])


def test_rounds_really_trail_the_counting(lib, ctx):
    """long records: the counting takes milliseconds, so most accepts must happen while it is still running"""
    nrec = 1200
    flat, off = lib.synth_host(77, nrec, 32, 2_000_000)
    ss = lib.SeqSet.upload(ctx, flat, off)
    order = np.random.default_rng(5).permutation(nrec).astype(np.uint32)
    kf = lib.KFreqs.count(ctx, ss, 6)
    ref = (kf,) + kf.select(order, lib.MODE_NMOST, 50, 50)
    best = 0
    for _ in range(3):
        got = lib.KFreqs.count_select(ctx, ss, 6, order, lib.MODE_NMOST, 50, 50, chunks=12)
        _same(lib, got, ref)
        # the trailing kernel is always launched (a host decision); how many accepts it makes before the counting ends
        # depends on the timing, so that number is only reported
        assert int(ctx._lib.dvs_select_last_trail_launches(ctx.handle)) >= 1
        assert int(ctx._lib.dvs_select_last_trail_sms(ctx.handle)) >= 1
        best = max(best, int(ctx._lib.dvs_select_last_trail_accepts(ctx.handle)))
    assert best <= int(ctx._lib.dvs_select_last_accepts(ctx.handle))
    if best == 0:
        import warnings
        warnings.warn("no accept was made while the counting was still running (timing dependent)")


def test_count_select_matches_the_oracle(lib, ctx, orc, families):
    ss, flat, off, nrec = families
    order = np.random.default_rng(9).permutation(nrec).astype(np.uint32)
    _, of, oe, ov = orc.count_batch(flat, off, 5, threads=max(1, orc.hardware_threads()), want_counts=False)
    exp = orc.select_rows(of, oe, order, "nmost", 30, 30, valid=ov)
    kf, idx, delta, stats = lib.KFreqs.count_select(ctx, ss, 5, order, lib.MODE_NMOST, 30, 30, chunks=6)
    assert idx.tolist() == exp.ids.tolist()
    assert np.array_equal(delta, exp.delta_jsd)
    assert stats[0] == exp.total_jsd and stats[1] == exp.mean_delta_jsd and stats[2] == exp.std_delta_jsd
    _, f, e, v = kf.download(counts=False)
    assert np.array_equal(f, of) and np.array_equal(e, oe) and np.array_equal(v, ov)


def test_subset_orders_duplicates_and_other_modes(lib, ctx, families):
    ss, _, _, nrec = families
    rng = np.random.default_rng(31)
    kf = lib.KFreqs.count(ctx, ss, 5)
    # `order` names only some of the records (the --include / limit flows), the rest is counted last
    sub = rng.permutation(nrec)[:700].astype(np.uint32)
    _same(lib, lib.KFreqs.count_select(ctx, ss, 5, sub, lib.MODE_NMOST, 25, 25, chunks=7),
          (kf,) + kf.select(sub, lib.MODE_NMOST, 25, 25))
    # a record named twice
    dup = np.concatenate([sub[:300], sub[100:200], sub[300:]]).astype(np.uint32)
    _same(lib, lib.KFreqs.count_select(ctx, ss, 5, dup, lib.MODE_NMOST, 10, 10, chunks=4),
          (kf,) + kf.select(dup, lib.MODE_NMOST, 10, 10))
    # the max modes and k above the shared-memory limit of the trailing kernel run the two steps back to back
    order = rng.permutation(nrec).astype(np.uint32)
    _same(lib, lib.KFreqs.count_select(ctx, ss, 5, order, lib.MODE_MAX_STDEV, 5, 30, chunks=8),
          (kf,) + kf.select(order, lib.MODE_MAX_STDEV, 5, 30))
    kf7 = lib.KFreqs.count(ctx, ss, 7)
    _same(lib, lib.KFreqs.count_select(ctx, ss, 7, order, lib.MODE_NMOST, 12, 12, chunks=8),
          (kf7,) + kf7.select(order, lib.MODE_NMOST, 12, 12))
    # min_size larger than the first chunk
    _same(lib, lib.KFreqs.count_select(ctx, ss, 5, order[:64], lib.MODE_NMOST, 60, 60, chunks=8),
          (kf,) + kf.select(order[:64], lib.MODE_NMOST, 60, 60))


def test_errors_are_the_ones_of_the_separate_calls(lib, ctx, families):
    ss, _, _, nrec = families
    order = np.arange(nrec, dtype=np.uint32)
    with pytest.raises(Exception, match="k cannot be 0"):
        lib.KFreqs.count_select(ctx, ss, 0, order, lib.MODE_NMOST, 5, 5)
    with pytest.raises(Exception, match="out of range"):
        lib.KFreqs.count_select(ctx, ss, 4, np.array([0, 1, nrec], dtype=np.uint32), lib.MODE_NMOST, 2, 2)
    # all records of the initial set invalid -> "records cannot be empty", and the context stays usable
    flat = np.full(4000, 4, dtype=np.uint8)
    off = np.array([0, 1000, 2000, 3000, 4000], dtype=np.uint64)
    bad = lib.SeqSet.upload(ctx, flat, off)
    with pytest.raises(Exception, match="records cannot be empty"):
        lib.KFreqs.count_select(ctx, bad, 4, np.arange(4, dtype=np.uint32), lib.MODE_NMOST, 2, 2, chunks=2)
    kf, idx, _, _ = lib.KFreqs.count_select(ctx, ss, 4, order, lib.MODE_NMOST, 5, 5)
    assert idx.size == 5
