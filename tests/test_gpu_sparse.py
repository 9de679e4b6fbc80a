"""Sparse k-mer counting (9 <= k <= 12) against the oracle's dense vectors: the sparse row of a record must be
exactly the non-zero entries of `count_kmers` (/root/reference/src/record.rs:41-84) in index order; totals and,
when requested, the sequential-order entropy (record.rs:86-106) are bit-identical.  The k=12 case at full
genome length is the north star's "k=12 counting" (dense rows of 134 MB per record are out of reach)."""
import numpy as np
import pytest

from conftest import random_seqs

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


def check_records(lib, ctx, orc, seqs, k, entropy=False):
    flat, off = lib.concat(seqs)
    sp = lib.KSparse.count(ctx, lib.SeqSet.upload(ctx, flat, off), k, want_entropy=entropy)
    nnz, tot, ent, valid = sp.stats()
    assert sp.nrec == len(seqs)
    for r, s in enumerate(seqs):
        dense = orc.kcounts(s, k)
        want_idx = np.flatnonzero(dense)
        idx, cnt = sp.record(r)
        assert int(nnz[r]) == want_idx.size and int(tot[r]) == int(dense.sum()) and bool(valid[r]) == bool(dense.sum())
        assert np.array_equal(idx, want_idx.astype(np.uint32))
        assert np.array_equal(cnt.astype(np.uint64), dense[want_idx].astype(np.uint64))
        if entropy and dense.sum():
            assert ent[r] == orc.entropy(dense / float(dense.sum()))  # bitwise
    return sp


@pytest.mark.parametrize("k", [9, 10, 11, 12])
def test_sparse_brca1(lib, ctx, orc, brca1, k):
    seqs = list(brca1.values())[:20]
    check_records(lib, ctx, orc, seqs, k, entropy=(k == 9))


def test_sparse_ragged_invalid_and_empty_records(lib, ctx, orc):
    rng = np.random.default_rng(12)
    seqs = random_seqs(rng, 30, 1, 70_000, invalid_rate=0.002)
    seqs[3] = np.zeros(0, dtype=np.uint8)                      # empty record
    seqs[5] = np.full(500, 4, dtype=np.uint8)                  # no valid k-mer
    seqs[7] = np.zeros(100_000, dtype=np.uint8)                # homopolymer: one bin, one bucket, > 32768 per item pair
    seqs[9] = rng.integers(0, 4, size=11, dtype=np.uint8)      # shorter than k = 12
    seqs[11] = np.tile(np.array([0, 1, 2, 3, 3, 1], dtype=np.uint8), 30_000)  # period-6 repeat
    check_records(lib, ctx, orc, seqs, 12, entropy=True)
    check_records(lib, ctx, orc, seqs, 10)


def test_sparse_k12_full_length_genomes(lib, ctx, orc):
    """64 synthetic genomes of ~4 Mbp at k=12: every record's distinct k-mers and counts"""
    nrec = 64
    ss = lib.SeqSet.synth(ctx, 1212, nrec, 8, 4_000_000)
    sp = lib.KSparse.count(ctx, ss, 12)
    nnz, tot, _ent, valid = sp.stats()
    off = ss.offsets()
    assert valid.all()
    for r in list(range(0, nrec, 7)) + [nrec - 1]:
        seq = ss.download(r, 1)
        dense = orc.kcounts(seq, 12)
        want_idx = np.flatnonzero(dense)
        idx, cnt = sp.record(r)
        assert int(tot[r]) == int(dense.sum()) and int(nnz[r]) == want_idx.size
        assert np.array_equal(idx, want_idx.astype(np.uint32)) and np.array_equal(cnt, dense[want_idx].astype(np.uint32))
    # every record: the counts add up to its number of valid 12-mers (checked against a k=6 style total is not
    # possible; the run-length identity total <= len - k + 1 is)
    lens = np.diff(off.astype(np.int64))
    assert (tot.astype(np.int64) <= lens - 11).all() and (nnz <= tot).all() and (nnz > 0.3 * tot).all()


def test_sparse_argument_errors(lib, ctx):
    ss = lib.SeqSet.from_seqs(ctx, [np.zeros(100, dtype=np.uint8)])
    with pytest.raises(TypeError):
        lib.KSparse.count(ctx, ss, 8)
    with pytest.raises(TypeError):
        lib.KSparse.count(ctx, ss, 12, num_states=5)
