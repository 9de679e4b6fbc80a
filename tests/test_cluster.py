"""Average-linkage cluster tree (diverse_seq/cluster.py:191-237).

CPU part: the oracle restatement (oracle/oracle.py::linkage_average) against scikit-learn's
AgglomerativeClustering(metric="precomputed", linkage="average") — the reference's own dependency,
installed here and on the GPU box — including heavy ties.  GPU part: dvs_linkage_average (csrc/cluster.cu)
against scikit-learn and the oracle, bit-exact children_ and heights, and the ctree app end to end.
"""
import numpy as np
import pytest

from oracle import oracle

gpu = pytest.mark.gpu


def _sklearn(d):
    from sklearn.cluster import AgglomerativeClustering

    m = AgglomerativeClustering(metric="precomputed", linkage="average", compute_distances=True).fit(d)
    return m.children_, m.distances_


def _matrices(rng, n):
    x = rng.random((n, 5))
    yield "euclid", np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1))
    a = np.triu(rng.integers(1, 4, (n, n)).astype(float), 1)
    yield "small-int ties", a + a.T
    yield "all equal", np.ones((n, n)) - np.eye(n)
    m = np.triu(np.where(rng.random((n, n)) < 0.7, 1.0, rng.random((n, n))), 1)
    yield "mash-like (many 1.0)", m + m.T
    b = np.triu(rng.integers(0, 2, (n, n)).astype(float), 1)
    yield "zeros and ones", b + b.T


@pytest.mark.parametrize("n", [2, 3, 4, 7, 33, 150])
def test_oracle_linkage_matches_sklearn(n):
    rng = np.random.default_rng(n)
    for name, d in _matrices(rng, n):
        c, h, cnt = oracle.linkage_average(d)
        sc, sh = _sklearn(d)
        assert np.array_equal(c, sc), name
        assert np.array_equal(h, sh), name
        assert cnt[-1] == n


def test_tree_string_like_reference():
    from diverseseq_b200.cluster import ClusterTree

    # children_ of 4 leaves: (0,1) -> 4, (2,3) -> 5, (4,5) -> 6  (cluster.py:222-232)
    t = ClusterTree(["a", "b", "c", "d"], np.array([[0, 1], [2, 3], [4, 5]]), np.zeros(3), np.zeros(3))
    assert t.treestring == "((a, b), (c, d))"
    assert t.get_tip_names() == ["a", "b", "c", "d"]
    t = ClusterTree(["a", "b", "c"], np.array([[1, 2], [0, 3]]), np.zeros(2), np.zeros(2))
    assert t.treestring == "(a, (b, c))"


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    from diverseseq_b200 import _lib

    return _lib.Context(0)


@gpu
@pytest.mark.parametrize("n", [2, 3, 5, 32, 33, 257, 1025, 1500])
def test_linkage_matches_sklearn(ctx, n):
    from diverseseq_b200 import _lib

    rng = np.random.default_rng(100 + n)
    for name, d in _matrices(rng, n):
        c, h, cnt = _lib.linkage_average(ctx, d)
        sc, sh = _sklearn(d)
        assert np.array_equal(c, sc), (name, n)
        assert np.array_equal(h, sh), (name, n)
        assert int(cnt[-1]) == n
        if n <= 257:
            oc, oh, ocnt = oracle.linkage_average(d)
            assert np.array_equal(c, oc) and np.array_equal(h, oh) and np.array_equal(cnt, ocnt)


@gpu
def test_linkage_uses_upper_triangle_and_trivial_sizes(ctx):
    from diverseseq_b200 import _lib

    rng = np.random.default_rng(4)
    n = 40
    u = np.triu(rng.random((n, n)), 1)
    d = u + u.T
    noisy = d + np.tril(rng.random((n, n)), -1)  # garbage below the diagonal must not matter (sklearn reads triu)
    c, h, _ = _lib.linkage_average(ctx, noisy)
    sc, sh = _sklearn(d)
    assert np.array_equal(c, sc) and np.array_equal(h, sh)
    for m in (0, 1):
        c, h, cnt = _lib.linkage_average(ctx, np.zeros((m, m)))
        assert c.shape == (0, 2) and h.size == 0
    with pytest.raises(ValueError):
        _lib.linkage_average(ctx, np.full((4, 4), np.nan))
    with pytest.raises(ValueError):
        _lib.linkage_average(ctx, np.zeros((3, 4)))


@gpu
def test_linkage_device_matrix_from_distance_kernels(ctx, brca1):
    """ctree end to end: distances written to device memory, linkage reads them there; the tree equals
    sklearn's on the host copy of the same matrix (cluster.py:152-188)"""
    import torch

    from diverseseq_b200 import _lib
    from diverseseq_b200.cluster import dvs_ctree, make_cluster_tree

    names = list(brca1)
    ss = _lib.SeqSet.from_seqs(ctx, [brca1[n] for n in names])
    n = len(names)
    for mode in ("mash", "euclidean"):
        if mode == "mash":
            sk = _lib.Sketches.sketch(ctx, ss, 8, 400, 4, True)
            host = sk.distances(8, 400)
        else:
            host = _lib.KFreqs.count(ctx, ss, 4).euclidean()
        sc, sh = _sklearn(host)
        app = dvs_ctree(k=8 if mode == "mash" else 4, sketch_size=400, distance_mode=mode,
                        mash_canonical_kmers=True, ctx=ctx)
        tree = app(names, [brca1[x] for x in names])
        assert np.array_equal(tree.children, sc), mode
        assert np.array_equal(tree.heights, sh), mode
        assert sorted(tree.get_tip_names()) == sorted(names)
        t2 = make_cluster_tree(names, host, ctx=ctx)
        assert t2.treestring == tree.treestring
        d = torch.from_numpy(host).cuda()
        t3 = make_cluster_tree(names, (d.data_ptr(), n), ctx=ctx)
        assert t3.treestring == tree.treestring


@gpu
def test_linkage_large_random(ctx):
    """n = 4000 with many ties (quantised distances) against scikit-learn"""
    from diverseseq_b200 import _lib

    rng = np.random.default_rng(8)
    n = 4000
    x = rng.random((n, 3)).astype(np.float32)
    d = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)).astype(np.float64)
    d = np.round(d, 2)  # ties
    d = np.triu(d, 1)
    d = d + d.T
    c, h, _ = _lib.linkage_average(ctx, d)
    sc, sh = _sklearn(d)
    assert np.array_equal(c, sc)
    assert np.array_equal(h, sh)
