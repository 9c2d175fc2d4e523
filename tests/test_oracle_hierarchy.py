"""Region hierarchy restatements of the oracle (I4 / L1): the product definition (saddle
region graph + single-linkage merge) on hand-checkable graphs, and the literal Cluster of
cluster.hpp:23-229 on the literal MST of ift.hpp:207-243 (informational)."""
import numpy as np

from oracle import pyoracle as orc
from test_oracle_ift_rg import csr_from_edges


def two_basins():
    #  path of 9 vertices, two basins separated by a ridge of height 9 at vertex 4
    w = np.array([0, 2, 3, 5, 9, 4, 1, 0, 6], np.float32)
    off, adj = csr_from_edges(9, [(i, i + 1) for i in range(8)])
    return off, adj, w


def test_saddle_and_attributes_known_answer():
    off, adj, w = two_basins()
    seeds = np.array([0, 7], np.int32)
    cost, label = orc.ift(off, adj, w, seeds, orc.POLICY_C)
    assert cost.tolist() == [0, 2, 3, 5, 9, 4, 1, 0, 6]
    assert label.tolist() == [0, 0, 0, 0, 0, 7, 7, 7, 7]          # tie at the ridge -> smaller seed
    a, b, s = orc.region_adjacency(off, adj, cost, label)
    assert (a.tolist(), b.tolist(), s.tolist()) == ([0], [7], [9])  # the pass is the ridge
    seed, area, height = orc.region_info(w, cost, label)
    assert seed.tolist() == [0, 7] and area.tolist() == [5, 4] and height.tolist() == [9.0, 6.0]
    cl, k = orc.cluster_saddle(cost, label, a, b, s, 2)
    assert k == 2 and np.array_equal(cl, label)
    cl, k = orc.cluster_saddle(cost, label, a, b, s, 1)
    assert k == 1 and np.all(cl == 0)


def test_single_linkage_merges_lowest_saddles_first():
    rng = np.random.default_rng(8)
    n = 24
    ii = lambda x, y: x * n + y
    edges = [(ii(x, y), ii(x + 1, y)) for x in range(n - 1) for y in range(n)] + \
            [(ii(x, y), ii(x, y + 1)) for x in range(n) for y in range(n - 1)]
    off, adj = csr_from_edges(n * n, edges)
    xx, yy = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    w = np.floor(30 + 20 * np.sin(xx / 2.5) * np.cos(yy / 3.0) + rng.integers(0, 3, (n, n))).astype(np.float32).ravel()
    seeds = orc.regmin(off, adj, w)
    cost, label = orc.ift(off, adj, w, seeds, orc.POLICY_C)
    a, b, s = orc.region_adjacency(off, adj, cost, label)
    S = len(np.unique(label))
    prev = None
    for k in (S, 12, 6, 3, 1):
        cl, got = orc.cluster_saddle(cost, label, a, b, s, k)
        assert got == max(k, 1) == len(np.unique(cl))
        # every cluster is a union of regions and is named by its smallest seed
        for c in np.unique(cl):
            assert c == label[cl == c].min()
        # the hierarchy is nested
        if prev is not None:
            assert len(set(zip(prev.tolist(), cl.tolist()))) == len(np.unique(prev))
        prev = cl
    # the literal Cluster runs and returns a partition into unions of regions (informational)
    c2, l2, ma, mb, mw = orc.ift_literal_mst(off, adj, w, seeds)
    rg, kk = orc.cluster_literal(l2, ma, mb, mw, 4)
    if kk > 0:
        for r in np.unique(rg[rg >= 0]):
            labs = np.unique(l2[rg == r])
            assert np.all(np.isin(l2[np.isin(l2, labs)], labs))
