"""Randomised GPU <-> oracle differential test: many small clouds of awkward shapes (duplicates,
lattices, lines, tight clusters next to sparse outliers, near-coincident points) with random k, so that
the rare paths — exact and near ties through the sorting network, crowded cells and capped queries in
the general k-NN path, levels of a few vertices under programmatic dependent launch — are all visited."""
import numpy as np
import pytest

from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from iftwt_rg_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def make(rng, kind, n):
    p = np.zeros((n, 4), np.float32)
    if kind == "uniform":
        p[:, :3] = rng.random((n, 3))
    elif kind == "lattice":
        p[:, :3] = rng.integers(0, 10, (n, 3)) * 0.1
    elif kind == "lattice_jitter":
        p[:, :3] = rng.integers(0, 10, (n, 3)) * 0.1 + rng.integers(-2, 3, (n, 3)) * 2.0 ** -23
    elif kind == "plane":
        p[:, :2] = rng.random((n, 2)) * 3
        p[:, 2] = rng.normal(0, 1e-3, n)
    elif kind == "line":
        t = rng.random(n)
        p[:, 0], p[:, 1], p[:, 2] = t, 2 * t, rng.normal(0, 1e-4, n)
    elif kind == "clusters":
        c = rng.random((8, 3)) * 10
        p[:, :3] = c[rng.integers(0, 8, n)] + rng.normal(0, 0.01, (n, 3))
        p[: n // 50, :3] = rng.random((n // 50, 3)) * 100          # sparse outliers around tight clusters
    elif kind == "duplicates":
        base = rng.random((max(1, n // 7), 3))
        p[:, :3] = base[rng.integers(0, len(base), n)]
    p[:, 3] = rng.integers(0, 12, n)                              # few weight values: ties in the IFT
    return p


KINDS = ["uniform", "lattice", "lattice_jitter", "plane", "line", "clusters", "duplicates"]


@pytest.mark.parametrize("seed", range(6))
def test_knn_graph_ift_random(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    for kind in KINDS:
        n = int(rng.integers(300, 30000))
        k = int(rng.choice([1, 2, 7, 16, 17, 24, 31, 32, 33, 48, 64]))
        pts = make(rng, kind, n)
        ctx.set_cloud(pts)
        idx, d2 = ctx.knn(k)
        ridx, rd2 = orc.knn_self(pts, k)
        assert np.array_equal(idx, ridx), (kind, n, k)
        assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32)), (kind, n, k)
        off, adj = ctx.knn_graph(k)
        roff, radj = orc.knn_graph(ridx)
        assert np.array_equal(off, roff) and np.array_equal(adj, radj), (kind, n, k)
        seeds = np.unique(rng.integers(0, n, int(rng.integers(1, 40)))).astype(np.int32)
        cost, label = ctx.ift(None, seeds)
        rcost, rlabel = orc.ift(roff, radj, pts[:, 3], seeds, orc.POLICY_C)
        assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel), (kind, n, k)


@pytest.mark.parametrize("seed", range(4))
def test_segment_and_voxel_random(ctx, seed):
    from iftwt_rg_b200 import api
    rng = np.random.default_rng(2000 + seed)
    for kind in ("plane", "uniform", "clusters"):
        n = int(rng.integers(2000, 40000))
        k = int(rng.choice([8, 16, 20, 32]))
        theta = float(rng.choice([0.05, 0.2, 0.5]))
        pts = make(rng, kind, n)
        ctx.set_cloud(pts)
        rg = api.RgParams(10, 100000, k, theta, 1.0)
        cost, label, ns = ctx.segment(k, k, rg)
        ridx, _ = orc.knn_self(pts, k)
        roff, radj = orc.knn_graph(ridx)
        rn = orc.normals(pts, ridx)
        rlab, _, kept, _ = orc.region_growing(rn, ridx, theta, 1.0, 10, 100000)
        rseeds, _ = orc.rg_seeds(pts, rlab, kept)
        rcost, rlabel = orc.ift(roff, radj, pts[:, 3], rseeds, orc.POLICY_C)
        assert ns == kept, (kind, n, k, theta)
        assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel), (kind, n, k, theta)
        # the reference's own graph builder on the same cloud
        ctx.set_cloud(pts)
        overlap = float(rng.choice([1.01, 2.0, 5.0]))
        res = ctx.cloud_resolution(overlap)
        assert res == orc.cloud_resolution(pts, overlap, 0), (kind, n, overlap)
        V = ctx.voxel_graph(res)
        cen, keys, voff, vadj, box = orc.voxel_graph(pts, res)
        assert V == len(cen) and np.array_equal(ctx.get_cloud().view(np.uint32), cen.view(np.uint32)), (kind, n, overlap)
        off, adj = ctx.graph_csr()
        assert np.array_equal(off, voff) and np.array_equal(adj, vadj), (kind, n, overlap)
        sd = ctx.regmin_seeds()
        assert np.array_equal(sd, orc.regmin(voff, vadj, cen[:, 3])), (kind, n, overlap)
        c2, l2 = ctx.ift(None, None)
        rc2, rl2 = orc.ift(voff, vadj, cen[:, 3], sd, orc.POLICY_C)
        assert np.array_equal(c2, rc2) and np.array_equal(l2, rl2), (kind, n, overlap)
