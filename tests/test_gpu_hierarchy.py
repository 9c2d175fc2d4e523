"""GPU <-> oracle parity of the region hierarchy (region adjacency + saddles, region
attributes, merge to num_regions) on both graph builders."""
import numpy as np
import pytest

from iftwt_rg_b200 import synth
from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from iftwt_rg_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def check(ctx, off, adj, w, cost, label, ks):
    a, b, s = ctx.region_adjacency()
    ra, rb, rs = orc.region_adjacency(off, adj, cost, label)
    assert np.array_equal(a, ra) and np.array_equal(b, rb) and np.array_equal(s, rs)
    seed, area, height = ctx.region_info()
    rseed, rarea, rheight = orc.region_info(w, cost, label)
    assert np.array_equal(seed, rseed) and np.array_equal(area, rarea) and np.array_equal(height, rheight)
    for k in ks:
        cl, nk = ctx.cluster(k)
        rcl, rk = orc.cluster_saddle(cost, label, ra, rb, rs, k)
        assert nk == rk and np.array_equal(cl, rcl), k


def test_hierarchy_on_knn_graph(ctx):
    pts = synth.scene_primitives(150_000, seed=31)
    rng = np.random.default_rng(1)
    ctx.set_cloud(pts)
    off, adj = ctx.knn_graph(12)
    seeds = rng.integers(0, len(pts), 400).astype(np.int32)
    w = pts[:, 3].copy()
    w[rng.integers(0, len(pts), 50)] = 900.0           # holes the IFT never conquers
    cost, label = ctx.ift(w, seeds)
    rcost, rlabel = orc.ift(off, adj, w, seeds, orc.POLICY_C)
    assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel)
    check(ctx, off, adj, w, cost, label, (10, 3, 1, 100000))


def test_hierarchy_on_voxel_graph_with_regmin_seeds(ctx):
    """main.cpp:124: Cluster(..., 10) after the IFT, here on the reference's own graph."""
    pts = synth.scene_statue(150_000)
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(2.0)
    ctx.voxel_graph(res)
    grad = ctx.morph(5)
    ctx.set_intensity(grad)
    off, adj = ctx.graph_csr()
    seeds = ctx.regmin_seeds()
    cost, label = ctx.ift(None, None)
    rcost, rlabel = orc.ift(off, adj, grad, seeds, orc.POLICY_C)
    assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel)
    check(ctx, off, adj, grad, cost, label, (10, 2))
