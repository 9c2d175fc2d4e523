"""Edge cases of the domain, GPU vs oracle through the C ABI: degenerate geometry, empty and
duplicate seeds, closed weights, non-finite input, capacity errors, stage ordering errors, and
the BASELINE full-size cloud through size-independent properties."""
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


def test_degenerate_geometry_normals(ctx):
    # exact plane, exact line, and a tight duplicate cluster: the computeRoots2 / zero-scale branches
    xs, ys = np.meshgrid(np.arange(40), np.arange(40))
    plane = np.stack([xs.ravel() * 0.125, ys.ravel() * 0.125, np.full(1600, 0.5)], 1)
    line = np.stack([np.arange(300) * 0.25, np.full(300, 7.0), np.full(300, -2.0)], 1)
    dup = np.tile(np.array([[3.0, 3.0, 3.0]]), (50, 1))
    pts = np.zeros((1950, 4), np.float32)
    pts[:, :3] = np.concatenate([plane, line, dup])
    ctx.set_cloud(pts)
    for k in (5, 12):
        out = ctx.normals(k)
        idx, _ = orc.knn_self(pts, k)
        ref = orc.normals(pts, idx)
        # 0/0 normals of the coincident points: NaN on both sides (x86 and CUDA differ in the NaN
        # payload, nothing else); every other value bit-exact
        nan = np.isnan(ref)
        assert np.array_equal(np.isnan(out), nan) and nan.any(), k
        assert np.array_equal(out.view(np.uint32)[~nan], ref.view(np.uint32)[~nan]), k
    two = pts[:2].copy()
    ctx.set_cloud(two)
    assert np.all(np.isnan(ctx.normals(4)))                  # fewer than 3 neighbours -> NaN (PCL)


def test_ift_without_seeds_duplicate_seeds_and_closed_seeds(ctx):
    pts = synth.scene_statue(30_000)
    ctx.set_cloud(pts)
    off, adj = ctx.knn_graph(8)
    w = pts[:, 3].copy()
    cost, label = ctx.ift(w, np.zeros(0, np.int32))
    assert np.all(cost == 500) and np.array_equal(label, np.arange(len(pts), dtype=np.uint32))
    seeds = np.array([5, 5, 17, 17, 17, 29999], np.int32)    # duplicates collapse (ift.hpp:132-142)
    w[17] = 800.0                                            # a seed's own value never matters (ift.hpp:138)
    cost, label = ctx.ift(w, seeds)
    rcost, rlabel = orc.ift(off, adj, w, seeds, orc.POLICY_C)
    assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel)
    assert cost[17] == 0 and label[17] == 17


def test_errors_are_loud(ctx):
    from iftwt_rg_b200 import api
    pts = synth.scene_primitives(5000, seed=2)
    bad = pts.copy()
    bad[7, 1] = np.nan
    ctx.set_cloud(bad)
    with pytest.raises(api.IftwtError) as e:
        ctx.knn(4)
    assert e.value.code == -4                                # IFTWT_ERR_NONFINITE
    ctx.set_cloud(pts)
    with pytest.raises(api.IftwtError) as e:
        ctx.ift(None, np.array([1], np.int32))               # no graph yet
    assert e.value.code == -3
    with pytest.raises(api.IftwtError) as e:
        ctx.region_seeds()                                   # no normals yet
    assert e.value.code == -3
    ctx.knn_graph(6, download=False)
    with pytest.raises(api.IftwtError) as e:
        ctx.ift(None, np.array([123456], np.int32))          # seed index out of range
    assert e.value.code == -1
    with pytest.raises(api.IftwtError) as e:
        ctx.knn(65)                                          # k out of range
    assert e.value.code == -1
    # the ctx is still usable after every error
    idx, _ = ctx.knn(6)
    assert np.array_equal(idx, orc.knn_self(pts, 6)[0])


def test_baseline_full_size_10m_properties(ctx):
    """cfg3 at its full size (10 M points, k = 32): no CPU k-NN at this size; size-independent
    properties instead — self first, sorted rows, the admissibility audit of the labelling over
    the exported graph, determinism across two runs."""
    from iftwt_rg_b200 import api
    n, k = 10_000_000, 32
    pts = np.concatenate([synth.scene_facade(2_000_000, start=s) for s in range(0, n, 2_000_000)])
    ctx.set_cloud(pts)
    rg = api.RgParams(50, 10000, k, 0.08, 1.0)
    cost, label, ns = ctx.segment(k, k, rg)
    assert ns > 10
    off, adj = ctx.graph_csr()
    assert off[-1] == len(adj) and len(off) == n + 1
    # rows ascending, self-free (sampled: the full check is O(E) in numpy)
    rng = np.random.default_rng(0)
    for v in rng.integers(0, n, 20000):
        row = adj[off[v]:off[v + 1]]
        assert np.all(np.diff(row.astype(np.int64)) > 0) and v not in row
    seeds = np.flatnonzero((cost == 0) & (label == np.arange(n, dtype=np.uint32))).astype(np.int32)
    assert len(seeds) == len(np.unique(label[cost < 500]))
    assert orc.ift_audit(off, adj, pts[:, 3], seeds, cost, label) == 0
    cost2, label2, ns2 = ctx.segment(k, k, rg)               # schedule independence of the whole path
    assert ns2 == ns and np.array_equal(cost, cost2) and np.array_equal(label, label2)
    st = ctx.stats()
    assert st["knn_fallback"] < 0.05 * n


def test_ift_lakes_and_plateaus_are_schedule_independent(ctx):
    """The IFT's lock-free level step (offers and hooks interleaving on one word per vertex) on
    the inputs that stress it: a few weight levels, so every level is made of huge plateaus;
    smooth basins that stay unconquered lakes for several levels and are then reached from many
    sides at once; thousands of seeds.  Repeated runs must give the same words, and those must be
    the oracle's (policy C)."""
    pts = synth.scene_facade(1_000_000, seed=0x1F77)
    n = len(pts)
    rng = np.random.default_rng(5)
    x, y = pts[:, 0], pts[:, 1]
    basins = np.floor(3.99 * (0.5 + 0.5 * np.sin(1.7 * x) * np.cos(2.3 * y)))           # 0..3, metre-scale lakes
    salt = (rng.random(n) < 0.15) * rng.integers(0, 3, n)
    w = (basins + salt).astype(np.float32)
    seeds = rng.integers(0, n, 3000).astype(np.int32)
    seeds = seeds[basins[seeds] >= 2]                                                    # keep the basins seedless: lakes
    ctx.set_cloud(pts)
    off, adj = ctx.knn_graph(16)
    runs = [ctx.ift(w, seeds) for _ in range(3)]
    for cost, label in runs[1:]:
        assert np.array_equal(cost, runs[0][0]) and np.array_equal(label, runs[0][1])
    rcost, rlabel = orc.ift(off, adj, w, seeds, orc.POLICY_C)
    assert np.array_equal(runs[0][0], rcost)
    assert np.array_equal(runs[0][1], rlabel)
    assert orc.ift_audit(off, adj, w, seeds, runs[0][0], runs[0][1]) == 0
    assert (rcost < 500).mean() > 0.99


def test_density_spread_sizes_the_cell_for_the_sparse_part():
    """grid.cu: a uniformly sampled surface leaves the cell alone; a cloud whose ground plane is sampled at a
    third of the density of its other parts gets a larger cell and (nearly) no query leaves the fast path.
    Either way the rows are exact (every parity test runs through this code)."""
    from iftwt_rg_b200 import api
    with api.Context(0) as ctx:
        fac = synth.scene_facade(600_000, seed=3)
        ctx.set_cloud(fac)
        ctx.knn(16, download=False)
        s_fac = ctx.stats()
        prim = synth.scene_primitives(600_000, seed=3)
        ctx.set_cloud(prim)
        idx, d2 = ctx.knn(16)
        s_prim = ctx.stats()
    assert 0.9 < s_fac["density_spread"] < 1.25 <= s_prim["density_spread"] < 2.5
    assert s_prim["knn_fallback"] < 0.02 * len(prim)
    ridx, rd2 = orc.knn_self(prim, 16)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)
