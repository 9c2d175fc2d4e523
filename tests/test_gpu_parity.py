"""GPU <-> oracle parity, through the C ABI (libiftwt.so via ctypes).

Bar: bit-exact for indices, adjacency, IFT costs and labels; curvature / normals
within 1e-5 relative (north_star) — in practice bit-exact, because the CUDA kernel
and the oracle evaluate the same FP32 operation order and the same specified
elementary functions."""
import numpy as np
import pytest

from iftwt_rg_b200 import synth
from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu

CURV_RTOL = 1e-5  # north_star tolerance


@pytest.fixture(scope="module")
def ctx():
    from iftwt_rg_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def clouds():
    return {
        "primitives200k": synth.scene_primitives(200_000),
        "statue100k": synth.scene_statue(100_000),
        "facade300k": synth.scene_facade(300_000),
    }


@pytest.mark.parametrize("name,k", [("primitives200k", 16), ("statue100k", 30), ("facade300k", 32),
                                    ("statue100k", 50), ("primitives200k", 2), ("statue100k", 1)])
def test_knn_bit_exact(ctx, name, k):
    pts = clouds()[name]
    ctx.set_cloud(pts)
    idx, d2 = ctx.knn(k)
    ridx, rd2 = orc.knn_self(pts, k)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))


def test_knn_ties_duplicates_and_small_clouds(ctx):
    rng = np.random.default_rng(3)
    g = rng.integers(0, 12, size=(20000, 3)).astype(np.float32) * 0.125   # lattice: massive ties + duplicates
    pts = np.zeros((20000, 4), np.float32)
    pts[:, :3] = g
    ctx.set_cloud(pts)
    for k in (8, 16, 24, 32, 40):   # 24 / 32: the sorting-network path must hand exact ties to the comparison ranking
        idx, d2 = ctx.knn(k)
        ridx, rd2 = orc.knn_self(pts, k)
        assert np.array_equal(idx, ridx), k
        assert np.array_equal(d2, rd2), k
    # near ties: the lattice with a jitter of a few ulp — distances that differ only in their last bits
    # share a quantised key in the sorting network and have to be ordered by their exact d2
    jit = pts.copy()
    jit[:, :3] += (rng.integers(-2, 3, size=(20000, 3)) * 2.0 ** -22).astype(np.float32)
    ctx.set_cloud(jit)
    for k in (20, 32):
        idx, d2 = ctx.knn(k)
        ridx, rd2 = orc.knn_self(jit, k)
        assert np.array_equal(idx, ridx), k
        assert np.array_equal(d2, rd2), k
    # fewer points than k
    small = synth.scene_primitives(5, seed=3)
    ctx.set_cloud(small)
    idx, d2 = ctx.knn(8)
    ridx, rd2 = orc.knn_self(small, 8)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)
    # one point
    one = synth.scene_primitives(1, seed=4)
    ctx.set_cloud(one)
    idx, d2 = ctx.knn(4)
    assert idx.tolist() == [[0, -1, -1, -1]]


def test_knn_outliers_and_mixed_density(ctx):
    pts = synth.scene_primitives(60_000, seed=9)
    rng = np.random.default_rng(2)
    far = np.zeros((40, 4), np.float32)
    far[:, :3] = rng.uniform(-60, 60, size=(40, 3))
    dense = synth.scene_statue(40_000)
    dense[:, :3] *= 0.05
    allp = np.concatenate([pts, far, dense]).astype(np.float32)
    ctx.set_cloud(allp)
    idx, d2 = ctx.knn(16)
    ridx, rd2 = orc.knn_self(allp, 16)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)
    assert ctx.stats()["knn_fallback"] > 0   # the general path really ran


def test_external_queries(ctx):
    pts = synth.scene_facade(100_000)
    ctx.set_cloud(pts)
    rng = np.random.default_rng(5)
    q = np.zeros((3000, 3), np.float32)
    q[:2000] = pts[rng.integers(0, len(pts), 2000), :3] + rng.normal(0, 0.02, (2000, 3)).astype(np.float32)
    q[2000:] = rng.uniform(-30, 70, size=(1000, 3))          # far outside the cloud's box too
    q4 = np.zeros((3000, 4), np.float32)
    q4[:, :3] = q
    for k in (1, 5):
        idx, d2 = ctx.knn_query(q, k)
        ridx, rd2 = orc.knn_query(pts, q4, k)
        assert np.array_equal(idx, ridx), k
        assert np.array_equal(d2, rd2), k


def test_pcl_record_layouts(ctx):
    """PointXYZI (32-byte records, intensity at 16) goes in without host repacking."""
    pts = synth.scene_primitives(30_000, seed=12)
    rec = np.zeros((30_000, 8), np.float32)
    rec[:, :3] = pts[:, :3]
    rec[:, 4] = pts[:, 3]
    ctx.set_cloud(rec, xyz_offset=0, intensity_offset=16)
    idx, _ = ctx.knn(10)
    assert np.array_equal(idx, orc.knn_self(pts, 10)[0])
    off, adj = ctx.knn_graph(10)
    seeds = np.array([5, 900, 20000], np.int32)
    cost, label = ctx.ift(None, seeds)                         # weights = the record's intensity
    rcost, rlabel = orc.ift(off, adj, pts[:, 3], seeds, orc.POLICY_C)
    assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel)


@pytest.mark.parametrize("name,k", [("primitives200k", 16), ("facade300k", 32)])
def test_graph_adjacency_bit_exact(ctx, name, k):
    pts = clouds()[name]
    ctx.set_cloud(pts)
    off, adj = ctx.knn_graph(k)
    ridx, _ = orc.knn_self(pts, k)
    roff, radj = orc.knn_graph(ridx)
    assert np.array_equal(off, roff)
    assert np.array_equal(adj, radj)


@pytest.mark.parametrize("name,k", [("primitives200k", 16), ("statue100k", 30), ("facade300k", 32)])
def test_normals_and_curvature(ctx, name, k):
    pts = clouds()[name]
    ctx.set_cloud(pts)
    out = ctx.normals(k)
    ridx, _ = orc.knn_self(pts, k)
    ref = orc.normals(pts, ridx, trig_mode=0)
    # north_star bar: curvature within 1e-5 relative
    assert np.allclose(out[:, 3], ref[:, 3], rtol=CURV_RTOL, atol=0.0)
    assert np.allclose(out[:, :3], ref[:, :3], rtol=CURV_RTOL, atol=1e-7)
    # what actually holds: bit-exact
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    # informational: distance to the libm flavour of the oracle
    lib = orc.normals(pts, ridx, trig_mode=1)
    assert np.max(np.abs(out[:, 3] - lib[:, 3])) < 5e-6


@pytest.mark.parametrize("name,k,theta,mn,mx", [
    ("primitives200k", 16, 3.0 / 180 * np.pi, 50, 10000),
    ("primitives200k", 16, 0.12, 30, 100000),
    ("facade300k", 32, 0.08, 50, 10000),
    ("statue100k", 30, 0.25, 10, 5000),
])
def test_region_growing_seeds_bit_exact(ctx, name, k, theta, mn, mx):
    from iftwt_rg_b200 import api
    pts = clouds()[name]
    ctx.set_cloud(pts)
    ctx.normals(k, download=False)
    lab, seeds = ctx.region_seeds(api.RgParams(mn, mx, k, theta, 1.0))
    ridx, _ = orc.knn_self(pts, k)
    rn = orc.normals(pts, ridx)
    rlab, _, kept, _ = orc.region_growing(rn, ridx, theta, 1.0, mn, mx)
    rseeds, _ = orc.rg_seeds(pts, rlab, kept)
    assert kept > 0
    assert np.array_equal(lab, rlab)
    assert np.array_equal(seeds, rseeds)


def test_region_growing_own_neighbourhood_size(ctx):
    """normals with k=30, region growing with k=50 (the reference's numbers, main.cpp:108/115)."""
    from iftwt_rg_b200 import api
    pts = clouds()["statue100k"]
    ctx.set_cloud(pts)
    ctx.normals(30, download=False)
    lab, seeds = ctx.region_seeds(api.RgParams(50, 10000, 50, 0.2, 1.0))
    i30, _ = orc.knn_self(pts, 30)
    i50, _ = orc.knn_self(pts, 50)
    rn = orc.normals(pts, i30)
    rlab, _, kept, _ = orc.region_growing(rn, i50, 0.2, 1.0, 50, 10000)
    rseeds, _ = orc.rg_seeds(pts, rlab, kept)
    assert np.array_equal(lab, rlab) and np.array_equal(seeds, rseeds)


def test_curvature_gate_that_could_fire_is_refused(ctx):
    from iftwt_rg_b200 import api
    pts = clouds()["statue100k"]
    ctx.set_cloud(pts)
    ctx.normals(16, download=False)
    with pytest.raises(api.IftwtError) as e:
        ctx.region_seeds(api.RgParams(50, 10000, 16, 0.1, 0.05))
    assert e.value.code == -5


@pytest.mark.parametrize("name,k,nseeds,wmax", [("primitives200k", 16, 40, 256), ("facade300k", 32, 300, 256),
                                                ("statue100k", 8, 5, 6), ("primitives200k", 16, 2000, 3)])
def test_ift_costs_and_labels_bit_exact_on_supplied_seeds_and_weights(ctx, name, k, nseeds, wmax):
    pts = clouds()[name]
    rng = np.random.default_rng(17)
    ctx.set_cloud(pts)
    off, adj = ctx.knn_graph(k)
    n = len(pts)
    seeds = rng.integers(0, n, nseeds).astype(np.int32)
    for w in (pts[:, 3].copy(), rng.integers(0, wmax, n).astype(np.float32)):
        if wmax <= 6:
            w = np.minimum(w, wmax)
        w[rng.integers(0, n, 20)] = 700.0                      # >= MAX_COST: never conquered
        cost, label = ctx.ift(w, seeds)
        rcost, rlabel = orc.ift(off, adj, w, seeds, orc.POLICY_C)
        assert np.array_equal(cost, rcost)
        assert np.array_equal(label, rlabel)
        # against the first-offer policy: costs identical, label differences are ties only
        fcost, flabel = orc.ift(off, adj, w, seeds, orc.POLICY_R)
        assert np.array_equal(cost, fcost)
        assert orc.ift_audit(off, adj, w, seeds, cost, label) == 0
        diff = int(np.count_nonzero(label != flabel))
        print(f"{name} k={k} seeds={nseeds}: tie-only label disagreements vs first-offer: {diff}/{n}")


def test_ift_rejects_non_integer_weights(ctx):
    from iftwt_rg_b200 import api
    pts = clouds()["statue100k"]
    ctx.set_cloud(pts)
    ctx.knn_graph(8, download=False)
    w = pts[:, 3].copy()
    w[7] = 12.7
    with pytest.raises(api.IftwtError) as e:
        ctx.ift(w, np.array([1, 2], np.int32))
    assert e.value.code == -7


@pytest.mark.parametrize("name,k,theta", [("primitives200k", 16, 0.1), ("facade300k", 32, 0.08)])
def test_segment_end_to_end_matches_oracle_pipeline(ctx, name, k, theta):
    from iftwt_rg_b200 import api
    pts = clouds()[name]
    ctx.set_cloud(pts)
    rg = api.RgParams(50, 10000, k, theta, 1.0)
    cost, label, ns = ctx.segment(k, k, rg)
    ridx, _ = orc.knn_self(pts, k)
    roff, radj = orc.knn_graph(ridx)
    rn = orc.normals(pts, ridx)
    rlab, _, kept, _ = orc.region_growing(rn, ridx, theta, 1.0, 50, 10000)
    rseeds, _ = orc.rg_seeds(pts, rlab, kept)
    rcost, rlabel = orc.ift(roff, radj, pts[:, 3], rseeds, orc.POLICY_C)
    assert ns == kept
    assert np.array_equal(cost, rcost)
    assert np.array_equal(label, rlabel)


def test_full_size_properties_1m(ctx):
    """cfg2 at full size (1 M, k=16): properties that need no CPU oracle run."""
    from iftwt_rg_b200 import api
    pts = synth.scene_primitives(1_000_000)
    ctx.set_cloud(pts)
    idx, d2 = ctx.knn(16)
    assert np.all(idx[:, 0] == np.arange(len(pts))) and np.all(d2[:, 0] == 0)
    assert np.all(np.diff(d2, axis=1) >= 0)
    # d2 recomputed on the host in the same FP32 order
    j = idx[:, 7]
    dx = pts[:, 0] - pts[j, 0]; dy = pts[:, 1] - pts[j, 1]; dz = pts[:, 2] - pts[j, 2]
    assert np.array_equal(((dx * dx + dy * dy) + dz * dz).astype(np.float32), d2[:, 7])
    off, adj = ctx.knn_graph(16)
    # symmetric, sorted, self-free
    src = np.repeat(np.arange(len(pts), dtype=np.uint32), np.diff(off).astype(np.int64))
    assert not np.any(src == adj)
    fwd = src.astype(np.uint64) << np.uint64(32) | adj.astype(np.uint64)
    rev = adj.astype(np.uint64) << np.uint64(32) | src.astype(np.uint64)
    assert np.all(np.diff(fwd.astype(np.int64)) > 0)
    assert np.array_equal(np.sort(rev), fwd)
    cost, label, ns = ctx.segment(16, 16, api.RgParams(50, 10000, 16, 0.1, 1.0))
    assert ns > 0
    seeds = np.unique(label[cost == 0])
    assert orc.ift_audit(off, adj, pts[:, 3], seeds.astype(np.int32), cost, label) == 0


def test_segment_with_gradient_weights(ctx):
    """iftwt_segment(weights = 1): the IFT runs on the morphological gradient of the intensity
    over the graph (graph.hpp:147-163) instead of the raw vertex values."""
    from iftwt_rg_b200 import api
    pts = clouds()["primitives200k"]
    k, theta = 16, 0.1
    ctx.set_cloud(pts)
    rg = api.RgParams(50, 10000, k, theta, 1.0)
    cost, label, ns = ctx.segment(k, k, rg, gradient_weights=True)
    ridx, _ = orc.knn_self(pts, k)
    roff, radj = orc.knn_graph(ridx)
    rn = orc.normals(pts, ridx)
    rlab, _, kept, _ = orc.region_growing(rn, ridx, theta, 1.0, 50, 10000)
    rseeds, _ = orc.rg_seeds(pts, rlab, kept)
    grad = orc.morph(roff, radj, pts[:, 3], orc.MORPH_GRADIENT)
    rcost, rlabel = orc.ift(roff, radj, grad, rseeds, orc.POLICY_C)
    assert ns == kept and np.array_equal(cost, rcost) and np.array_equal(label, rlabel)


@pytest.mark.parametrize("name", ["knn8_3k", "knn16_6k"])
def test_ift_against_the_reference_own_code(ctx, name):
    """GPU vs the REFERENCE's IFT_PCD (ift.hpp compiled unmodified, outputs committed as
    tests/golden/ref_ift_*.npz by tests/golden/make_ref_ift_golden.py): the graph the GPU builds is the
    fixture's, costs are bit-exact, labels are an admissible tie-break; disagreements are counted."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"ref_ift_{name}.npz"))
    pts, k = z["pts"], int(z["k"])
    ctx.set_cloud(pts)
    off, adj = ctx.knn_graph(k)
    assert np.array_equal(off, z["off"]) and np.array_equal(adj, z["adj"])
    q = np.zeros((len(z["seeds_xyz"]), 4), np.float32)
    q[:, :3] = z["seeds_xyz"]
    sidx, _ = ctx.knn_query(q, 1)                     # the seed snap of ift.hpp:135
    seeds = sidx[:, 0].astype(np.int32)
    cost, label = ctx.ift(None, seeds)
    assert np.array_equal(cost, z["ref_cost"])
    assert orc.ift_audit(off, adj, pts[:, 3], seeds, cost, label) == 0
    seed_set = np.zeros(len(pts), bool)
    seed_set[seeds] = True
    assert np.array_equal(label[seed_set], np.flatnonzero(seed_set).astype(np.uint32))
    assert np.array_equal(label[seed_set], z["ref_root"][seed_set])
    diff = int(np.count_nonzero(label != z["ref_root"]))
    print(f"{name}: tie-only label disagreements GPU (policy C) vs the reference's own run: {diff}/{len(pts)}")
