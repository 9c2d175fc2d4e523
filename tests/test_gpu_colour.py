"""GPU <-> oracle parity for the colour path: what the reference's live
main() runs — GraCo gc(cloud_t, 5.0); gc.morph_erode(cloud_g) (main.cpp:87-89) — and the grey /
Otsu preparation (main.cpp:16-76) that feeds the IFT-WT block."""
import os

import numpy as np
import pytest

from iftwt_rg_b200 import synth
from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    from iftwt_rg_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def coloured(pts, seed):
    """pcl::PointXYZRGB-style cloud: the 4th float carries the packed 0xFFRRGGBB word."""
    rng = np.random.default_rng(seed)
    n = len(pts)
    base = np.clip(pts[:, 3].astype(np.int64), 0, 255)
    r = np.clip(base + rng.integers(-20, 20, n), 0, 255)
    g = np.clip(255 - base + rng.integers(-20, 20, n), 0, 255)
    b = rng.integers(0, 4, n) * 64                              # few distinct values: many equal r+g+b sums
    word = (0xFF << 24) | (r << 16) | (g << 8) | b
    out = pts.copy()
    out[:, 3] = word.astype(np.uint32).view(np.float32)
    return out


def _ref_ply():
    xyz = np.load(os.path.join(GOLD, "ref_test_ply_xyz.npy"))
    pts = np.zeros((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    pts[:, 3] = np.floor(127.5 + 127.5 * np.sin(40.0 * xyz[:, 0]))
    return pts


@pytest.mark.parametrize("name", ["ref_test_ply", "primitives"])
def test_graco_colour_morphology_bit_exact(ctx, name):
    pts = coloured(_ref_ply() if name == "ref_test_ply" else synth.scene_primitives(150_000), 3)
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(5.0 if name == "ref_test_ply" else 1.5)        # main.cpp:87: overlap 5.0
    ctx.voxel_graph(res)
    cen, keys, roff, radj, box = orc.voxel_graph(pts, res)
    got = ctx.get_cloud()
    rgb = cen[:, 3].copy().view(np.uint32)
    assert np.array_equal(got[:, 3].copy().view(np.uint32), rgb)              # colours snapped untouched, alpha kept
    for op in (orc.MORPH_DILATE, orc.MORPH_ERODE):
        assert np.array_equal(ctx.morph_rgb(op), orc.morph_rgb(roff, radj, rgb, op)), op
    from iftwt_rg_b200 import api
    with pytest.raises(api.IftwtError):
        ctx.morph_rgb(orc.MORPH_GRADIENT)                                      # GraCo has erode and dilate only


@pytest.mark.parametrize("otsu", [False, True])
def test_grey_and_otsu_prep_then_ift(ctx, otsu):
    pts = coloured(synth.scene_primitives(120_000, seed=11), 4)
    ctx.set_cloud(pts)
    thr = ctx.rgb_to_gray(otsu)
    gray, rthr = orc.rgb_to_gray(pts[:, 3].copy().view(np.uint32), otsu)
    assert thr == rthr
    # the cloud is an XYZI cloud from here on: the graph stages read the grey values
    k = 12
    off, adj = ctx.knn_graph(k)
    ridx, _ = orc.knn_self(pts, k)
    roff, radj = orc.knn_graph(ridx)
    assert np.array_equal(off, roff) and np.array_equal(adj, radj)
    assert np.array_equal(ctx.morph(orc.MORPH_ERODE), orc.morph(roff, radj, gray, orc.MORPH_ERODE))
    seeds = np.arange(0, len(pts), 4001, dtype=np.int32)
    cost, label = ctx.ift(None, seeds)
    rcost, rlabel = orc.ift(roff, radj, gray, seeds, orc.POLICY_C)
    assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel)


@pytest.mark.parametrize("name", ["graco_ref_test_ply_overlap5", "graco_primitives20k"])
def test_against_the_reference_own_graco(ctx, name):
    """GPU vs the REFERENCE's GraCo (graco.hpp compiled unmodified against oracle/shim; outputs in
    tests/golden/ref_graph.json, tests/golden/make_ref_graph_golden.py): resolution, voxel centres, snapped
    colours and colour erode / dilate, bit-exact."""
    import hashlib
    import json

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    g = json.load(open(os.path.join(GOLD, "ref_graph.json")))[name]
    if name == "graco_ref_test_ply_overlap5":
        pts = coloured(_ref_ply(), 3)
    else:
        pts = coloured(synth.scene_primitives(20_000, seed=0x1F88), 4)
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(g["overlap"])
    assert res == g["resolution"]
    assert ctx.voxel_graph(res) == g["n_voxels"]
    got = ctx.get_cloud()
    assert sha(got[:, :3]) == g["centres_xyz_sha256"]
    assert sha(got[:, 3].copy().view(np.uint32) & 0xFFFFFF) == g["rgb_sha256"]
    for op in (orc.MORPH_DILATE, orc.MORPH_ERODE):
        assert sha(ctx.morph_rgb(op) & 0xFFFFFF) == g["morph_rgb_sha256"][str(op)], op
