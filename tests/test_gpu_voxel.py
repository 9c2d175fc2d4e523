"""GPU <-> oracle parity for graph builder (A) — the reference's own voxel-26 graph —
graph morphology, regional-minima seeds and the reference-shaped pipeline of
main.cpp:97-123 run on the graph's vertices."""
import json
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


def ref_cloud():
    xyz = np.load(os.path.join(GOLD, "ref_test_ply_xyz.npy"))
    pts = np.zeros((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    # test.ply carries intensity 0 everywhere: give it a deterministic integer texture
    pts[:, 3] = np.floor(127.5 + 127.5 * np.sin(40.0 * xyz[:, 0]) * np.cos(31.0 * xyz[:, 1]))
    return pts


CLOUDS = {
    "ref_test_ply": (ref_cloud, 1.01),
    "ref_test_ply_overlap5": (ref_cloud, 5.0),            # what the live main() passes (main.cpp:87)
    "statue": (lambda: synth.scene_statue(120_000), 2.0),
    "primitives": (lambda: synth.scene_primitives(150_000), 1.5),
}


@pytest.mark.parametrize("name", sorted(CLOUDS))
def test_resolution_and_voxel_graph_bit_exact(ctx, name):
    gen, overlap = CLOUDS[name]
    pts = gen()
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(overlap)
    assert res == orc.cloud_resolution(pts, overlap, 0)
    V = ctx.voxel_graph(res)
    cen, keys, roff, radj, box = orc.voxel_graph(pts, res)
    assert V == len(cen)
    got = ctx.get_cloud()
    assert np.array_equal(got.view(np.uint32), cen.view(np.uint32))     # centres + snapped intensity
    off, adj = ctx.graph_csr()
    assert np.array_equal(off, roff) and np.array_equal(adj, radj)
    if name == "ref_test_ply":
        gold = json.load(open(os.path.join(GOLD, "ref_test_ply_voxel.json")))
        assert V == gold["n_voxels"] and len(adj) == gold["n_arcs"] and res == gold["resolution"]


@pytest.mark.parametrize("name", ["ref_test_ply_overlap5", "primitives"])
def test_morphology_regmin_and_ift_on_the_voxel_graph(ctx, name):
    gen, overlap = CLOUDS[name]
    pts = gen()
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(overlap)
    ctx.voxel_graph(res)
    cen, keys, roff, radj, box = orc.voxel_graph(pts, res)
    w = cen[:, 3]
    for op in (orc.MORPH_ERODE, orc.MORPH_DILATE, orc.MORPH_GRADIENT, orc.MORPH_BIN_ERODE, orc.MORPH_BIN_DILATE):
        assert np.array_equal(ctx.morph(op), orc.morph(roff, radj, w, op)), op
    # IFT_PCD(Graph&): regional-minima seeds (ift.hpp:38-44) on the morphological gradient
    grad = ctx.morph(orc.MORPH_GRADIENT)
    ctx.set_intensity(grad)                                  # Graph::pc_to_g write-back
    seeds = ctx.regmin_seeds()
    rseeds = orc.regmin(roff, radj, grad)
    assert np.array_equal(seeds, rseeds) and len(seeds) > 0
    cost, label = ctx.ift(None, None)                        # weights = vertex values, seeds = regmin
    rcost, rlabel = orc.ift(roff, radj, grad, rseeds, orc.POLICY_C)
    assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel)
    assert orc.ift_audit(roff, radj, grad, rseeds, cost, label) == 0


def test_reference_shaped_pipeline_on_the_voxel_graph(ctx):
    """main.cpp:97-123 with the reference's neighbourhood sizes: voxel graph, normals k=30,
    region growing k=50 on the graph's vertices, IFT seeded by the snapped centroids."""
    from iftwt_rg_b200 import api
    pts = synth.scene_primitives(200_000, seed=77)
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(2.5)
    V = ctx.voxel_graph(res)
    cen, keys, roff, radj, box = orc.voxel_graph(pts, res)
    nrm = ctx.normals(30)
    i30, _ = orc.knn_self(cen, 30)
    rn = orc.normals(cen, i30)
    assert np.array_equal(nrm.view(np.uint32), rn.view(np.uint32))
    theta = 0.05
    lab, seeds = ctx.region_seeds(api.RgParams(20, 100000, 50, theta, 1.0))
    i50, _ = orc.knn_self(cen, 50)
    rlab, _, kept, _ = orc.region_growing(rn, i50, theta, 1.0, 20, 100000)
    rseeds, _ = orc.rg_seeds(cen, rlab, kept)
    assert kept > 0 and np.array_equal(lab, rlab) and np.array_equal(seeds, rseeds)
    cost, label = ctx.ift(None, None)
    rcost, rlabel = orc.ift(roff, radj, cen[:, 3], rseeds, orc.POLICY_C)
    assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel)


def _sha(a) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _binary_cloud():
    pts = synth.scene_primitives(20_000, seed=0x1F77)
    pts[:, 3] = np.where(np.sin(3.0 * pts[:, 0]) * np.cos(2.0 * pts[:, 1]) > 0, 255.0, 0.0)
    return pts


REF_GRAPH_CLOUDS = {
    "ref_test_ply": ref_cloud,
    "ref_test_ply_overlap5": ref_cloud,
    "primitives20k": lambda: synth.scene_primitives(20_000, seed=0x1F66),
    "binary20k": _binary_cloud,
}


@pytest.mark.parametrize("name", sorted(REF_GRAPH_CLOUDS))
def test_against_the_reference_own_graph_and_ift(ctx, name):
    """GPU vs the REFERENCE's graph.hpp + ift.hpp (compiled unmodified against oracle/shim; outputs
    committed as tests/golden/ref_graph.json by tests/golden/make_ref_graph_golden.py): resolution, voxel
    centres and snapped intensities, adjacency, morphology, regional minima and IFT costs are
    bit-exact; IFT labels are an admissible tie-break of the reference's rule."""
    g = json.load(open(os.path.join(GOLD, "ref_graph.json")))[name]
    pts = REF_GRAPH_CLOUDS[name]()
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(g["overlap"])
    assert res == g["resolution"]
    assert ctx.voxel_graph(res) == g["n_voxels"]
    cen = ctx.get_cloud()
    assert _sha(cen) == g["centres_sha256"]
    off, adj = ctx.graph_csr()
    assert _sha(off) == g["off_sha256"] and _sha(adj) == g["adj_sha256"]
    for op in (orc.MORPH_BIN_ERODE, orc.MORPH_DILATE, orc.MORPH_ERODE, orc.MORPH_GRADIENT):
        assert _sha(ctx.morph(op)) == g["morph_sha256"][str(op)], op
    rm = ctx.regmin_seeds()
    assert len(rm) == g["regmin_count"] and _sha(np.asarray(rm, np.int32)) == g["regmin_sha256"]
    seeds = np.array(g["ift"]["seed_vertices"], np.int32)
    cost, label = ctx.ift(None, seeds)
    assert _sha(cost) == g["ift"]["cost_sha256"]
    assert int(np.count_nonzero(cost < 500)) == g["ift"]["reached"]
    assert orc.ift_audit(off, adj, cen[:, 3], seeds, cost, label) == 0
    grad = ctx.morph(orc.MORPH_GRADIENT)
    ctx.set_intensity(grad)                                  # Graph::pc_to_g write-back
    assert _sha(ctx.get_cloud()[:, 3]) == g["values_after_pc_to_g_sha256"]
    assert _sha(ctx.morph(orc.MORPH_ERODE)) == g["erode_gradient_sha256"]
