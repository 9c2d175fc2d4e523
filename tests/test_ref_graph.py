"""The oracle against the REFERENCE's own Graph (graph.hpp) and IFT_PCD (ift.hpp) on it.

oracle/_ref/libref_graph.so is graph.hpp + ift.hpp + globals.hpp of /root/reference compiled
UNMODIFIED (oracle/ref_graph_driver.cpp) against the API stand-ins of oracle/shim.  Its outputs on
the reference's test.ply (two overlaps) and two synthetic scenes are committed as
tests/golden/ref_graph.json (tests/golden/make_ref_graph_golden.py; big arrays as SHA-256).  Pinned here:
computeCloudResolution, the voxel graph (centres, snapped intensities, adjacency), the four
morphology operators + gradient, pc_to_g write-back, getRegmin, and IFT_PCD on that graph."""
import hashlib
import json
import os

import numpy as np
import pytest

from iftwt_rg_b200 import synth
from oracle import pyoracle as orc
from oracle import pyref

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = json.load(open(os.path.join(GOLD_DIR, "ref_graph.json")))
NAMES = [k for k in GOLD if not k.startswith("_") and not k.startswith("graco_")]
GRACO = [k for k in GOLD if k.startswith("graco_")]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_cloud():
    xyz = np.load(os.path.join(GOLD_DIR, "ref_test_ply_xyz.npy"))
    pts = np.zeros((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    pts[:, 3] = np.floor(127.5 + 127.5 * np.sin(40.0 * xyz[:, 0]) * np.cos(31.0 * xyz[:, 1]))
    return pts


def binary_cloud():
    pts = synth.scene_primitives(20_000, seed=0x1F77)
    pts[:, 3] = np.where(np.sin(3.0 * pts[:, 0]) * np.cos(2.0 * pts[:, 1]) > 0, 255.0, 0.0)
    return pts


CLOUDS = {
    "ref_test_ply": ref_cloud,
    "ref_test_ply_overlap5": ref_cloud,
    "primitives20k": lambda: synth.scene_primitives(20_000, seed=0x1F66),
    "binary20k": binary_cloud,
}


@pytest.mark.parametrize("name", NAMES)
def test_oracle_graph_morphology_regmin_ift_equal_reference(name):
    g = GOLD[name]
    pts = CLOUDS[name]()
    assert len(pts) == g["n_points"]
    # computeCloudResolution: the literal sequential double sum AND the order-free exact sum (the
    # GPU contract) both reproduce the reference's double, bit for bit
    assert orc.cloud_resolution(pts, g["overlap"], 1) == g["resolution"]
    assert orc.cloud_resolution(pts, g["overlap"], 0) == g["resolution"]
    cen, keys, off, adj, box = orc.voxel_graph(pts, g["resolution"])
    assert len(cen) == g["n_voxels"] and len(adj) == g["n_arcs"]
    assert g["self_loops"] == g["n_voxels"]              # closed neighbourhoods
    assert sha(cen) == g["centres_sha256"]               # voxel centres + optimizeGraph intensities
    assert sha(off) == g["off_sha256"] and sha(adj) == g["adj_sha256"]
    w = cen[:, 3]
    for op in (orc.MORPH_BIN_ERODE, orc.MORPH_DILATE, orc.MORPH_ERODE, orc.MORPH_GRADIENT):
        assert sha(orc.morph(off, adj, w, op)) == g["morph_sha256"][str(op)], op
    # bin_dilate: the reference discards its result (shadowed local, graph.hpp:345-349) and writes 0
    # everywhere; the oracle and the GPU implement the evident intent
    assert g["bin_dilate_all_zero"]
    bd = orc.morph(off, adj, w, orc.MORPH_BIN_DILATE)
    assert (sha(bd) == g["morph_sha256"]["2"]) == bool(np.all(bd == 0))
    rm = orc.regmin(off, adj, w)
    assert len(rm) == g["regmin_count"] and sha(rm.astype(np.int32)) == g["regmin_sha256"]
    seeds = np.array(g["ift"]["seed_vertices"], np.int32)
    cost, label, a, b, mw = orc.ift_literal_mst(off, adj, w, seeds)
    assert sha(cost) == g["ift"]["cost_sha256"] and sha(label) == g["ift"]["root_sha256"]
    assert [[int(x), int(y), float(z)] for x, y, z in zip(a, b, mw)] == g["ift"]["mst"]
    ccost, clabel = orc.ift(off, adj, w, seeds, orc.POLICY_C)
    assert sha(ccost) == g["ift"]["cost_sha256"]
    assert int(np.count_nonzero(ccost < 500)) == g["ift"]["reached"]
    assert orc.ift_audit(off, adj, w, seeds, ccost, clabel) == 0
    grad = orc.morph(off, adj, w, orc.MORPH_GRADIENT)
    assert sha(grad) == g["values_after_pc_to_g_sha256"]
    assert sha(orc.morph(off, adj, grad, orc.MORPH_ERODE)) == g["erode_gradient_sha256"]


def test_reference_reproduces_survey_numbers_on_test_ply():
    g = GOLD["ref_test_ply"]
    old = json.load(open(os.path.join(GOLD_DIR, "ref_test_ply_voxel.json")))
    assert g["n_voxels"] == old["n_voxels"] == 71368 and g["n_arcs"] == old["n_arcs"]
    assert g["resolution"] == old["resolution"]


def coloured(pts, seed):
    rng = np.random.default_rng(seed)
    n = len(pts)
    base = np.clip(pts[:, 3].astype(np.int64), 0, 255)
    r = np.clip(base + rng.integers(-20, 20, n), 0, 255)
    g = np.clip(255 - base + rng.integers(-20, 20, n), 0, 255)
    b = rng.integers(0, 4, n) * 64
    word = (0xFF << 24) | (r << 16) | (g << 8) | b
    out = pts.copy()
    out[:, 3] = word.astype(np.uint32).view(np.float32)
    return out


def ref_cloud_colour():
    xyz = np.load(os.path.join(GOLD_DIR, "ref_test_ply_xyz.npy"))
    pts = np.zeros((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    pts[:, 3] = np.floor(127.5 + 127.5 * np.sin(40.0 * xyz[:, 0]))
    return coloured(pts, 3)


GRACO_CLOUDS = {
    "graco_ref_test_ply_overlap5": ref_cloud_colour,
    "graco_primitives20k": lambda: coloured(synth.scene_primitives(20_000, seed=0x1F88), 4),
}


@pytest.mark.parametrize("name", GRACO)
def test_oracle_colour_graph_and_morphology_equal_reference_graco(name):
    """GraCo gc(cloud_t, overlap); gc.morph_erode / morph_dilate — graco.hpp, what the live main() runs"""
    g = GOLD[name]
    pts = GRACO_CLOUDS[name]()
    assert orc.cloud_resolution(pts, g["overlap"], 0) == g["resolution"]
    cen, keys, off, adj, box = orc.voxel_graph(pts, g["resolution"])
    assert len(cen) == g["n_voxels"] and sha(cen[:, :3]) == g["centres_xyz_sha256"]
    rgb = cen[:, 3].copy().view(np.uint32)
    assert sha(rgb & 0xFFFFFF) == g["rgb_sha256"]              # optimizeGraph: colour of the nearest point
    for op in (orc.MORPH_DILATE, orc.MORPH_ERODE):
        assert sha(orc.morph_rgb(off, adj, rgb, op) & 0xFFFFFF) == g["morph_rgb_sha256"][str(op)], op


needs_ref = pytest.mark.skipif(
    not (pyref.graph_available() or os.path.exists(os.path.join(pyref.REFERENCE, "graph.hpp"))),
    reason="oracle/_ref/libref_graph.so is built where /root/reference exists")


@needs_ref
@pytest.mark.parametrize("name", ["primitives20k", "binary20k"])
def test_reference_live_equals_fixture(name):
    pyref.build()
    g = GOLD[name]
    with pyref.RefGraph(CLOUDS[name](), g["overlap"]) as r:
        assert r.resolution() == g["resolution"] and r.n == g["n_voxels"]
        assert sha(r.cloud()) == g["centres_sha256"]
        off, adj, loops = r.csr()
        assert sha(off) == g["off_sha256"] and sha(adj) == g["adj_sha256"] and loops == g["self_loops"]
        for op in (1, 2, 3, 4, 5):
            assert sha(r.morph(op)) == g["morph_sha256"][str(op)]
        assert sha(r.regmin().astype(np.int32)) == g["regmin_sha256"]
        cost, root, a, b, w = r.ift(r.cloud()[np.array(g["ift"]["seed_vertices"]), :3])
        assert sha(cost) == g["ift"]["cost_sha256"] and sha(root) == g["ift"]["root_sha256"]


@needs_ref
def test_reference_graco_live_equals_fixture():
    pyref.build()
    name = "graco_primitives20k"
    g = GOLD[name]
    with pyref.RefGraCo(GRACO_CLOUDS[name](), g["overlap"]) as r:
        assert r.resolution() == g["resolution"] and r.n == g["n_voxels"]
        xyz, rgb = r.cloud()
        assert sha(xyz) == g["centres_xyz_sha256"] and sha(rgb & 0xFFFFFF) == g["rgb_sha256"]
        for op in (3, 4):
            assert sha(r.morph(op) & 0xFFFFFF) == g["morph_rgb_sha256"][str(op)]


def test_shim_headers_unit(tmp_path):
    """oracle/shim: the container, k-NN and octree stand-ins behave as the reference's results assume
    (tests/cpp/shim_test.cpp)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "shim_test")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(root, "oracle", "shim"),
                           os.path.join(root, "tests", "cpp", "shim_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "shim ok" in out.stdout, out.stdout + out.stderr
