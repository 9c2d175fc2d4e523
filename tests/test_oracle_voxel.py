"""Oracle restatement of the reference's voxel-26 graph (graph.hpp:228-308) against the
committed golden fixture (the reference's own test.ply geometry) and known answers."""
import json
import os

import numpy as np

from oracle import pyoracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ply_cloud():
    xyz = np.load(os.path.join(GOLD, "ref_test_ply_xyz.npy"))
    pts = np.zeros((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    return pts


def test_voxel_graph_matches_golden_and_survey_probe():
    gold = json.load(open(os.path.join(GOLD, "ref_test_ply_voxel.json")))
    pts = _ply_cloud()
    assert len(pts) == gold["n_points"]
    res = orc.cloud_resolution(pts, gold["overlap"], 0)
    assert res == gold["resolution"]
    assert orc.cloud_resolution(pts, gold["overlap"], 1) == gold["resolution_literal"]
    cen, keys, off, adj, box = orc.voxel_graph(pts, res)
    deg = np.diff(off)
    assert int(box[3]) == gold["depth"] == gold["survey_app_d"]["depth"]
    assert len(cen) == gold["n_voxels"] == gold["survey_app_d"]["n_voxels"]
    assert int((deg == 0).sum()) == gold["isolated"] == gold["survey_app_d"]["isolated"]
    assert abs(deg.mean() - gold["survey_app_d"]["mean_degree"]) < 5e-3
    assert abs(res - gold["survey_app_d"]["resolution"]) < 1e-7


def test_voxel_graph_known_answer_lattice():
    # 3 x 3 x 1 block of unit-spaced points + one far point: every voxel holds one point
    g = np.array([[x, y, 0] for x in range(3) for y in range(3)] + [[10, 10, 10]], np.float32)
    pts = np.zeros((len(g), 4), np.float32)
    pts[:, :3] = g
    pts[:, 3] = np.arange(len(g))
    cen, keys, off, adj, box = orc.voxel_graph(pts, 1.0)
    assert len(cen) == 10
    deg = np.diff(off)
    # corner voxels see 3 neighbours, edge voxels 5, the centre 8, the far point none
    assert sorted(deg.tolist()) == [0, 3, 3, 3, 3, 5, 5, 5, 5, 8]
    # symmetric, sorted, self-free
    src = np.repeat(np.arange(10), deg)
    assert not np.any(src == adj)
    pairs = set(zip(src.tolist(), adj.tolist()))
    assert all((b, a) in pairs for a, b in pairs)
    # intensities are snapped from the nearest input point: a bijection here
    assert sorted(cen[:, 3].tolist()) == list(range(10))
    # ascending Morton order of the keys, x most significant
    codes = [int(''.join(f"{(int(k[0]) >> b) & 1}{(int(k[1]) >> b) & 1}{(int(k[2]) >> b) & 1}"
                         for b in range(20, -1, -1)), 2) for k in keys]
    assert codes == sorted(codes)


def test_morphology_and_regmin_known_answer():
    # path 0-1-2-3-4 with values 5 3 3 7 0
    off = np.array([0, 1, 3, 5, 7, 8], np.uint32)
    adj = np.array([1, 0, 2, 1, 3, 2, 4, 3], np.uint32)
    w = np.array([5, 3, 3, 7, 0], np.float32)
    assert orc.morph(off, adj, w, orc.MORPH_ERODE).tolist() == [3, 3, 3, 0, 0]
    assert orc.morph(off, adj, w, orc.MORPH_DILATE).tolist() == [5, 5, 7, 7, 7]
    assert orc.morph(off, adj, w, orc.MORPH_GRADIENT).tolist() == [2, 2, 4, 7, 7]
    assert orc.morph(off, adj, w, orc.MORPH_BIN_ERODE).tolist() == [255, 255, 255, 0, 0]
    # plateau 1-2 is a (non-strict) regional minimum: both points are seeds (graph.hpp:217-222)
    assert orc.regmin(off, adj, w).tolist() == [1, 2, 4]
