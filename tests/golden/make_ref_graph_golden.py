#!/usr/bin/env python
"""Writes tests/golden/ref_graph.json: what the REFERENCE's own Graph (graph.hpp) and IFT_PCD
(ift.hpp) — compiled unmodified against oracle/shim, driver oracle/ref_graph_driver.cpp — give
on the reference's test.ply (with the deterministic intensity texture the tests use) and on a
synthetic scene.  Large arrays are stored as SHA-256 of their bytes.  Run in the development
container, where /root/reference exists:

    python tests/golden/make_ref_graph_golden.py
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from iftwt_rg_b200 import synth  # noqa: E402
from oracle import pyref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_cloud():
    xyz = np.load(os.path.join(GOLD, "ref_test_ply_xyz.npy"))
    pts = np.zeros((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    pts[:, 3] = np.floor(127.5 + 127.5 * np.sin(40.0 * xyz[:, 0]) * np.cos(31.0 * xyz[:, 1]))
    return pts


def binary_cloud():
    """intensities in {0, 255}: what the live main() feeds the binary morphology (main.cpp:16-50)"""
    pts = synth.scene_primitives(20_000, seed=0x1F77)
    pts[:, 3] = np.where(np.sin(3.0 * pts[:, 0]) * np.cos(2.0 * pts[:, 1]) > 0, 255.0, 0.0)
    return pts


CASES = {
    "ref_test_ply": (ref_cloud, 1.01),
    "ref_test_ply_overlap5": (ref_cloud, 5.0),
    "primitives20k": (lambda: synth.scene_primitives(20_000, seed=0x1F66), 1.5),
    "binary20k": (binary_cloud, 1.5),
}
N_SEEDS = 24


def coloured(pts, seed):
    """pcl::PointXYZRGB-style cloud: the 4th float carries the packed 0xFFRRGGBB word (tests/test_gpu_colour.py)"""
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
    xyz = np.load(os.path.join(GOLD, "ref_test_ply_xyz.npy"))
    pts = np.zeros((len(xyz), 4), np.float32)
    pts[:, :3] = xyz
    pts[:, 3] = np.floor(127.5 + 127.5 * np.sin(40.0 * xyz[:, 0]))
    return coloured(pts, 3)


GRACO_CASES = {
    "graco_ref_test_ply_overlap5": (ref_cloud_colour, 5.0),          # main.cpp:87: GraCo gc(cloud_t, 5.0)
    "graco_primitives20k": (lambda: coloured(synth.scene_primitives(20_000, seed=0x1F88), 4), 1.5),
}


def seed_vertices(V):
    return (np.arange(N_SEEDS, dtype=np.int64) * 7919 + 13) % V


def main():
    assert pyref.build() and pyref.graph_available(), "needs /root/reference (development container)"
    out = {"_comment": "outputs of the reference's own graph.hpp / ift.hpp (tests/golden/make_ref_graph_golden.py); "
                       "vertex order = ascending voxel Morton key; op ids = enum MORPH of graph.hpp:13, 5 = gradient; "
                       "graco_* = GraCo of graco.hpp, colours as 0x00RRGGBB"}
    for name, (gen, overlap) in CASES.items():
        pts = gen()
        t0 = time.time()
        with pyref.RefGraph(pts, overlap) as g:
            cen = g.cloud()
            off, adj, loops = g.csr()
            e = {"n_points": int(len(pts)), "overlap": overlap, "resolution": g.resolution(), "n_voxels": g.n,
                 "n_arcs": int(len(adj)), "self_loops": loops, "centres_sha256": sha(cen), "off_sha256": sha(off),
                 "adj_sha256": sha(adj)}
            e["morph_sha256"] = {str(op): sha(g.morph(op)) for op in (1, 2, 3, 4, 5)}
            e["bin_dilate_all_zero"] = bool(np.all(g.morph(2) == 0))     # graph.hpp:345-349 discards its result
            rm = g.regmin()
            e["regmin_count"] = int(len(rm))
            e["regmin_sha256"] = sha(rm.astype(np.int32))
            sv = seed_vertices(g.n)
            cost, root, a, b, w = g.ift(cen[sv, :3])
            e["ift"] = {"seed_vertices": [int(x) for x in sv], "cost_sha256": sha(cost), "root_sha256": sha(root),
                        "reached": int(np.count_nonzero(cost < 500)), "cost_sum": int(cost.astype(np.int64).sum()),
                        "mst": [[int(x), int(y), float(z)] for x, y, z in zip(a, b, w)]}
            # morph_erode_gradient (graph.hpp:203-208): gradient written back through pc_to_g, then eroded
            eg = g.morph(6)
            e["erode_gradient_sha256"] = sha(eg)
            e["values_after_pc_to_g_sha256"] = sha(g.cloud()[:, 3])
        out[name] = e
        print(f"{name}: V={e['n_voxels']} arcs={e['n_arcs']} res={e['resolution']:.9g} regmin={e['regmin_count']} "
              f"reached={e['ift']['reached']} ({time.time() - t0:.1f} s)")
    for name, (gen, overlap) in GRACO_CASES.items():
        pts = gen()
        with pyref.RefGraCo(pts, overlap) as g:
            xyz, rgb = g.cloud()
            e = {"n_points": int(len(pts)), "overlap": overlap, "resolution": g.resolution(), "n_voxels": g.n,
                 "centres_xyz_sha256": sha(xyz), "rgb_sha256": sha(rgb & 0xFFFFFF),
                 "morph_rgb_sha256": {str(op): sha(g.morph(op) & 0xFFFFFF) for op in (3, 4)}}
        out[name] = e
        print(f"{name}: V={e['n_voxels']} res={e['resolution']:.9g}")
    json.dump(out, open(os.path.join(GOLD, "ref_graph.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
