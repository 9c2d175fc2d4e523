#!/usr/bin/env python
"""Every ABI path once on small clouds — meant to be run under compute-sanitizer
(`compute-sanitizer --tool memcheck python tools/sanitize_small.py`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iftwt_rg_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
with api.Context(0) as ctx:
    for gen, k in ((synth.scene_primitives, 16), (synth.scene_statue, 30), (synth.scene_facade, 50)):
        pts = gen(n)
        ctx.set_cloud(pts)
        ctx.knn(k)
        ctx.knn_query(pts[::97, :3] + 0.01, 3)
        ctx.knn_graph(k)
        ctx.normals(k)
        ctx.region_seeds(api.RgParams(10, 100000, k, 0.3, 1.0))
        ctx.ift(None, None)
        ctx.region_adjacency(); ctx.region_info(); ctx.cluster(5)
        ctx.segment(k, k, api.RgParams(10, 100000, k, 0.3, 1.0))
    pts = synth.scene_primitives(n)
    ctx.set_cloud(pts)
    res = ctx.cloud_resolution(2.0)
    ctx.voxel_graph(res)
    ctx.graph_csr(); ctx.get_cloud()
    g = ctx.morph(5); ctx.set_intensity(g)
    ctx.regmin_seeds(); ctx.ift(None, None); ctx.cluster(10)
    ctx.normals(30); ctx.region_seeds(api.RgParams(10, 100000, 50, 0.3, 1.0))
    lattice = np.zeros((4000, 4), np.float32)
    lattice[:, :3] = np.random.default_rng(0).integers(0, 8, (4000, 3)) * 0.25
    ctx.set_cloud(lattice); ctx.knn(8); ctx.knn(40)
    ctx.set_cloud(pts[:5].copy()); ctx.knn(8); ctx.normals(8)
print("sanitize_small ok")
