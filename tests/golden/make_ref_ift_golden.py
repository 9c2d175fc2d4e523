#!/usr/bin/env python
"""Writes tests/golden/ref_ift_*.npz: inputs and outputs of the REFERENCE's own IFT
(IFT_PCD of /root/reference/ift.hpp, compiled unmodified against oracle/shim — see
oracle/ref_ift_driver.cpp) on small graphs.  Run in the development container, where
/root/reference exists; the fixtures travel to the GPU box, the reference does not.

    python tests/golden/make_ref_ift_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from iftwt_rg_b200 import synth  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402
from oracle import pyref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def csr_from_edges(n, edges):
    nb = [set() for _ in range(n)]
    for a, b in edges:
        nb[a].add(b)
        nb[b].add(a)
    off = np.zeros(n + 1, np.uint32)
    adj = []
    for i in range(n):
        adj += sorted(nb[i])
        off[i + 1] = len(adj)
    return off, np.array(adj, np.uint32)


def cases():
    # (a), (b): symmetric k-NN graphs of synthetic scenes — the graphs the GPU path builds itself
    for name, n, k, ns, seed in (("knn8_3k", 3000, 8, 12, 7), ("knn16_6k", 6000, 16, 25, 11)):
        pts = synth.scene_primitives(n, seed=seed)
        idx, _ = orc.knn_self(pts, k)
        off, adj = orc.knn_graph(idx)
        rng = np.random.default_rng(seed)
        sv = np.sort(rng.choice(n, ns, replace=False))
        # seed positions a little off their vertex: IFT_PCD snaps them with a 1-NN (ift.hpp:135);
        # two seeds share a vertex (duplicates collapse)
        sxyz = pts[sv, :3] + rng.normal(0, 1e-4, (ns, 3)).astype(np.float32)
        sxyz = np.vstack([sxyz, pts[sv[:1], :3]])
        yield name, dict(pts=pts, off=off, adj=adj, seeds_xyz=sxyz.astype(np.float32), k=np.int32(k))
    # (c) 40 x 40 grid, 4-connected, six intensity values: plateaus and ties everywhere
    g = 40
    rng = np.random.default_rng(3)
    pts = np.zeros((g * g, 4), np.float32)
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    pts[:, 0], pts[:, 1] = ii.ravel(), jj.ravel()
    pts[:, 3] = rng.integers(0, 6, g * g)
    edges = [(i * g + j, i * g + j + 1) for i in range(g) for j in range(g - 1)] + \
            [(i * g + j, (i + 1) * g + j) for i in range(g - 1) for j in range(g)]
    off, adj = csr_from_edges(g * g, edges)
    sv = np.array([0, g * g - 1, g * (g // 2) + g // 2, 17, 17 * g])
    yield "grid40_ties", dict(pts=pts, off=off, adj=adj, seeds_xyz=pts[sv, :3].copy(), k=np.int32(0))
    # (d) two components, seeds in one only; a wall of closed vertices (intensity >= 500); a seed on a
    # vertex whose own intensity is high (its cost is 0 all the same, ift.hpp:138)
    n = 60
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = np.arange(n)
    pts[:, 3] = (np.arange(n) * 37) % 23
    pts[25, 3] = 700.0
    pts[5, 3] = 450.0
    edges = [(i, i + 1) for i in range(39)] + [(i, i + 1) for i in range(40, 59)] + [(3, 9), (12, 30)]
    off, adj = csr_from_edges(n, edges)
    yield "path_components", dict(pts=pts, off=off, adj=adj, seeds_xyz=pts[[5, 20], :3].copy(), k=np.int32(0))
    # (e) Petersen graph (the shape of boost_helper/exercices_toolbox.h:135-185)
    pts = np.zeros((10, 4), np.float32)
    pts[:, 0] = np.arange(10)
    pts[:, 3] = [3, 1, 4, 1, 5, 9, 2, 6, 5, 3]
    edges = [(i, (i + 1) % 5) for i in range(5)] + [(i, i + 5) for i in range(5)] + \
            [(5 + i, 5 + (i + 2) % 5) for i in range(5)]
    off, adj = csr_from_edges(10, edges)
    yield "petersen", dict(pts=pts, off=off, adj=adj, seeds_xyz=pts[[0, 7], :3].copy(), k=np.int32(0))


def main():
    assert pyref.build() and pyref.available(), "needs /root/reference (development container)"
    for name, c in cases():
        cost, root, a, b, w = pyref.ift(c["pts"], c["off"], c["adj"], c["seeds_xyz"])
        path = os.path.join(OUT, f"ref_ift_{name}.npz")
        np.savez_compressed(path, ref_cost=cost, ref_root=root, ref_mst_a=a, ref_mst_b=b, ref_mst_w=w, **c)
        print(f"{path}: n={len(cost)} seeds={len(c['seeds_xyz'])} reached={(cost < 500).mean():.3f} "
              f"mst={len(a)} {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
