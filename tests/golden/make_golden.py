"""Generates the golden fixtures.  Runs only in the development container (it reads the
reference tree, which does not exist on the GPU box); the outputs are committed.

  ref_test_ply_xyz.npy   geometry of /root/reference/cmake-build-debug/test.ply — the only
                         data file the reference ships (85,227 XYZI points written by PCL,
                         intensity identically 0).
  ref_test_ply_voxel.json  what the reference-style voxel graph must give on it at the
                         reference's default overlap 1.01: the numbers SURVEY App. D measured
                         with an independent throw-away script (res 1.1706e-3, depth 10,
                         V = 71,368, mean 26-degree 1.21, 19,052 isolated voxels) reproduced
                         by the oracle.
  ift_known_answers.json hand-derived IFT results on path / K2 / Petersen graphs.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as orc  # noqa: E402


def load_ply(path):
    with open(path) as f:
        n = 0
        for line in f:
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line.strip() == "end_header":
                break
        return np.loadtxt(f, max_rows=n, dtype=np.float32)


def main():
    a = load_ply("/root/reference/cmake-build-debug/test.ply")
    assert a.shape == (85227, 4) and np.all(a[:, 3] == 0)
    np.save(os.path.join(HERE, "ref_test_ply_xyz.npy"), a[:, :3].copy())
    pts = np.ascontiguousarray(a)
    res_lit = orc.cloud_resolution(pts, 1.01, 1)
    res = orc.cloud_resolution(pts, 1.01, 0)
    cen, keys, off, adj, box = orc.voxel_graph(pts, res)
    deg = np.diff(off)
    gold = {"n_points": int(len(pts)), "overlap": 1.01, "resolution": res, "resolution_literal": res_lit,
            "depth": int(box[3]), "n_voxels": int(len(cen)), "n_arcs": int(off[-1]),
            "mean_degree": float(deg.mean()), "isolated": int((deg == 0).sum()),
            "survey_app_d": {"resolution": 1.1706e-3, "depth": 10, "n_voxels": 71368, "mean_degree": 1.21,
                             "isolated": 19052}}
    json.dump(gold, open(os.path.join(HERE, "ref_test_ply_voxel.json"), "w"), indent=1)
    print(gold)


if __name__ == "__main__":
    main()
