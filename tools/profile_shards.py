#!/usr/bin/env python
"""One sharded k-NN / normals pass with shards time-sharing device 0 — the target of
`ncu -k regex:shard_ ...` for profiles/ (every kernel of csrc/shard.cu runs in this mode)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iftwt_rg_b200 import api, sharded, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=8_000_000)
ap.add_argument("--world", type=int, default=4)
ap.add_argument("--k", type=int, default=16)
a = ap.parse_args()
pts = synth.scene_facade(a.points)
dev = torch.device("cuda", 0)
bounds = [sharded.slice_bounds(a.points, a.world, r) for r in range(a.world)]
slabs = [torch.from_numpy(np.ascontiguousarray(pts[lo:hi])).to(dev) for lo, hi in bounds]
torch.cuda.synchronize()
with api.Comm.all([0] * a.world) as cm:
    sl = [(t.data_ptr(), t.shape[0], lo) for t, (lo, hi) in zip(slabs, bounds)]
    o = cm.sharded_knn_normals(sl, a.k, True, 0.0)           # estimate + allocations
    o = cm.sharded_knn_normals(sl, a.k, True, 1.05 * o[0].kth_distance_max)
    print({kk: v for kk, v in o[0].as_dict().items()})
