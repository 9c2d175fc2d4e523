#!/usr/bin/env python
"""BASELINE config 4 driver: Morton-range sharded k-NN + normals on all ranks (NCCL halo
exchange), rows gathered to rank 0, graph / seeds / IFT on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tools/sharded_run.py --points 20000000 --k 16 [--verify]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iftwt_rg_b200 import api, sharded, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=20_000_000)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--tiles", type=int, default=10)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--no-ift", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, k = args.points, args.k
    lo, hi = rank * n // world, (rank + 1) * n // world
    # generation order is random in space, so a contiguous id range is a uniform subsample
    pts_np = synth.facade_tiled(n, tiles=args.tiles, start=lo, count=hi - lo)
    radius = sharded.estimate_halo_radius(pts_np, k, 1)   # k-th neighbour distance inside the slice
    r = torch.tensor([radius], device=dev)
    if world > 1:
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
    # the slice is a 1/world subsample: its k-th neighbour is sqrt(world) farther than the cloud's
    radius = float(r) / (world ** 0.5)
    pts = torch.from_numpy(pts_np).to(dev)
    ctx = api.Context(local)
    hook = sharded.gpu_local_knn(ctx)

    def run():
        return sharded.sharded_knn(pts, lo, k, hook, radius, want_normals=True)

    run()  # warm-up (allocations, NCCL channels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_knn = time.perf_counter() - t0
    stats = res.stats
    gathered = sharded.gather_rows_to_rank0(res, n, k)
    out = {"config": "cfg4 (facade tiled)", "points": n, "k": k, "n_gpus": world, "sharded_knn_normals_s": t_knn,
           "points_per_s": n / t_knn, "halo_radius": stats["halo_radius"], "attempts": stats["attempts"]}
    owned = torch.tensor([stats["owned"], stats["halo"]], device=dev, dtype=torch.int64)
    allo = [torch.zeros_like(owned) for _ in range(world)]
    if world > 1:
        dist.all_gather(allo, owned)
    else:
        allo = [owned]
    out["owned_per_rank"] = [int(a[0]) for a in allo]
    out["halo_per_rank"] = [int(a[1]) for a in allo]
    if rank == 0:
        all_pts, all_idx, all_d2 = gathered
        if not args.no_ift:
            torch.cuda.synchronize()   # the ctx has its own stream
            t1 = time.perf_counter()
            ctx.set_cloud_device(all_pts.data_ptr(), n)
            ctx.set_knn_rows_device(k, all_idx.data_ptr(), all_d2.data_ptr())
            ctx.knn_graph(k, download=False)
            ctx.normals(k, download=False)
            lab, seeds = ctx.region_seeds(api.RgParams(50, 10000, k, 0.08, 1.0), download=False)
            ctx.ift(None, None, download=False)
            ctx.synchronize()
            out["rank0_graph_seeds_ift_s"] = time.perf_counter() - t1
            out["stats"] = {kk: v for kk, v in ctx.stats().items() if kk in ("n_arcs", "n_seeds", "ift_levels")}
        if args.verify:
            torch.cuda.synchronize()
            c2 = api.Context(local)
            c2.set_cloud_device(all_pts.data_ptr(), n)
            vi = torch.empty((n, k), dtype=torch.int32, device=dev)
            vd = torch.empty((n, k), dtype=torch.float32, device=dev)
            c2.knn_device(k, vi.data_ptr(), vd.data_ptr())
            out["verify_rows_equal"] = bool(torch.equal(vi, all_idx) and torch.equal(vd, all_d2))
            c2.close()
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
