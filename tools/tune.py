#!/usr/bin/env python
"""Tuning sweep over the compile-time variants the library keeps selectable through environment
variables (IFTWT_NRM_VARIANT, IFTWT_RGA_VARIANT, IFTWT_IFT_WAVES, IFTWT_KNN_MARGIN, ...): one GPU call
measures them all on the benchmark cloud, the winner becomes the default in the source.

    python tools/tune.py --workload cfg3 --sweep IFTWT_NRM_VARIANT=0,1,2,3,4,5,6 IFTWT_RGA_VARIANT=0,1,2,3,4,5

The library reads its switches at every call, so the variants run one after the other in one process:
3 warm-up + `--steps` timed segmentations of the same device-resident cloud each; prints one JSON line per variant with the
whole-step time and the per-stage CUDA-event times, and a checksum of cost / label so a variant that
changes the result is caught on the spot."""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(args, ctx, api, torch, dev, n, k, env):
    import numpy as np
    for name in list(os.environ):
        if name.startswith("IFTWT_"):
            del os.environ[name]
    os.environ.update(env)
    rg = api.RgParams(50, 10000, k, 0.08, 1.0)
    stream = torch.cuda.ExternalStream(ctx.stream())
    for _ in range(3):
        ctx.set_cloud_device(dev.data_ptr(), n)
        ctx.segment(k, k, rg, download=False)
    acc = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.set_cloud_device(dev.data_ptr(), n)
        ctx.segment(k, k, rg, download=False)
        for s, v in ctx.stage_ms().items():
            acc[s] = acc.get(s, 0.0) + v
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ctx.set_cloud_device(dev.data_ptr(), n)
    cost, label, ns = ctx.segment(k, k, rg)
    h = hashlib.sha256(cost.tobytes() + label.tobytes()).hexdigest()[:16]
    print(json.dumps({"env": env, "ms": round(ms, 4),
                      "stage_ms": {s: round(v / args.steps, 4) for s, v in acc.items() if v}, "seeds": ns, "sha": h,
                      "stats": {kk: ctx.stats()[kk] for kk in ("cell_size", "density_spread", "knn_fallback")}}),
          flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg5"])
    ap.add_argument("--points", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--sweep", nargs="*", default=[], help="NAME=v1,v2,... (each NAME swept on its own)")
    args = ap.parse_args()
    import torch
    from iftwt_rg_b200 import api, synth
    n = args.points or (10_000_000 if args.workload == "cfg3" else 2_000_000)
    pts = synth.scene_facade(n) if args.workload == "cfg3" else synth.scene_primitives(n)
    k = args.k or (32 if args.workload == "cfg3" else 16)
    dev = torch.from_numpy(pts).cuda()
    torch.cuda.synchronize()
    runs = [{}]
    for spec in args.sweep:
        name, vals = spec.split("=")
        runs += [{name: v} for v in vals.split(",")]
    runs.append({})   # the default once more at the end: drift check
    with api.Context(0) as ctx:
        for env in runs:
            measure(args, ctx, api, torch, dev, n, k, env)


if __name__ == "__main__":
    main()
