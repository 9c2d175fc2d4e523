#!/usr/bin/env python
"""bench.py — IFT-WT segmentation throughput (BASELINE.json metric) on N B200s.

A step = one pass of the hot path (k-NN graph, normals/curvature, region-growing
seeds, IFT watershed) over one synthetic cloud; every rank runs K steps on its own clouds.

  value     points/s, all ranks, clouds already resident in HBM when the timed region starts,
            results left on the device (CUDA events, max over ranks).  The K clouds go through
            iftwt_batch_segment — 2 lanes (ctx + stream + host thread) on the GPU pulling from
            one queue, so the latency-bound stretches of one cloud (227 small IFT levels) are filled
            by the kernels of the other; every cloud still runs the whole pipeline.  The one-ctx,
            one-cloud-at-a-time figure that the roofline and the stage times refer to is
            detail.serial_one_ctx (--serial-headline makes it `value`).
  e2e       the same through the C ABI with HOST buffers: pinned host cloud in,
            host cost/label out, every step, inside the timed region
  roofline  the dominant kernel (k-NN fast path) against the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference's CPU path; the reference as a
            whole cannot be built here) on a bounded sample of the same cloud, plus the
            reference's own ift.hpp (oracle/_ref) on a 20 k-point sample
  detail.parity   GPU vs oracle on that sample, counted (cost / label mismatches, tie-only
            disagreements with the reference's first-offer rule, the reference's own fixture)
  detail.cfg5     BASELINE config 5 on the same GPUs: 64 clouds of 2 M points split over the ranks
  detail.sharded  (N > 1) BASELINE config 4's shape: k-NN + normals sharded by Morton range with an
            NCCL halo exchange inside libiftwt.so, rows verified against one GPU, then the whole
            sharded segmentation (gather to rank 0, IFT there)

`--workload cfg4` runs config 4 itself (100 M points, one iftwt_sharded_segment per step).
`--impl reference` times the CPU path as the reference arm (thread count pinned to the cores).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from iftwt_rg_b200 import synth  # noqa: E402

METRIC = "ift_wt_points_per_sec"
UNIT = "points/s"
WORKLOADS = {
    # name: (generator, default n, k, description)
    "cfg3": ("facade", 10_000_000, 32, "cfg3: synthetic 10M-point architectural facade, k=32, single B200"),
    "cfg2": ("primitives", 1_000_000, 16, "cfg2: synthetic 1M-point plane+cylinder+sphere, k=16, single B200"),
    # BASELINE config 5: the batch is split over the ranks (64 // N clouds each, distinct seeds), a step
    # segments the next cloud of the rank's share
    "cfg5": ("primitives", 2_000_000, 16, "cfg5: batch of 64 synthetic 2M-point clouds, k=16, one cloud per GPU at a time"),
}
# BASELINE config 4: the facade tiled 10x along x (100 M points), k = 16; k-NN + normals sharded by Morton
# range over the ranks with an NCCL halo exchange, rows gathered to rank 0, graph / seeds / IFT there.
# One step = one iftwt_sharded_segment over the whole cloud; total work is fixed as N grows ("strong").
CFG4 = {"points": 100_000_000, "k": 16, "tiles": 10,
        "desc": "cfg4: synthetic 100M-point cloud (facade tiled x10), k=16, k-NN/curvature sharded by Morton range "
                "with NVLink halo exchange (NCCL), IFT on one GPU"}
CFG4_PER_GPU = 12_500_000   # detail.sharded of the default line: cfg4's per-GPU share at 8 GPUs, on every N > 1
CFG5_BATCH = 64
THETA = 0.08  # smoothness threshold (rad) used for the synthetic scenes; see DESIGN.md
RG_MIN, RG_MAX = 50, 10000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg4"])
    ap.add_argument("--no-batch", action="store_true", help="skip the iftwt_batch_segment measurement (profiling runs)")
    ap.add_argument("--serial-headline", action="store_true",
                    help="report the one-ctx figure as `value` (default: the iftwt_batch_segment figure)")
    ap.add_argument("--lanes", type=int, default=0,
                    help="lanes of the iftwt_batch measurement (0: 4 for cfg5, 2 otherwise; swept on B200)")
    ap.add_argument("--no-cfg5", action="store_true",
                    help="skip detail.cfg5 (BASELINE config 5's batch of 64 2M-point clouds split over the ranks)")
    ap.add_argument("--no-sharded", action="store_true",
                    help="N > 1: skip detail.sharded (the cfg4-style Morton-range sharded pass)")
    ap.add_argument("--sharded-per-gpu", type=int, default=CFG4_PER_GPU,
                    help="points per GPU of detail.sharded (default: cfg4's share at 8 GPUs)")
    ap.add_argument("--points", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000,
                    help="points of the x-slab the cpu_baseline leg of the GPU arm runs once")
    ap.add_argument("--ref-sample", type=int, default=500_000,
                    help="points of the x-slab every step of --impl reference runs")
    ap.add_argument("--cpu-threads", type=int, default=0,
                    help="OpenMP threads of the CPU arm (0 = every core this process may run on)")
    ap.add_argument("--no-full-size-cpu", action="store_true",
                    help="--impl reference: skip the single full-size pass (N = 1 only) that checks the slab figure")
    ap.add_argument("--no-parity", action="store_true", help="skip detail.parity (GPU vs oracle on the cpu sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-lanes", type=int, default=2,
                    help="lanes of the e2e measurement (iftwt_batch_segment with host clouds; each lane has its own "
                         "upload / download stream)")
    return ap.parse_args()


def make_cloud(gen: str, n: int, seed_off: int = 0) -> np.ndarray:
    if gen == "facade":
        return synth.scene_facade(n, seed=0x1F03 + seed_off)
    return synth.scene_primitives(n, seed=0x1F02 + seed_off)


def cpu_sample(gen: str, n_full: int, n_sample: int) -> np.ndarray:
    """A spatial slab of the full cloud holding ~n_sample points: same density, same scene."""
    if n_sample >= n_full:
        return make_cloud(gen, n_full)
    # generate a prefix large enough, cut an x-slab out of it scaled to the full density
    frac = n_sample / n_full
    pts = make_cloud(gen, n_full if n_full <= 4_000_000 else 0) if n_full <= 4_000_000 else None
    if pts is None:
        # stream the full cloud in chunks, keep the slab
        keep = []
        lo, hi = (0.0, 40.0 * frac) if gen == "facade" else (-5.0, -5.0 + 10.0 * frac * 1.6)
        chunk = 2_000_000
        fn = synth.scene_facade if gen == "facade" else synth.scene_primitives
        seed = 0x1F03 if gen == "facade" else 0x1F02
        for s in range(0, n_full, chunk):
            c = fn(min(chunk, n_full - s), seed=seed, start=s)
            keep.append(c[(c[:, 0] >= lo) & (c[:, 0] < hi)])
        return np.ascontiguousarray(np.concatenate(keep))
    xs = np.sort(pts[:, 0])
    hi = xs[min(len(xs) - 1, n_sample)]
    return np.ascontiguousarray(pts[pts[:, 0] < hi])


def cpu_threads(args) -> int:
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm would silently run on one core at N > 1.  The
    thread count is set explicitly, to every core the process may run on, and reported."""
    from oracle import pyoracle as orc
    t = args.cpu_threads or len(os.sched_getaffinity(0))
    orc.set_threads(t)
    return orc.num_threads()


def run_cpu_path(pts: np.ndarray, k: int) -> float:
    """The reference's CPU path as restated by the oracle; returns seconds."""
    from oracle import pyoracle as orc
    t0 = time.perf_counter()
    idx, _ = orc.knn_self(pts, k)
    off, adj = orc.knn_graph(idx)
    nrm = orc.normals(pts, idx)
    lab, _, kept, _ = orc.region_growing(nrm, idx, THETA, 1.0, RG_MIN, RG_MAX)
    seeds, _ = orc.rg_seeds(pts, lab, kept)
    orc.ift(off, adj, pts[:, 3], seeds, orc.POLICY_R)   # faithful-fast first-offer Dial == ift.hpp:170-250
    return time.perf_counter() - t0


def reference_ift_sample(gen: str, n_full: int, k: int, n_sample: int = 20_000):
    """The REFERENCE's own IFT_PCD (ift.hpp compiled unmodified against oracle/shim — oracle/_ref,
    built where /root/reference exists and shipped prebuilt) on the k-NN graph of a small slab of the
    workload.  Its heap has an O(|Q|) remove() per decrease-key and its region bookkeeping spawns one
    OpenMP task per vertex per region contact, so its cost grows faster than linearly: the figure is
    reported for the stated sample only and is not extrapolated."""
    try:
        from oracle import pyoracle as orc
        from oracle import pyref
        if not pyref.available():
            return None
        pts = cpu_sample(gen, n_full, n_sample)
        idx, _ = orc.knn_self(pts, k)
        off, adj = orc.knn_graph(idx)
        nrm = orc.normals(pts, idx)
        lab, _, kept, _ = orc.region_growing(nrm, idx, THETA, 1.0, RG_MIN, RG_MAX)
        seeds, _ = orc.rg_seeds(pts, lab, kept)
        if len(seeds) == 0:
            seeds = np.arange(0, len(pts), max(1, len(pts) // 16), dtype=np.int32)
        t0 = time.perf_counter()
        cost, root, *_ = pyref.ift(pts, off, adj, pts[np.asarray(seeds, np.int64), :3])
        t = time.perf_counter() - t0
        fcost, _ = orc.ift(off, adj, pts[:, 3], np.asarray(seeds, np.int32), orc.POLICY_R)
        return {"kind": "reference", "what": "IFT_PCD of ift.hpp (unmodified) on the symmetric k-NN graph of the sample",
                "points": int(len(pts)), "seeds": int(len(seeds)), "seconds": t, "points_per_s": len(pts) / t,
                "costs_equal_port": bool(np.array_equal(cost, fcost)),
                "note": "superlinear in the sample size (O(|Q|) heap remove, task-per-vertex region bookkeeping)"}
    except Exception as e:  # the checker must never take the bench down
        return {"kind": "reference", "error": repr(e)[:200]}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            hi = sorted(sm)[len(sm) // 2:]          # upper half = under load
            out["sm_mhz"] = float(np.median(hi))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def place_rank(local: int, world: int) -> dict:
    """Host placement of a rank before it allocates pinned memory: the cores NVML reports as close to its
    GPU, split evenly between the ranks that share them, so that the e2e copy threads of eight ranks do not
    migrate over one another (first-touch then places the pinned buffers on that node).  On a box whose GPUs
    all report the same CPU set (the B200 pool: one NUMA node, `nvidia-smi topo -m`) only the split applies."""
    info = {"cpus_before": len(os.sched_getaffinity(0))}
    try:
        allowed = sorted(os.sched_getaffinity(0))
        near = allowed
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
            if ideal & set(allowed):
                near = sorted(ideal & set(allowed))
            info["gpu_cpu_affinity"] = f"{min(ideal)}-{max(ideal)}" if ideal else None
        except Exception:
            pass
        if world > 1 and len(near) // world >= 3:
            per = len(near) // world
            mine = near[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            info["pinned_to"] = f"{mine[0]}-{mine[-1]}"
        elif near != allowed:
            os.sched_setaffinity(0, near)
            info["pinned_to"] = f"{near[0]}-{near[-1]}"
    except Exception as e:
        info["error"] = repr(e)[:100]
    info["cpus_after"] = len(os.sched_getaffinity(0))
    return info


class ShardedRun:
    """BASELINE config 4 through the C ABI (iftwt_comm_*, iftwt_sharded_knn_normals, iftwt_sharded_segment):
    one process per GPU, the slice of this rank resident in HBM, torch.distributed only for the barrier,
    the max-over-ranks reductions and the 128-byte communicator id."""

    def __init__(self, torch, dist, api, world, rank, local, n_total, k, tiles):
        from iftwt_rg_b200 import sharded as host
        self.torch, self.dist, self.api, self.host = torch, dist, api, host
        self.world, self.rank, self.local = world, rank, local
        self.n_total, self.k = n_total, k
        self.lo, self.hi = host.slice_bounds(n_total, world, rank)
        # generation order is random in space, so a contiguous id range is a uniform subsample of 1-2 tiles
        self.host_pts = torch.from_numpy(synth.facade_tiled(n_total, tiles=tiles, start=self.lo,
                                                            count=self.hi - self.lo)).pin_memory()
        self.dev_pts = self.host_pts.cuda()
        torch.cuda.synchronize()
        self.comm = host.make_comm(local)
        self.ctx = self.comm.ctx(0)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", local))
        self.rg = api.RgParams(RG_MIN, RG_MAX, k, THETA, 1.0)
        self.radius = 0.0

    def slices(self):
        return [(self.dev_pts.data_ptr(), self.hi - self.lo, self.lo)]

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op):
        t = self.torch.tensor(vals, device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return [float(x) for x in t.tolist()]

    def timed(self, fn, steps=1):
        """CUDA events on the shard's stream around `steps` calls, max over ranks (ms)."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        r = None
        for _ in range(steps):
            r = fn()
        e1.record(self.stream)
        self.barrier()
        return self.reduce([e0.elapsed_time(e1)], self.dist.ReduceOp.MAX)[0], r

    def knn_normals(self, radius):
        return self.comm.sharded_knn_normals(self.slices(), self.k, True, radius)[0]

    def segment(self, download=False):
        cost = label = None
        if download and self.rank == 0:
            if not hasattr(self, "host_cost"):   # pinned, allocated once
                self.host_cost = self.torch.empty(self.n_total, dtype=self.torch.int32).pin_memory()
                self.host_label = self.torch.empty(self.n_total, dtype=self.torch.int32).pin_memory()
            cost = self.host_cost.numpy().view(np.uint32)
            label = self.host_label.numpy().view(np.uint32)
        return self.comm.sharded_segment(self.slices(), self.k, self.rg, root=0, halo_radius=self.radius,
                                         n_total=self.n_total, download=download, cost=cost, label=label)

    def warm(self):
        """First call: allocations, NCCL channels, and the halo-radius estimate.  Later calls pass a little
        more than the largest k-th neighbour distance the first one measured — what a caller that segments
        similar clouds repeatedly does (iftwt.h: iftwt_shard_out.kth_distance_max)."""
        ms, o = self.timed(lambda: self.knn_normals(0.0))
        self.first = {"ms": ms, "attempts": o.attempts, "estimated_radius": o.halo_radius,
                      "kth_distance_max": o.kth_distance_max}
        self.radius = 1.05 * o.kth_distance_max
        return o

    def phase_stats(self, o):
        dist = self.dist
        mx = self.reduce([o.ms_estimate, o.ms_partition, o.ms_exchange, o.ms_compute, o.ms_export, o.knn_cell_ms],
                         dist.ReduceOp.MAX)
        sm = self.reduce([float(o.halo_bytes_sent), float(o.exchange_bytes_sent), float(o.n_owned), float(o.n_halo)],
                         dist.ReduceOp.SUM)
        per = self.torch.zeros(self.world * 2, device="cuda", dtype=self.torch.float64)
        per[2 * self.rank], per[2 * self.rank + 1] = float(o.n_owned), float(o.n_halo)
        if self.world > 1:
            dist.all_reduce(per)
        per = [int(x) for x in per.tolist()]
        return {"phase_ms_max_over_ranks": {"estimate": mx[0], "partition_route_pack": mx[1], "nccl_exchange": mx[2],
                                            "grid_knn_normals": mx[3], "export_rows_and_check": mx[4],
                                            "knn_cell_kernel": mx[5]},
                "halo_bytes_over_nvlink": int(sm[0]), "exchange_bytes_over_nvlink": int(sm[1]),
                "owned_per_rank": per[0::2], "halo_per_rank": per[1::2], "halo_fraction": sm[3] / max(sm[2], 1.0),
                "halo_radius": o.halo_radius, "attempts": o.attempts, "peer_to_peer": int(o.peer_to_peer)}

    def verify(self, o, target_rows=1_000_000):
        """A sample of every rank's rows against ONE GPU holding the whole cloud: rank 0 answers the sampled
        points as external queries (the general k-NN path, not the code that produced the rows)."""
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", self.local)
        v = self.host.as_torch(o, self.local)
        stride = max(1, self.n_total // target_rows)
        sel = slice(0, None, stride)
        if o.n_owned:
            gid, knn, d2 = v["gid"][sel].clone(), v["knn"][sel].clone(), v["d2"][sel].clone()
        else:
            gid = torch.zeros(0, dtype=torch.uint32, device=dev)
            knn = torch.zeros((0, self.k), dtype=torch.int32, device=dev)
            d2 = torch.zeros((0, self.k), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        if self.world > 1:
            per = (self.n_total + self.world - 1) // self.world
            cap = (per + 2 * self.world) // stride + 4096   # owned counts are balanced to a bin, not exactly
            cnt = torch.tensor([gid.shape[0]], device=dev, dtype=torch.int64)
            mx = cnt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            cap = max(cap, int(mx))
            def pad(t, shape, dtype):
                out = torch.zeros(shape, dtype=dtype, device=dev)
                out[: t.shape[0]] = t
                return out
            g_all = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(self.world)]
            k_all = [torch.empty((cap, self.k), dtype=torch.int32, device=dev) for _ in range(self.world)]
            d_all = [torch.empty((cap, self.k), dtype=torch.float32, device=dev) for _ in range(self.world)]
            c_all = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
            dist.all_gather(g_all, pad(gid.view(torch.int32), (cap,), torch.int32))
            dist.all_gather(k_all, pad(knn, (cap, self.k), torch.int32))
            dist.all_gather(d_all, pad(d2, (cap, self.k), torch.float32))
            dist.all_gather(c_all, cnt)
            # the whole cloud on every rank (only rank 0 uses it): slices padded to one size
            slab = torch.zeros((per, 4), dtype=torch.float32, device=dev)
            slab[: self.hi - self.lo] = self.dev_pts
            s_all = [torch.empty_like(slab) for _ in range(self.world)]
            dist.all_gather(s_all, slab)
            if self.rank != 0:
                return None
            whole = torch.cat([s_all[r][: self.host.slice_bounds(self.n_total, self.world, r)[1] -
                                        self.host.slice_bounds(self.n_total, self.world, r)[0]] for r in range(self.world)])
            gid = torch.cat([g_all[r][: int(c_all[r])] for r in range(self.world)])
            knn = torch.cat([k_all[r][: int(c_all[r])] for r in range(self.world)])
            d2 = torch.cat([d_all[r][: int(c_all[r])] for r in range(self.world)])
            del s_all, g_all, k_all, d_all
        else:
            whole = self.dev_pts
            gid = gid.view(torch.int32)
        q = whole[gid.long(), :3].cpu().numpy()
        torch.cuda.synchronize()
        with self.api.Context(self.local) as c1:
            c1.set_cloud_device(whole.data_ptr(), self.n_total)
            ridx, rd2 = c1.knn_query(q, self.k)
        ok = bool(np.array_equal(ridx, knn.cpu().numpy()) and
                  np.array_equal(rd2.view(np.uint32), d2.cpu().numpy().view(np.uint32)))
        return {"verify_rows_equal": ok, "verified_rows": int(len(q)),
                "verified_against": "one GPU holding the whole cloud, general k-NN path (iftwt_knn_query)"}

    def close(self):
        self.comm.close()


def sharded_detail(args, torch, dist, api, world, rank, local, n_total, k, tiles, steps):
    """The cfg4-shaped pass reported as detail.sharded in the default line when N > 1."""
    run = ShardedRun(torch, dist, api, world, rank, local, n_total, k, tiles)
    try:
        run.warm()
        for _ in range(2):   # the second call maps the peers' receive buffers (cudaIpcOpenMemHandle: ~0.1 s, once)
            run.knn_normals(run.radius)
        ms, o = run.timed(lambda: run.knn_normals(run.radius), steps)
        ms /= steps
        out = {"workload": f"cfg4-shaped: facade tiled, {n_total} points over {world} GPUs "
                           f"({(n_total + world - 1) // world} per GPU), k={k}; slices resident in HBM",
               "what": "iftwt_sharded_knn_normals: Morton-range partition, owner + halo exchange (grouped ncclSend/"
                       "ncclRecv), grid + k-NN + normals on owned + halo points, rows exported with global ids",
               "points": n_total, "k": k, "n_gpus": world, "steps": steps, "knn_normals_ms": ms,
               "knn_normals_points_per_s": n_total / (ms * 1e-3), "first_call": run.first}
        out.update(run.phase_stats(o))
        alg = (16 + 4 * k + 16) * n_total      # fused k-NN + curvature figure of SURVEY 8(d)
        out["algorithmic_gbs_all_gpus"] = alg / (ms * 1e-3) / 1e9
        ver = run.verify(o)
        if ver:
            out.update(ver)
        run.segment()                                   # warm-up: allocations on the root
        run.segment()                                   # ... which the peers map at the start of this call
        ms_seg, res = run.timed(lambda: run.segment())
        st = run.ctx.stats() if rank == 0 else {}
        out["segment_ms"] = ms_seg
        out["segment_peer_to_peer"] = int(res[3][0].peer_to_peer)
        out["segment_points_per_s"] = n_total / (ms_seg * 1e-3)
        out["segment_what"] = "iftwt_sharded_segment: the above + rows gathered to rank 0 (NCCL), graph / seeds / IFT there"
        if rank == 0:
            out["segment_seeds"] = int(res[2])
            out["segment_root_stage_ms"] = {s: v for s, v in run.ctx.stage_ms().items()
                                            if s in ("grid", "graph", "graph_rg", "seeds", "ift")}
            out["segment_stats"] = {kk: st[kk] for kk in ("n_arcs", "n_seeds", "ift_levels")}
        return out
    finally:
        run.close()


def reference_arm(args, gen, n_full, k, desc):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = cpu_threads(args)
    pts = cpu_sample(gen, n_full, args.ref_sample)
    for _ in range(min(args.warmup, 1)):
        run_cpu_path(pts[: max(1000, len(pts) // 20)], k)
    times = [run_cpu_path(pts, k) for _ in range(max(1, args.steps))]
    t = float(np.mean(times))
    v = len(pts) / t
    sample = f"x-slab of the {n_full}-point cloud holding {len(pts)} points, full pipeline per step"
    full = None
    if world == 1 and not args.no_full_size_cpu and len(pts) < n_full:
        # the slab figure is an extrapolation: check it once against the whole cloud
        whole = make_cloud(gen, n_full)
        tf = run_cpu_path(whole, k)
        full = {"points": n_full, "seconds": tf, "points_per_s": n_full / tf, "cores": cores,
                "slab_over_full": v / (n_full / tf)}
        del whole
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": max(1, args.steps), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
        "config": {"workload": desc, "k": k, "points": n_full, "sample_points": len(pts), "same_config": False,
                   "note": "reference CPU path restated by oracle/ (the reference as a whole needs PCL/Boost/VTK "
                           "and cannot be built in this image): OpenMP k-NN + normals, sequential region growing "
                           "and IFT as in the reference, on an x-slab of the workload's cloud (same scene, same "
                           "density).  Its own graph.hpp / ift.hpp do compile against "
                           "oracle/shim and pin this port bit for bit (tests/test_ref_*.py); ift.hpp itself is "
                           "timed on a small sample in cpu_baseline.reference_ift"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "full_size_once": full,
                         "reference_ift": reference_ift_sample(gen, n_full, k)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def gpu_vs_oracle_parity(ctx, api, pts: np.ndarray, k: int) -> dict:
    """GPU vs the oracle on the cpu sample (the slab the CPU baseline ran on), counted — the numbers the
    north star asks for next to the throughput: costs and policy-C labels must be equal; labels under
    the reference's abstract first-offer rule (policy R, ift.hpp:196-205) may differ at ties only, and
    those are counted separately.  Plus the GPU against the REFERENCE's own IFT_PCD on the committed
    fixture (tests/golden/ref_ift_knn16_6k.npz, ift.hpp compiled unmodified)."""
    from oracle import pyoracle as orc
    out = {"sample_points": int(len(pts)), "k": k}
    rg = api.RgParams(RG_MIN, RG_MAX, k, THETA, 1.0)
    ctx.set_cloud(pts)
    cost, label, ns = ctx.segment(k, k, rg)
    idx, _ = orc.knn_self(pts, k)
    off, adj = orc.knn_graph(idx)
    nrm = orc.normals(pts, idx)
    lab, _, kept, _ = orc.region_growing(nrm, idx, THETA, 1.0, RG_MIN, RG_MAX)
    seeds, _ = orc.rg_seeds(pts, lab, kept)
    ccost, clabel = orc.ift(off, adj, pts[:, 3], seeds, orc.POLICY_C)
    _, rlabel = orc.ift(off, adj, pts[:, 3], seeds, orc.POLICY_R)
    out.update({
        "seeds": int(ns), "seeds_oracle": int(kept),
        "cost_mismatches_vs_oracle": int(np.count_nonzero(cost != ccost)),
        "label_mismatches_vs_oracle_policy_C": int(np.count_nonzero(label != clabel)),
        "tie_only_label_disagreements_vs_policy_R": int(np.count_nonzero(label != rlabel)),
        "tie_audit_violations": int(orc.ift_audit(off, adj, pts[:, 3], seeds, cost, label)),
    })
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "ref_ift_knn16_6k.npz"))
        fp, fk = z["pts"], int(z["k"])
        ctx.set_cloud(fp)
        foff, fadj = ctx.knn_graph(fk)
        q = np.zeros((len(z["seeds_xyz"]), 4), np.float32)
        q[:, :3] = z["seeds_xyz"]
        sidx, _ = ctx.knn_query(q, 1)
        fcost, flabel = ctx.ift(None, sidx[:, 0].astype(np.int32))
        out["reference_fixture"] = {
            "name": "ref_ift_knn16_6k (the reference's own IFT_PCD, ift.hpp unmodified)", "points": int(len(fp)),
            "adjacency_equal": bool(np.array_equal(foff, z["off"]) and np.array_equal(fadj, z["adj"])),
            "cost_mismatches": int(np.count_nonzero(fcost != z["ref_cost"])),
            "tie_only_label_disagreements": int(np.count_nonzero(flabel != z["ref_root"])),
            "tie_audit_violations": int(orc.ift_audit(foff, fadj, fp[:, 3], sidx[:, 0].astype(np.int32), fcost, flabel)),
        }
    except Exception as e:  # the checker must never take the bench down
        out["reference_fixture"] = {"error": repr(e)[:200]}
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE line, the JSON: everything libraries print there (NCCL's version
    banner, the reference's own progress prints) goes to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    quiet_stdout()
    if args.workload == "cfg4":
        if args.impl == "reference":
            # the CPU arm of config 4 is the same pipeline on the same scene: a slab of one tile
            reference_arm(args, "facade", 10_000_000, args.k or CFG4["k"], CFG4["desc"])
            return
        main_cfg4(args)
        return
    gen, n_default, k_default, desc = WORKLOADS[args.workload]
    n = args.points or n_default
    k = args.k or k_default
    if args.impl == "reference":
        reference_arm(args, gen, n, k, desc)
        return

    import torch
    import torch.distributed as dist
    from iftwt_rg_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    placement = place_rank(local, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one cloud per rank (batch data-parallel, cfg5 style): no data-path collective
    # cfg5: 64 // N clouds per rank; at most 8 DISTINCT clouds are generated per rank (the batch cycles through
    # them), which bounds the start-up time at N = 1
    per_rank = min(8, max(1, CFG5_BATCH // world)) if args.workload == "cfg5" else 1
    host_ins = [torch.from_numpy(make_cloud(gen, n, seed_off=(0x4E if per_rank > 1 else 0) + rank * per_rank + i)).pin_memory()
                for i in range(per_rank)]
    dev_ins = [h.cuda(non_blocking=False) for h in host_ins]
    in_nps = [h.numpy() for h in host_ins]
    host_in, dev_in = host_ins[0], dev_ins[0]
    turn = {"resident": 0, "e2e": 0}
    host_cost = torch.empty(n, dtype=torch.int32).pin_memory()
    host_label = torch.empty(n, dtype=torch.int32).pin_memory()
    cost_np = host_cost.numpy().view(np.uint32)
    label_np = host_label.numpy().view(np.uint32)
    in_np = host_in.numpy()

    ctx = api.Context(local)
    rg = api.RgParams(RG_MIN, RG_MAX, k, THETA, 1.0)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def step_resident():
        d = dev_ins[turn["resident"] % per_rank]
        turn["resident"] += 1
        ctx.set_cloud_device(d.data_ptr(), n)
        ctx.segment(k, k, rg, download=False)

    def step_e2e():
        h = in_nps[turn["e2e"] % per_rank]
        turn["e2e"] += 1
        ctx.set_cloud(h)
        ctx.segment(k, k, rg, cost=cost_np, label=label_np)

    def timed(fn, steps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.reset_counters()
    # per-stage device times, accumulated over the timed steps (events recorded inside the library)
    stage_acc: dict[str, float] = {}

    def step_resident_acc():
        step_resident()
        for s, v in ctx.stage_ms().items():
            stage_acc[s] = stage_acc.get(s, 0.0) + v

    ms_total = timed(step_resident_acc, args.steps)
    stats = ctx.stats()
    launches = int(stats["kernel_launches"])
    value = world * n * args.steps / (ms_total * 1e-3)

    # iftwt_batch_segment: the same clouds through `lanes` ctxs (one stream + one host thread each) that pull
    # from one queue; every cloud still goes through the whole pipeline, results are iftwt_segment's
    lanes = args.lanes or (4 if args.workload == "cfg5" else 2)
    batch_info = None
    try:
        if args.no_batch:
            raise RuntimeError("skipped (--no-batch)")
        with api.Batch(local, lanes) as bt:
            nb = args.steps                                           # EXACTLY K steps = K clouds
            clouds = [(dev_ins[i % per_rank].data_ptr(), n) for i in range(max(nb, 2 * lanes))]
            for _ in range(max(2, (max(args.warmup, 3) + 2 * lanes - 1) // (2 * lanes))):
                bt.segment(clouds[: 2 * lanes], k, rg, on_device=True)   # warm-up: >= W clouds, every lane allocates
            clouds = clouds[:nb]
            for l in range(lanes):
                bt.ctx(l).reset_counters()
            st0 = torch.cuda.ExternalStream(bt.ctx(0).stream(), device=torch.device("cuda", local))
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(st0)
            seeds_b = bt.segment(clouds, k, rg, on_device=True)    # returns when every lane has synchronised
            e1.record(st0)
            barrier()
            msb = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(msb, op=dist.ReduceOp.MAX)
            msb = float(msb.item())
            lane_launches = sum(int(bt.ctx(l).stats()["kernel_launches"]) for l in range(lanes))
            batch_info = {"lanes": lanes, "clouds": nb, "ms_per_cloud": msb / nb, "ms_total": msb,
                          "value": world * n * nb / (msb * 1e-3), "seeds_first_cloud": int(seeds_b[0]),
                          "gpu_launches": lane_launches,
                          "what": "iftwt_batch_segment, clouds resident in HBM, results left on the device"}
    except Exception as e:
        batch_info = {"error": repr(e)[:300]}
    clocks = sampler.stop() if rank == 0 else {}

    e2e = None
    if not args.no_e2e:
        # serial: one cloud at a time through one ctx (the latency a single caller sees)
        for _ in range(2):
            step_e2e()
        ms_serial = timed(step_e2e, args.steps)
        # pipelined: iftwt_batch_segment with HOST clouds and host results — `e2e_lanes` ctxs, each with its own
        # upload and download stream (csrc/capi.cu: LaneIO): the lane's next cloud is uploaded while the current
        # one is segmented and the results are downloaded while the next one runs.  Every step still uploads
        # its cloud from pinned memory and downloads cost + label.
        e2e_lanes = max(1, args.e2e_lanes)
        ms_e2e = ms_serial
        keep = []
        with api.Batch(local, e2e_lanes) as bt:
            nb = args.steps
            n_out = min(nb, 2 * e2e_lanes + 2)      # result buffers are reused round-robin by the caller
            for _ in range(n_out):
                hc = torch.empty(n, dtype=torch.int32).pin_memory()
                hl = torch.empty(n, dtype=torch.int32).pin_memory()
                keep.append((hc.numpy().view(np.uint32), hl.numpy().view(np.uint32), hc, hl))
            h_clouds = [in_nps[j % per_rank] for j in range(nb)]
            costs = [keep[j % n_out][0] for j in range(nb)]
            labels = [keep[j % n_out][1] for j in range(nb)]
            w = max(2 * e2e_lanes, 3)
            bt.segment(h_clouds[:w] if nb >= w else (h_clouds * w)[:w], k, rg, cost=(costs * w)[:w],
                       label=(labels * w)[:w])
            st0 = torch.cuda.ExternalStream(bt.ctx(0).stream(), device=torch.device("cuda", local))
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(st0)
            bt.segment(h_clouds, k, rg, cost=costs, label=labels)   # returns when every copy has landed
            e1.record(st0)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms_e2e = float(ms.item())
            if per_rank == 1:  # the same cloud went through every lane: same result as the serial ctx
                assert np.array_equal(costs[0], cost_np) and np.array_equal(labels[0], label_np)
                assert np.array_equal(costs[-1], cost_np) and np.array_equal(labels[-1], label_np)
        inflight = e2e_lanes
        e2e = {"value": world * n * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n * 16),
               "d2h_bytes_per_step": int(n * 8), "ms_per_step": ms_e2e / args.steps,
               "entry_point": f"iftwt_batch_segment, host clouds + host results, {inflight} lanes, copies in the "
                              "lanes' own upload / download streams",
               "clouds_in_flight": inflight, "serial_ms_per_step": ms_serial / args.steps,
               "serial_value": world * n * args.steps / (ms_serial * 1e-3)}

    if rank == 0:
        peaks = {}
        which = "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            which = "measured"
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        stage_ms = {s: v / args.steps for s, v in stage_acc.items()}
        # dominant kernel: the k-NN fast path; algorithmic bytes (16 + 4k) * N per launch
        knn_ms = stage_ms.get("knn_cell", 0.0) or stage_ms.get("knn", 0.0)
        alg_bytes = (16 + 4 * k) * n
        achieved = alg_bytes / (knn_ms * 1e-3) / 1e9 if knn_ms > 0 else 0.0
        traffic = None
        issue_slot = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json"))).get(args.workload)
            if tr and n == n_default and k == k_default:
                traffic = int(tr["dram_bytes_read"]) + int(tr["dram_bytes_write"])
                if tr.get("warp_instructions") and knn_ms > 0:
                    # the limiter that is actually active: warp instructions issued (ncu smsp__inst_executed.sum of
                    # the committed capture) against what the SMs can issue — 4 schedulers x 1 instruction per clock
                    wi = float(tr["warp_instructions"])
                    sm_clock = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
                    rate = 148 * 4 * sm_clock
                    issue_slot = {"warp_instructions_per_launch": int(wi), "warp_instructions_per_query": wi / n,
                                  "issue_rate_warp_inst_per_s": rate, "floor_ms": 1e3 * wi / rate,
                                  "frac": (1e3 * wi / rate) / knn_ms, "source": tr.get("source")}
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "knn_cell_kernel (first launch + the crowded-cell launch, one event pair)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic, "issue_slot": issue_slot,
                    "active_limiter": "instruction issue (ncu: 78.4 % issue-active, ALU pipe 61 %, LSU pipe 53 %, DRAM 8.8 % busy), not HBM: "
                                      "see issue_slot",
                    "peak_source": which, "algorithmic_bytes_per_launch": alg_bytes,
                    "avg_launch_ms": knn_ms}
        nrm_ms = stage_ms.get("normals", 0.0)
        # IFT: algorithmic bytes 20 V + 12 E-> over the whole stage, and the global-atomic
        # rate — atomics counted by ncu at L2 (atomicMin offers + atomicCAS hooks of all level kernels of
        # one run, profiles/kernel_traffic.json) over the stage time measured in THIS run
        ift_ms = stage_ms.get("ift", 0.0)
        ift = None
        if ift_ms > 0:
            ift_bytes = 20 * n + 12 * int(stats["n_arcs"])
            ift = {"ms": ift_ms, "levels": int(stats["ift_levels"]), "algorithmic_bytes": ift_bytes,
                   "achieved_gbs": ift_bytes / (ift_ms * 1e-3) / 1e9,
                   "frac_of_hbm": ift_bytes / (ift_ms * 1e-3) / 1e9 / peak if peak else None,
                   "bound": "latency of dependent gathers / atomics, not bandwidth"}
            try:
                ta = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json"))).get(args.workload + "_ift")
                if ta and n == n_default and k == k_default:
                    atoms = int(ta["l2_atomic_alu_requests"]) + int(ta["l2_atomic_cas_requests"])
                    ift["atomics_per_run"] = atoms
                    ift["atomics_per_s"] = atoms / (ift_ms * 1e-3)
                    ift["atomics_source"] = ta["source"]
            except Exception:
                pass
        extra = {
            "stage_ms": stage_ms,
            "normals_gbs": ((32 + 4 * k) * n) / (nrm_ms * 1e-3) / 1e9 if nrm_ms > 0 else None,
            "normals_frac_of_hbm": ((32 + 4 * k) * n) / (nrm_ms * 1e-3) / 1e9 / peak if nrm_ms > 0 and peak else None,
            "ift": ift,
            "batch": batch_info,
            "serial_one_ctx": {"value": value, "ms_per_step": ms_total / args.steps},
            "stats": {kk: stats[kk] for kk in ("n_cells", "cell_size", "density_spread", "knn_fallback", "n_arcs",
                                                "rg_segments", "rg_sweeps", "n_seeds", "ift_levels")},
        }
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = cpu_threads(args)
            sample = cpu_sample(gen, n, args.cpu_sample)
            t = run_cpu_path(sample, k)
            cpu = {"value": len(sample) / t, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"x-slab of the cloud holding {len(sample)} points, full pipeline once ({t:.1f} s)",
                   "reference_ift": reference_ift_sample(gen, n, k)}
            if not args.no_parity:
                try:
                    extra["parity"] = gpu_vs_oracle_parity(ctx, api, sample, k)
                except Exception as e:  # reported, never fatal: the timed numbers above stand on their own
                    extra["parity"] = {"error": repr(e)[:300]}
        head_value, head_ms, head_launches = value, ms_total / args.steps, launches
        entry = "iftwt_segment, one ctx, one cloud at a time"
        if batch_info and "value" in batch_info and not args.serial_headline:
            # whole-job throughput: the K clouds of the timed region go through iftwt_batch_segment (`lanes` ctxs,
            # one stream + one host thread each, pulling from one queue).  Every cloud runs the whole pipeline;
            # the one-ctx figure, which the roofline and the stage times refer to, stays in detail.serial_one_ctx
            head_value, head_ms, head_launches = batch_info["value"], batch_info["ms_per_cloud"], batch_info["gpu_launches"]
            entry = f"iftwt_batch_segment, {lanes} lanes on the GPU"
        line = {
            "metric": METRIC, "value": head_value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": desc if (n == n_default and k == k_default) else f"{gen} n={n} k={k}",
                       "k": k, "points_per_gpu": n, "entry_point": entry,
                       "parallelism": f"independent clouds, {world} GPU(s), no data-path collective", "clouds_per_rank": per_rank,
                       "host_placement": placement,
                       "l2": "inputs (16 B/pt) and every intermediate exceed the 126 MB L2; cloud re-uploaded "
                             "device-to-device each step"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": head_launches, "roofline": roofline,
            "cpu_baseline": cpu, "detail": extra,
        }
    ctx.close()
    if not args.no_cfg5 and args.workload == "cfg3":
        # BASELINE config 5 on the same GPUs: 64 clouds of 2 M points, 64 // N per rank through iftwt_batch_segment
        # (4 lanes), no data-path collective.  At most 8 distinct clouds per rank are generated; the batch cycles.
        c5 = None
        try:
            n5, k5 = 2_000_000, 16
            mine = max(1, CFG5_BATCH // world)
            distinct = min(8, mine)
            dev5 = [torch.from_numpy(synth.scene_primitives(n5, seed=0x1F50 + rank * mine + i)).cuda()
                    for i in range(distinct)]
            rg5 = api.RgParams(RG_MIN, RG_MAX, k5, THETA, 1.0)
            with api.Batch(local, 4) as bt:
                clouds = [(dev5[i % distinct].data_ptr(), n5) for i in range(mine)]
                bt.segment(clouds[:8], k5, rg5, on_device=True)
                bt.segment(clouds[:8], k5, rg5, on_device=True)
                st0 = torch.cuda.ExternalStream(bt.ctx(0).stream(), device=torch.device("cuda", local))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e0.record(st0)
                seeds5 = bt.segment(clouds, k5, rg5, on_device=True)
                e1.record(st0)
                barrier()
                ms5 = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(ms5, op=dist.ReduceOp.MAX)
                ms5 = float(ms5.item())
            c5 = {"workload": WORKLOADS["cfg5"][3], "clouds": mine * world, "clouds_per_gpu": mine,
                  "distinct_clouds_per_gpu": distinct, "points_per_cloud": n5, "k": k5, "lanes": 4, "batch_ms": ms5,
                  "clouds_per_s": mine * world / (ms5 * 1e-3), "points_per_s": mine * world * n5 / (ms5 * 1e-3),
                  "points_per_s_per_gpu": mine * n5 / (ms5 * 1e-3), "seeds_first_cloud": int(seeds5[0]),
                  "what": "iftwt_batch_segment per rank, clouds resident in HBM, results left on the device"}
            del dev5
        except Exception as e:
            c5 = {"error": repr(e)[:300]}
        if rank == 0:
            line["detail"]["cfg5"] = c5
    if world > 1 and not args.no_sharded and args.workload == "cfg3":
        # BASELINE config 4's shape on the same GPUs: k-NN + normals sharded by Morton range with an NCCL halo
        # exchange, inside the library.  A collective that hangs must not cost the line above: a watchdog on
        # every rank lets rank 0 print what it has and ends the process.
        import threading as _th

        def _bail():
            if rank == 0:
                line["detail"]["sharded"] = {"error": "timed out after 240 s"}
                emit(line)
            os._exit(0)

        wd = _th.Timer(240.0, _bail)
        wd.daemon = True
        wd.start()
        try:
            sh = sharded_detail(args, torch, dist, api, world, rank, local, args.sharded_per_gpu * world, CFG4["k"],
                                max(1, round(args.sharded_per_gpu * world / 10_000_000)), steps=5)
        except Exception as e:
            sh = {"error": repr(e)[:400]}
        wd.cancel()
        if rank == 0:
            line["detail"]["sharded"] = sh
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main_cfg4(args):
    """--workload cfg4: a step = one iftwt_sharded_segment over the whole 100 M-point cloud."""
    import torch
    import torch.distributed as dist
    from iftwt_rg_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total = args.points or CFG4["points"]
    k = args.k or CFG4["k"]
    tiles = max(1, round(n_total / 10_000_000))
    run = ShardedRun(torch, dist, api, world, rank, local, n_total, k, tiles)
    o = run.warm()
    for _ in range(max(args.warmup, 3) - 1):
        run.knn_normals(run.radius)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_knn, o = run.timed(lambda: run.knn_normals(run.radius), args.steps)
    ms_knn /= args.steps
    phases = run.phase_stats(o)
    ver = run.verify(o)
    run.segment()
    run.segment()
    run.ctx.reset_counters()
    ms_seg, res = run.timed(lambda: run.segment(), args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    launches = run.reduce([float(run.ctx.stats()["kernel_launches"])], dist.ReduceOp.SUM)[0]
    root_stage = run.ctx.stage_ms() if rank == 0 else {}
    root_stats = run.ctx.stats() if rank == 0 else {}
    e2e = None
    if not args.no_e2e:
        # slices from pinned host memory every step, cost + label of the whole cloud back to the host on rank 0
        # (the copy runs on torch's own stream and is waited for before the library is called: torch must never
        # see the library's stream as a user of its pinned blocks — it would record an event on that stream when
        # the block is freed, after iftwt_comm_destroy has destroyed it)
        def step_e2e():
            run.dev_pts.copy_(run.host_pts, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return run.segment(download=True)
        step_e2e()
        ms_e2e, _ = run.timed(step_e2e, args.steps)
        e2e = {"value": n_total * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n_total * 16),
               "d2h_bytes_per_step": int(n_total * 8), "ms_per_step": ms_e2e / args.steps,
               "note": "every rank uploads its slice from pinned memory; rank 0 downloads cost + label of all points "
                       "into pinned buffers"}
    if rank == 0:
        peaks, which = {}, "fallback"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            which = "measured"
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        cell_ms = phases["phase_ms_max_over_ranks"]["knn_cell_kernel"]
        per_gpu = (n_total + world - 1) // world
        alg = (16 + 4 * k) * per_gpu
        ach = alg / (cell_ms * 1e-3) / 1e9 if cell_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": n_total * args.steps / (ms_seg * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_seg / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": CFG4["desc"] if n_total == CFG4["points"] and k == CFG4["k"] else
                       f"facade tiled n={n_total} k={k}", "k": k, "points": n_total,
                       "parallelism": f"k-NN + normals sharded by Morton range over {world} GPUs (NCCL halo exchange), "
                                      "rows gathered to rank 0, graph / seeds / IFT on rank 0",
                       "l2": "every slice (>= 200 MB) and every intermediate exceed the 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "knn_cell_kernel on the slowest shard (owned + halo points)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                         "traffic": None, "peak_source": which, "algorithmic_bytes_per_launch": alg,
                         "avg_launch_ms": cell_ms,
                         "active_limiter": "instruction issue, not HBM (see the cfg3 line and profiles/)"},
            "cpu_baseline": None,
            "detail": {"sharded_knn_normals": dict({"ms": ms_knn, "points_per_s": n_total / (ms_knn * 1e-3),
                                                    "first_call": run.first}, **phases, **(ver or {})),
                       "root_stage_ms": root_stage, "segment_peer_to_peer": int(res[3][0].peer_to_peer),
                       "root_stats": {kk: root_stats[kk] for kk in ("n_arcs", "n_seeds", "ift_levels", "rg_segments")},
                       "seeds": int(res[2])},
        }
        emit(line)
    run.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
