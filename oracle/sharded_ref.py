"""Reference restatement, in plain torch + torch.distributed, of the Morton-range sharded k-NN that
the product implements in CUDA + NCCL (iftwt_rg_b200/csrc/shard.cu, iftwt_sharded_knn_normals).

TEST INFRASTRUCTURE (it was the round-1 product path; the product now does all of this inside
libiftwt.so).  It runs on CPU tensors with the gloo backend, which is how tests/test_sharded_gloo.py
covers the world_size > 1 logic without a GPU: partition by contiguous ranges of a coarse Morton key,
owner redistribution, halo = the points of the 26 coarse cells around an owned cell, local k-NN on
owned + halo in global-id order, exactness check (k-th distance <= halo radius, redo with 2 s).

  1. global bounding box                                 all_reduce MIN / MAX
  2. coarse grid of side `s`, histogram of the Morton key's top bits      all_reduce SUM
     -> contiguous, balanced Morton ranges, one per rank
  3. every point goes to the rank that owns its range    all_to_all
  4. halo copies                                         all_to_all
  5. local k-NN through the `local_knn` hook (the CPU oracle in the tests), rows of the owned points
     translated to global ids.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist


def _spread10(v: torch.Tensor) -> torch.Tensor:
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def morton30(ix: torch.Tensor, iy: torch.Tensor, iz: torch.Tensor) -> torch.Tensor:
    """x is the most significant axis bit of every triple (same convention as the kernels)."""
    return (_spread10(ix) << 2) | (_spread10(iy) << 1) | _spread10(iz)


@dataclass
class Partition:
    origin: torch.Tensor          # (3,) float64 global bbox min
    cell: float                   # coarse cell side = halo radius
    dims: tuple                   # coarse grid dims
    shift: int                    # histogram bin = key >> shift
    owner_of_bin: torch.Tensor    # (nbins,) int64

    def cells(self, xyz: torch.Tensor):
        c = torch.floor((xyz.double() - self.origin) / self.cell).long()
        for a in range(3):
            c[:, a].clamp_(0, self.dims[a] - 1)
        return c

    def owner_of_cells(self, c: torch.Tensor) -> torch.Tensor:
        key = morton30(c[:, 0], c[:, 1], c[:, 2])
        return self.owner_of_bin[key >> self.shift]


def _all_reduce(t, op, group):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def make_partition(xyz: torch.Tensor, cell: float, world: int, group=None) -> Partition:
    dev = xyz.device
    lo = xyz.min(dim=0).values.float() if len(xyz) else torch.full((3,), float("inf"), device=dev)
    hi = xyz.max(dim=0).values.float() if len(xyz) else torch.full((3,), float("-inf"), device=dev)
    _all_reduce(lo, dist.ReduceOp.MIN, group)
    _all_reduce(hi, dist.ReduceOp.MAX, group)
    ext = (hi - lo).double().clamp_min(1e-30)
    # at most 1024 cells per axis: a coarser grid only makes the halo thicker, never wrong
    cell = max(float(cell), float(ext.max()) / 1023.0)
    dims = tuple(int(math.floor(float(e) / cell)) + 1 for e in ext)
    bits = max(1, max(int(math.ceil(math.log2(max(d, 2)))) for d in dims))
    shift = max(0, 3 * bits - 18)
    nbins = 1 << (3 * bits - shift)
    part = Partition(lo.double(), cell, dims, shift, torch.zeros(nbins, dtype=torch.long, device=dev))
    c = part.cells(xyz)
    key = morton30(c[:, 0], c[:, 1], c[:, 2])
    hist = torch.bincount(key >> shift, minlength=nbins).to(torch.int64)
    _all_reduce(hist, dist.ReduceOp.SUM, group)
    cum = torch.cumsum(hist, 0)
    total = int(cum[-1])
    # bin b goes to rank floor(points_before_b * world / total): contiguous Morton ranges
    before = cum - hist
    part.owner_of_bin = torch.clamp((before * world) // max(total, 1), max=world - 1)
    return part


def _exchange(payloads, dest: torch.Tensor, world: int, group=None):
    """Send row i of every payload to rank dest[i]; returns the received payloads (ordered by
    source rank, then by the sender's order) ."""
    if world == 1:
        return [p.clone() for p in payloads]
    order = torch.argsort(dest, stable=True)
    counts = torch.bincount(dest, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    out = []
    send_split = counts.tolist()
    recv_split = recv_counts.tolist()
    for p in payloads:
        src = p[order].contiguous()
        dst = torch.empty((sum(recv_split),) + tuple(p.shape[1:]), dtype=p.dtype, device=p.device)
        dist.all_to_all_single(dst, src, output_split_sizes=recv_split, input_split_sizes=send_split, group=group)
        out.append(dst)
    return out


def halo_destinations(part: Partition, xyz: torch.Tensor, rank: int, world: int):
    """(point index, destination rank) pairs: ranks other than `rank` owning one of the 26
    coarse cells around each point's own cell."""
    c = part.cells(xyz)
    m = len(xyz)
    need = torch.zeros((m, world), dtype=torch.bool, device=xyz.device)
    dims = torch.tensor(part.dims, device=xyz.device)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                n = c + torch.tensor([dx, dy, dz], device=xyz.device)
                ok = ((n >= 0) & (n < dims)).all(dim=1)
                own = part.owner_of_cells(torch.where(ok[:, None], n, c))
                need[torch.arange(m, device=xyz.device)[ok], own[ok]] = True
    need[:, rank] = False
    idx, dst = torch.nonzero(need, as_tuple=True)
    return idx, dst


@dataclass
class ShardResult:
    gid: torch.Tensor            # (m,) global ids of the owned points
    pts: torch.Tensor            # (m, 4) owned points
    knn_gid: torch.Tensor        # (m, k) global ids of the neighbours, canonical order
    d2: torch.Tensor             # (m, k)
    normals: torch.Tensor | None  # (m, 4) {nx,ny,nz,curvature}
    stats: dict = field(default_factory=dict)


def sharded_knn(local_pts: torch.Tensor, gid0: int, k: int, local_knn, halo_radius: float,
                want_normals: bool = True, group=None, max_retries: int = 10) -> ShardResult:
    """local_pts: (m, 4) float32 {x,y,z,intensity} slice held by this rank, global ids
    gid0 .. gid0+m-1.  Returns the exact global k-NN rows (and normals) of the points this rank
    ends up owning."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = local_pts.device
    gid = gid0 + torch.arange(local_pts.shape[0], device=dev, dtype=torch.int64)
    s = float(halo_radius)
    for attempt in range(max_retries + 1):
        part = make_partition(local_pts[:, :3], s, world, group)
        owner = part.owner_of_cells(part.cells(local_pts[:, :3]))
        own_pts, own_gid = _exchange([local_pts, gid], owner, world, group)
        h_idx, h_dst = halo_destinations(part, own_pts[:, :3], rank, world)
        halo_pts, halo_gid = _exchange([own_pts[h_idx], own_gid[h_idx]], h_dst, world, group)
        m = own_pts.shape[0]
        cloud = torch.cat([own_pts, halo_pts], 0).contiguous()
        cloud_gid = torch.cat([own_gid, halo_gid], 0)
        # ties are broken by index: make the LOCAL index order agree with the GLOBAL id order
        order = torch.argsort(cloud_gid, stable=True)
        inv_order = torch.empty_like(order)
        inv_order[order] = torch.arange(len(order), device=dev)
        cloud_s = cloud[order].contiguous()
        gid_s = cloud_gid[order]
        idx, d2, nrm = local_knn(cloud_s, k, want_normals)
        rows = inv_order[:m]                                   # local position of each owned point
        idx_o = idx[rows].long()
        d2_o = d2[rows]
        knn_gid = torch.where(idx_o >= 0, gid_s[idx_o.clamp_min(0)], torch.full_like(idx_o, -1))
        # exactness check: k-th distance inside the halo radius (inf rows = fewer than k points overall)
        kth = d2_o[:, k - 1]
        finite = torch.isfinite(kth)
        # an infinite k-th distance only says that THIS rank's owned + halo set holds fewer than k points:
        # the row is exact only if that set is the whole cloud
        total = torch.tensor([local_pts.shape[0]], device=dev, dtype=torch.int64)
        _all_reduce(total, dist.ReduceOp.SUM, group)
        short = (~finite) if cloud.shape[0] < int(total) else torch.zeros_like(finite)
        viol = torch.tensor([int((((kth > part.cell * part.cell) & finite) | short).sum())], device=dev,
                            dtype=torch.int64)
        _all_reduce(viol, dist.ReduceOp.SUM, group)
        if int(viol) == 0:
            stats = {"rank": rank, "owned": int(m), "halo": int(halo_pts.shape[0]), "halo_radius": part.cell,
                     "attempts": attempt + 1, "violations": int(viol)}
            return ShardResult(own_gid, own_pts, knn_gid, d2_o, None if nrm is None else nrm[rows], stats)
        s = 2.0 * part.cell
    raise RuntimeError(f"sharded k-NN: halo radius still too small after {max_retries + 1} exchanges "
                       f"({int(viol)} owned points with a k-th neighbour beyond {part.cell:g})")


def gather_rows_to_rank0(res: ShardResult, n_total: int, k: int, group=None):
    """Rows of every rank, scattered to global-id order on rank 0 (config 4: the IFT runs on
    one GPU).  Returns (pts, knn_gid int32, d2) on rank 0, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = res.gid.device
    dest = torch.zeros(len(res.gid), dtype=torch.int64, device=dev)
    gid, pts, knn, d2 = _exchange([res.gid, res.pts, res.knn_gid.int(), res.d2], dest, world, group)
    if rank != 0:
        return None
    out_pts = torch.empty((n_total, 4), dtype=torch.float32, device=dev)
    out_idx = torch.empty((n_total, k), dtype=torch.int32, device=dev)
    out_d2 = torch.empty((n_total, k), dtype=torch.float32, device=dev)
    out_pts[gid] = pts
    out_idx[gid] = knn
    out_d2[gid] = d2
    return out_pts, out_idx, out_d2


def estimate_halo_radius(local_pts: np.ndarray, k: int, world: int, sample: int = 256, seed: int = 0) -> float:
    """A rank's slice is a uniform 1/world subsample of the cloud, so its ceil(k/world)+1-th
    neighbour distance estimates the cloud's k-th.  Brute force on a small sample; the result
    only sizes the halo (exactness is verified afterwards)."""
    rng = np.random.default_rng(seed)
    n = local_pts.shape[0]
    q = local_pts[rng.integers(0, n, size=min(sample, n)), :3].astype(np.float64)
    base = local_pts[rng.integers(0, n, size=min(50_000, n)), :3].astype(np.float64)
    kk = min(len(base) - 1, int(math.ceil(k / world)) + 1)
    scale = (len(base) / n) ** 0.5  # the base is itself a subsample of the slice (surface: r ~ density^-1/2)
    d = ((q[:, None, :] - base[None, :, :]) ** 2).sum(-1)
    kth = np.sqrt(np.partition(d, kk, axis=1)[:, kk])
    return float(np.quantile(kth, 0.99) * scale * 1.5)
