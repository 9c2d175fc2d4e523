"""Host side of the sharded path in the one-process-per-GPU mode (torchrun): slice bookkeeping, the
hand-over of the NCCL unique id, and zero-copy torch views of the device-resident results.

Everything that touches the data — partition, owner routing, halo selection, the NCCL exchange, the
k-NN / normals kernels, the row export, the gather to the IFT rank — lives in libiftwt.so
(csrc/shard.cu: iftwt_comm_*, iftwt_sharded_knn_normals, iftwt_sharded_segment).  `torch.distributed`
is the side channel for 128 bytes and nothing else.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import api


def slice_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of rank's slice: consecutive global-id ranges, ascending with the rank, covering
    0 .. n_total (what iftwt_shard_in asks for)."""
    return rank * n_total // world, (rank + 1) * n_total // world


def broadcast_id(uid: bytes | None, src: int = 0, group=None) -> bytes:
    """Rank `src` passes the id it got from Comm.unique_id(); every rank returns it.  Works on any
    backend (a CPU tensor under gloo, a CUDA tensor under nccl)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert uid is not None
        return bytes(uid)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8)
    if dist.get_rank(group) == src:
        assert uid is not None and len(uid) == api.COMM_ID_BYTES
        t = torch.frombuffer(bytearray(uid), dtype=torch.uint8).clone()
    t = t.to(dev)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def make_comm(device: int, group=None) -> api.Comm:
    """The comm of this process (one GPU): ncclGetUniqueId on rank 0, ncclCommInitRank everywhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return api.Comm.all([device])
    uid = api.Comm.unique_id() if rank == 0 else None
    return api.Comm.rank(broadcast_id(uid, 0, group), world, rank, device)


def as_torch(out: api.ShardOut, device: int) -> dict:
    """Zero-copy torch views of one shard's result (valid until the comm's next call)."""
    dev = torch.device("cuda", device)
    res = {}
    for name, a in api.Comm.arrays(out).items():
        if a is None or out.n_owned == 0:
            res[name] = None
        else:
            res[name] = torch.as_tensor(a, device=dev)
    return res
