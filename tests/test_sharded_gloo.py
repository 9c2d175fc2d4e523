"""The world_size > 1 logic of the Morton-range sharded k-NN on CPU: world_size 2 and 3, gloo backend.

The product does the partition / halo / exchange in CUDA + NCCL (csrc/shard.cu); what can run
without a GPU is (a) its reference restatement in plain torch.distributed (oracle/sharded_ref.py),
the oracle standing in for the GPU kernels — the rows every rank returns must be the exact global
k-NN (ids, order and distances) of the points it owns, and the ranks together must own every point
exactly once — and (b) the host side of the product's process-per-GPU mode
(iftwt_rg_b200/sharded.py: slice bounds, hand-over of the communicator id)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iftwt_rg_b200 import synth
from oracle import sharded_ref as sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_hook(cloud: torch.Tensor, k: int, want_normals: bool):
    from oracle import pyoracle as orc
    pts = cloud.numpy()
    idx, d2 = orc.knn_self(pts, k)
    nrm = orc.normals(pts, idx) if want_normals else None
    return torch.from_numpy(idx), torch.from_numpy(d2), None if nrm is None else torch.from_numpy(nrm)


def _worker(rank, world, port, n, k, radius, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts = synth.scene_primitives(n, seed=0x77)
        lo, hi = rank * n // world, (rank + 1) * n // world     # contiguous slice of the generation order
        res = sharded.sharded_knn(torch.from_numpy(pts[lo:hi].copy()), lo, k, _oracle_hook, radius)
        g = sharded.gather_rows_to_rank0(res, n, k)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), gid=res.gid.numpy(), knn=res.knn_gid.numpy(),
                 d2=res.d2.numpy(), nrm=res.normals.numpy(), halo=res.stats["halo"], owned=res.stats["owned"],
                 attempts=res.stats["attempts"], violations=res.stats["violations"])
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), pts=g[0].numpy(), idx=g[1].numpy(), d2=g[2].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,radius", [(2, 0.2), (3, 0.2), (2, 0.01)])
def test_sharded_knn_equals_global_knn(tmp_path, world, radius):
    from oracle import pyoracle as orc
    n, k = 30_000, 12
    mp.spawn(_worker, args=(world, _free_port(), n, k, radius, str(tmp_path)), nprocs=world, join=True)
    pts = synth.scene_primitives(n, seed=0x77)
    ridx, rd2 = orc.knn_self(pts, k)
    rn = orc.normals(pts, ridx)
    seen = np.zeros(n, np.int32)
    halo_total = 0
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"r{r}.npz"))
        gid = z["gid"]
        seen[gid] += 1
        assert np.array_equal(z["knn"], ridx[gid].astype(np.int64)), r
        assert np.array_equal(z["d2"], rd2[gid]), r
        assert np.array_equal(z["nrm"].view(np.uint32), rn[gid].view(np.uint32)), r
        assert int(z["violations"]) == 0
        halo_total += int(z["halo"])
        if radius < 0.05:
            assert int(z["attempts"]) > 1          # the too-small halo was detected and widened
    assert np.all(seen == 1)
    assert 0 < halo_total < n                        # a thin shell, not a replica
    g = np.load(os.path.join(tmp_path, "gathered.npz"))
    assert np.array_equal(g["pts"], pts) and np.array_equal(g["idx"], ridx) and np.array_equal(g["d2"], rd2)


def test_partition_is_contiguous_in_morton_order_and_balanced():
    pts = torch.from_numpy(synth.scene_facade(50_000))
    part = sharded.make_partition(pts[:, :3], 0.25, 4)
    ob = part.owner_of_bin
    assert torch.all(ob[1:] >= ob[:-1])                                  # contiguous Morton ranges
    owner = part.owner_of_cells(part.cells(pts[:, :3]))
    counts = torch.bincount(owner, minlength=4)
    assert counts.min() > 0.5 * 50_000 / 4 and counts.max() < 1.6 * 50_000 / 4


def _id_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from iftwt_rg_b200 import sharded as host
        uid = bytes(range(128)) if rank == 0 else None   # stands in for ncclGetUniqueId's 128 bytes
        got = host.broadcast_id(uid)
        with open(os.path.join(out_dir, f"id{rank}.bin"), "wb") as f:
            f.write(got)
    finally:
        dist.destroy_process_group()


def test_host_side_of_the_product_path(tmp_path):
    """iftwt_rg_b200/sharded.py: the slices are the contiguous ascending id ranges iftwt_shard_in asks for,
    and the communicator id reaches every rank intact (world 2, gloo)."""
    from iftwt_rg_b200 import sharded as host
    for n, world in ((10, 3), (100_000_001, 8), (5, 8)):
        b = [host.slice_bounds(n, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
    mp.spawn(_id_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(tmp_path, f"id{r}.bin"), "rb").read() == bytes(range(128))
