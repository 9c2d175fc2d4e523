"""The Morton-range sharded k-NN / normals / segmentation (csrc/shard.cu) on ONE GPU: a comm whose
shards time-share device 0 (iftwt_comm_create_all with the same device named several times) runs
every kernel of the path — partition histogram, owner table, routing + stable multi-split, row
export, gather + install on the root — with device-to-device copies where the multi-GPU mode has
grouped ncclSend / ncclRecv.  tests/test_gpu_multi.py repeats it over NCCL when two GPUs are there.

Bar: rows (ids, order, squared distances) and normals bit-identical to ONE ctx on the whole cloud,
every point owned exactly once, the owner of every point as the partition rule says, cost / label of
iftwt_sharded_segment bit-identical to iftwt_segment."""
import numpy as np
import pytest

from iftwt_rg_b200 import synth

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _single_gpu(pts, k):
    from iftwt_rg_b200 import api
    with api.Context(0) as ctx:
        ctx.set_cloud(pts)
        idx, d2 = ctx.knn(k)
        nrm = ctx.normals(k)
    return idx, d2, nrm


def _run_sharded(pts, k, world, radius=0.0, bounds=None):
    """Returns per-shard dicts of numpy arrays + the ShardOut stats."""
    torch = _torch()
    from iftwt_rg_b200 import api, sharded
    n = len(pts)
    dev = torch.device("cuda", 0)
    bounds = bounds or [sharded.slice_bounds(n, world, r) for r in range(world)]
    slabs = [torch.from_numpy(np.ascontiguousarray(pts[lo:hi])).to(dev) for lo, hi in bounds]
    torch.cuda.synchronize()
    def collect(outs):
        res = []
        for o in outs:
            v = sharded.as_torch(o, 0)
            res.append({name: (None if t is None else t.cpu().numpy()) for name, t in v.items()} | {"o": o.as_dict()})
        return res

    with api.Comm.all([0] * world) as cm:
        assert cm.info() == {"nranks": world, "n_local": world, "first_local_rank": 0}
        sl = [(t.data_ptr() if t.numel() else 0, t.shape[0], lo) for t, (lo, hi) in zip(slabs, bounds)]
        # the first call on a comm sizes the receive buffers and moves the points with copies (NCCL between GPUs);
        # from the second call on the pack kernel writes straight into the peers' receive buffers.  Same rows.
        first = collect(cm.sharded_knn_normals(sl, k, True, radius))
        assert all(r["o"]["peer_to_peer"] == 0 for r in first)
        res = collect(cm.sharded_knn_normals(sl, k, True, radius))
        for a, b in zip(first, res):
            assert a["o"]["n_owned"] == b["o"]["n_owned"] and a["o"]["n_halo"] == b["o"]["n_halo"]
            for name in ("gid", "points", "knn", "d2", "normals"):
                if a[name] is not None:
                    assert np.array_equal(a[name].view(np.uint32), b[name].view(np.uint32)), name
            if b["o"]["attempts"] == 1:
                assert b["o"]["peer_to_peer"] == 1
    return res


def _check_rows(pts, k, res, ridx, rd2, rn):
    n = len(pts)
    seen = np.zeros(n, np.int32)
    for r in res:
        if r["o"]["n_owned"] == 0:
            continue
        gid = r["gid"]
        assert np.all(np.diff(gid.astype(np.int64)) > 0)               # ascending global ids
        seen[gid] += 1
        assert np.array_equal(r["points"], pts[gid])
        assert np.array_equal(r["knn"], ridx[gid]), "rows differ from the single-GPU rows"
        assert np.array_equal(r["d2"].view(np.uint32), rd2[gid].view(np.uint32))
        assert np.array_equal(r["normals"].view(np.uint32), rn[gid].view(np.uint32))
    assert np.all(seen == 1), "every point is owned exactly once"


@pytest.mark.parametrize("scene,n,k,world", [("facade", 300_000, 16, 2), ("primitives", 200_000, 16, 3),
                                             ("statue", 100_000, 30, 4), ("facade", 250_000, 32, 8)])
def test_sharded_rows_equal_single_gpu(scene, n, k, world):
    pts = getattr(synth, "scene_" + scene)(n, seed=0x51A + world)
    ridx, rd2, rn = _single_gpu(pts, k)
    res = _run_sharded(pts, k, world)
    _check_rows(pts, k, res, ridx, rd2, rn)
    o = res[0]["o"]
    assert o["attempts"] == 1 and o["n_total"] == n
    owned = [r["o"]["n_owned"] for r in res]
    halo = [r["o"]["n_halo"] for r in res]
    assert min(owned) > 0.5 * n / world and max(owned) < 1.6 * n / world     # balanced Morton ranges
    # a shell, not a replica (eight shards of a 250 k-point cloud at k = 32 are mostly shell: 31 k points each,
    # a halo 2 x the 32nd-neighbour distance thick)
    assert 0 < sum(halo) < (n if world <= 4 else 3 * n)
    # the reported largest k-th distance is the true one
    assert abs(o["kth_distance_max"] - float(np.sqrt(rd2[:, k - 1].max()))) <= 1e-6 * o["kth_distance_max"]
    assert o["halo_radius"] >= o["kth_distance_max"]


def test_one_shard_is_the_single_gpu_path():
    """A comm of one rank (bench.py --workload cfg4 at N = 1): no peer, no halo, same rows."""
    pts = synth.scene_primitives(120_000, seed=0x11)
    k = 12
    ridx, rd2, rn = _single_gpu(pts, k)
    res = _run_sharded(pts, k, 1)
    _check_rows(pts, k, res, ridx, rd2, rn)
    assert res[0]["o"]["n_owned"] == len(pts) and res[0]["o"]["n_halo"] == 0 and res[0]["o"]["attempts"] == 1


def test_host_slices_through_iftwt_comm_upload():
    """A host without CUDA of its own hands over host memory (iftwt_comm_upload); rows fetched back with plain
    cudaMemcpy-free means are out of this test's scope — the segmentation result is compared instead."""
    from iftwt_rg_b200 import api, sharded
    pts = synth.scene_facade(150_000, seed=0x33)
    n, k = len(pts), 16
    rg = api.RgParams(50, 10000, k, 0.1, 1.0)
    with api.Context(0) as ctx:
        ctx.set_cloud(pts)
        rcost, rlabel, rns = ctx.segment(k, k, rg)
    with api.Comm.all([0, 0, 0]) as cm:
        sl = [cm.upload(r, pts[lo:hi], lo) for r, (lo, hi) in enumerate(sharded.slice_bounds(n, 3, r) for r in range(3))]
        cost, label, ns, _ = cm.sharded_segment(sl, k, rg, root=0, n_total=n)
        assert ns == rns and np.array_equal(cost, rcost) and np.array_equal(label, rlabel)
        with pytest.raises(api.IftwtError):
            cm.upload(3, pts[:10], 0)


def test_owner_follows_the_partition_rule():
    """Numpy restatement of the rule (shard.cu phases B, C): coarse cells of side s over the global box,
    Morton key (x most significant), bins = the key's top <= 18 bits, bin b goes to rank
    floor(points_before_b * W / N)."""
    pts = synth.scene_facade(120_000, seed=0x77)
    k, W, s = 8, 4, 0.35
    res = _run_sharded(pts, k, W, radius=s)
    assert res[0]["o"]["attempts"] == 1 and res[0]["o"]["halo_radius"] == s
    lo = pts[:, :3].min(axis=0).astype(np.float64)
    hi = pts[:, :3].max(axis=0).astype(np.float64)
    dims = np.floor((hi - lo) / s).astype(np.int64) + 1
    cell = np.floor((pts[:, :3].astype(np.float64) - lo) * (1.0 / s)).astype(np.int64)
    cell = np.minimum(np.maximum(cell, 0), dims - 1)
    bits = 1
    while (1 << bits) < dims.max():
        bits += 1

    def spread(v):
        out = np.zeros_like(v)
        for b in range(bits):
            out |= ((v >> b) & 1) << (3 * b)
        return out
    key = (spread(cell[:, 0]) << 2) | (spread(cell[:, 1]) << 1) | spread(cell[:, 2])
    shift = max(0, 3 * bits - 18)
    b = key >> shift
    hist = np.bincount(b, minlength=1 << (3 * bits - shift)).astype(np.int64)
    before = np.cumsum(hist) - hist
    owner_of_bin = np.minimum(before * W // len(pts), W - 1)
    owner = owner_of_bin[b]
    for r, sh in enumerate(res):
        assert np.array_equal(np.sort(sh["gid"]), np.flatnonzero(owner == r).astype(np.uint32)), r
    # halo = exactly the non-owned points whose 3x3x3 coarse block touches a cell of the rank's range
    # (cells without points have an owner too: it follows from the key)
    def owner_at(c):
        kk = (int(spread(np.array([c[0]]))[0]) << 2) | (int(spread(np.array([c[1]]))[0]) << 1) | int(spread(np.array([c[2]]))[0])
        return int(owner_of_bin[kk >> shift])
    expect_halo = np.zeros(W, np.int64)
    got_halo = np.array([sh["o"]["n_halo"] for sh in res])
    # exact count on the whole cloud through the unique cells (cheap: few thousand cells)
    ucell, inv_c, cnt_c = np.unique(cell, axis=0, return_inverse=True, return_counts=True)
    for ci, c in enumerate(ucell):
        dests = set()
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    q = (c[0] + dx, c[1] + dy, c[2] + dz)
                    if all(0 <= q[a] < dims[a] for a in range(3)):
                        dests.add(owner_at(q))
        dests.discard(owner_at(c))
        for d in dests:
            expect_halo[d] += cnt_c[ci]
    assert np.array_equal(got_halo, expect_halo)


def test_too_small_radius_is_detected_and_widened():
    pts = synth.scene_primitives(150_000, seed=0x99)
    k = 16
    ridx, rd2, rn = _single_gpu(pts, k)
    res = _run_sharded(pts, k, 3, radius=0.004)      # the 16th neighbour is ~0.05 away
    _check_rows(pts, k, res, ridx, rd2, rn)
    assert res[0]["o"]["attempts"] > 1
    assert res[0]["o"]["halo_radius"] >= res[0]["o"]["kth_distance_max"]


def test_tiny_clouds_empty_slices_and_fewer_points_than_k():
    k = 8
    # fewer points than k: rows are -1 / +inf padded and still equal to one GPU's
    pts = synth.scene_primitives(5, seed=3)
    ridx, rd2, rn = _single_gpu(pts, k)
    res = _run_sharded(pts, k, 2)
    seen = np.zeros(5, np.int32)
    for r in res:
        if r["o"]["n_owned"]:
            seen[r["gid"]] += 1
            assert np.array_equal(r["knn"], ridx[r["gid"]])
            assert np.array_equal(r["d2"].view(np.uint32), rd2[r["gid"]].view(np.uint32))
    assert np.all(seen == 1)
    # a rank that contributes nothing, another that contributes almost nothing
    pts = synth.scene_statue(40_000, seed=5)
    ridx, rd2, rn = _single_gpu(pts, k)
    res = _run_sharded(pts, k, 3, bounds=[(0, 0), (0, 7), (7, 40_000)])
    _check_rows(pts, k, res, ridx, rd2, rn)
    # far outliers: the estimated radius cannot cover them, the check must
    far = pts.copy()
    far[1:4, :3] += 40.0      # indices the evenly spaced sample of the estimate does not visit
    ridx, rd2, rn = _single_gpu(far, k)
    res = _run_sharded(far, k, 2)
    _check_rows(far, k, res, ridx, rd2, rn)
    assert res[0]["o"]["attempts"] > 1


def test_bad_slices_are_refused():
    torch = _torch()
    from iftwt_rg_b200 import api
    pts = torch.from_numpy(synth.scene_primitives(1000, seed=1)).cuda()
    with api.Comm.all([0, 0]) as cm:
        with pytest.raises(api.IftwtError) as e:   # gap between the slices
            cm.sharded_knn_normals([(pts.data_ptr(), 500, 0), (pts[500:].data_ptr(), 500, 600)], 4)
        assert e.value.code == -1 and "contiguous" in str(e.value)
        bad = pts.clone()
        bad[3, 1] = float("nan")
        with pytest.raises(api.IftwtError) as e:
            cm.sharded_knn_normals([(bad.data_ptr(), 500, 0), (bad[500:].data_ptr(), 500, 500)], 4)
        assert e.value.code == -4
        # and the comm still works afterwards
        outs = cm.sharded_knn_normals([(pts.data_ptr(), 500, 0), (pts[500:].data_ptr(), 500, 500)], 4)
        assert sum(o.n_owned for o in outs) == 1000
    with pytest.raises(api.IftwtError):
        api.Comm.all([0] * 9)


@pytest.mark.parametrize("world,gradient", [(2, False), (4, True)])
def test_sharded_segment_equals_single_gpu_segment(world, gradient):
    """BASELINE config 4's shape at test size: k-NN + normals sharded, rows gathered to the root, graph /
    seeds / IFT there — cost and label bit-identical to iftwt_segment on one GPU."""
    torch = _torch()
    from iftwt_rg_b200 import api, sharded
    pts = synth.scene_facade(200_000, seed=0x4F)
    n, k, theta = len(pts), 16, 0.1
    rg = api.RgParams(50, 10000, k, theta, 1.0)
    with api.Context(0) as ctx:
        ctx.set_cloud(pts)
        rcost, rlabel, rns = ctx.segment(k, k, rg, gradient_weights=gradient)
        roff, radj = ctx.graph_csr()
    assert rns >= 10
    dev = torch.device("cuda", 0)
    bounds = [sharded.slice_bounds(n, world, r) for r in range(world)]
    slabs = [torch.from_numpy(np.ascontiguousarray(pts[lo:hi])).to(dev) for lo, hi in bounds]
    torch.cuda.synchronize()
    root = world - 1
    with api.Comm.all([0] * world) as cm:
        sl = [(t.data_ptr(), t.shape[0], lo) for t, (lo, hi) in zip(slabs, bounds)]
        for call in range(3):
            # call 0: exchange and gather by copies (they size the receive buffers and the root's buffers); from
            # call 1 on the pack kernel writes into the peers' receive buffers and the export kernels write the rows
            # straight into the root's buffers
            cost, label, ns, outs = cm.sharded_segment(sl, k, rg, root=root, n_total=n, gradient_weights=gradient)
            assert ns == rns, call
            assert np.array_equal(cost, rcost) and np.array_equal(label, rlabel), call
            assert all(o.peer_to_peer == (0, 3, 3)[call] for o in outs), (call, [o.peer_to_peer for o in outs])
        # the root's ctx now holds the whole cloud and its graph: the follow-up calls of the ABI apply
        rctx = cm.ctx(root)
        rctx.n = n
        off, adj = rctx.graph_csr()
        assert np.array_equal(off, roff) and np.array_equal(adj, radj)
        assert np.array_equal(rctx.get_cloud(), pts)
