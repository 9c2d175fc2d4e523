"""Multi-GPU paths on one node (skipped unless >= 2 GPUs are visible): the Morton-range
sharded k-NN with NCCL halo exchange against the single-GPU rows, and bench.py under
torchrun (one cloud per rank, no data-path collective)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), script, *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    if out.returncode != 0:   # keep the whole story for the post-mortem (gpurun brings gpurun_out/ back)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "torchrun_failure.log"), "a") as f:
            f.write(" ".join(cmd) + "\n" + out.stdout + "\n" + out.stderr + "\n")
    assert out.returncode == 0, out.stderr[:1500] + "\n...\n" + out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:]
    return json.loads(lines[-1])


def test_sharded_two_gpus_nccl_single_process():
    """iftwt_comm_create_all over two DIFFERENT devices = ncclCommInitAll: the grouped ncclSend / ncclRecv
    exchange instead of the device-to-device copies of the single-GPU test mode; rows and the segmentation
    bit-identical to one GPU."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    import torch
    from iftwt_rg_b200 import api, sharded, synth
    pts = synth.scene_facade(400_000, seed=0x2A)
    n, k = len(pts), 16
    rg = api.RgParams(50, 10000, k, 0.1, 1.0)
    with api.Context(0) as ctx:
        ctx.set_cloud(pts)
        ridx, rd2 = ctx.knn(k)
        rn = ctx.normals(k)
        rcost, rlabel, rns = ctx.segment(k, k, rg)
    bounds = [sharded.slice_bounds(n, 2, r) for r in range(2)]
    slabs = [torch.from_numpy(np.ascontiguousarray(pts[lo:hi])).to(torch.device("cuda", r))
             for r, (lo, hi) in enumerate(bounds)]
    for r in range(2):
        torch.cuda.synchronize(r)
    with api.Comm.all([0, 1]) as cm:
        sl = [(t.data_ptr(), t.shape[0], lo) for t, (lo, hi) in zip(slabs, bounds)]
        cm.sharded_knn_normals(sl, k, True, 0.0)          # sizes the receive buffers, NCCL exchange
        outs = cm.sharded_knn_normals(sl, k, True, 0.0)   # pack kernel writes into the peer GPU over NVLink
        assert all(o.peer_to_peer == 1 for o in outs)
        seen = np.zeros(n, np.int32)
        for r, o in enumerate(outs):
            v = {name: t.cpu().numpy() for name, t in sharded.as_torch(o, r).items()}
            gid = v["gid"]
            seen[gid] += 1
            assert np.array_equal(v["knn"], ridx[gid]) and np.array_equal(v["d2"].view(np.uint32), rd2[gid].view(np.uint32))
            assert np.array_equal(v["normals"].view(np.uint32), rn[gid].view(np.uint32))
            assert o.halo_bytes_sent > 0 and o.attempts == 1
        assert np.all(seen == 1)
        for call in range(2):   # the second call's export kernels write the rows into GPU 1 directly
            cost, label, ns, so = cm.sharded_segment(sl, k, rg, root=1, n_total=n)
            assert ns == rns and np.array_equal(cost, rcost) and np.array_equal(label, rlabel), call
        assert all(o.peer_to_peer == 3 for o in so)


def test_bench_cfg4_two_gpus():
    """bench.py --workload cfg4 under torchrun: one process per GPU, ncclCommInitRank, iftwt_sharded_segment."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--workload", "cfg4", "--points",
                  "4000000")
    assert r["n_gpus"] == 2 and r["scaling"] == "strong" and r["value"] > 0 and r["gpu_launches"] > 0
    d = r["detail"]["sharded_knn_normals"]
    assert d["verify_rows_equal"] is True and d["verified_rows"] > 100000
    assert sum(d["owned_per_rank"]) == 4000000 and min(d["owned_per_rank"]) > 1200000
    assert 0 < sum(d["halo_per_rank"]) < 800000 and d["halo_bytes_over_nvlink"] > 0
    assert r["detail"]["seeds"] > 0 and r["e2e"]["value"] > 0
    assert d["peer_to_peer"] == 1 and r["detail"]["segment_peer_to_peer"] == 3   # cudaIpc between the two processes


def test_bench_two_gpus_default_line_carries_the_sharded_pass():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--points", "1000000",
                  "--sharded-per-gpu", "1500000", "--no-e2e")
    assert r["n_gpus"] == 2 and r["scaling"] == "weak" and r["value"] > 0 and r["gpu_launches"] > 0
    sh = r["detail"]["sharded"]
    assert "error" not in sh, sh
    assert sh["verify_rows_equal"] is True and sh["knn_normals_ms"] > 0 and sh["segment_seeds"] > 0


def test_bench_two_gpus_weak_scaling_line():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--workload", "cfg2",
                  "--no-cpu-baseline")
    assert r["n_gpus"] == 2 and r["scaling"] == "weak" and r["value"] > 0 and r["gpu_launches"] > 0
