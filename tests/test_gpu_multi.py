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
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:]
    return json.loads(lines[-1])


def test_sharded_knn_two_gpus_equals_single_gpu():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "tools/sharded_run.py", "--points", "2000000", "--k", "16", "--verify")
    assert r["verify_rows_equal"] is True
    assert sum(r["owned_per_rank"]) == 2000000 and min(r["owned_per_rank"]) > 600000
    assert 0 < sum(r["halo_per_rank"]) < 400000
    assert r["stats"]["n_seeds"] > 0


def test_bench_two_gpus_weak_scaling_line():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--workload", "cfg2",
                  "--no-cpu-baseline")
    assert r["n_gpus"] == 2 and r["scaling"] == "weak" and r["value"] > 0 and r["gpu_launches"] > 0
