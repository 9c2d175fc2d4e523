"""Specialised kernels against the general ones they stand in for.  The library reads its IFTWT_* switches at call
time, so one process can run both forms on the same ctx: the ballot-form region-growing classify (k = 16 / 32) against
the generic kernel, the hook kernel with and without its lockstep pointer jumps, the one-block-per-query nearest-point
kernel (the centroid snap of the seed stage) against the general k-NN path, and the batch entry point with pinned host
buffers (uploads / downloads in the lanes' own copy streams) against separate iftwt_segment calls."""
import contextlib
import os

import numpy as np
import pytest

from iftwt_rg_b200 import synth

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("k", [16, 32])
def test_ballot_classify_and_hook_variants_equal_generic(k):
    from iftwt_rg_b200 import api
    pts = synth.scene_facade(150_000, seed=0x72 + k)
    rg = api.RgParams(50, 1_000_000, k, 3.0 / 180.0 * np.pi, 1.0)
    with api.Context(0) as ctx:
        ctx.set_cloud(pts)
        ref = ctx.segment(k, k, rg)
        with env(IFTWT_RGA_GENERIC=1):
            ctx.set_cloud(pts)
            gen = ctx.segment(k, k, rg)
        with env(IFTWT_HOOK_JUMPS=0):
            ctx.set_cloud(pts)
            nojump = ctx.segment(k, k, rg)
        with env(IFTWT_RGA_VARIANT=1):
            ctx.set_cloud(pts)
            var1 = ctx.segment(k, k, rg)
    assert ref[2] >= 1, "the scene must yield seeds for the comparison to mean something"
    for other in (gen, nojump, var1):
        assert other[2] == ref[2]
        assert np.array_equal(other[0], ref[0]) and np.array_equal(other[1], ref[1])


def test_nearest_block_kernel_equals_general_path():
    from iftwt_rg_b200 import api
    pts = synth.scene_primitives(120_000, seed=0x71)
    lo, hi = pts[:, :3].min(0), pts[:, :3].max(0)
    rng = np.random.default_rng(5)
    near = pts[rng.integers(0, len(pts), 300), :3] + rng.normal(0, 0.05, (300, 3)).astype(np.float32)
    inside = (lo + (hi - lo) * rng.random((200, 3))).astype(np.float32)       # off the surfaces: the block grows
    outside = (lo - 3 * (hi - lo) + 7 * (hi - lo) * rng.random((60, 3))).astype(np.float32)
    exact = pts[:40, :3]
    far = np.array([[1e3, 1e3, 1e3], [-1e3, 0, 0], [0, 0, 5e2]], np.float32)
    q = np.ascontiguousarray(np.concatenate([near, inside, outside, exact, far]).astype(np.float32))
    assert len(q) <= 2048
    with api.Context(0) as ctx:
        ctx.set_cloud(pts)
        with env(IFTWT_NEAREST_BLOCK=0):
            i0, d0 = ctx.knn_query(q, 1)
        i1, d1 = ctx.knn_query(q, 1)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    # and both are the nearest point: binary64 distances, ties and rounding aside
    d64 = ((pts[None, :, :3].astype(np.float64) - q[-103::7, None, :].astype(np.float64)) ** 2).sum(-1)
    got = d64[np.arange(d64.shape[0]), i1[-103::7, 0]]
    assert np.all(got <= d64.min(1) * (1 + 1e-5) + 1e-12)


def test_batch_with_pinned_host_buffers_equals_separate_calls():
    import torch
    from iftwt_rg_b200 import api
    k = 16
    rg = api.RgParams(50, 10000, k, 3.0 / 180.0 * np.pi, 1.0)
    sizes = [90_000, 40_000, 120_000, 60_000, 75_000]
    keep = [torch.from_numpy(synth.scene_facade(n, seed=0xC0 + i)).pin_memory() for i, n in enumerate(sizes)]
    clouds = [t.numpy() for t in keep]
    ref = []
    with api.Context(0) as ctx:
        for c in clouds:
            ctx.set_cloud(c)
            ref.append(ctx.segment(k, k, rg))
    assert max(r[2] for r in ref) >= 5
    with api.Batch(0, 2) as bt:
        for _ in range(2):  # the second call reuses the lanes' staging and result buffers
            ct = [torch.empty(n, dtype=torch.int32).pin_memory() for n in sizes]
            lt = [torch.empty(n, dtype=torch.int32).pin_memory() for n in sizes]
            cost = [t.numpy().view(np.uint32) for t in ct]
            label = [t.numpy().view(np.uint32) for t in lt]
            ns = bt.segment(clouds, k, rg, cost=cost, label=label)
            for i, r in enumerate(ref):
                assert ns[i] == r[2], i
                assert np.array_equal(cost[i], r[0]) and np.array_equal(label[i], r[1]), i
