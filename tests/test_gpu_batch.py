"""iftwt_batch_segment (BASELINE config 5's entry point): several lanes — ctx + stream + host thread —
pull clouds from one queue.  Cost, label and seed count of every cloud are bit-identical to a
separate iftwt_segment call; an error in one cloud is reported, not swallowed."""
import numpy as np
import pytest

from iftwt_rg_b200 import synth

pytestmark = pytest.mark.gpu


def test_batch_equals_separate_segment_calls():
    import torch
    from iftwt_rg_b200 import api
    k = 16
    rg = api.RgParams(50, 10000, k, 3.0 / 180.0 * np.pi, 1.0)
    clouds = [synth.scene_facade(60_000 + 7_000 * i, seed=0xB0 + i) for i in range(4)] + \
             [synth.scene_primitives(80_000, seed=0xB9), synth.scene_statue(50_000, seed=0xBA),
              synth.scene_primitives(300, seed=0xBB)]
    ref = []
    with api.Context(0) as ctx:
        for c in clouds:
            ctx.set_cloud(c)
            ref.append(ctx.segment(k, k, rg))
    assert max(r[2] for r in ref) >= 10
    for lanes in (1, 3):
        with api.Batch(0, lanes) as bt:
            cost = [np.empty(len(c), np.uint32) for c in clouds]
            label = [np.empty(len(c), np.uint32) for c in clouds]
            ns = bt.segment(clouds, k, rg, cost=cost, label=label)
            for i, r in enumerate(ref):
                assert ns[i] == r[2], (lanes, i)
                assert np.array_equal(cost[i], r[0]) and np.array_equal(label[i], r[1]), (lanes, i)
            # device-resident clouds, results left on the lanes: only the seed counts come back
            dev = [torch.from_numpy(c).cuda() for c in clouds]
            torch.cuda.synchronize()
            ns2 = bt.segment([(d.data_ptr(), d.shape[0]) for d in dev], k, rg, on_device=True)
            assert ns2 == [r[2] for r in ref]
            assert sum(bt.ctx(l).stats()["kernel_launches"] for l in range(lanes)) > 0


def test_batch_reports_the_failing_cloud():
    from iftwt_rg_b200 import api
    k = 8
    rg = api.RgParams(50, 10000, k, 0.1, 1.0)
    good = synth.scene_primitives(20_000, seed=1)
    bad = good.copy()
    bad[17, 2] = np.inf
    with api.Batch(0, 2) as bt:
        with pytest.raises(api.IftwtError) as e:
            bt.segment([good, bad, good], k, rg)
        assert e.value.code == -4 and "cloud 1" in str(e.value)
        assert bt.segment([good, good], k, rg)[0] == bt.segment([good], k, rg)[0]   # still usable
    with pytest.raises(api.IftwtError):
        api.Batch(0, 0)
