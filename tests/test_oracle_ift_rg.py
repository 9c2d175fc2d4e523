"""Pins the oracle's IFT (ift.hpp:170-250) and region-growing
(flat_point.hpp:23-42 via pcl::RegionGrowing) restatements: known answers on
path / K2 / Petersen graphs (the shapes of boost_helper/exercices_toolbox.h:135-185),
an independent pure-Python minimax, schedule independence of the cost map, and the
tie-admissibility of every label policy."""
import heapq

import numpy as np
import pytest

from iftwt_rg_b200 import synth
from oracle import pyoracle as orc


def csr_from_edges(n, edges):
    rows = [set() for _ in range(n)]
    for a, b in edges:
        if a != b:
            rows[a].add(b)
            rows[b].add(a)
    off = np.zeros(n + 1, np.uint32)
    adj = []
    for i, r in enumerate(rows):
        adj.extend(sorted(r))
        off[i + 1] = len(adj)
    return off, np.array(adj, np.uint32)


def py_minimax(off, adj, w, seeds):
    """cost(t) = min over seeds, over paths, of max(w[v]) for v after the seed."""
    n = len(off) - 1
    cost = np.full(n, 500, np.int64)
    h = []
    for s in set(int(x) for x in seeds):
        cost[s] = 0
        heapq.heappush(h, (0, s))
    while h:
        c, s = heapq.heappop(h)
        if c != cost[s]:
            continue
        for t in adj[off[s]:off[s + 1]]:
            nc = max(int(w[t]), c)
            if nc < cost[t]:
                cost[t] = nc
                heapq.heappush(h, (nc, int(t)))
    return cost.astype(np.uint32)


def py_policy_C(off, adj, w, seeds):
    """Level-synchronous label-min fixpoint, written directly from the contract."""
    n = len(off) - 1
    cost = py_minimax(off, adj, w, seeds).astype(np.int64)
    seeds = set(int(x) for x in seeds)
    label = np.arange(n, dtype=np.int64)
    INF = 1 << 40
    for v in range(n):
        if v not in seeds and cost[v] < 500:
            label[v] = INF
    for c in sorted(set(cost[cost < 500].tolist())):
        changed = True
        while changed:
            changed = False
            for t in np.nonzero(cost == c)[0]:
                if int(t) in seeds:
                    continue
                for s in adj[off[t]:off[t + 1]]:
                    if cost[s] <= c and max(int(w[t]), int(cost[s])) == c and label[s] < label[t]:
                        label[t] = label[s]
                        changed = True
    return cost.astype(np.uint32), label.astype(np.uint32)


def test_path_graph_known_answer():
    #   0 - 1 - 2 - 3 - 4 - 5 - 6 ;  seeds at 0 and 6
    w = np.array([9, 3, 5, 7, 2, 4, 1], np.float32)
    off, adj = csr_from_edges(7, [(i, i + 1) for i in range(6)])
    for pol in (orc.POLICY_R, orc.POLICY_C, orc.POLICY_LITERAL):
        cost, label = orc.ift(off, adj, w, [0, 6], pol)
        # seed cost forced to 0 (ift.hpp:138) even though w[0] = 9
        assert cost.tolist() == [0, 3, 5, 7, 4, 4, 0]
        # vertex 3 is a tie: offered cost 7 by vertex 4 (cost 4, label 6) and by
        # vertex 2 (cost 5, label 0).  First-offer (R / literal) takes the offer of
        # the lower-cost neighbour; min-seed (C) takes the smaller seed id.
        if pol == orc.POLICY_C:
            assert label.tolist() == [0, 0, 0, 0, 6, 6, 6]
        else:
            assert label.tolist() == [0, 0, 0, 6, 6, 6, 6]


def test_k2_and_unreachable():
    off, adj = csr_from_edges(4, [(0, 1)])          # 2,3 isolated
    w = np.array([5, 7, 1, 600], np.float32)
    for pol in (0, 1, 2):
        cost, label = orc.ift(off, adj, w, [0], pol)
        assert cost.tolist() == [0, 7, 500, 500]     # MAX_COST kept (ift.hpp:1, App. B #7)
        assert label.tolist() == [0, 0, 2, 3]        # unreachable keep label = self (ift.hpp:179)


def test_weight_at_or_above_max_cost_is_never_conquered():
    off, adj = csr_from_edges(3, [(0, 1), (1, 2)])
    w = np.array([0, 500, 3], np.float32)
    for pol in (0, 1, 2):
        cost, label = orc.ift(off, adj, w, [0], pol)
        assert cost.tolist() == [0, 500, 500]


def petersen():
    e = [(i, (i + 1) % 5) for i in range(5)] + [(i, i + 5) for i in range(5)] + \
        [(5 + i, 5 + (i + 2) % 5) for i in range(5)]
    return csr_from_edges(10, e)


def test_petersen_tie_policies():
    off, adj = petersen()
    w = np.full(10, 4, np.float32)
    seeds = [7, 2]
    cR, lR = orc.ift(off, adj, w, seeds, orc.POLICY_R)
    cC, lC = orc.ift(off, adj, w, seeds, orc.POLICY_C)
    assert np.array_equal(cR, cC)
    assert cC[2] == 0 and cC[7] == 0 and np.all(np.delete(cC, [2, 7]) == 4)
    # policy C: the whole plateau takes the smallest seed id, seeds keep their own
    expect = np.full(10, 2, np.uint32)
    expect[7] = 7
    assert np.array_equal(lC, expect)
    for c, l in ((cR, lR), (cC, lC)):
        assert orc.ift_audit(off, adj, w, seeds, c, l) == 0


@pytest.mark.parametrize("seed", range(12))
def test_random_graphs_costs_and_policies(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(20, 400))
    m = int(n * rng.uniform(1.0, 4.0))
    edges = rng.integers(0, n, size=(m, 2))
    off, adj = csr_from_edges(n, edges.tolist())
    w = rng.integers(0, int(rng.integers(2, 30)), size=n).astype(np.float32)
    if seed % 3 == 0:
        w[rng.integers(0, n, 3)] = 510.0
    seeds = rng.integers(0, n, size=int(rng.integers(1, 8))).astype(np.int32)
    ref_cost = py_minimax(off, adj, w, seeds)
    results = {}
    for pol in (0, 1, 2):
        cost, label = orc.ift(off, adj, w, seeds, pol)
        assert np.array_equal(cost, ref_cost), f"policy {pol}"
        assert orc.ift_audit(off, adj, w, seeds, cost, label) == 0, f"policy {pol}"
        results[pol] = label
    cC, lC = py_policy_C(off, adj, w, seeds)
    assert np.array_equal(results[1], lC)
    # the audit must actually be able to fail
    bad = results[1].copy()
    reached = np.nonzero((ref_cost < 500) & (ref_cost > 0))[0]
    if len(reached) and len(set(seeds.tolist())) > 1:
        bad[reached[0]] = n + 5
        assert orc.ift_audit(off, adj, w, seeds, ref_cost, bad) > 0


def test_literal_mst_block_runs_and_is_bounded_by_seed_count():
    rng = np.random.default_rng(5)
    n = 300
    edges = rng.integers(0, n, size=(900, 2))
    off, adj = csr_from_edges(n, edges.tolist())
    w = rng.integers(0, 40, size=n).astype(np.float32)
    seeds = np.array([3, 50, 120, 250], np.int32)
    cost, label, a, b, wv = orc.ift_literal_mst(off, adj, w, seeds)
    assert np.array_equal(cost, py_minimax(off, adj, w, seeds))
    # `status` lets each region record at most one edge as the s side (ift.hpp:220-222)
    assert len(a) <= len(seeds) and len(set(a.tolist())) == len(a)
    assert np.all(wv >= 0)


def test_region_growing_literal_equals_min_rank_reachability():
    pts = synth.scene_primitives(6000, seed=21)
    idx, _ = orc.knn_self(pts, 12)
    nrm = orc.normals(pts, idx)
    for theta in (3.0 / 180.0 * np.pi, 0.15, 0.6):
        lab, raw, kept, nseg = orc.region_growing(nrm, idx, theta, 1.0, 5, 100000)
        mr = orc.rg_minrank(nrm, idx, theta)
        # same partition: raw segment id <-> min rank is a bijection
        pairs = set(zip(raw.tolist(), mr.tolist()))
        assert len(pairs) == nseg == len(set(mr.tolist()))
        # PCL emits segments in ascending seed rank
        first = {}
        for r, m in pairs:
            first[r] = m
        ranks = [first[s] for s in range(nseg)]
        assert ranks == sorted(ranks)
        sizes = np.bincount(raw, minlength=nseg)
        keep = (sizes >= 5) & (sizes <= 100000)
        assert kept == int(keep.sum())
        assert np.array_equal(lab >= 0, keep[raw])


def test_rg_seeds_are_nearest_points_to_fp32_centroids():
    pts = synth.scene_primitives(8000, seed=22)
    idx, _ = orc.knn_self(pts, 16)
    nrm = orc.normals(pts, idx)
    lab, raw, kept, nseg = orc.region_growing(nrm, idx, 0.2, 1.0, 20, 10000)
    assert kept > 0
    seeds, cen = orc.rg_seeds(pts, lab, kept)
    for c in range(min(kept, 10)):
        m = pts[lab == c, :3]
        acc = np.zeros(3, np.float32)
        for row in m:                     # sequential FP32 sum in ascending index order
            acc = (acc + row).astype(np.float32)
        assert np.array_equal(cen[c, :3], (acc / np.float32(len(m))).astype(np.float32))
        d = ((pts[:, :3].astype(np.float32) - cen[c, :3]) ** 2)
        d2 = (d[:, 0] + d[:, 1]) + d[:, 2]
        assert seeds[c] == int(np.flatnonzero(d2 == d2.min())[0])
