"""ctypes wrapper of the CPU oracle (oracle/iftwt_oracle.cpp).

TEST INFRASTRUCTURE.  Importable only from tests/, __graft_entry__.smoke() and
the cpu_baseline / --impl reference legs of bench.py.  The product package
(iftwt_rg_b200) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "iftwt_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_cloud_resolution.restype = C.c_double
        for f in ("orc_knn_graph", "orc_region_growing", "orc_ift_literal_mst", "orc_ift_audit",
                  "orc_voxel_graph", "orc_cluster_literal"):
            if hasattr(_lib, f):
                getattr(_lib, f).restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _pts(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] == 4
    return pts


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(t: int):
    lib().orc_set_threads(C.c_int(t))


def knn_self(pts, k):
    pts = _pts(pts)
    n = pts.shape[0]
    idx = np.empty((n, k), np.int32)
    d2 = np.empty((n, k), np.float32)
    lib().orc_knn_self(_p(pts), C.c_int64(n), C.c_int(k), _p(idx), _p(d2))
    return idx, d2


def knn_query(pts, q, k, brute=False):
    pts, q = _pts(pts), _pts(q)
    idx = np.empty((q.shape[0], k), np.int32)
    d2 = np.empty((q.shape[0], k), np.float32)
    f = lib().orc_knn_brute if brute else lib().orc_knn_query
    f(_p(pts), C.c_int64(pts.shape[0]), _p(q), C.c_int64(q.shape[0]), C.c_int(k), _p(idx), _p(d2))
    return idx, d2


def normals(pts, nbr, trig_mode=0):
    pts = _pts(pts)
    nbr = np.ascontiguousarray(nbr, np.int32)
    out = np.empty((pts.shape[0], 4), np.float32)
    lib().orc_normals(_p(pts), C.c_int64(pts.shape[0]), _p(nbr), C.c_int(nbr.shape[1]),
                      C.c_int(trig_mode), _p(out))
    return out


def knn_graph(nbr):
    nbr = np.ascontiguousarray(nbr, np.int32)
    n, k = nbr.shape
    off = np.empty(n + 1, np.uint32)
    adj = np.empty(2 * n * k, np.uint32)   # upper bound: one pass instead of count + fill
    tot = lib().orc_knn_graph(_p(nbr), C.c_int64(n), C.c_int(k), _p(off), _p(adj))
    return off, adj[:tot].copy()


def region_growing(nrm, nbr, theta, curv_thr=1.0, min_sz=50, max_sz=10000):
    nrm = np.ascontiguousarray(nrm, np.float32)
    nbr = np.ascontiguousarray(nbr, np.int32)
    n, k = nbr.shape
    lab = np.empty(n, np.int32)
    raw = np.empty(n, np.int32)
    nseg = C.c_int64(0)
    kept = lib().orc_region_growing(_p(nrm), _p(nbr), C.c_int64(n), C.c_int(k), C.c_float(theta),
                                    C.c_float(curv_thr), C.c_int(min_sz), C.c_int(max_sz), _p(lab),
                                    _p(raw), C.byref(nseg))
    return lab, raw, int(kept), int(nseg.value)


def rg_minrank(nrm, nbr, theta):
    nrm = np.ascontiguousarray(nrm, np.float32)
    nbr = np.ascontiguousarray(nbr, np.int32)
    n, k = nbr.shape
    out = np.empty(n, np.int32)
    lib().orc_rg_minrank(_p(nrm), _p(nbr), C.c_int64(n), C.c_int(k), C.c_float(theta), _p(out))
    return out


def rg_seeds(pts, point_label, n_kept):
    pts = _pts(pts)
    point_label = np.ascontiguousarray(point_label, np.int32)
    seeds = np.empty(n_kept, np.int32)
    cen = np.empty((n_kept, 4), np.float32)
    lib().orc_rg_seeds(_p(pts), C.c_int64(pts.shape[0]), _p(point_label), C.c_int64(n_kept),
                       _p(seeds), _p(cen))
    return seeds, cen


POLICY_R, POLICY_C, POLICY_LITERAL = 0, 1, 2


def ift(off, adj, w, seeds, policy=POLICY_C):
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    w = np.ascontiguousarray(w, np.float32)
    seeds = np.ascontiguousarray(seeds, np.int32)
    n = off.shape[0] - 1
    cost = np.empty(n, np.uint32)
    label = np.empty(n, np.uint32)
    lib().orc_ift(_p(off), _p(adj), C.c_int64(n), _p(w), _p(seeds), C.c_int64(seeds.shape[0]),
                  C.c_int(policy), _p(cost), _p(label))
    return cost, label


def ift_literal_mst(off, adj, w, seeds):
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    w = np.ascontiguousarray(w, np.float32)
    seeds = np.ascontiguousarray(seeds, np.int32)
    n = off.shape[0] - 1
    cost = np.empty(n, np.uint32)
    label = np.empty(n, np.uint32)
    cap = max(1, seeds.shape[0])
    a = np.empty(cap, np.uint32)
    b = np.empty(cap, np.uint32)
    wv = np.empty(cap, np.float64)
    m = lib().orc_ift_literal_mst(_p(off), _p(adj), C.c_int64(n), _p(w), _p(seeds),
                                  C.c_int64(seeds.shape[0]), _p(cost), _p(label), _p(a), _p(b),
                                  _p(wv), C.c_int64(cap))
    m = int(m)
    return cost, label, a[:m], b[:m], wv[:m]


def ift_audit(off, adj, w, seeds, cost, label) -> int:
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    w = np.ascontiguousarray(w, np.float32)
    seeds = np.ascontiguousarray(seeds, np.int32)
    cost = np.ascontiguousarray(cost, np.uint32)
    label = np.ascontiguousarray(label, np.uint32)
    n = off.shape[0] - 1
    return int(lib().orc_ift_audit(_p(off), _p(adj), C.c_int64(n), _p(w), _p(seeds),
                                   C.c_int64(seeds.shape[0]), _p(cost), _p(label)))


def cloud_resolution(pts, overlap=1.01, mode=0) -> float:
    pts = _pts(pts)
    return float(lib().orc_cloud_resolution(_p(pts), C.c_int64(pts.shape[0]), C.c_double(overlap),
                                            C.c_int(mode)))


def voxel_graph(pts, resolution):
    pts = _pts(pts)
    n = pts.shape[0]
    cen = np.empty((n, 4), np.float32)
    keys = np.empty((n, 3), np.uint32)
    off = np.empty(n + 1, np.uint32)
    adj = np.empty(26 * n, np.uint32)
    box = np.empty(4, np.float64)
    lib().orc_voxel_graph.restype = C.c_int64
    V = int(lib().orc_voxel_graph(_p(pts), C.c_int64(n), C.c_double(resolution), _p(cen), _p(keys), _p(off),
                                  _p(adj), _p(box)))
    return cen[:V].copy(), keys[:V].copy(), off[:V + 1].copy(), adj[:off[V]].copy(), box


MORPH_BIN_ERODE, MORPH_BIN_DILATE, MORPH_DILATE, MORPH_ERODE, MORPH_GRADIENT = 1, 2, 3, 4, 5


def morph(off, adj, inten, op):
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    inten = np.ascontiguousarray(inten, np.float32)
    out = np.empty_like(inten)
    lib().orc_morph(_p(off), _p(adj), C.c_int64(len(inten)), _p(inten), C.c_int(op), _p(out))
    return out


def regmin(off, adj, inten):
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    inten = np.ascontiguousarray(inten, np.float32)
    seeds = np.empty(len(inten), np.int32)
    lib().orc_regmin.restype = C.c_int64
    m = int(lib().orc_regmin(_p(off), _p(adj), C.c_int64(len(inten)), _p(inten), _p(seeds)))
    return seeds[:m].copy()


def region_adjacency(off, adj, cost, label):
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    cost = np.ascontiguousarray(cost, np.uint32)
    label = np.ascontiguousarray(label, np.uint32)
    n = len(cost)
    f = lib().orc_region_adjacency
    f.restype = C.c_int64
    m = int(f(_p(off), _p(adj), C.c_int64(n), _p(cost), _p(label), None, None, None, C.c_int64(0)))
    a = np.empty(m, np.uint32)
    b = np.empty(m, np.uint32)
    s = np.empty(m, np.uint32)
    f(_p(off), _p(adj), C.c_int64(n), _p(cost), _p(label), _p(a), _p(b), _p(s), C.c_int64(m))
    return a, b, s


def region_info(w, cost, label):
    w = np.ascontiguousarray(w, np.float32)
    cost = np.ascontiguousarray(cost, np.uint32)
    label = np.ascontiguousarray(label, np.uint32)
    n = len(cost)
    f = lib().orc_region_info
    f.restype = C.c_int64
    m = int(f(C.c_int64(n), _p(w), _p(cost), _p(label), None, None, None, C.c_int64(0)))
    seed = np.empty(m, np.uint32)
    area = np.empty(m, np.uint32)
    height = np.empty(m, np.float32)
    f(C.c_int64(n), _p(w), _p(cost), _p(label), _p(seed), _p(area), _p(height), C.c_int64(m))
    return seed, area, height


def cluster_saddle(cost, label, ra, rb, saddle, num_regions):
    cost = np.ascontiguousarray(cost, np.uint32)
    label = np.ascontiguousarray(label, np.uint32)
    ra = np.ascontiguousarray(ra, np.uint32)
    rb = np.ascontiguousarray(rb, np.uint32)
    saddle = np.ascontiguousarray(saddle, np.uint32)
    out = np.empty(len(cost), np.uint32)
    f = lib().orc_cluster_saddle
    f.restype = C.c_int64
    k = int(f(C.c_int64(len(cost)), _p(cost), _p(label), _p(ra), _p(rb), _p(saddle), C.c_int64(len(ra)),
              C.c_int(num_regions), _p(out)))
    return out, k


def cluster_literal(label, mst_a, mst_b, mst_w, num_regions):
    label = np.ascontiguousarray(label, np.uint32)
    mst_a = np.ascontiguousarray(mst_a, np.uint32)
    mst_b = np.ascontiguousarray(mst_b, np.uint32)
    mst_w = np.ascontiguousarray(mst_w, np.float64)
    out = np.empty(len(label), np.int32)
    f = lib().orc_cluster_literal
    f.restype = C.c_int64
    k = int(f(C.c_int64(len(label)), _p(label), _p(mst_a), _p(mst_b), _p(mst_w), C.c_int64(len(mst_a)),
              C.c_int(num_regions), _p(out)))
    return out, k


def morph_rgb(off, adj, rgb, op):
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    rgb = np.ascontiguousarray(rgb, np.uint32)
    out = np.empty_like(rgb)
    lib().orc_morph_rgb(_p(off), _p(adj), C.c_int64(len(rgb)), _p(rgb), C.c_int(op), _p(out))
    return out


def rgb_to_gray(rgb, otsu=False):
    rgb = np.ascontiguousarray(rgb, np.uint32)
    gray = np.empty(len(rgb), np.float32)
    f = lib().orc_rgb_to_gray
    f.restype = C.c_double
    thr = float(f(_p(rgb), C.c_int64(len(rgb)), C.c_int(1 if otsu else 0), _p(gray)))
    return gray, thr
