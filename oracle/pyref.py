"""ctypes wrapper of oracle/_ref/libref_ift.so — the REFERENCE's own IFT (ift.hpp +
globals.hpp compiled unmodified from /root/reference against oracle/shim, driver
oracle/ref_ift_driver.cpp).

TEST INFRASTRUCTURE, same rules as pyoracle: tests/ and tools/ only.  The library is
built in the development container (where /root/reference exists) and travels to the GPU
box prebuilt; nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libref_ift.so")
REFERENCE = os.environ.get("IFTWT_REFERENCE", "/root/reference")


def build(force: bool = False) -> str | None:
    """Builds _ref/libref_ift.so when the reference sources are present; None otherwise."""
    if not os.path.exists(os.path.join(REFERENCE, "ift.hpp")):
        return _LIB if os.path.exists(_LIB) else None
    if force and os.path.exists(_LIB):
        os.remove(_LIB)
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REFERENCE={REFERENCE}"])
    return _LIB


def available() -> bool:
    return os.path.exists(_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB)
        _lib.ref_ift_run.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ift(pts, off, adj, seeds_xyz):
    """IFT_PCD(Graph&, Cloud::Ptr seeds) of ift.hpp:31-37 on the graph (pts, off, adj).
    pts: n x {x, y, z, intensity}; seeds_xyz: ns x 3.  Returns cost[n], root[n] (vertex index of the
    conquering seed, self when unreached) and the reference's `mst` map as (a, b, weight) rows."""
    pts = np.ascontiguousarray(pts, np.float32)
    off = np.ascontiguousarray(off, np.uint32)
    adj = np.ascontiguousarray(adj, np.uint32)
    seeds_xyz = np.ascontiguousarray(seeds_xyz, np.float32).reshape(-1, 3)
    n = pts.shape[0]
    cost = np.empty(n, np.uint32)
    root = np.empty(n, np.uint32)
    cap = max(1, seeds_xyz.shape[0])
    a = np.empty(cap, np.uint32)
    b = np.empty(cap, np.uint32)
    w = np.empty(cap, np.float64)
    m = int(lib().ref_ift_run(_p(pts), C.c_int64(n), _p(off), _p(adj), _p(seeds_xyz), C.c_int64(seeds_xyz.shape[0]),
                              _p(cost), _p(root), _p(a), _p(b), _p(w), C.c_int64(cap)))
    return cost, root, a[:m], b[:m], w[:m]


# ---- the reference's own Graph (graph.hpp) + IFT_PCD on it: oracle/_ref/libref_graph.so ----
_LIBG = os.path.join(_HERE, "_ref", "libref_graph.so")
_libg = None


def graph_available() -> bool:
    return os.path.exists(_LIBG)


def libg():
    global _libg
    if _libg is None:
        _libg = C.CDLL(_LIBG)
        _libg.ref_graph_create.restype = C.c_void_p
        _libg.ref_graph_resolution.restype = C.c_double
        for f in ("ref_graph_cloud", "ref_graph_csr", "ref_graph_regmin", "ref_graph_ift"):
            getattr(_libg, f).restype = C.c_int64
    return _libg


class RefGraph:
    """`Graph gg(cloud, overlap)` of graph.hpp:43-50 — vertices are numbered in the graph's own
    iteration order (ascending voxel Morton key with the shim's octree)."""

    def __init__(self, pts, overlap: float = 1.01):
        pts = np.ascontiguousarray(pts, np.float32)
        self._h = C.c_void_p(libg().ref_graph_create(_p(pts), C.c_int64(pts.shape[0]), C.c_double(overlap)))
        self.n = int(libg().ref_graph_cloud(self._h, None))

    def close(self):
        if self._h:
            libg().ref_graph_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def resolution(self) -> float:
        return float(libg().ref_graph_resolution(self._h))

    def cloud(self):
        out = np.empty((self.n, 4), np.float32)
        libg().ref_graph_cloud(self._h, _p(out))
        return out

    def csr(self):
        off = np.empty(self.n + 1, np.uint32)
        loops = C.c_int64(0)
        m = int(libg().ref_graph_csr(self._h, _p(off), None, C.byref(loops)))
        adj = np.empty(max(m, 1), np.uint32)
        libg().ref_graph_csr(self._h, _p(off), _p(adj), C.byref(loops))
        return off, adj[:m], int(loops.value)

    def morph(self, op: int):
        out = np.empty(self.n, np.float32)
        rc = libg().ref_graph_morph(self._h, C.c_int(op), _p(out))
        assert rc == 0
        return out

    def regmin(self):
        idx = np.empty(self.n, np.int32)
        m = int(libg().ref_graph_regmin(self._h, _p(idx)))
        return idx[:m].copy()

    def ift(self, seeds_xyz):
        seeds_xyz = np.ascontiguousarray(seeds_xyz, np.float32).reshape(-1, 3)
        cost = np.empty(self.n, np.uint32)
        root = np.empty(self.n, np.uint32)
        cap = max(1, seeds_xyz.shape[0])
        a = np.empty(cap, np.uint32)
        b = np.empty(cap, np.uint32)
        w = np.empty(cap, np.float64)
        m = int(libg().ref_graph_ift(self._h, _p(seeds_xyz), C.c_int64(seeds_xyz.shape[0]), _p(cost), _p(root),
                                     _p(a), _p(b), _p(w), C.c_int64(cap)))
        return cost, root, a[:m], b[:m], w[:m]


class RefGraCo:
    """`GraCo gc(cloud_t, overlap)` of graco.hpp:47-56.  pts: n x {x, y, z, bits of the packed colour}."""

    def __init__(self, pts, overlap: float = 5.0):
        pts = np.ascontiguousarray(pts, np.float32)
        g = libg()
        g.ref_graco_create.restype = C.c_void_p
        g.ref_graco_resolution.restype = C.c_double
        g.ref_graco_cloud.restype = C.c_int64
        g.ref_graco_morph.restype = C.c_int64
        self._h = C.c_void_p(g.ref_graco_create(_p(pts), C.c_int64(pts.shape[0]), C.c_double(overlap)))
        self._res = float(g.ref_graco_resolution(self._h))
        self.n = int(g.ref_graco_cloud(self._h, None, None))

    def close(self):
        if self._h:
            libg().ref_graco_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def resolution(self) -> float:
        return self._res

    def cloud(self):
        xyz = np.empty((self.n, 3), np.float32)
        rgb = np.empty(self.n, np.uint32)
        libg().ref_graco_cloud(self._h, _p(xyz), _p(rgb))
        return xyz, rgb

    def morph(self, op: int):
        rgb = np.empty(self.n, np.uint32)
        assert libg().ref_graco_morph(self._h, C.c_int(op), _p(rgb)) == self.n
        return rgb
