"""ctypes binding of include/iftwt.h — the Python mirror of the C ABI.

The product is libiftwt.so (hand-written sm_100a CUDA behind a C ABI); this module
only marshals numpy arrays across that boundary for the test-suite and bench.py.
There is no CPU path here: if the library cannot be loaded or no B200 is visible,
every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

IFTWT_MAX_COST = 500

STAGES = ("upload", "grid", "knn", "graph", "normals", "seeds", "ift", "download", "knn_cell", "graph_rg")


class IftwtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"iftwt error {code}: {msg}")
        self.code = code


class RgParams(C.Structure):
    """pcl::RegionGrowing setters of main.cpp:111-119 (defaults are the reference's)."""
    _fields_ = [("min_cluster_size", C.c_int), ("max_cluster_size", C.c_int),
                ("number_of_neighbours", C.c_int), ("smoothness_threshold", C.c_float),
                ("curvature_threshold", C.c_float)]

    def __init__(self, min_cluster_size=50, max_cluster_size=10000, number_of_neighbours=50,
                 smoothness_threshold=3.0 / 180.0 * np.pi, curvature_threshold=1.0):
        super().__init__(min_cluster_size, max_cluster_size, number_of_neighbours,
                         smoothness_threshold, curvature_threshold)


class SegmentParams(C.Structure):
    _fields_ = [("k_graph", C.c_int), ("k_normals", C.c_int), ("rg", RgParams), ("weights", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("n_cells", C.c_uint64), ("cell_size", C.c_double),
                ("knn_fallback", C.c_uint64), ("n_arcs", C.c_uint64), ("rg_segments", C.c_uint64),
                ("rg_sweeps", C.c_uint64), ("n_seeds", C.c_uint64), ("ift_levels", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_bytes", C.c_uint64), ("density_spread", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class ShardIn(C.Structure):
    """A rank's slice of the cloud (include/iftwt.h: iftwt_shard_in)."""
    _fields_ = [("d_points", C.c_void_p), ("n", C.c_size_t), ("gid0", C.c_uint64)]


class ShardOut(C.Structure):
    """include/iftwt.h: iftwt_shard_out.  Device pointers are owned by the comm."""
    _fields_ = [("n_owned", C.c_size_t), ("n_halo", C.c_size_t), ("n_total", C.c_size_t), ("k", C.c_int),
                ("attempts", C.c_int), ("d_gid", C.c_void_p), ("d_points", C.c_void_p), ("d_knn", C.c_void_p),
                ("d_d2", C.c_void_p), ("d_normals", C.c_void_p), ("halo_radius", C.c_double),
                ("kth_distance_max", C.c_double), ("halo_bytes_sent", C.c_uint64),
                ("exchange_bytes_sent", C.c_uint64), ("ms_estimate", C.c_float), ("ms_partition", C.c_float),
                ("ms_exchange", C.c_float), ("ms_compute", C.c_float), ("ms_export", C.c_float),
                ("knn_cell_ms", C.c_float), ("peer_to_peer", C.c_int)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if not f.startswith("d_")}


COMM_ID_BYTES = 128

_lib = None


def library_path() -> str:
    return _build.LIB


def load_library(build_if_missing: bool = True):
    """Loads libiftwt.so; fails loudly (no fallback) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        if not build_if_missing:
            raise ImportError(f"{path} is missing: run `python -m iftwt_rg_b200._build`")
        _build.build()
    lib = C.CDLL(path)
    vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    lib.iftwt_version.restype = i32
    lib.iftwt_last_error.restype = C.c_char_p
    lib.iftwt_last_error.argtypes = [vp]
    lib.iftwt_create.argtypes = [C.POINTER(vp), i32]
    lib.iftwt_destroy.argtypes = [vp]
    lib.iftwt_destroy.restype = None
    lib.iftwt_set_cloud.argtypes = [vp, vp, sz, sz, i32, i32]
    lib.iftwt_set_cloud_device.argtypes = [vp, vp, sz]
    lib.iftwt_knn.argtypes = [vp, i32, vp, vp]
    lib.iftwt_knn_query.argtypes = [vp, vp, sz, sz, i32, vp, vp]
    lib.iftwt_knn_graph.argtypes = [vp, i32, vp, vp, sz, C.POINTER(sz)]
    lib.iftwt_normals.argtypes = [vp, i32, vp, sz, i32, i32]
    lib.iftwt_region_seeds.argtypes = [vp, C.POINTER(RgParams), vp, vp, sz, C.POINTER(sz)]
    lib.iftwt_ift.argtypes = [vp, vp, vp, sz, vp, vp]
    lib.iftwt_segment.argtypes = [vp, C.POINTER(SegmentParams), vp, vp, C.POINTER(sz)]
    lib.iftwt_cloud_resolution.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
    lib.iftwt_voxel_graph.argtypes = [vp, C.c_double, C.POINTER(sz)]
    lib.iftwt_get_cloud.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.iftwt_graph_csr.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    lib.iftwt_morph.argtypes = [vp, i32, vp]
    lib.iftwt_set_intensity.argtypes = [vp, vp]
    lib.iftwt_regmin_seeds.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.iftwt_region_adjacency.argtypes = [vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.iftwt_region_info.argtypes = [vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.iftwt_cluster.argtypes = [vp, i32, vp, C.POINTER(sz)]
    lib.iftwt_knn_device.argtypes = [vp, i32, vp, vp]
    lib.iftwt_normals_device.argtypes = [vp, i32, vp]
    lib.iftwt_set_knn_rows_device.argtypes = [vp, i32, vp, vp]
    lib.iftwt_morph_rgb.argtypes = [vp, i32, vp]
    lib.iftwt_rgb_to_gray.argtypes = [vp, i32, C.POINTER(C.c_double)]
    lib.iftwt_stage_ms.argtypes = [vp, i32, C.POINTER(f32)]
    lib.iftwt_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.iftwt_reset_counters.argtypes = [vp]
    lib.iftwt_synchronize.argtypes = [vp]
    lib.iftwt_stream.argtypes = [vp]
    lib.iftwt_stream.restype = vp
    lib.iftwt_batch_create.argtypes = [C.POINTER(vp), i32, i32]
    lib.iftwt_batch_destroy.argtypes = [vp]
    lib.iftwt_batch_destroy.restype = None
    lib.iftwt_batch_last_error.argtypes = [vp]
    lib.iftwt_batch_last_error.restype = C.c_char_p
    lib.iftwt_batch_ctx.argtypes = [vp, i32]
    lib.iftwt_batch_ctx.restype = vp
    lib.iftwt_batch_segment.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), sz, i32, C.POINTER(SegmentParams),
                                        C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)]
    lib.iftwt_comm_unique_id.argtypes = [vp]
    lib.iftwt_comm_create.argtypes = [C.POINTER(vp), vp, i32, i32, i32]
    lib.iftwt_comm_create_all.argtypes = [C.POINTER(vp), i32, C.POINTER(i32)]
    lib.iftwt_comm_destroy.argtypes = [vp]
    lib.iftwt_comm_destroy.restype = None
    lib.iftwt_comm_last_error.argtypes = [vp]
    lib.iftwt_comm_last_error.restype = C.c_char_p
    lib.iftwt_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.iftwt_comm_ctx.argtypes = [vp, i32]
    lib.iftwt_comm_ctx.restype = vp
    lib.iftwt_comm_upload.argtypes = [vp, i32, vp, sz, C.c_uint64, C.POINTER(ShardIn)]
    lib.iftwt_sharded_knn_normals.argtypes = [vp, C.POINTER(ShardIn), i32, i32, C.c_double, C.POINTER(ShardOut)]
    lib.iftwt_sharded_segment.argtypes = [vp, C.POINTER(ShardIn), C.POINTER(SegmentParams), C.c_double, i32, vp, vp,
                                          C.POINTER(sz), C.POINTER(ShardOut)]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One ctx = one GPU, one stream (include/iftwt.h)."""

    def __init__(self, device: int = 0, _borrowed=None):
        self._lib = load_library()
        self._owned = _borrowed is None
        if _borrowed is not None:  # a ctx that belongs to a Comm (iftwt_comm_ctx)
            self._h = C.c_void_p(_borrowed)
            self.n = 0
            return
        h = C.c_void_p()
        rc = self._lib.iftwt_create(C.byref(h), device)
        if rc != 0:
            raise IftwtError(rc, self._lib.iftwt_last_error(None).decode())
        self._h = h
        self.n = 0

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                self._lib.iftwt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise IftwtError(rc, self._lib.iftwt_last_error(self._h).decode())

    # -- input ---------------------------------------------------------------
    def set_cloud(self, pts: np.ndarray, xyz_offset: int = 0, intensity_offset: int | None = None):
        """pts: (n, m) float32 rows (m >= 3) or a structured / byte array with explicit offsets."""
        assert pts.flags.c_contiguous
        n = pts.shape[0]
        stride = pts.strides[0]
        if intensity_offset is None:
            intensity_offset = 12 if stride >= 16 else -1
        self._keep = pts
        self._check(self._lib.iftwt_set_cloud(self._h, _ptr(pts), stride, n, xyz_offset, intensity_offset))
        self.n = n

    def set_cloud_device(self, device_ptr: int, n: int):
        self._check(self._lib.iftwt_set_cloud_device(self._h, C.c_void_p(device_ptr), n))
        self.n = n

    # -- stages --------------------------------------------------------------
    def knn(self, k: int, want_d2: bool = True, download: bool = True):
        if not download:
            self._check(self._lib.iftwt_knn(self._h, k, None, None))
            return None, None
        idx = np.empty((self.n, k), np.int32)
        d2 = np.empty((self.n, k), np.float32) if want_d2 else None
        self._check(self._lib.iftwt_knn(self._h, k, _ptr(idx), _ptr(d2)))
        return idx, d2

    def knn_query(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, np.float32)
        m = q.shape[0]
        idx = np.empty((m, k), np.int32)
        d2 = np.empty((m, k), np.float32)
        self._check(self._lib.iftwt_knn_query(self._h, _ptr(q), q.strides[0], m, k, _ptr(idx), _ptr(d2)))
        return idx, d2

    def knn_graph(self, k: int, download: bool = True):
        narcs = C.c_size_t(0)
        if not download:
            self._check(self._lib.iftwt_knn_graph(self._h, k, None, None, 0, C.byref(narcs)))
            return None, None
        self._check(self._lib.iftwt_knn_graph(self._h, k, None, None, 0, C.byref(narcs)))
        off = np.empty(self.n + 1, np.uint32)
        adj = np.empty(narcs.value, np.uint32)
        self._check(self._lib.iftwt_knn_graph(self._h, k, _ptr(off), _ptr(adj), adj.shape[0], C.byref(narcs)))
        return off, adj

    def normals(self, k: int, download: bool = True):
        if not download:
            self._check(self._lib.iftwt_normals(self._h, k, None, 16, 0, 12))
            return None
        out = np.empty((self.n, 4), np.float32)
        self._check(self._lib.iftwt_normals(self._h, k, _ptr(out), 16, 0, 12))
        return out

    def region_seeds(self, params: RgParams | None = None, download: bool = True):
        params = params or RgParams()
        ns = C.c_size_t(0)
        if not download:
            self._check(self._lib.iftwt_region_seeds(self._h, C.byref(params), None, None, 0, C.byref(ns)))
            return None, None
        lab = np.empty(self.n, np.int32)
        self._check(self._lib.iftwt_region_seeds(self._h, C.byref(params), _ptr(lab), None, 0, C.byref(ns)))
        seeds = np.empty(ns.value, np.int32)
        if ns.value:
            self._check(self._lib.iftwt_region_seeds(self._h, C.byref(params), None, _ptr(seeds), ns.value,
                                                     C.byref(ns)))
        return lab, seeds

    def ift(self, weights: np.ndarray | None = None, seeds: np.ndarray | None = None, download: bool = True):
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        s = None if seeds is None else np.ascontiguousarray(seeds, np.int32)
        cost = np.empty(self.n, np.uint32) if download else None
        label = np.empty(self.n, np.uint32) if download else None
        self._check(self._lib.iftwt_ift(self._h, _ptr(w), _ptr(s), 0 if s is None else s.shape[0], _ptr(cost),
                                        _ptr(label)))
        return cost, label

    def segment(self, k_graph: int, k_normals: int | None = None, rg: RgParams | None = None,
                cost: np.ndarray | None = None, label: np.ndarray | None = None, download: bool = True,
                gradient_weights: bool = False):
        """main.cpp:97-123 in one call."""
        p = SegmentParams(k_graph, k_normals or k_graph, rg or RgParams(number_of_neighbours=k_graph),
                          1 if gradient_weights else 0)
        ns = C.c_size_t(0)
        if download:
            cost = np.empty(self.n, np.uint32) if cost is None else cost
            label = np.empty(self.n, np.uint32) if label is None else label
        self._check(self._lib.iftwt_segment(self._h, C.byref(p), _ptr(cost), _ptr(label), C.byref(ns)))
        return cost, label, ns.value

    # -- graph builder (A), morphology, regional minima ------------------------
    def cloud_resolution(self, overlap: float = 1.01) -> float:
        """Graph::computeCloudResolution (graph.hpp:228-259)."""
        v = C.c_double(0)
        self._check(self._lib.iftwt_cloud_resolution(self._h, overlap, C.byref(v)))
        return float(v.value)

    def voxel_graph(self, voxel_size: float) -> int:
        """Graph::initialize (graph.hpp:293-308); the voxel centres become the cloud."""
        nv = C.c_size_t(0)
        self._check(self._lib.iftwt_voxel_graph(self._h, voxel_size, C.byref(nv)))
        self.n = nv.value
        return self.n

    def get_cloud(self) -> np.ndarray:
        out = np.empty((self.n, 4), np.float32)
        m = C.c_size_t(0)
        self._check(self._lib.iftwt_get_cloud(self._h, _ptr(out), self.n, C.byref(m)))
        return out

    def graph_csr(self):
        narcs = C.c_size_t(0)
        self._check(self._lib.iftwt_graph_csr(self._h, None, None, 0, C.byref(narcs)))
        off = np.empty(self.n + 1, np.uint32)
        adj = np.empty(narcs.value, np.uint32)
        self._check(self._lib.iftwt_graph_csr(self._h, _ptr(off), _ptr(adj), adj.shape[0], C.byref(narcs)))
        return off, adj

    def morph(self, op: int) -> np.ndarray:
        out = np.empty(self.n, np.float32)
        self._check(self._lib.iftwt_morph(self._h, op, _ptr(out)))
        return out

    def set_intensity(self, values: np.ndarray):
        v = np.ascontiguousarray(values, np.float32)
        assert v.shape[0] == self.n
        self._check(self._lib.iftwt_set_intensity(self._h, _ptr(v)))

    def regmin_seeds(self) -> np.ndarray:
        ns = C.c_size_t(0)
        self._check(self._lib.iftwt_regmin_seeds(self._h, None, 0, C.byref(ns)))
        seeds = np.empty(ns.value, np.int32)
        if ns.value:
            self._check(self._lib.iftwt_regmin_seeds(self._h, _ptr(seeds), ns.value, C.byref(ns)))
        return seeds

    # -- colour path (graco.hpp, main.cpp:16-89) ---------------------------------
    def morph_rgb(self, op: int) -> np.ndarray:
        out = np.empty(self.n, np.uint32)
        self._check(self._lib.iftwt_morph_rgb(self._h, op, _ptr(out)))
        return out

    def rgb_to_gray(self, otsu: bool = False) -> float:
        t = C.c_double(0)
        self._check(self._lib.iftwt_rgb_to_gray(self._h, 1 if otsu else 0, C.byref(t)))
        return float(t.value)

    # -- device-pointer variants (multi-GPU plumbing) --------------------------
    def knn_device(self, k: int, d_idx: int, d_d2: int = 0):
        self._check(self._lib.iftwt_knn_device(self._h, k, C.c_void_p(d_idx), C.c_void_p(d_d2) if d_d2 else None))

    def normals_device(self, k: int, d_out4: int):
        self._check(self._lib.iftwt_normals_device(self._h, k, C.c_void_p(d_out4)))

    def set_knn_rows_device(self, k: int, d_idx: int, d_d2: int):
        self._check(self._lib.iftwt_set_knn_rows_device(self._h, k, C.c_void_p(d_idx), C.c_void_p(d_d2)))

    # -- region hierarchy ------------------------------------------------------
    def region_adjacency(self):
        m = C.c_size_t(0)
        self._check(self._lib.iftwt_region_adjacency(self._h, None, None, None, 0, C.byref(m)))
        a = np.empty(m.value, np.uint32)
        b = np.empty(m.value, np.uint32)
        s = np.empty(m.value, np.uint32)
        if m.value:
            self._check(self._lib.iftwt_region_adjacency(self._h, _ptr(a), _ptr(b), _ptr(s), m.value, C.byref(m)))
        return a, b, s

    def region_info(self):
        m = C.c_size_t(0)
        self._check(self._lib.iftwt_region_info(self._h, None, None, None, 0, C.byref(m)))
        seed = np.empty(m.value, np.uint32)
        area = np.empty(m.value, np.uint32)
        height = np.empty(m.value, np.float32)
        if m.value:
            self._check(self._lib.iftwt_region_info(self._h, _ptr(seed), _ptr(area), _ptr(height), m.value,
                                                    C.byref(m)))
        return seed, area, height

    def cluster(self, num_regions: int = 10):
        """Cluster(..., num_regions) + getLabelCloud (cluster.hpp:25-81; main.cpp:124-126)."""
        out = np.empty(self.n, np.uint32)
        k = C.c_size_t(0)
        self._check(self._lib.iftwt_cluster(self._h, num_regions, _ptr(out), C.byref(k)))
        return out, k.value

    # -- instrumentation -----------------------------------------------------
    def stage_ms(self) -> dict:
        out = {}
        v = C.c_float(0)
        for i, name in enumerate(STAGES):
            self._check(self._lib.iftwt_stage_ms(self._h, i, C.byref(v)))
            out[name] = float(v.value)
        return out

    def stats(self) -> dict:
        s = Stats()
        self._check(self._lib.iftwt_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_counters(self):
        self._check(self._lib.iftwt_reset_counters(self._h))

    def synchronize(self):
        self._check(self._lib.iftwt_synchronize(self._h))

    def stream(self) -> int:
        return int(self._lib.iftwt_stream(self._h) or 0)


class Batch:
    """A batch of clouds on one GPU, several lanes (ctx + stream + host thread) pulling from one queue
    (include/iftwt.h: iftwt_batch_*; BASELINE config 5)."""

    def __init__(self, device: int = 0, lanes: int = 3):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.iftwt_batch_create(C.byref(h), device, lanes)
        if rc != 0:
            raise IftwtError(rc, self._lib.iftwt_batch_last_error(None).decode())
        self._h = h
        self.lanes = lanes

    def close(self):
        if getattr(self, "_h", None):
            self._lib.iftwt_batch_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ctx(self, lane: int) -> Context:
        return Context(_borrowed=self._lib.iftwt_batch_ctx(self._h, lane))

    def segment(self, clouds, k: int, rg: RgParams, ns=None, on_device: bool = False, cost=None, label=None,
                gradient_weights: bool = False):
        """clouds: host (n_i, 4) float32 arrays, or (device_ptr, n_i) pairs with on_device=True.
        cost / label: lists of uint32 arrays (or None: results stay on the lanes' devices).  Returns n_seeds."""
        B = len(clouds)
        if on_device:
            ptrs = (C.c_void_p * B)(*[C.c_void_p(int(p)) for p, _ in clouds])
            sizes = (C.c_size_t * B)(*[int(m) for _, m in clouds])
        else:
            for a in clouds:
                assert a.flags.c_contiguous and a.dtype == np.float32 and a.shape[1] == 4
            ptrs = (C.c_void_p * B)(*[a.ctypes.data_as(C.c_void_p) for a in clouds])
            sizes = (C.c_size_t * B)(*[a.shape[0] for a in clouds])
        cp = (C.c_void_p * B)(*[None if cost is None else cost[i].ctypes.data_as(C.c_void_p) for i in range(B)])
        lp = (C.c_void_p * B)(*[None if label is None else label[i].ctypes.data_as(C.c_void_p) for i in range(B)])
        seeds = (C.c_size_t * B)()
        p = SegmentParams(k, k, rg, 1 if gradient_weights else 0)
        rc = self._lib.iftwt_batch_segment(self._h, ptrs, sizes, B, 1 if on_device else 0, C.byref(p), cp, lp, seeds)
        if rc != 0:
            raise IftwtError(rc, self._lib.iftwt_batch_last_error(self._h).decode())
        return list(seeds)


class DeviceArray:
    """A device buffer owned by the library, seen through __cuda_array_interface__ (zero copy:
    torch.as_tensor(x, device="cuda") / cupy.asarray(x) wrap it without moving data)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str, owner=None):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr or 0), False), "version": 2, "strides": None}
        self._owner = owner


class Comm:
    """Morton-range sharded k-NN / normals over the GPUs of one node (include/iftwt.h, csrc/shard.cu).

    Comm.rank(id, nranks, rank, device): one process per GPU (ncclCommInitRank); the id comes from
    Comm.unique_id() on rank 0 and travels over any side channel.
    Comm.all(devices): one process driving several GPUs (ncclCommInitAll); the same device named several
    times = shards time-sharing one GPU (the single-GPU test mode)."""

    def __init__(self, handle, nlocal):
        self._lib = load_library()
        self._h = handle
        self.n_local = nlocal

    @staticmethod
    def unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = lib.iftwt_comm_unique_id(buf)
        if rc != 0:
            raise IftwtError(rc, lib.iftwt_comm_last_error(None).decode())
        return buf.raw

    @classmethod
    def rank(cls, uid: bytes, nranks: int, rank: int, device: int) -> "Comm":
        lib = load_library()
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(uid), COMM_ID_BYTES)
        rc = lib.iftwt_comm_create(C.byref(h), buf, nranks, rank, device)
        if rc != 0:
            raise IftwtError(rc, lib.iftwt_comm_last_error(None).decode())
        return cls(h, 1)

    @classmethod
    def all(cls, devices) -> "Comm":
        lib = load_library()
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = lib.iftwt_comm_create_all(C.byref(h), len(devices), arr)
        if rc != 0:
            raise IftwtError(rc, lib.iftwt_comm_last_error(None).decode())
        return cls(h, len(devices))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.iftwt_comm_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise IftwtError(rc, self._lib.iftwt_comm_last_error(self._h).decode())

    def info(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self._lib.iftwt_comm_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"nranks": a.value, "n_local": b.value, "first_local_rank": c.value}

    def ctx(self, local_index: int = 0) -> Context:
        h = self._lib.iftwt_comm_ctx(self._h, local_index)
        if not h:
            raise IftwtError(-1, "no such local shard")
        return Context(_borrowed=h)

    def upload(self, local_index: int, pts: np.ndarray, gid0: int):
        """Host slice -> a device buffer the comm owns; returns the (device_ptr, n, gid0) triple the calls below take."""
        pts = np.ascontiguousarray(pts, np.float32)
        assert pts.ndim == 2 and pts.shape[1] == 4
        si = ShardIn()
        self._check(self._lib.iftwt_comm_upload(self._h, local_index, _ptr(pts), pts.shape[0], gid0, C.byref(si)))
        return (si.d_points or 0, si.n, si.gid0)

    def _ins(self, slices):
        """slices: one (device_ptr, n, gid0) per local shard."""
        assert len(slices) == self.n_local
        return (ShardIn * self.n_local)(*[ShardIn(C.c_void_p(int(p) or None), int(n), int(g)) for p, n, g in slices])

    def sharded_knn_normals(self, slices, k: int, want_normals: bool = True, halo_radius: float = 0.0):
        ins = self._ins(slices)
        outs = (ShardOut * self.n_local)()
        self._check(self._lib.iftwt_sharded_knn_normals(self._h, ins, k, 1 if want_normals else 0,
                                                        float(halo_radius), outs))
        return list(outs)

    def sharded_segment(self, slices, k: int, rg: RgParams, root: int = 0, halo_radius: float = 0.0,
                        n_total: int | None = None, download: bool = True, gradient_weights: bool = False,
                        cost: np.ndarray | None = None, label: np.ndarray | None = None):
        """BASELINE config 4.  Returns (cost, label, n_seeds, outs); cost / label are None unless this
        process holds `root`, downloads and knows n_total (or passes its own n_total-entry buffers)."""
        ins = self._ins(slices)
        outs = (ShardOut * self.n_local)()
        p = SegmentParams(k, k, rg, 1 if gradient_weights else 0)
        ns = C.c_size_t(0)
        info = self.info()
        holds_root = info["first_local_rank"] <= root < info["first_local_rank"] + info["n_local"]
        if not (download and holds_root):
            cost = label = None
        elif n_total and cost is None:
            cost = np.empty(n_total, np.uint32)
            label = np.empty(n_total, np.uint32)
        self._check(self._lib.iftwt_sharded_segment(self._h, ins, C.byref(p), float(halo_radius), root, _ptr(cost),
                                                    _ptr(label), C.byref(ns), outs))
        return cost, label, ns.value, list(outs)

    @staticmethod
    def arrays(o: ShardOut):
        """The device-resident result of one shard as __cuda_array_interface__ objects."""
        n, k = o.n_owned, o.k
        return {"gid": DeviceArray(o.d_gid, (n,), "<u4"), "points": DeviceArray(o.d_points, (n, 4), "<f4"),
                "knn": DeviceArray(o.d_knn, (n, k), "<i4"), "d2": DeviceArray(o.d_d2, (n, k), "<f4"),
                "normals": DeviceArray(o.d_normals, (n, 4), "<f4") if o.d_normals else None}
