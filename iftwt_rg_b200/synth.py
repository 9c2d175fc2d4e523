"""Synthetic point clouds for the BASELINE.json configurations.

The reference ships no usable input (cloud/lucy.pcd and cloud/lucy_gray.ply are
absent, `.MISSING_LARGE_BLOBS:1-2`), so every measured cloud is synthetic.

RNG: counter-based SplitMix64.  Draw number ``n`` of stream ``seed`` is
``mix(seed + (n+1)*GAMMA)``; ``u = (z >> 40) * 2**-24`` is an FP32-exact uniform
in [0,1).  Point ``i`` owns draws ``8*i .. 8*i+7`` so generation is vectorised
and any sub-range of a cloud can be produced independently (this is what the
multi-GPU bench uses to build per-rank shards without materialising 100 M points
on one host).  Points are emitted in generation order (unsorted in space), so
the Morton sort is part of the measured work.  Rows are float32
``{x, y, z, intensity}`` with integer-valued intensity in [0, 255].
"""
from __future__ import annotations

import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)
DRAWS = 8


def _mix(z: np.ndarray) -> np.ndarray:
    z = z.copy()
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def uniforms(seed: int, start: int, count: int) -> np.ndarray:
    """(count, DRAWS) float64 uniforms in [0,1) for points start..start+count."""
    with np.errstate(over="ignore"):
        n = np.arange(start * DRAWS + 1, (start + count) * DRAWS + 1, dtype=np.uint64)
        z = _mix(np.uint64(seed) + n * GAMMA)
    return ((z >> np.uint64(40)).astype(np.float64) * 2.0**-24).reshape(count, DRAWS)


def _gauss(u1: np.ndarray, u2: np.ndarray) -> np.ndarray:
    return np.sqrt(-2.0 * np.log(u1 + 2.0**-24)) * np.cos(2.0 * np.pi * u2)


def _finish(xyz: np.ndarray, base: np.ndarray, g_int: np.ndarray) -> np.ndarray:
    inten = np.clip(np.rint(base + 6.0 * g_int), 0, 255)
    out = np.empty((xyz.shape[0], 4), dtype=np.float32)
    out[:, :3] = xyz.astype(np.float32)
    out[:, 3] = inten.astype(np.float32)
    return out


def scene_primitives(n: int, seed: int = 0x1F02, start: int = 0, sigma: float = 2e-3) -> np.ndarray:
    """cfg2 / cfg5: 40 % plane z=0 on [-5,5]^2, 30 % cylinder r=1 h=4 axis || z at
    (2,0), 30 % sphere r=1.5 at (-2,0,1.5); noise along the surface normal."""
    u = uniforms(seed, start, n)
    g = _gauss(u[:, 3], u[:, 4]) * sigma
    gi = _gauss(u[:, 5], u[:, 6])
    xyz = np.empty((n, 3))
    base = np.empty(n)
    sel = u[:, 0]
    # plane
    m = sel < 0.4
    xyz[m, 0] = -5.0 + 10.0 * u[m, 1]
    xyz[m, 1] = -5.0 + 10.0 * u[m, 2]
    xyz[m, 2] = g[m]
    base[m] = 60.0
    # cylinder
    m = (sel >= 0.4) & (sel < 0.7)
    phi = 2.0 * np.pi * u[m, 1]
    r = 1.0 + g[m]
    xyz[m, 0] = 2.0 + r * np.cos(phi)
    xyz[m, 1] = r * np.sin(phi)
    xyz[m, 2] = 4.0 * u[m, 2]
    base[m] = 140.0
    # sphere (area-uniform: z uniform)
    m = sel >= 0.7
    phi = 2.0 * np.pi * u[m, 1]
    cz = 2.0 * u[m, 2] - 1.0
    sr = np.sqrt(np.maximum(0.0, 1.0 - cz * cz))
    r = 1.5 + g[m]
    xyz[m, 0] = -2.0 + r * sr * np.cos(phi)
    xyz[m, 1] = r * sr * np.sin(phi)
    xyz[m, 2] = 1.5 + r * cz
    base[m] = 220.0
    return _finish(xyz, base, gi)


# facade geometry (cfg3 / cfg4), metres
_W, _H = 40.0, 15.0
_WIN_W, _WIN_H, _WIN_D = 2.0, 2.5, 0.3
_COL_R, _CORN_R, _GROUND_D = 0.4, 0.5, 5.0


def scene_facade(n: int, seed: int = 0x1F03, start: int = 0, sigma: float = 3e-3,
                 x_offset: float = 0.0) -> np.ndarray:
    """cfg3 / cfg4: 40 x 15 m facade — wall, 6x3 window recesses 0.3 m deep,
    8 half-columns r=0.4, cornice half-cylinder r=0.5, ground plane.  Primitives
    are sampled in proportion to their area (about 9.4e3 pts/m^2 at 10 M)."""
    u = uniforms(seed, start, n)
    g = _gauss(u[:, 3], u[:, 4]) * sigma
    gi = _gauss(u[:, 5], u[:, 6])
    areas = np.array([
        _W * _H,                                        # wall (+ window backs)
        18 * 2.0 * (_WIN_W + _WIN_H) * _WIN_D,          # recess side faces
        8 * np.pi * _COL_R * _H,                        # half columns
        np.pi * _CORN_R * _W,                           # cornice
        _W * _GROUND_D,                                 # ground
    ])
    cum = np.cumsum(areas) / areas.sum()
    sel = np.searchsorted(cum, u[:, 0], side="right").clip(0, 4)
    xyz = np.empty((n, 3))
    base = np.empty(n)
    wcx = _W * (np.arange(6) + 0.5) / 6.0
    wcz = _H * (np.arange(3) + 0.5) / 3.0
    # wall / window backs
    m = sel == 0
    x = _W * u[m, 1]
    z = _H * u[m, 2]
    in_win = (np.abs(x[:, None] - wcx[None, :]).min(axis=1) < _WIN_W / 2) & \
             (np.abs(z[:, None] - wcz[None, :]).min(axis=1) < _WIN_H / 2)
    xyz[m, 0] = x
    xyz[m, 1] = np.where(in_win, -_WIN_D, 0.0) + g[m]
    xyz[m, 2] = z
    base[m] = np.where(in_win, 40.0, 120.0)
    # recess side faces: u7 picks window (18) and face (4)
    m = sel == 1
    pick = np.minimum((u[m, 7] * 72.0).astype(np.int64), 71)
    win, face = pick // 4, pick % 4
    cx, cz_ = wcx[win % 6], wcz[win // 6]
    depth = -_WIN_D * u[m, 2]
    t = u[m, 1]
    horiz = face < 2  # bottom / top faces span x, normal along z
    sgn = np.where(face % 2 == 0, -1.0, 1.0)
    xyz[m, 0] = np.where(horiz, cx - _WIN_W / 2 + _WIN_W * t, cx + sgn * _WIN_W / 2 + g[m])
    xyz[m, 1] = depth
    xyz[m, 2] = np.where(horiz, cz_ + sgn * _WIN_H / 2 + g[m], cz_ - _WIN_H / 2 + _WIN_H * t)
    base[m] = 80.0
    # half columns
    m = sel == 2
    col = np.minimum((u[m, 7] * 8.0).astype(np.int64), 7)
    phi = np.pi * u[m, 1]
    r = _COL_R + g[m]
    xyz[m, 0] = _W * col / 7.0 + r * np.cos(phi)
    xyz[m, 1] = r * np.sin(phi)
    xyz[m, 2] = _H * u[m, 2]
    base[m] = 180.0
    # cornice
    m = sel == 3
    phi = np.pi * u[m, 1]
    r = _CORN_R + g[m]
    xyz[m, 0] = _W * u[m, 2]
    xyz[m, 1] = r * np.sin(phi)
    xyz[m, 2] = _H + r * np.cos(phi)
    base[m] = 210.0
    # ground
    m = sel == 4
    xyz[m, 0] = _W * u[m, 1]
    xyz[m, 1] = _GROUND_D * u[m, 2]
    xyz[m, 2] = g[m]
    base[m] = 90.0
    xyz[:, 0] += x_offset
    return _finish(xyz, base, gi)


def scene_statue(n: int, seed: int = 0x1F01, start: int = 0, sigma: float = 1e-3) -> np.ndarray:
    """cfg1 substitute: union of 6 ellipsoid shells ("statue"), direction-uniform
    sampling, so the surface density varies (exercises the k-NN ring fallback)."""
    u = uniforms(seed, start, n)
    g = _gauss(u[:, 3], u[:, 4]) * sigma
    gi = _gauss(u[:, 5], u[:, 6])
    centres = np.array([[0, 0, 0.9], [0, 0, 1.75], [-0.35, 0, 1.1], [0.35, 0, 1.1],
                        [-0.12, 0, 0.3], [0.12, 0, 0.3]], dtype=np.float64)
    radii = np.array([[0.22, 0.15, 0.35], [0.11, 0.12, 0.14], [0.07, 0.07, 0.3],
                      [0.07, 0.07, 0.3], [0.08, 0.08, 0.32], [0.08, 0.08, 0.32]])
    part = np.minimum((u[:, 0] * 6.0).astype(np.int64), 5)
    phi = 2.0 * np.pi * u[:, 1]
    cz = 2.0 * u[:, 2] - 1.0
    sr = np.sqrt(np.maximum(0.0, 1.0 - cz * cz))
    d = np.stack([sr * np.cos(phi), sr * np.sin(phi), cz], axis=1)
    xyz = centres[part] + d * (radii[part] + g[:, None])
    base = 40.0 + 35.0 * part
    return _finish(xyz, base, gi)


def facade_tiled(n: int, tiles: int = 10, seed0: int = 0x1F04, start: int = 0, count: int | None = None
                 ) -> np.ndarray:
    """cfg4: the facade generator tiled `tiles` times along x with distinct seeds.
    Point j of tile t has global index t*(n//tiles)+j; [start, start+count) may
    be requested to build one rank's share."""
    per = n // tiles
    if count is None:
        count = n - start
    out = []
    i = start
    while i < start + count:
        t = min(i // per, tiles - 1)
        j0 = i - t * per
        tile_n = per if t < tiles - 1 else n - per * (tiles - 1)
        c = min(tile_n - j0, start + count - i)
        out.append(scene_facade(c, seed=seed0 + t, start=j0, x_offset=_W * t))
        i += c
    return np.concatenate(out, axis=0)


SCENES = {"primitives": scene_primitives, "facade": scene_facade, "statue": scene_statue}
