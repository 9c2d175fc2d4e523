"""Oracle restatement of the colour path: GraCo colour morphology
(graco.hpp:301-351) and the grey / Otsu preparation (main.cpp:16-76), on known answers and
against independent numpy restatements."""
import numpy as np

from oracle import pyoracle as orc


def _rgb(r, g, b):
    return (int(r) << 16) | (int(g) << 8) | int(b)


def test_colour_morphology_known_answer_path():
    # path 0-1-2-3 ; sums 30, 600, 30 (tie with vertex 0, different colour), 765
    off = np.array([0, 1, 3, 5, 6], np.uint32)
    adj = np.array([1, 0, 2, 1, 3, 2], np.uint32)
    rgb = np.array([_rgb(10, 10, 10), _rgb(200, 200, 200), _rgb(0, 30, 0), _rgb(255, 255, 255)], np.uint32)
    ero = orc.morph_rgb(off, adj, rgb, 4)
    dil = orc.morph_rgb(off, adj, rgb, 3)
    # vertex 1 sees {0,1,2}: equal smallest sums 30 at 0 and 2 -> first in ascending id (strict update)
    assert ero.tolist() == [_rgb(10, 10, 10), _rgb(10, 10, 10), _rgb(0, 30, 0), _rgb(0, 30, 0)]
    assert dil.tolist() == [_rgb(200, 200, 200), _rgb(200, 200, 200), _rgb(255, 255, 255), _rgb(255, 255, 255)]


def test_colour_morphology_initial_values_and_alpha():
    # an isolated vertex keeps its own colour through the closed neighbourhood; the alpha byte
    # PCL stores (0xFF) is ignored and cleared; all-white / all-black stay at the initial value
    off = np.array([0, 0, 0, 0], np.uint32)
    adj = np.zeros(0, np.uint32)
    rgb = np.array([0xFF102030, 0xFFFFFFFF, 0xFF000000], np.uint32)
    assert orc.morph_rgb(off, adj, rgb, 4).tolist() == [0x102030, 0xFFFFFF, 0x000000]
    assert orc.morph_rgb(off, adj, rgb, 3).tolist() == [0x102030, 0xFFFFFF, 0x000000]


def test_colour_morphology_matches_numpy_on_random_graph():
    rng = np.random.default_rng(5)
    n = 400
    a = rng.integers(0, n, 1500)
    b = rng.integers(0, n, 1500)
    keep = a != b
    pairs = np.unique(np.concatenate([np.stack([a[keep], b[keep]], 1), np.stack([b[keep], a[keep]], 1)]), axis=0)
    off = np.zeros(n + 1, np.uint32)
    np.add.at(off, pairs[:, 0] + 1, 1)
    off = np.cumsum(off).astype(np.uint32)
    adj = pairs[:, 1].astype(np.uint32)
    rgb = rng.integers(0, 1 << 24, n).astype(np.uint32)
    rgb[rng.integers(0, n, 60)] = rgb[rng.integers(0, n, 60)]          # equal sums with equal colours
    sums = ((rgb >> 16) & 255) + ((rgb >> 8) & 255) + (rgb & 255)
    for op, sign in ((4, 1), (3, -1)):
        want = np.empty(n, np.uint32)
        for v in range(n):
            nb = np.sort(np.concatenate([[v], adj[off[v]:off[v + 1]]])).astype(np.int64)
            j = nb[np.argmin(sign * sums[nb].astype(np.int64))]   # argmin returns the first: the first smallest / largest in id order
            want[v] = rgb[j]
        assert np.array_equal(orc.morph_rgb(off, adj, rgb, op), want)


def test_rgb_to_gray_fixed_point_and_otsu():
    # cv::cvtColor RGB2GRAY on 8-bit data: primaries and grey ramp known answers
    rgb = np.array([_rgb(255, 0, 0), _rgb(0, 255, 0), _rgb(0, 0, 255), _rgb(255, 255, 255), 0, _rgb(128, 128, 128)],
                   np.uint32)
    gray, thr = orc.rgb_to_gray(rgb, False)
    assert gray.tolist() == [76, 150, 29, 255, 0, 128] and thr == -1.0
    # bimodal sample: Otsu picks a threshold between the modes; binarisation is v < t -> 0 else 255
    rng = np.random.default_rng(9)
    lo = rng.integers(20, 60, 3000)
    hi = rng.integers(170, 230, 2000)
    v = np.concatenate([lo, hi]).astype(np.uint32)
    rgb = (v << 16) | (v << 8) | v
    gray, thr = orc.rgb_to_gray(rgb, True)
    assert 59 <= thr < 170
    assert np.array_equal(gray, np.where(v < thr, 0, 255).astype(np.float32))
    # independent between-class-variance maximiser (float64, first maximum)
    h = np.bincount(v, minlength=256).astype(np.float64)
    p = h / h.sum()
    best, arg = 0.0, 0
    for t in range(256):
        q1 = p[:t + 1].sum()
        q2 = 1 - q1
        if min(q1, q2) < 1.1920929e-07:
            continue
        m1 = (np.arange(t + 1) * p[:t + 1]).sum() / q1
        m2 = (np.arange(t + 1, 256) * p[t + 1:]).sum() / q2
        s = q1 * q2 * (m1 - m2) ** 2
        if s > best * (1 + 1e-12):
            best, arg = s, t
    assert thr == arg
