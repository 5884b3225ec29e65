"""Shared helpers of the test suite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def ulp_diff(a, b):
    """Distance in float32 units-in-the-last-place between two float32 arrays (NaN==NaN counts as 0)."""
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()
    b = np.ascontiguousarray(b, dtype=np.float32).ravel()
    both_nan = np.isnan(a) & np.isnan(b)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)
    d[both_nan] = 0
    return d


def fresh_layers(rows, cols, elevation=None):
    """AerialGridMap::initialize values (aerial-mapper-grid-map.cc:40-48) as float32 F-order arrays."""
    L = {"elevation": np.full((rows, cols), np.nan, np.float32, order="F"),
         "elevation_angle": np.zeros((rows, cols), np.float32, order="F"),
         "observation_index": np.full((rows, cols), np.nan, np.float32, order="F"),
         "ortho": np.full((rows, cols), 255.0, np.float32, order="F"),
         "colored_ortho": np.full((rows, cols), np.nan, np.float32, order="F")}
    if elevation is not None:
        L["elevation"][...] = elevation
    return L


def brute_dsm(rows, cols, res, xyz, radius=1, ce=0.0, cn=0.0, pos=(0.0, 0.0)):
    """Independent O(cells x points) numpy evaluation of the DSM definition (tiny inputs only).
    numpy element-wise double arithmetic is un-fused, i.e. rounds like the reference's."""
    from aerial_mapper_b200 import synth
    qx, qy = synth.grid_positions(rows, cols, res, pos[0], pos[1])
    px = xyz[:, 0] - cn
    py = xyz[:, 1] - ce
    pz = xyz[:, 2]
    thr = []
    lam = 1.0
    while True:
        thr.append(lam * radius)
        lam *= 1.1
        if lam * radius > 7.0:
            break
    elev = np.full((rows, cols), np.nan, np.float32, order="F")
    cnt = np.zeros((rows, cols), np.int32, order="F")
    lvl = np.full((rows, cols), -1, np.int8, order="F")
    for j in range(cols):
        dy2 = (qy[j] - py) * (qy[j] - py)
        for i in range(rows):
            dx = qx[i] - px
            d2 = dx * dx + dy2
            for k, t in enumerate(thr):
                m = d2 < t
                if m.any():
                    w = 1.0 / d2[m]
                    num = np.sum(pz[m] / d2[m])
                    elev[i, j] = np.float32(num / np.sum(w))
                    cnt[i, j] = int(m.sum())
                    lvl[i, j] = k
                    break
    return elev, cnt, lvl
