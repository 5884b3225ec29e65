"""The oracle reproduces the committed golden fixtures bit for bit (tests/golden/make_golden.py wrote them; the DSM
one with oracle/_ref = the reference's nanoflann.hpp compiled verbatim)."""
import os

import numpy as np

from common import GOLDEN, fresh_layers, ulp_diff
from aerial_mapper_b200 import synth
from oracle import pyoracle as po


def test_dsm_golden():
    z = np.load(os.path.join(GOLDEN, "dsm_96x80.npz"))
    rows, cols, res = int(z["rows"]), int(z["cols"]), float(z["res"])
    e = np.full((rows, cols), np.nan, np.float32, order="F")
    st, cnt, lvl, _ = po.dsm_process(po.make_geometry(rows, cols, res), e, z["xyz"], debug=True)
    assert st == 0
    assert np.array_equal(cnt, z["neighbour_count"]) and np.array_equal(lvl, z["threshold_index"])
    assert np.array_equal(np.isnan(e), np.isnan(z["elevation"]))
    d = ulp_diff(e, z["elevation"])
    assert d.max() <= 1 and (d != 0).sum() <= 2   # golden carries the kd-tree's summation order
    if po.have_ref():
        e2 = np.full((rows, cols), np.nan, np.float32, order="F")
        st, _, _, _ = po.dsm_process(po.make_geometry(rows, cols, res), e2, z["xyz"], use_ref=True)
        assert st == 0 and np.array_equal(e2.view(np.uint32), z["elevation"].view(np.uint32))


def _ortho(colored):
    z = np.load(os.path.join(GOLDEN, "ortho_color_96x80.npz" if colored else "ortho_gray_96x80.npz"))
    rows, cols, res = int(z["rows"]), int(z["cols"]), float(z["res"])
    camd = synth.scaled_camera(float(z["cam_scale"]), dist_type=1)
    ch = 3 if colored else 1
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], ch) for k in range(len(z["poses"]))]
    L = fresh_layers(rows, cols, z["elevation"])
    st, _ = po.ortho_process(po.make_geometry(rows, cols, res), L, po.make_camera(**camd), z["poses"], imgs,
                             colored=colored)
    assert st == 0
    assert np.array_equal(L["observation_index"], z["observation_index"], equal_nan=True)
    assert np.array_equal(L["elevation_angle"].view(np.uint32), z["elevation_angle"].view(np.uint32))
    assert np.array_equal(L["colored_ortho" if colored else "ortho"].view(np.uint32), z["out"])


def test_ortho_gray_golden():
    _ortho(False)


def test_ortho_color_golden():
    _ortho(True)
