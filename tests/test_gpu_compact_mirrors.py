"""Opt-in one-byte transport of `ortho` / `observation_index` to their host mirrors (amb_set_host_mirror_compact,
csrc/mirror_compact.cu): the mirror must receive exactly the bits of the layer — through the codes when every value has
one, through the plain float32 download otherwise.
"""
import numpy as np
import pytest

import aerial_mapper_b200 as amb
from aerial_mapper_b200 import synth
from test_gpu_ortho import make_inputs

pytestmark = pytest.mark.gpu

NAMES = ("ortho", "elevation", "elevation_angle", "observation_index")


def reference(rows, cols, res, xyz, camd, poses, imgs):
    ref = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.Dsm(amb.DsmSettings(), ref).process(xyz, ref)
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), ref).process(poses, imgs, ref)
    return ref


def mirrored(monkeypatch, rows, cols, res, xyz, camd, poses, imgs, rounds=2):
    monkeypatch.setenv("AMB_COMPACT_MIRRORS", "1")
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res), pinned=True).getMutable()
    gm.to_device(0)
    gm.set_mirrors(NAMES)
    for _ in range(rounds):          # the second round reuses the code buffers and joins the first round's threads
        amb.check(amb.lib().amb_init_layers(gm.context()), gm.context())
        for k in NAMES:
            gm[k][...] = -7.0          # whatever was in the host layer must be overwritten by the mirror
        amb.Dsm(amb.DsmSettings(), gm).process(xyz, gm)
        amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gm).process(poses, imgs, gm)
        gm.sync()
    return gm


def test_compact_mirrors_deliver_the_same_bits(monkeypatch):
    rows, cols, res = 128, 96, 0.5
    xyz = synth.point_cloud(30000, 33.0, 25.0, seed=44, holes=2, hole_sides=(4.0, 9.0))   # NaN elevation -> NaN index cells
    camd, poses, imgs = make_inputs(rows, cols, res, 2, 3, 50.0, 0.08, False)
    ref = reference(rows, cols, res, xyz, camd, poses, imgs)
    gm = mirrored(monkeypatch, rows, cols, res, xyz, camd, poses, imgs)
    assert np.isnan(ref["observation_index"]).any()
    for k in NAMES:
        assert np.array_equal(gm[k].view(np.uint32), ref[k].view(np.uint32)), k


def test_large_map_two_chunks_with_a_ragged_end(monkeypatch):
    # 2300 x 2000 = 4.6 M cells: one full 4 M-cell chunk and a ragged second one (the emulated build uses 3000-cell
    # chunks, so every other test of this file already spans many chunks there)
    import os
    if os.environ.get("AMB_TEST_EMU", "0") not in ("", "0"):
        pytest.skip("too large for the fiber emulation")
    rows, cols, res = 2300, 2000, 0.5
    rng = np.random.default_rng(3)
    xyz = np.c_[rng.uniform(-rows * res / 2, rows * res / 2, 400000), rng.uniform(-cols * res / 2, cols * res / 2, 400000),
                rng.normal(100.0, 2.0, 400000)]
    camd, poses, imgs = make_inputs(rows, cols, res, 2, 2, 600.0, 0.05, False)
    ref = reference(rows, cols, res, xyz, camd, poses, imgs)
    gm = mirrored(monkeypatch, rows, cols, res, xyz, camd, poses, imgs, rounds=1)
    for k in NAMES:
        assert np.array_equal(gm[k].view(np.uint32), ref[k].view(np.uint32)), k


def test_values_without_a_code_travel_as_float32(monkeypatch):
    rows, cols, res = 64, 64, 1.0
    monkeypatch.setenv("AMB_COMPACT_MIRRORS", "1")
    # (a) OrthoFromPcl writes IDW floats into `ortho`
    rng = np.random.default_rng(4)
    xyz = np.c_[rng.uniform(-30, 30, 4000), rng.uniform(-30, 30, 4000), rng.uniform(0, 5, 4000)]
    inten = rng.integers(0, 256, 4000).astype(np.int32)
    ref = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.OrthoFromPcl(amb.OrthoFromPclSettings()).process(xyz, inten, ref)
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res), pinned=True).getMutable()
    gm.to_device(0)
    gm.set_mirrors(("ortho",))
    gm["ortho"][...] = -7.0
    amb.OrthoFromPcl(amb.OrthoFromPclSettings()).process(xyz, inten, gm)
    gm.sync()
    assert np.array_equal(gm["ortho"].view(np.uint32), ref["ortho"].view(np.uint32))
    assert (ref["ortho"] != np.floor(ref["ortho"])).any()
    # (b) 300 frames: observation indices beyond 254 have no code
    rows, cols, res = 96, 64, 0.5
    camd, poses, imgs = make_inputs(rows, cols, res, 10, 30, 60.0, 0.03, False)
    elev = synth.analytic_elevation(rows, cols, res)
    ref = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    ref["elevation"] = elev
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), ref).process(poses, imgs, ref)
    assert np.nanmax(ref["observation_index"]) >= 255
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res), pinned=True).getMutable()
    gm["elevation"] = elev
    gm.to_device(0)
    gm.set_mirrors(("ortho", "elevation_angle", "observation_index"))
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gm).process(poses, imgs, gm)
    gm.sync()
    for k in ("ortho", "elevation_angle", "observation_index"):
        assert np.array_equal(gm[k].view(np.uint32), ref[k].view(np.uint32)), k


def test_only_the_two_integer_layers_accept_the_switch():
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, 8.0, 8.0, 1.0)).getMutable()
    ctx = gm.context()
    L = amb.lib()
    assert L.amb_set_host_mirror_compact(ctx, amb.LAYER_ID["ortho"], 1) == 0
    assert L.amb_set_host_mirror_compact(ctx, amb.LAYER_ID["observation_index"], 1) == 0
    assert L.amb_set_host_mirror_compact(ctx, amb.LAYER_ID["elevation"], 1) == -5          # AMB_ERR_INVALID_ARGUMENT
    assert L.amb_set_host_mirror_compact(ctx, amb.LAYER_ID["colored_ortho"], 1) == -5       # packed colour bits
