"""GPU parity against the REFERENCE'S OWN code: the CUDA library through the C ABI vs dsm::Dsm::process,
ortho::OrthoBackwardGrid::process and ortho::OrthoFromPcl::process from the reference's dsm.cc /
ortho-backward-grid.cc / ortho-from-pcl.cc, compiled verbatim into oracle/_ref (see tests/test_oracle_refsrc.py and
oracle/refsrc_stubs/amb_refsrc_deps.h).  The libraries are built where /root/reference exists and travel to the GPU
box prebuilt; nothing here reads /root/reference at run time.

Bars: NaN mask, frame indices and pixel values bit-exact; heights and angles within one float32 ulp (the CUDA path
sums a cell's neighbours in point-index order, the reference in kd-tree traversal order)."""
import numpy as np
import pytest

from common import fresh_layers, ulp_diff
import aerial_mapper_b200 as amb
from aerial_mapper_b200 import synth
from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not po.have_refsrc(), reason="oracle/_ref/libamb_refsrc_*.so not present")]


@pytest.mark.parametrize("rows,cols,res,n,holes,radius,ce,cn", [(256, 256, 1.0, 100000, 0, 1, 0.0, 0.0),   # config C1
                                                                (300, 200, 0.25, 40000, 6, 1, 0.0, 0.0),
                                                                (100, 130, 0.5, 20000, 2, 2, 12.5, -7.25)])
def test_dsm_equals_reference_dsm_cc(rows, cols, res, n, holes, radius, ce, cn):
    xyz = synth.point_cloud(n, rows * res / 2, cols * res / 2, seed=61, holes=holes, hole_sides=(2.0, 10.0))
    xyz[:, 0] += cn
    xyz[:, 1] += ce
    gm = amb.AerialGridMap(amb.GridMapSettings(0.0, 0.0, rows * res, cols * res, res)).getMutable()
    amb.Dsm(amb.DsmSettings(interpolation_radius=radius, center_easting=ce, center_northing=cn), gm).process(xyz, gm)
    e = np.full((rows, cols), np.nan, np.float32, order="F")
    st, _ = po.refsrc_dsm_process(po.make_geometry(rows, cols, res), e, xyz, radius, ce, cn)
    assert st == 0, po.refsrc_last_error()
    assert np.array_equal(np.isnan(gm["elevation"]), np.isnan(e))
    assert ulp_diff(gm["elevation"], e).max() <= 1
    if holes:
        assert np.isnan(e).any()


@pytest.mark.parametrize("colored,dist_type", [(False, 1), (True, 2), (True, 0)])
def test_dsm_then_ortho_equals_reference_sources(colored, dist_type):
    # the batch demo's order (main-ortho-backward-grid.cc:129-141) on both sides: DSM, then the orthomosaic over it
    rows, cols, res = 160, 128, 0.5
    dist = {0: (0, 0, 0, 0), 1: (-0.05, 0.01, 1e-4, 1e-4), 2: (0.01, -0.002, 0.0005, -0.0001)}[dist_type]
    camd = synth.scaled_camera(0.08, dist_type=dist_type, dist=dist)
    poses = synth.lawnmower_poses(2, 4, rows * res / 2, cols * res / 2, 50.0, 63, jitter_pos=0.5)
    ch = 3 if colored else 1
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], ch) for k in range(len(poses))]
    xyz = synth.point_cloud(60000, rows * res / 2 + 1, cols * res / 2 + 1, seed=62, holes=2, hole_sides=(4.0, 12.0))

    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.Dsm(amb.DsmSettings(), gm).process(xyz, gm)
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(colored_ortho=colored), gm).process(poses, imgs, gm)

    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)
    L = fresh_layers(rows, cols)
    assert po.refsrc_dsm_process(g, L["elevation"], xyz)[0] == 0
    assert ulp_diff(gm["elevation"], L["elevation"]).max() <= 1
    # the orthomosaic reads float32 elevation: give the reference the GPU's layer so that a 1-ulp height difference
    # cannot move a frame decision, then every output must be exact
    L["elevation"][...] = gm["elevation"]
    st, _ = po.refsrc_ortho_process(g, L, cam, poses, imgs, colored=colored)
    assert st == 0, po.refsrc_last_error()
    a, b = gm["observation_index"], L["observation_index"]
    assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()
    for k in ("ortho", "colored_ortho"):
        assert np.array_equal(gm[k].view(np.uint32), L[k].view(np.uint32)), k
    assert ulp_diff(gm["elevation_angle"], L["elevation_angle"]).max() <= 1
    assert (~np.isnan(a)).mean() > 0.5


def test_ortho_from_pcl_equals_reference_cc():
    rows, cols, res = 120, 90, 0.5
    xyz = synth.point_cloud(15000, rows * res / 2, cols * res / 2, seed=64, holes=2, hole_sides=(3.0, 9.0))
    inten = np.random.default_rng(6).integers(0, 256, len(xyz)).astype(np.int32)
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.OrthoFromPcl(amb.OrthoFromPclSettings(interpolation_radius=2)).process(xyz, inten, gm)
    o = np.full((rows, cols), 255.0, np.float32, order="F")
    assert po.refsrc_ortho_from_pcl_process(po.make_geometry(rows, cols, res), o, xyz, inten, 2, False) == 0
    assert np.array_equal(gm["ortho"] == 255.0, o == 255.0)
    assert ulp_diff(gm["ortho"], o).max() <= 1


def test_full_size_c3_stripe_equals_reference_sources():
    """BASELINE config C3 at FULL size — 250 frames of 4000x3000 over 8000x8000 @ 0.5 m — on a 24-column stripe: the CUDA
    path (stripe context, host frames) against the reference's own ortho-backward-grid.cc restricted to the same cells.
    Every frame index and every pixel must be identical."""
    rows = cols = 8000
    res = 0.5
    c0, c1 = 3988, 4012
    camd = dict(synth.C3_CAMERA)
    poses = synth.lawnmower_poses(10, 25, rows * res / 2, cols * res / 2, 600.0, seed=4)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(len(poses))]
    qx, qy = synth.grid_positions(rows, cols, res)
    elev_stripe = np.asfortranarray(synth.terrain(qx[:, None], qy[None, c0:c1]).astype(np.float32))

    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res),
                           layer_names=("ortho", "elevation", "elevation_angle", "observation_index")).getMutable()
    gm.layers["elevation"][:, c0:c1] = elev_stripe
    gm.context(0, (c0, c1))
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gm).process(poses, imgs, gm)

    L = {"elevation": np.full((rows, cols), np.nan, np.float32, order="F"),
         "elevation_angle": np.zeros((rows, cols), np.float32, order="F"),
         "observation_index": np.full((rows, cols), np.nan, np.float32, order="F"),
         "ortho": np.full((rows, cols), 255.0, np.float32, order="F")}
    L["elevation"][:, c0:c1] = elev_stripe
    st, _ = po.refsrc_ortho_process(po.make_geometry(rows, cols, res), L, po.make_camera(**camd), poses, imgs,
                                    cell_range=(rows * c0, rows * c1))
    assert st == 0, po.refsrc_last_error()
    a, b = gm["observation_index"][:, c0:c1], L["observation_index"][:, c0:c1]
    assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()
    assert np.array_equal(gm["ortho"][:, c0:c1].view(np.uint32), L["ortho"][:, c0:c1].view(np.uint32))
    assert ulp_diff(gm["elevation_angle"][:, c0:c1], L["elevation_angle"][:, c0:c1]).max() <= 1
    assert (~np.isnan(b)).all() and len(np.unique(b)) > 20      # every cell seen; many different winners along the stripe
