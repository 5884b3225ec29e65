"""Pageable caller memory (numpy arrays, like the reference's std::vector / cv::Mat / Eigen storage) goes through the
library's own staging (csrc/host_staging.cu: worker pool + pinned slots, chunked) once a transfer is >= 4 MB: the results
must be bit-identical to the same job with the inputs already on the device / in pinned memory."""
import numpy as np
import pytest

import aerial_mapper_b200 as amb
from aerial_mapper_b200 import synth

pytestmark = pytest.mark.gpu


def test_pageable_cloud_layers_and_frames_equal_the_device_path():
    import torch
    rows, cols, res = 1200, 1100, 0.25          # 5.3 MB layers, staged in both directions
    n = 700_000                                 # 16.8 MB cloud: several chunks would need > 32 MB; one partial chunk here
    xyz = synth.point_cloud(n, rows * res / 2, cols * res / 2, seed=71, holes=3, hole_sides=(2.0, 6.0))
    n = len(xyz)                                # (the holes removed some)
    camd = synth.scaled_camera(0.5)             # 2000 x 1500 frames (3 MB each); rectangles are packed by the pool
    poses = synth.lawnmower_poses(2, 3, rows * res / 2, cols * res / 2, 150.0, seed=72)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], 1) for k in range(len(poses))]

    # pageable: numpy everywhere, host layers authoritative (upload + download around every process())
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.Dsm(amb.DsmSettings(), gm).process(xyz, gm)
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gm).process(poses, imgs, gm)

    # device path: inputs resident in HBM, layers resident, one download at the end
    dev = torch.device("cuda:0")
    xyz_d = torch.from_numpy(xyz).to(dev)
    imgs_d = [torch.from_numpy(im).to(dev) for im in imgs]
    gd = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    gd.to_device(0)
    amb.Dsm(amb.DsmSettings(), gd).process_device(xyz_d.data_ptr(), n, gd)
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gd).process_device(
        poses, [t.data_ptr() for t in imgs_d], camd["width"], gd)
    gd.sync()
    gd.download(("elevation", "elevation_angle", "observation_index", "ortho"))
    for name in ("elevation", "elevation_angle", "observation_index", "ortho"):
        assert np.array_equal(gm[name].view(np.uint32), gd[name].view(np.uint32)), name
    assert np.isfinite(gm["elevation"]).mean() > 0.99 and (~np.isnan(gm["observation_index"])).mean() > 0.5


def test_large_pageable_cloud_spans_several_staging_chunks():
    import torch
    rows, cols, res = 600, 600, 0.5
    n = 3_000_000                               # 72 MB: three 32 MB chunks, the slots are reused
    xyz = synth.point_cloud(n, rows * res / 2, cols * res / 2, seed=73)
    n = len(xyz)
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.Dsm(amb.DsmSettings(), gm).process(xyz, gm)
    gd = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    gd.to_device(0)
    xyz_d = torch.from_numpy(xyz).to(torch.device("cuda:0"))
    amb.Dsm(amb.DsmSettings(), gd).process_device(xyz_d.data_ptr(), n, gd)
    gd.sync()
    gd.download(("elevation",))
    assert np.array_equal(gm["elevation"].view(np.uint32), gd["elevation"].view(np.uint32))
