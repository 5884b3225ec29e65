"""The C++ drop-in headers (aerial_mapper_b200/shim: dsm::Dsm, ortho::OrthoBackwardGrid with the reference's
signatures) compile against the C ABI; on a GPU the demo's call sequence through them gives the same layers as the
Python mirror."""
import os
import subprocess

import numpy as np
import pytest

from common import ROOT
import aerial_mapper_b200 as amb
from aerial_mapper_b200 import synth


def build_demo(tmp_path):
    exe = str(tmp_path / "shim_demo")
    libdir = os.path.join(ROOT, "aerial_mapper_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-DAMB_SHIM_MINI",
                           "-I" + os.path.join(libdir, "shim"), os.path.join(ROOT, "tests", "cpp", "shim_demo.cc"),
                           "-o", exe, "-L" + libdir, "-laerial_mapper_b200", "-Wl,-rpath," + libdir])
    return exe


def test_shim_headers_compile_and_link(tmp_path):
    amb.lib()
    assert os.path.exists(build_demo(tmp_path))
    # the OrthoFromPcl drop-in has its own ortho::Settings: separate translation unit, like in the reference
    src = tmp_path / "pcl.cc"
    src.write_text("#include <aerial-mapper-ortho/ortho-from-pcl.h>\n"
                   "int main() { ortho::Settings s; ortho::OrthoFromPcl o(s); (void)o; return 0; }\n")
    libdir = os.path.join(ROOT, "aerial_mapper_b200")
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-DAMB_SHIM_MINI",
                           "-I" + os.path.join(libdir, "shim"), str(src), "-o", str(tmp_path / "pcl"),
                           "-L" + libdir, "-laerial_mapper_b200", "-Wl,-rpath," + libdir])


@pytest.mark.gpu
@pytest.mark.parametrize("colored", [False, True])
def test_demo_sequence_through_the_cpp_shim(tmp_path, colored):
    exe = build_demo(tmp_path)
    rows, cols, res = 120, 90, 0.5
    xyz = synth.point_cloud(30000, rows * res / 2, cols * res / 2, seed=51, holes=2, hole_sides=(3.0, 9.0))
    camd = synth.scaled_camera(0.06)
    poses = synth.lawnmower_poses(2, 3, rows * res / 2, cols * res / 2, 50.0, seed=52, jitter_pos=0.5)
    ch = 3 if colored else 1
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], ch) for k in range(len(poses))]
    scen = tmp_path / "scenario.bin"
    with open(scen, "wb") as f:
        np.array([rows * res, cols * res, res, 0.0, 0.0], np.float64).tofile(f)
        np.array([len(xyz), len(poses), camd["width"], camd["height"], ch, camd["dist_type"]], np.int64).tofile(f)
        np.array([camd["fu"], camd["fv"], camd["cu"], camd["cv"]] + list(camd["dist"]), np.float64).tofile(f)
        xyz.tofile(f)
        poses.tofile(f)
        for im in imgs:
            im.tofile(f)
    out = tmp_path / "layers.bin"
    r = subprocess.run([exe, str(scen), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out, np.float32).reshape(5, cols, rows).transpose(0, 2, 1)  # column-major layers

    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.Dsm(amb.DsmSettings(), gm).process(xyz, gm)
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(colored_ortho=colored), gm).process(poses, imgs, gm)
    for k, name in enumerate(["elevation", "elevation_angle", "observation_index", "ortho", "colored_ortho"]):
        assert np.array_equal(got[k].view(np.uint32), gm[name].view(np.uint32)), name
    assert np.isnan(got[0]).any() and (~np.isnan(got[2])).any()
