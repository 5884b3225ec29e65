"""The C++ drop-in headers (aerial_mapper_b200/shim: dsm::Dsm, ortho::OrthoBackwardGrid with the reference's
signatures) compile against the C ABI; on a GPU the demo's call sequence through them gives the same layers as the
Python mirror.

Source-level drop-in: tests/cpp/shim_demo.cc — caller code that only uses the reference's public API, the batch demo's
call sequence — is ONE source with two builds: against the drop-in headers + CUDA library (here), and against the
reference's OWN headers and sources (dsm.cc, ortho-backward-grid.cc compiled verbatim; oracle/_ref/
libamb_reference_demo.so).  Swapping the include path and the link line is the whole integration."""
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


LAYERS = ["elevation", "elevation_angle", "observation_index", "ortho", "colored_ortho"]


def make_scenario(tmp_path, colored, rows=120, cols=90, res=0.5):
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
    return scen, (rows, cols, res, xyz, camd, poses, imgs)


def read_layers(path, rows, cols):
    return np.fromfile(path, np.float32).reshape(5, cols, rows).transpose(0, 2, 1)  # column-major layers


def run_reference_demo(tmp_path, scen, rows, cols):
    from oracle import pyoracle as po
    out = tmp_path / "layers_reference.bin"
    assert po.reference_demo_main(str(scen), str(out)) == 0
    return read_layers(out, rows, cols)


@pytest.mark.parametrize("colored", [False, True])
def test_same_demo_source_runs_on_the_reference_sources(tmp_path, colored):
    # CPU: the demo source built against the reference's own dsm.cc / ortho-backward-grid.cc gives, bit for bit, what
    # the reference classes give when driven through the oracle glue (and hence what the restated oracle gives)
    from common import fresh_layers
    from oracle import pyoracle as po
    if not po.have_reference_demo():
        pytest.skip("oracle/_ref/libamb_reference_demo.so not built (needs /root/reference)")
    scen, (rows, cols, res, xyz, camd, poses, imgs) = make_scenario(tmp_path, colored)
    got = run_reference_demo(tmp_path, scen, rows, cols)
    g = po.make_geometry(rows, cols, res)
    L = fresh_layers(rows, cols)
    assert po.refsrc_dsm_process(g, L["elevation"], xyz)[0] == 0
    assert po.refsrc_ortho_process(g, L, po.make_camera(**camd), poses, imgs, colored=colored)[0] == 0
    for k, name in enumerate(LAYERS):
        assert np.array_equal(got[k].view(np.uint32), L[name].view(np.uint32)), name
    assert np.isnan(got[0]).any() and (~np.isnan(got[2])).mean() > 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("colored", [False, True])
def test_one_demo_source_two_builds_cuda_vs_reference(tmp_path, colored):
    # the same caller source: drop-in headers + CUDA library vs the reference's headers + sources
    from common import ulp_diff
    from oracle import pyoracle as po
    if not po.have_reference_demo():
        pytest.skip("oracle/_ref/libamb_reference_demo.so not present")
    exe = build_demo(tmp_path)
    scen, (rows, cols, res, xyz, camd, poses, imgs) = make_scenario(tmp_path, colored)
    out = tmp_path / "layers.bin"
    r = subprocess.run([exe, str(scen), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cuda = read_layers(out, rows, cols)
    ref = run_reference_demo(tmp_path, scen, rows, cols)
    assert np.array_equal(np.isnan(cuda[0]), np.isnan(ref[0]))
    assert ulp_diff(cuda[0], ref[0]).max() <= 1                       # elevation: summation order only
    # each side's orthomosaic runs on its own elevation; a 1-ulp height difference can move a frame decision only in
    # a cell where two frames tie to ~1e-8 in observation angle: allow a couple of such cells, everything else exact
    a, b = cuda[2], ref[2]
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    assert (~same).sum() <= 2
    for k in (3, 4):
        assert (cuda[k].view(np.uint32) != ref[k].view(np.uint32))[same].sum() == 0, LAYERS[k]
    assert ulp_diff(cuda[1][same], ref[1][same]).max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("colored", [False, True])
def test_demo_sequence_through_the_cpp_shim(tmp_path, colored):
    exe = build_demo(tmp_path)
    scen, (rows, cols, res, xyz, camd, poses, imgs) = make_scenario(tmp_path, colored)
    out = tmp_path / "layers.bin"
    r = subprocess.run([exe, str(scen), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = read_layers(out, rows, cols)

    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    amb.Dsm(amb.DsmSettings(), gm).process(xyz, gm)
    amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(colored_ortho=colored), gm).process(poses, imgs, gm)
    for k, name in enumerate(["elevation", "elevation_angle", "observation_index", "ortho", "colored_ortho"]):
        assert np.array_equal(got[k].view(np.uint32), gm[name].view(np.uint32)), name
    assert np.isnan(got[0]).any() and (~np.isnan(got[2])).any()


@pytest.mark.gpu
@pytest.mark.parametrize("colored", [False, True])
def test_demo_through_the_shim_on_several_gpus_of_one_process(tmp_path, colored, gpu_count):
    """AMB_SHIM_GPUS=N: the drop-in classes spread the map's column stripes over N GPUs of the demo process (amb_multi_*:
    one context per device, the whole cloud to every device, all frames per stripe) — every layer bit-identical to the
    single-GPU run of the same binary.  Needs >= 2 devices (e.g. `gpurun --gpus 2`)."""
    if gpu_count < 2:
        pytest.skip("needs at least two CUDA devices")
    exe = build_demo(tmp_path)
    scen, (rows, cols, res, xyz, camd, poses, imgs) = make_scenario(tmp_path, colored, rows=200, cols=160)
    outs = []
    for gpus in (1, min(gpu_count, 4)):
        out = tmp_path / ("layers_%d.bin" % gpus)
        env = dict(os.environ, AMB_SHIM_GPUS=str(gpus))
        r = subprocess.run([exe, str(scen), str(out)], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs.append(read_layers(out, rows, cols))
    for k, name in enumerate(LAYERS):
        assert np.array_equal(outs[0][k].view(np.uint32), outs[1][k].view(np.uint32)), name
    assert np.isnan(outs[0][0]).any() and (~np.isnan(outs[0][2])).any()
