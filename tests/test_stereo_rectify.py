""""Next" row N3, second half: stereo::Rectifier::rectifyStereoPair (rectifier.cpp:36-107) — Fusiello's
compact rectification on the host, then the per-pixel fill of the four CV_32FC1 rectification maps.
CPU: the oracle restatement against numpy linear algebra and the geometric properties rectification must have
(common rows, positive disparity), and the product's host half (plain C++ inside the CUDA library, runs without a
GPU) bit for bit against the oracle.  GPU: the map-fill kernel against the oracle, bit-identical floats.
"""
import numpy as np
import pytest

import aerial_mapper_b200 as amb
from aerial_mapper_b200 import _lib
from oracle import pyoracle as po


def rig(seed, w=640, h=480):
    """Two nadir-looking frames of a fixed-wing survey: ~25 m apart along track, small attitude differences."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(seed)
    K = np.array([[0.9 * w, 0.0, w / 2 + 0.4], [0.0, 0.92 * w, h / 2 - 0.3], [0.0, 0.0, 1.0]])
    nadir = R.from_euler("x", np.pi)                       # camera z-axis = -Z world
    yaw = rng.uniform(-np.pi, np.pi)
    R1 = (R.from_euler("z", yaw) * nadir * R.from_euler("xyz", rng.normal(0, 0.03, 3))).as_matrix()
    R2 = (R.from_euler("z", yaw) * nadir * R.from_euler("xyz", rng.normal(0, 0.03, 3))).as_matrix()
    t1 = np.array([464000.0, 5248000.0, 900.0]) + rng.normal(0, 5.0, 3)
    # camera x-axis in the world: moving along it makes camera 1 the LEFT camera (rectifier.cpp:44-46)
    t2 = t1 + R1[:, 0] * rng.uniform(15.0, 35.0) + rng.normal(0, 0.8, 3)
    return K, R1, R2, t1, t2


def numpy_setup(K, R1, R2, t1, t2):
    x = t2 - t1
    y = np.cross(R1[:, 2], x)
    z = np.cross(x, y)
    Rr = np.stack([x / np.linalg.norm(x), y / np.linalg.norm(y), z / np.linalg.norm(z)])
    T1 = (K @ Rr) @ np.linalg.inv(K @ R1.T)
    T2 = (K @ Rr) @ np.linalg.inv(K @ R2.T)
    return np.linalg.norm(x), Rr, T1, T2


@pytest.mark.parametrize("seed", range(6))
def test_oracle_setup_matches_numpy_and_rectifies(seed):
    K, R1, R2, t1, t2 = rig(seed)
    st, base, Rr, T1i, T2i = po.stereo_rectify_setup(K, R1, R2, t1, t2)
    assert st == 0
    nb, nR, nT1, nT2 = numpy_setup(K, R1, R2, t1, t2)
    assert abs(base - nb) <= 1e-12 * nb
    assert np.allclose(Rr, nR, rtol=0, atol=1e-13)
    assert np.allclose(Rr @ Rr.T, np.eye(3), atol=1e-13) and abs(np.linalg.det(Rr) - 1.0) < 1e-12
    assert np.allclose(Rr @ (t2 - t1) / base, [1.0, 0.0, 0.0], atol=1e-12)     # new x axis = baseline direction
    assert T1i.dtype == np.float32
    assert np.allclose(T1i, np.linalg.inv(nT1), rtol=2e-6, atol=1e-9)
    assert np.allclose(T2i, np.linalg.inv(nT2), rtol=2e-6, atol=1e-9)
    # what rectification is for: a world point lands on the SAME ROW in both rectified images, left of itself in the
    # right image (disparity = fx * baseline / depth > 0 — the quantity Densifier::computePointCloud inverts)
    rng = np.random.default_rng(100 + seed)
    X = np.c_[t1[0] + rng.uniform(-150, 150, 50), t1[1] + rng.uniform(-150, 150, 50), rng.uniform(380, 420, 50)]
    for Xw in X:
        r = []
        for Ri, ti, Ti in ((R1, t1, nT1), (R2, t2, nT2)):
            p = Ti @ (K @ (Ri.T @ (Xw - ti)))
            r.append(p[:2] / p[2])
        depth = (Rr @ (Xw - t1))[2]
        assert abs(r[0][1] - r[1][1]) < 1e-6
        assert abs((r[0][0] - r[1][0]) - K[0, 0] * base / depth) < 1e-6 and r[0][0] > r[1][0]


def test_oracle_setup_contract_violations():
    K, R1, R2, t1, t2 = rig(0)
    assert po.stereo_rectify_setup(K, R1, R2, t1, t1)[0] == _lib.AMB_ERR_CHECK_FAILED       # zero baseline
    assert po.stereo_rectify_setup(np.zeros((3, 3)), R1, R2, t1, t2)[0] == _lib.AMB_ERR_CHECK_FAILED  # singular K


@pytest.mark.parametrize("w,h,seed", [(64, 48, 1), (640, 480, 2), (333, 77, 3)])
def test_oracle_maps_match_a_float64_evaluation(w, h, seed):
    K, R1, R2, t1, t2 = rig(seed, w, h)
    st, _, _, T1i, T2i = po.stereo_rectify_setup(K, R1, R2, t1, t2)
    assert st == 0
    st, maps = po.stereo_rectify_maps(T1i, T2i, w, h)
    assert st == 0
    v, u = np.mgrid[0:h, 0:w].astype(np.float64)
    for Ti, mx, my in ((T1i, maps[0], maps[1]), (T2i, maps[2], maps[3])):
        T = Ti.astype(np.float64)
        x = T[0, 0] * u + T[0, 1] * v + T[0, 2]
        y = T[1, 0] * u + T[1, 1] * v + T[1, 2]
        ww = T[2, 0] * u + T[2, 1] * v + T[2, 2]
        assert mx.dtype == np.float32 and mx.shape == (h, w)
        assert np.abs(mx - x / ww).max() < 4e-7 * max(w, h) and np.abs(my - y / ww).max() < 4e-7 * max(w, h)
    # the maps send rectified pixels back into the original frame: near-identity for a nearly-rectified pair
    assert np.abs(maps[0] - u).max() < 0.25 * w and np.abs(maps[1] - v).max() < 0.25 * w


def test_oracle_maps_report_a_vanishing_w():
    T = np.eye(3, dtype=np.float32)
    T[2] = (1.0, 0.0, -5.0)                       # w = u - 5: zero at u == 5
    st, _ = po.stereo_rectify_maps(T, np.eye(3, dtype=np.float32), 16, 4)
    assert st == _lib.AMB_ERR_CHECK_FAILED        # CHECK_NE(xyw_1(2), 0.0), rectifier.cpp:92


@pytest.mark.parametrize("seed", range(8))
def test_product_host_half_equals_the_oracle_bit_for_bit(seed):
    K, R1, R2, t1, t2 = rig(seed)
    st, base, Rr, T1i, T2i = po.stereo_rectify_setup(K, R1, R2, t1, t2)
    assert st == 0
    b, R, T1, T2 = amb.rectify_stereo_setup(K, R1, R2, t1, t2)    # plain C++ in libaerial_mapper_b200.so: no GPU needed
    assert b == base and np.array_equal(R, Rr)
    assert np.array_equal(T1.view(np.uint32), T1i.view(np.uint32)) and np.array_equal(T2.view(np.uint32), T2i.view(np.uint32))
    with pytest.raises(amb.AmbError) as ei:
        amb.rectify_stereo_setup(K, R1, R2, t1, t1)
    assert ei.value.status == _lib.AMB_ERR_CHECK_FAILED


def test_map_fill_fails_loudly_without_a_gpu(gpu_count):
    if gpu_count > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(amb.AmbError) as ei:
        amb.rectify_stereo_maps(np.eye(3), np.eye(3), 32, 16)
    assert ei.value.status == _lib.AMB_ERR_NO_DEVICE


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,seed", [(640, 480, 2), (333, 77, 3), (4000, 3000, 4), (5, 3, 5)])
def test_gpu_maps_bit_identical_to_the_oracle(w, h, seed):
    K, R1, R2, t1, t2 = rig(seed, w, h)
    _, _, T1, T2 = amb.rectify_stereo_setup(K, R1, R2, t1, t2)
    got = amb.rectify_stereo_maps(T1, T2, w, h)
    st, want = po.stereo_rectify_maps(T1, T2, w, h)
    assert st == 0
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    Tz = np.eye(3, dtype=np.float32)
    Tz[2] = (1.0, 0.0, -2.0)
    with pytest.raises(amb.AmbError) as ei:
        amb.rectify_stereo_maps(Tz, T2, w, h)
    assert ei.value.status == _lib.AMB_ERR_CHECK_FAILED
