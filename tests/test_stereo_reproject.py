""""Next" row N3: stereo::Densifier::computePointCloud (densifier.cpp:25-108), disparity -> world points.
CPU: the oracle restatement against a vectorised numpy evaluation.  GPU: the CUDA kernels (count / scan / ordered
write) against the oracle — bit-identical doubles, identical raster order."""
import numpy as np
import pytest

from oracle import pyoracle as po


def scenario(h, w, seed):
    rng = np.random.default_rng(seed)
    disp = rng.uniform(-2.0, 40.0, (h, w)).astype(np.float32)
    disp[rng.random((h, w)) < 0.2] = -1.0            # block matcher's invalid marker
    disp[rng.random((h, w)) < 0.05] = 1.0            # exactly kMaxInvalidDisparity: not valid (strict >)
    img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    k4 = np.array([1200.5, 1190.25, w / 2 + 0.3, h / 2 - 0.7])
    from scipy.spatial.transform import Rotation as R
    Rm = R.from_euler("xyz", [3.05, 0.04, -0.6]).as_matrix()
    t = np.array([464000.5, 5248000.25, 820.0])
    return disp, img, k4, 0.37, Rm, t


def numpy_reproject(disp, img, k4, baseline, Rm, t):
    h, w = disp.shape
    fx, fy, cx, cy = k4
    v, u = np.mgrid[0:h, 0:w]
    valid = disp > np.float32(1.0)
    d = disp.astype(np.float64)
    with np.errstate(all="ignore"):
        ww = (1.0 / baseline) * d
        x1 = (u + (-cx)) / ww
        y1 = ((fx / fy) * v + (-cy * (fx / fy))) / ww
        z1 = fx / ww
        X = ((Rm[0, 0] * x1 + Rm[0, 1] * y1) + Rm[0, 2] * z1) + t[0]
        Y = ((Rm[1, 0] * x1 + Rm[1, 1] * y1) + Rm[1, 2] * z1) + t[1]
        Z = ((Rm[2, 0] * x1 + Rm[2, 1] * y1) + Rm[2, 2] * z1) + t[2]
        valid &= ~np.isinf(Z.astype(np.float32))
    return np.stack([X[valid], Y[valid], Z[valid]], -1), img[valid].astype(np.int32)


def test_oracle_matches_numpy():
    disp, img, k4, b, Rm, t = scenario(37, 53, 1)
    st, xyz, inten = po.stereo_reproject(disp, img, k4, b, Rm, t)
    nx, ni = numpy_reproject(disp, img, k4, b, Rm, t)
    assert st == 0 and xyz.shape == nx.shape and len(xyz) > 500
    assert np.array_equal(xyz.view(np.uint64), nx.view(np.uint64))   # same IEEE operations in the same order
    assert np.array_equal(inten, ni)
    assert po.stereo_reproject(disp, img, k4, 0.0, Rm, t)[0] == -6    # CHECK_NE(baseline, 0.0)


@pytest.mark.parametrize("h,w,seed", [(37, 53, 1), (240, 376, 6), (1, 7, 4), (64, 300, 7)])
def test_oracle_equals_the_reference_densifier_cpp(h, w, seed):
    # oracle/_ref/libamb_refsrc_stereo.so: the reference's own densifier.cpp compiled verbatim against stand-in
    # third-party headers (oracle/refsrc_stubs/amb_refsrc_stereo_deps.h): same points, same order, same bits
    if not po.have_refsrc_stereo():
        pytest.skip("oracle/_ref/libamb_refsrc_stereo.so not built (needs /root/reference)")
    disp, img, k4, b, Rm, t = scenario(h, w, seed)
    disp[0, :3] = (np.float32(1.0), np.nextafter(np.float32(1.0), np.float32(2.0)), np.float32(0.0))
    st, xyz, inten = po.stereo_reproject(disp, img, k4, b, Rm, t)
    st2, rxyz, rinten = po.refsrc_stereo_reproject(disp, img, k4, b, Rm, t)
    assert st == 0 and st2 == 0 and xyz.shape == rxyz.shape and len(xyz) > 0
    assert np.array_equal(xyz.view(np.uint64), rxyz.view(np.uint64))
    assert np.array_equal(inten, rinten)
    # an (almost) infinite depth: the float copy of z overflows -> the reference drops the point (:78)
    d2 = np.full((2, 3), 2.0, np.float32)
    d2[0, 0] = np.float32(1e-38)                      # valid only if > kMaxInvalidDisparity: it is not
    d2[1, 1] = np.float32(1.0000001)
    i2 = np.arange(6, dtype=np.uint8).reshape(2, 3)
    a = po.stereo_reproject(d2, i2, k4, b, Rm * 1e37, t)
    r = po.refsrc_stereo_reproject(d2, i2, k4, b, Rm * 1e37, t)
    assert a[0] == 0 and r[0] == 0 and np.array_equal(a[1].view(np.uint64), r[1].view(np.uint64))
    assert np.array_equal(a[2], r[2]) and len(a[2]) < 5   # z overflowed float32 for the kept disparities


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,seed", [(37, 53, 2), (480, 752, 3), (1, 7, 4), (300, 1000, 5)])
def test_gpu_matches_oracle_bit_for_bit(h, w, seed):
    import aerial_mapper_b200 as amb
    disp, img, k4, b, Rm, t = scenario(h, w, seed)
    xyz, inten = amb.compute_point_cloud(disp, img, k4, b, Rm, t)
    st, oxyz, ointen = po.stereo_reproject(disp, img, k4, b, Rm, t)
    assert st == 0 and xyz.shape == oxyz.shape
    assert np.array_equal(xyz.view(np.uint64), oxyz.view(np.uint64))
    assert np.array_equal(inten, ointen)
    # and the points feed the DSM exactly like a host cloud does
    if h >= 300:
        K = np.array([[k4[0], 0, k4[2]], [0, k4[1], k4[3]], [0, 0, 1]])
        x2, i2 = amb.compute_point_cloud(disp, img, K, b, Rm, t)
        assert np.array_equal(x2, xyz) and np.array_equal(i2, inten)
