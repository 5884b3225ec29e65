"""CPU tests of the orthomosaic oracle (ortho-backward-grid.cc:128-239 restated): camera model against OpenCV,
pose algebra against scipy, the cell loop against an independent numpy evaluation, and the path's invariants."""
import numpy as np
import pytest

from common import fresh_layers, ulp_diff
from aerial_mapper_b200 import synth
from oracle import pyoracle as po

RADTAN = (-0.05, 0.01, 1e-4, 1e-4)
EQUI = (0.01, -0.002, 0.0005, -0.0001)


def test_radtan_projection_matches_opencv():
    import cv2
    cam = po.make_camera(640, 480, 480.0, 470.0, 321.5, 239.25, 1, RADTAN)
    rng = np.random.default_rng(0)
    P = np.c_[rng.uniform(-2, 2, 400), rng.uniform(-1.5, 1.5, 400), rng.uniform(1.0, 6.0, 400)]
    K = np.array([[480.0, 0, 321.5], [0, 470.0, 239.25], [0, 0, 1]])
    ref, _ = cv2.projectPoints(P, np.zeros(3), np.zeros(3), K, np.array(RADTAN))
    ref = ref.reshape(-1, 2)
    for p, r in zip(P, ref):
        vis, kp = po.project3(cam, p)
        assert np.allclose(kp, r, rtol=0, atol=1e-9)
        assert vis == (0 <= kp[0] < 640 and 0 <= kp[1] < 480)


def test_equidistant_projection_matches_opencv_fisheye():
    import cv2
    cam = po.make_camera(640, 480, 300.0, 300.0, 320.0, 240.0, 2, EQUI)
    rng = np.random.default_rng(1)
    P = np.c_[rng.uniform(-2, 2, 300), rng.uniform(-1.5, 1.5, 300), rng.uniform(0.5, 4.0, 300)]
    K = np.array([[300.0, 0, 320.0], [0, 300.0, 240.0], [0, 0, 1]])
    ref, _ = cv2.fisheye.projectPoints(P.reshape(-1, 1, 3), np.zeros(3), np.zeros(3), K, np.array(EQUI))
    ref = ref.reshape(-1, 2)
    for p, r in zip(P, ref):
        _, kp = po.project3(cam, p)
        assert np.allclose(kp, r, rtol=0, atol=1e-9)


def test_projection_status_rules():
    cam = po.make_camera(100, 80, 50.0, 50.0, 50.0, 40.0)
    assert po.project3(cam, [0.0, 0.0, 1.0])[0]
    assert not po.project3(cam, [0.0, 0.0, -1.0])[0]       # POINT_BEHIND_CAMERA
    assert not po.project3(cam, [0.0, 0.0, 5e-11])[0]      # PROJECTION_INVALID (z <= 1e-10)
    assert not po.project3(cam, [-1.0001, 0.0, 1.0])[0]    # kx < 0
    assert po.project3(cam, [-1.0, 0.0, 1.0])[0]           # kx == 0 is inside
    assert not po.project3(cam, [1.0, 0.0, 1.0])[0]        # kx == width is outside


def test_pose_algebra_matches_scipy():
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(2)
    for _ in range(50):
        qb = rng.normal(size=4); qb /= np.linalg.norm(qb)
        qc = rng.normal(size=4); qc /= np.linalg.norm(qc)
        tb, tc, p = rng.normal(size=3) * 50, rng.normal(size=3), rng.normal(size=3) * 100
        cam = po.make_camera(10, 10, 1, 1, 1, 1, q_C_B=qc, t_C_B=tc)
        got = po.transform_to_camera(cam, np.r_[tb, qb], p)
        R_G_B = R.from_quat([qb[1], qb[2], qb[3], qb[0]]).as_matrix()
        R_C_B = R.from_quat([qc[1], qc[2], qc[3], qc[0]]).as_matrix()
        p_B = R_G_B.T @ (p - tb)          # T_G_B^-1
        want = R_C_B @ p_B + tc           # T_C_B
        assert np.allclose(got, want, rtol=0, atol=1e-10)


def test_color_packing_is_0x00RRGGBB_for_every_byte():
    for c in range(256):
        f = np.float32(np.float64(np.float32(c)) / 255.0)
        assert int(np.float32(f * np.float32(255.0))) == c  # the cast chain of colorVectorToValue is lossless
        assert po.pack_color(c, 0, 0) == c
        assert po.pack_color(0, c, 0) == c << 8
        assert po.pack_color(0, 0, c) == c << 16
    assert po.pack_color(0x12, 0x34, 0x56) == 0x563412


def numpy_ortho(rows, cols, res, elevation, camd, poses, imgs, colored, state=None):
    """Independent vectorised evaluation: rotation matrices from scipy, per-frame sequential update."""
    from scipy.spatial.transform import Rotation as R
    qx, qy = synth.grid_positions(rows, cols, res)
    X = np.broadcast_to(qx[:, None], (rows, cols)).astype(np.float64)
    Y = np.broadcast_to(qy[None, :], (rows, cols)).astype(np.float64)
    Z = elevation.astype(np.float64)
    L = state if state is not None else fresh_layers(rows, cols, elevation)
    k1, k2, p1, p2 = camd["dist"]
    for f, pose in enumerate(poses):
        Rm = R.from_quat([pose[4], pose[5], pose[6], pose[3]]).as_matrix()
        d = np.stack([X - pose[0], Y - pose[1], Z - pose[2]], -1)
        c = d @ Rm  # R^T d
        x, y, z = c[..., 0], c[..., 1], c[..., 2]
        with np.errstate(all="ignore"):
            u, v = x / z, y / z
            r2 = u * u + v * v
            rad = k1 * r2 + k2 * r2 * r2
            ud = u + u * rad + 2 * p1 * u * v + p2 * (r2 + 2 * u * u)
            vd = v + v * rad + 2 * p2 * u * v + p1 * (r2 + 2 * v * v)
            kx = camd["fu"] * ud + camd["cu"]
            ky = camd["fv"] * vd + camd["cv"]
            vis = (kx >= 0) & (ky >= 0) & (kx < camd["width"]) & (ky < camd["height"]) & (z > 1e-10)
            alpha = np.arcsin(np.abs(z) / np.sqrt(x * x + y * y + z * z))
            upd = vis & (alpha > L["elevation_angle"].astype(np.float64))
        kxs, kys = np.where(upd, kx, 0.0), np.where(upd, ky, 0.0)
        px = np.minimum(np.floor(kxs + 0.5).astype(np.int64), camd["width"] - 1)   # round half away (kx >= 0)
        py = np.minimum(np.floor(kys + 0.5).astype(np.int64), camd["height"] - 1)
        ii, jj = np.nonzero(upd)
        L["elevation_angle"][ii, jj] = alpha[ii, jj].astype(np.float32)
        L["observation_index"][ii, jj] = np.float32(f)
        if colored:
            pix = imgs[f][py[ii, jj], px[ii, jj]].astype(np.uint32)
            packed = (pix[:, 2] << 16) | (pix[:, 1] << 8) | pix[:, 0]
            L["colored_ortho"].view(np.uint32)[ii, jj] = packed
        else:
            L["ortho"][ii, jj] = imgs[f][py[ii, jj], px[ii, jj]].astype(np.float32)
    return L


@pytest.mark.parametrize("colored", [False, True])
def test_cell_loop_matches_independent_numpy(colored):
    rows, cols, res = 70, 50, 0.5
    camd = synth.scaled_camera(0.05)
    poses = synth.lawnmower_poses(2, 3, rows * res / 2, cols * res / 2, agl=45.0, seed=31, jitter_pos=1.0)
    ch = 3 if colored else 1
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], ch) for k in range(len(poses))]
    elev = synth.analytic_elevation(rows, cols, res)
    elev[3:6, 4:9] = np.nan
    L = fresh_layers(rows, cols, elev)
    st, _ = po.ortho_process(po.make_geometry(rows, cols, res), L, po.make_camera(**camd), poses, imgs,
                             colored=colored, num_threads=3)
    assert st == 0
    N = numpy_ortho(rows, cols, res, elev, camd, poses, imgs, colored)
    assert np.array_equal(L["observation_index"], N["observation_index"], equal_nan=True)
    key = "colored_ortho" if colored else "ortho"
    assert np.array_equal(L[key].view(np.uint32), N[key].view(np.uint32))
    assert ulp_diff(L["elevation_angle"], N["elevation_angle"]).max() <= 1
    assert np.isnan(L["observation_index"][3:6, 4:9]).all() and (L["ortho"][3:6, 4:9] == 255).all()
    assert (~np.isnan(L["observation_index"])).sum() > 0.8 * rows * cols


def test_batch_split_invariance_and_batch_relative_index():
    # state lives in the layers; observation_index is the index WITHIN each call.
    rows, cols, res = 60, 60, 0.5
    camd = synth.scaled_camera(0.05)
    poses = synth.lawnmower_poses(2, 4, 15.0, 15.0, agl=45.0, seed=33, jitter_pos=1.0)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(8)]
    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)
    elev = synth.analytic_elevation(rows, cols, res)
    A = fresh_layers(rows, cols, elev)
    po.ortho_process(g, A, cam, poses, imgs)
    B = fresh_layers(rows, cols, elev)
    po.ortho_process(g, B, cam, poses[:5], imgs[:5])
    po.ortho_process(g, B, cam, poses[5:], imgs[5:])
    assert np.array_equal(A["ortho"], B["ortho"])
    assert np.array_equal(A["elevation_angle"], B["elevation_angle"])
    late = A["observation_index"] >= 5
    assert np.array_equal(B["observation_index"][late], A["observation_index"][late] - 5)
    assert np.array_equal(B["observation_index"][~late], A["observation_index"][~late], equal_nan=True)


def test_threads_and_single_thread_twins_agree():
    rows, cols, res = 64, 40, 0.5
    camd = synth.scaled_camera(0.05, dist_type=2, dist=EQUI)
    poses = synth.lawnmower_poses(2, 2, 16.0, 10.0, agl=40.0, seed=35)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(4)]
    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)
    outs = []
    for t in (-1, 1, 5):
        L = fresh_layers(rows, cols, synth.analytic_elevation(rows, cols, res))
        st, _ = po.ortho_process(g, L, cam, poses, imgs, num_threads=t)
        assert st == 0
        outs.append(L)
    for L in outs[1:]:
        for k in ("ortho", "elevation_angle", "observation_index"):
            assert np.array_equal(L[k], outs[0][k], equal_nan=True)


def test_argument_checks():
    g, cam = po.make_geometry(4, 4, 1.0), po.make_camera(8, 8, 4, 4, 4, 4)
    L = fresh_layers(4, 4, np.zeros((4, 4), np.float32))
    assert po.ortho_process(g, L, cam, np.zeros((0, 7)), [])[0] == -1  # CHECK(!T_G_Bs.empty())


def test_fov_branch_is_the_stated_formula():
    """The oracle's FOV branch against the SAME stated formula evaluated with numpy: r_d = atan(2 tan(w/2) r) / w, limits
    w*w < 1e-5 -> r and r*r < 1e-5 -> 2 tan(w/2) r / w.  This checks the C++ transcription only.  It is NOT an independent
    cross-check of the model (unlike rad-tan / equidistant vs OpenCV above): the formula itself is recalled from upstream
    aslam_cv2 distortion-fisheye.cc, which is not available offline."""
    rng = np.random.default_rng(8)
    for w in (0.9, 1.2, 0.3, 1e-3):
        cam = po.make_camera(width=4000, height=3000, fu=1500.0, fv=1490.0, cu=2000.0, cv=1500.0, dist_type=3,
                             dist=(w, 0.0, 0.0, 0.0))
        for _ in range(200):
            p = np.array([rng.normal(0, 2.0), rng.normal(0, 2.0), rng.uniform(0.5, 8.0)])
            if rng.random() < 0.1:
                p[:2] *= 1e-4          # small-radius limit branch
            u, v = p[0] / p[2], p[1] / p[2]
            r = np.hypot(u, v)
            t = np.tan(w / 2.0)
            s = 1.0 if w * w < 1e-5 else (2 * t / w if r * r < 1e-5 else np.arctan(2 * t * r) / (r * w))
            want = np.array([1500.0 * u * s + 2000.0, 1490.0 * v * s + 1500.0])
            got = po.project3(cam, p)[1]
            assert np.allclose(got, want, rtol=1e-13, atol=1e-10), (w, p, got, want)
