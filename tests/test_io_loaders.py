""""Next" row N2: the reference's input formats (aerial_mapper_io/src/aerial-mapper-io.cc) as read by
aerial_mapper_b200.io — host-side parsing only."""
import numpy as np
import pytest

from aerial_mapper_b200 import io as aio


def test_poses_stream_semantics(tmp_path):
    p = tmp_path / "poses.txt"
    p.write_text("1 2 3 1 0 0 0\n4.5 5 6\n0.5 0.5 0.5 0.5   7 8 9 0 1 0 0 garbage 1 2 3 4 5 6 7\n")
    T = aio.load_poses_from_file_standard(str(p))
    assert T.shape == (3, 7)                      # records may span lines; parsing stops at the first bad token
    assert T[1].tolist() == [4.5, 5, 6, 0.5, 0.5, 0.5, 0.5] and T[2, 4] == 1.0
    (tmp_path / "empty.txt").write_text("\n")
    with pytest.raises(ValueError):
        aio.load_poses_from_file_standard(str(tmp_path / "empty.txt"))


def test_point_cloud_filter_and_int_intensity(tmp_path):
    p = tmp_path / "cloud.txt"
    p.write_text("0 0 10 5\n1 1 -100 7\n2 2 -99.5 9\n3 3 -250 1\n4 4 4 255\n5 5 5 12.5\n6 6 6 3\n")
    xyz, inten = aio.load_point_cloud_from_file(str(p), with_intensities=True)
    # z > -100 keeps -99.5 and drops -100 / -250; "12.5" is not an int: extraction fails, the loop ends
    assert xyz.tolist() == [[0, 0, 10], [2, 2, -99.5], [4, 4, 4]] and inten.tolist() == [5, 9, 255]
    assert inten.dtype == np.int32 and aio.load_point_cloud_from_file(str(p)).shape == (3, 3)


def test_images_prefix_numbering_gray_and_bgr(tmp_path):
    import cv2
    rng = np.random.default_rng(0)
    for i in range(3):
        cv2.imwrite(str(tmp_path / ("image_%d.jpg" % i)), rng.integers(0, 255, (24, 32, 3), dtype=np.uint8))
    gray = aio.load_images_from_file(str(tmp_path / "image_"), 3)
    col = aio.load_images_from_file(str(tmp_path / "image_"), 3, load_colored_images=True)
    assert len(gray) == 3 and gray[0].shape == (24, 32) and gray[0].dtype == np.uint8
    assert col[2].shape == (24, 32, 3) and col[2].flags.c_contiguous
    with pytest.raises(IOError):
        aio.load_images_from_file(str(tmp_path / "image_"), 4)


def test_ncamera_yaml_to_amb_camera(tmp_path):
    from scipy.spatial.transform import Rotation as R
    Rbc = R.from_euler("xyz", [3.0, -0.2, 0.1]).as_matrix()
    tbc = np.array([0.1, -0.05, 0.02])
    T = np.eye(4); T[:3, :3] = Rbc; T[:3, 3] = tbc
    y = tmp_path / "rig.yaml"
    y.write_text("""label: ncamera
id: 4c07c22b7a5e46bdf1a2a5dc3bd5e4a3
cameras:
- camera:
    label: cam0
    id: 54812562fa109c40fe90b29a59dd7798
    line-delay-nanoseconds: 0
    image_height: 3000
    image_width: 4000
    type: pinhole
    intrinsics:
      cols: 1
      rows: 4
      data: [3000.0, 2990.0, 2000.5, 1500.25]
    distortion:
      type: radial-tangential
      parameters:
        cols: 1
        rows: 4
        data: [-0.05, 0.01, 0.0001, 0.0002]
  T_B_C:
    cols: 4
    rows: 4
    data: %s
""" % np.array2string(T.ravel(), separator=", ", max_line_width=10000))
    nc = aio.load_camera_rig_from_file(str(y))
    c = nc.camera
    assert (c.width, c.height, c.fu, c.fv, c.cu, c.cv) == (4000, 3000, 3000.0, 2990.0, 2000.5, 1500.25)
    assert c.dist_type == 1 and list(c.dist) == [-0.05, 0.01, 0.0001, 0.0002]
    q = np.array(c.q_C_B)
    Rcb = R.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()
    assert np.allclose(Rcb, Rbc.T, atol=1e-12) and np.allclose(np.array(c.t_C_B), -Rbc.T @ tbc, atol=1e-12)
