"""ctypes binding of the CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module
(see oracle/amb_oracle.h).  The product package `aerial_mapper_b200` never does.

Pinning (details in amb_oracle.h): the reference holds no golden vectors for this path; the
oracle is pinned to oracle/_ref — the reference's own dsm.cc / ortho-backward-grid.cc / ortho-from-pcl.cc compiled
verbatim here (refsrc_* below) and checked bit-for-bit against the restatement.  PARITY UNPINNED only for the
arithmetic of the absent third-party dependencies (grid_map, aslam_cv2, minkindr), restated in thirdparty_math.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libamb_oracle_ref.so")
_REFSRC_MAIN_PATH = os.path.join(_HERE, "_ref", "libamb_refsrc_main.so")
_REFSRC_PCL_PATH = os.path.join(_HERE, "_ref", "libamb_refsrc_pcl.so")
_REFERENCE_DEMO_PATH = os.path.join(_HERE, "_ref", "libamb_reference_demo.so")
_REFSRC_STEREO_PATH = os.path.join(_HERE, "_ref", "libamb_refsrc_stereo.so")
_REFERENCE_ROOT = "/root/reference"


class Geometry(C.Structure):
    """amb_geometry (include/aerial_mapper_b200.h)."""
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("resolution", C.c_double),
                ("length_x", C.c_double), ("length_y", C.c_double), ("pos_x", C.c_double), ("pos_y", C.c_double)]


class Camera(C.Structure):
    """amb_camera (include/aerial_mapper_b200.h)."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fu", C.c_double), ("fv", C.c_double),
                ("cu", C.c_double), ("cv", C.c_double), ("dist_type", C.c_int32), ("reserved_", C.c_int32),
                ("dist", C.c_double * 4), ("q_C_B", C.c_double * 4), ("t_C_B", C.c_double * 3)]


def build(force=False):
    """Compile liboracle.so (always possible) and oracle/_ref (only where /root/reference exists).  `make` is
    incremental; on the GPU box (no /root/reference) the prebuilt oracle/_ref/*.so that travelled are used as is."""
    flags = ["-B"] if force else []
    subprocess.check_call(["make", "-C", _HERE] + flags + ["liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir(_REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", _HERE] + flags + ["ref"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)


_lib = None
_ref = None

_DSM_ARGS = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_double, C.c_double,
             C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]


_PCL_ARGS = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32,
             C.c_int64, C.c_int64, C.c_void_p]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ambo_dsm_process.argtypes = _DSM_ARGS
        L.ambo_dsm_process.restype = C.c_int
        L.ambo_ortho_process.argtypes = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.POINTER(Camera), C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_int32, C.c_size_t, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                         C.c_void_p]
        L.ambo_ortho_process.restype = C.c_int
        L.ambo_ortho_from_pcl_process.argtypes = _PCL_ARGS
        L.ambo_ortho_from_pcl_process.restype = C.c_int
        L.ambo_stereo_reproject.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ambo_stereo_reproject.restype = C.c_int
        L.ambo_stereo_rectify_setup.argtypes = [C.c_void_p] * 9
        L.ambo_stereo_rectify_setup.restype = C.c_int
        L.ambo_stereo_rectify_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_size_t, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
        L.ambo_stereo_rectify_maps.restype = C.c_int
        L.ambo_project3.argtypes = [C.POINTER(Camera), C.c_void_p, C.c_void_p]
        L.ambo_project3.restype = C.c_int
        L.ambo_transform_to_camera.argtypes = [C.POINTER(Camera), C.c_void_p, C.c_void_p, C.c_void_p]
        L.ambo_transform_to_camera.restype = C.c_int
        L.ambo_pack_color.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8]
        L.ambo_pack_color.restype = C.c_uint32
        L.ambo_hardware_concurrency.restype = C.c_int
        _lib = L
    return _lib


def have_ref():
    if os.path.isdir(_REFERENCE_ROOT):
        build()
    return os.path.exists(_REF_PATH)


def ref():
    """oracle/_ref: the cell loop around the reference's own nanoflann.hpp (None if it was never built)."""
    global _ref
    if _ref is None:
        if not have_ref():
            return None
        L = C.CDLL(_REF_PATH)
        L.ambo_ref_dsm_process.argtypes = _DSM_ARGS
        L.ambo_ref_dsm_process.restype = C.c_int
        L.ambo_ref_ortho_from_pcl_process.argtypes = _PCL_ARGS
        L.ambo_ref_ortho_from_pcl_process.restype = C.c_int
        _ref = L
    return _ref


_refsrc = None


def have_refsrc():
    if os.path.isdir(_REFERENCE_ROOT):
        build()
    return os.path.exists(_REFSRC_MAIN_PATH) and os.path.exists(_REFSRC_PCL_PATH)


def refsrc():
    """oracle/_ref: the reference's OWN dsm.cc, ortho-backward-grid.cc (main) and ortho-from-pcl.cc (pcl) compiled
    verbatim against the stand-in third-party headers (refsrc_stubs/).  Returns (main, pcl) or None."""
    global _refsrc
    if _refsrc is None:
        if not have_refsrc():
            return None
        M = C.CDLL(_REFSRC_MAIN_PATH)
        M.ambo_refsrc_dsm_process.argtypes = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32,
                                              C.c_double, C.c_double, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
        M.ambo_refsrc_dsm_process.restype = C.c_int
        M.ambo_refsrc_ortho_process.argtypes = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.POINTER(Camera), C.c_void_p, C.c_void_p,
                                                C.c_size_t, C.c_int32, C.c_size_t, C.c_int32, C.c_int32, C.c_int64,
                                                C.c_int64, C.c_void_p]
        M.ambo_refsrc_ortho_process.restype = C.c_int
        M.ambo_refsrc_last_error.restype = C.c_char_p
        M.ambo_refsrc_hardware_concurrency.restype = C.c_int
        P = C.CDLL(_REFSRC_PCL_PATH)
        P.ambo_refsrc_ortho_from_pcl_process.argtypes = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_void_p,
                                                         C.c_size_t, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                                         C.c_void_p]
        P.ambo_refsrc_ortho_from_pcl_process.restype = C.c_int
        P.ambo_refsrc_pcl_last_error.restype = C.c_char_p
        _refsrc = (M, P)
    return _refsrc


def have_refsrc_stereo():
    if os.path.isdir(_REFERENCE_ROOT):
        build()
    return os.path.exists(_REFSRC_STEREO_PATH)


def refsrc_stereo_reproject(disparity, image_left, k4, baseline, R_G_C, t_G_C1):
    """stereo::Densifier::computePointCloud — the reference's densifier.cpp compiled verbatim (oracle/_ref).
    Returns (status, xyz [n,3] float64, intensities [n] int32); kMaxInvalidDisparity is the reference's constant 1."""
    L = C.CDLL(_REFSRC_STEREO_PATH)
    L.ambo_refsrc_stereo_reproject.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_size_t, C.POINTER(C.c_size_t)]
    L.ambo_refsrc_stereo_reproject.restype = C.c_int
    disp = np.ascontiguousarray(disparity, dtype=np.float32)
    img = np.ascontiguousarray(image_left, dtype=np.uint8)
    h, w = disp.shape
    k4 = np.ascontiguousarray(k4, dtype=np.float64)
    R = np.ascontiguousarray(R_G_C, dtype=np.float64).reshape(9)
    t = np.ascontiguousarray(t_G_C1, dtype=np.float64).reshape(3)
    xyz = np.empty((h * w, 3), np.float64)
    inten = np.empty(h * w, np.int32)
    n = C.c_size_t(0)
    st = L.ambo_refsrc_stereo_reproject(_ptr(disp), w, _ptr(img), w, w, h, _ptr(k4), float(baseline), _ptr(R), _ptr(t),
                                        _ptr(xyz), _ptr(inten), h * w, C.byref(n))
    return st, xyz[:n.value].copy(), inten[:n.value].copy()


def have_reference_demo():
    if os.path.isdir(_REFERENCE_ROOT):
        build()
    return os.path.exists(_REFERENCE_DEMO_PATH)


def reference_demo_main(*argv):
    """tests/cpp/shim_demo.cc (the batch demo's call sequence) built against the reference's own headers + sources:
    runs its main(argc, argv) in-process and returns its exit code."""
    L = C.CDLL(_REFERENCE_DEMO_PATH)
    L.amb_demo_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.amb_demo_main.restype = C.c_int
    args = [b"reference_demo"] + [os.fsencode(a) for a in argv]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    return L.amb_demo_main(len(args), arr)


def make_geometry(rows, cols, resolution, pos_x=0.0, pos_y=0.0):
    return Geometry(int(rows), int(cols), float(resolution), rows * float(resolution), cols * float(resolution),
                    float(pos_x), float(pos_y))


def make_camera(width, height, fu, fv, cu, cv, dist_type=0, dist=(0, 0, 0, 0), q_C_B=(1, 0, 0, 0),
                t_C_B=(0, 0, 0)):
    cam = Camera()
    cam.width, cam.height = int(width), int(height)
    cam.fu, cam.fv, cam.cu, cam.cv = float(fu), float(fv), float(cu), float(cv)
    cam.dist_type = int(dist_type)
    cam.dist = (C.c_double * 4)(*[float(d) for d in dist])
    cam.q_C_B = (C.c_double * 4)(*[float(d) for d in q_C_B])
    cam.t_C_B = (C.c_double * 3)(*[float(d) for d in t_C_B])
    return cam


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dsm_process(geom, elevation, xyz, radius=1, center_easting=0.0, center_northing=0.0, num_threads=0,
                cell_range=None, debug=False, use_ref=False):
    """Run the oracle DSM on `elevation` (float32 F-order rows x cols, modified in place).

    Returns (status, neighbour_count|None, threshold_index|None, seconds[2])."""
    assert elevation.dtype == np.float32 and elevation.flags.f_contiguous
    assert elevation.shape == (geom.rows, geom.cols)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n = xyz.shape[0] if xyz.ndim == 2 else xyz.size // 3
    total = geom.rows * geom.cols
    lo, hi = (0, total) if cell_range is None else cell_range
    cnt = np.full(total, -1, np.int32) if debug else None
    lvl = np.full(total, -1, np.int8) if debug else None
    sec = np.zeros(2, np.float64)
    fn = ref().ambo_ref_dsm_process if use_ref else lib().ambo_dsm_process
    st = fn(C.byref(geom), _ptr(elevation), _ptr(xyz), n, int(radius), float(center_easting),
            float(center_northing), int(num_threads), lo, hi, _ptr(cnt), _ptr(lvl), _ptr(sec))
    if debug:
        cnt = cnt.reshape((geom.rows, geom.cols), order="F")
        lvl = lvl.reshape((geom.rows, geom.cols), order="F")
    return st, cnt, lvl, sec


def ortho_from_pcl_process(geom, ortho, xyz, intensities, radius=2, adaptive=False, num_threads=0, cell_range=None,
                           use_ref=False):
    """Oracle OrthoFromPcl on `ortho` (float32 F-order rows x cols, modified in place).  Returns status."""
    assert ortho.dtype == np.float32 and ortho.flags.f_contiguous and ortho.shape == (geom.rows, geom.cols)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    inten = np.ascontiguousarray(intensities, dtype=np.int32)
    n = xyz.size // 3
    assert inten.size >= n
    total = geom.rows * geom.cols
    lo, hi = (0, total) if cell_range is None else cell_range
    fn = ref().ambo_ref_ortho_from_pcl_process if use_ref else lib().ambo_ortho_from_pcl_process
    return fn(C.byref(geom), _ptr(ortho), _ptr(xyz), _ptr(inten), n, int(radius), 1 if adaptive else 0,
              int(num_threads), lo, hi, None)


def stereo_reproject(disparity, image_left, k4, baseline, R_G_C, t_G_C1, max_invalid_disparity=1.0):
    """Oracle Densifier::computePointCloud.  Returns (status, xyz [n,3] float64, intensities [n] int32)."""
    disp = np.ascontiguousarray(disparity, dtype=np.float32)
    img = np.ascontiguousarray(image_left, dtype=np.uint8)
    h, w = disp.shape
    k4 = np.ascontiguousarray(k4, dtype=np.float64)
    R = np.ascontiguousarray(R_G_C, dtype=np.float64).reshape(9)
    t = np.ascontiguousarray(t_G_C1, dtype=np.float64).reshape(3)
    xyz = np.empty((h * w, 3), np.float64)
    inten = np.empty(h * w, np.int32)
    n = C.c_size_t(0)
    st = lib().ambo_stereo_reproject(_ptr(disp), w, _ptr(img), w, w, h, _ptr(k4), float(baseline), _ptr(R), _ptr(t),
                                     float(max_invalid_disparity), _ptr(xyz), _ptr(inten), h * w, C.byref(n))
    return st, xyz[:n.value].copy(), inten[:n.value].copy()


def stereo_rectify_setup(K, R_G_C1, R_G_C2, t_G_C1, t_G_C2):
    """Oracle Rectifier::rectifyStereoPair, host half.  Returns (status, baseline, R_G_C_rect, T1_inv, T2_inv)."""
    a = [np.ascontiguousarray(m, dtype=np.float64).reshape(n) for m, n in
         ((K, 9), (R_G_C1, 9), (R_G_C2, 9), (t_G_C1, 3), (t_G_C2, 3))]
    base = np.zeros(1, np.float64)
    R = np.zeros(9, np.float64)
    T1, T2 = np.zeros(9, np.float32), np.zeros(9, np.float32)
    st = lib().ambo_stereo_rectify_setup(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(a[4]), _ptr(base),
                                         _ptr(R), _ptr(T1), _ptr(T2))
    return st, float(base[0]), R.reshape(3, 3), T1.reshape(3, 3), T2.reshape(3, 3)


def stereo_rectify_maps(T1_inv, T2_inv, width, height):
    """Oracle per-pixel map fill.  Returns (status, (map1_x, map1_y, map2_x, map2_y)) float32 [H, W]."""
    T1 = np.ascontiguousarray(T1_inv, dtype=np.float32).reshape(9)
    T2 = np.ascontiguousarray(T2_inv, dtype=np.float32).reshape(9)
    maps = [np.empty((int(height), int(width)), np.float32) for _ in range(4)]
    st = lib().ambo_stereo_rectify_maps(_ptr(T1), _ptr(T2), int(width), int(height), int(width), _ptr(maps[0]),
                                        _ptr(maps[1]), _ptr(maps[2]), _ptr(maps[3]))
    return st, tuple(maps)


def ortho_process(geom, layers, camera, T_G_B, images, colored=False, num_threads=0, cell_range=None):
    """Run the oracle orthomosaic.  layers: dict of float32 F-order arrays (elevation, elevation_angle,
    observation_index, ortho, colored_ortho), modified in place.  images: list of uint8 arrays HxW or HxWx3 (BGR).

    Returns (status, seconds)."""
    T = np.ascontiguousarray(T_G_B, dtype=np.float64).reshape(-1, 7)
    n = T.shape[0]
    assert len(images) == n
    if n == 0:
        return -1, 0.0  # AMB_ERR_EMPTY: CHECK(!T_G_Bs.empty()), ortho-backward-grid.cc:225
    imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
    channels = 3 if colored else 1
    h, w = imgs[0].shape[:2]
    for im in imgs:
        assert im.shape[:2] == (h, w) and (im.ndim == 3 and im.shape[2] == 3 if colored else im.ndim == 2)
    row_step = w * channels
    ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
    total = geom.rows * geom.cols
    lo, hi = (0, total) if cell_range is None else cell_range
    for k in ("elevation", "elevation_angle", "observation_index"):
        a = layers[k]
        assert a.dtype == np.float32 and a.flags.f_contiguous and a.shape == (geom.rows, geom.cols), k
    sec = np.zeros(1, np.float64)
    st = lib().ambo_ortho_process(C.byref(geom), _ptr(layers["elevation"]), _ptr(layers["elevation_angle"]),
                                  _ptr(layers["observation_index"]), _ptr(layers.get("ortho")),
                                  _ptr(layers.get("colored_ortho")), C.byref(camera), _ptr(T),
                                  C.cast(ptrs, C.c_void_p), n, channels, row_step, 1 if colored else 0,
                                  int(num_threads), lo, hi, _ptr(sec))
    return st, float(sec[0])


def project3(camera, p_C):
    p = np.ascontiguousarray(p_C, dtype=np.float64)
    kp = np.zeros(2, np.float64)
    vis = lib().ambo_project3(C.byref(camera), _ptr(p), _ptr(kp))
    return bool(vis), kp


def transform_to_camera(camera, T_G_B_row, p_G):
    T = np.ascontiguousarray(T_G_B_row, dtype=np.float64)
    p = np.ascontiguousarray(p_G, dtype=np.float64)
    out = np.zeros(3, np.float64)
    lib().ambo_transform_to_camera(C.byref(camera), _ptr(T), _ptr(p), _ptr(out))
    return out


def pack_color(b, g, r):
    return int(lib().ambo_pack_color(int(b), int(g), int(r)))


def hardware_concurrency():
    return int(lib().ambo_hardware_concurrency())


# ---- the reference's own translation units (oracle/_ref/libamb_refsrc_*.so) ------------------------------------
def refsrc_dsm_process(geom, elevation, xyz, radius=1, center_easting=0.0, center_northing=0.0, multi_thread=True,
                       cell_range=None):
    """dsm::Dsm(settings, &map).process(cloud, &map) — the reference's dsm.cc.  Returns (status, seconds[2])."""
    assert elevation.dtype == np.float32 and elevation.flags.f_contiguous
    assert elevation.shape == (geom.rows, geom.cols)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n = xyz.shape[0] if xyz.ndim == 2 else xyz.size // 3
    lo, hi = (0, geom.rows * geom.cols) if cell_range is None else cell_range
    sec = np.zeros(2, np.float64)
    st = refsrc()[0].ambo_refsrc_dsm_process(C.byref(geom), _ptr(elevation), _ptr(xyz), n, int(radius),
                                             float(center_easting), float(center_northing),
                                             1 if multi_thread else 0, lo, hi, _ptr(sec))
    return st, sec


def refsrc_ortho_process(geom, layers, camera, T_G_B, images, colored=False, multi_thread=True, cell_range=None):
    """ortho::OrthoBackwardGrid(ncameras, settings, &map).process(T_G_Bs, images, &map) — the reference's
    ortho-backward-grid.cc.  layers as for ortho_process, plus optional "num_observations".
    Returns (status, seconds[2])."""
    T = np.ascontiguousarray(T_G_B, dtype=np.float64).reshape(-1, 7)
    n = T.shape[0]
    assert len(images) == n
    imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
    channels = 3 if colored else 1
    row_step = camera.width * channels
    for im in imgs:
        assert im.shape[:2] == (camera.height, camera.width)
        assert (im.ndim == 3 and im.shape[2] == 3) if colored else im.ndim == 2
    ptrs = (C.c_void_p * max(n, 1))(*[im.ctypes.data for im in imgs])
    lo, hi = (0, geom.rows * geom.cols) if cell_range is None else cell_range
    for k in ("elevation", "elevation_angle", "observation_index"):
        a = layers[k]
        assert a.dtype == np.float32 and a.flags.f_contiguous and a.shape == (geom.rows, geom.cols), k
    sec = np.zeros(2, np.float64)
    st = refsrc()[0].ambo_refsrc_ortho_process(
        C.byref(geom), _ptr(layers["elevation"]), _ptr(layers["elevation_angle"]),
        _ptr(layers["observation_index"]), _ptr(layers.get("num_observations")), _ptr(layers.get("ortho")),
        _ptr(layers.get("colored_ortho")), C.byref(camera), _ptr(T), C.cast(ptrs, C.c_void_p), n, channels, row_step,
        1 if colored else 0, 1 if multi_thread else 0, lo, hi, _ptr(sec))
    return st, sec


def refsrc_ortho_from_pcl_process(geom, ortho, xyz, intensities, radius=2, adaptive=False, cell_range=None):
    """ortho::OrthoFromPcl(settings).process(cloud, intensities, &map) — the reference's ortho-from-pcl.cc."""
    assert ortho.dtype == np.float32 and ortho.flags.f_contiguous and ortho.shape == (geom.rows, geom.cols)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    inten = np.ascontiguousarray(intensities, dtype=np.int32)
    n = xyz.size // 3
    assert inten.size >= n
    lo, hi = (0, geom.rows * geom.cols) if cell_range is None else cell_range
    return refsrc()[1].ambo_refsrc_ortho_from_pcl_process(C.byref(geom), _ptr(ortho), _ptr(xyz), _ptr(inten), n,
                                                          int(radius), 1 if adaptive else 0, lo, hi, None)


def refsrc_last_error():
    M, P = refsrc()
    return (M.ambo_refsrc_last_error() or b"").decode(), (P.ambo_refsrc_pcl_last_error() or b"").decode()
