"""Python mirror of the reference's interface for the grid-mapping hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference's C++ classes, so the parity tests read like the
reference's callers (aerial_mapper_demos/src/dsm/main-dsm.cc:94-107,
src/ortho/main-ortho-backward-grid.cc:119-141, main-ortho-backward-grid-incremental.cc:99-163):

    settings = GridMapSettings(...);  map = AerialGridMap(settings)           # aerial-mapper-grid-map.h:23-54
    Dsm(DsmSettings(...), map.get_mutable()).process(point_cloud, map.get_mutable())     # dsm.h:25-47
    OrthoBackwardGrid(ncameras, OrthoSettings(...), map.get_mutable()).process(T_G_Bs, images, map.get_mutable())

The C++ drop-in headers with the literal reference signatures live in aerial_mapper_b200/shim/.

All compute happens in libaerial_mapper_b200.so on the GPU; there is no CPU path in this package.
"""
import ctypes as C
import os
import logging

import numpy as np

from . import _lib
from ._lib import AmbError, Camera, Geometry, LAYER_ID, LAYER_NAMES, check, lib

log = logging.getLogger("aerial_mapper_b200")

# Layers the hot path reads or writes (ortho-backward-grid.cc:133-138, dsm.cc:116).
HOT_LAYERS = ("ortho", "elevation", "elevation_angle", "observation_index", "colored_ortho")


class GridMapSettings(object):
    """grid_map::Settings (aerial-mapper-grid-map.h:23-29)."""

    def __init__(self, center_easting=0.0, center_northing=0.0, delta_easting=100.0, delta_northing=100.0,
                 resolution=1.0):
        self.center_easting = center_easting
        self.center_northing = center_northing
        self.delta_easting = delta_easting
        self.delta_northing = delta_northing
        self.resolution = resolution


class GridMap(object):
    """The slice of grid_map::GridMap the hot path uses: geometry + float32 column-major layers.

    `layers[name]` is a numpy float32 array of shape (rows, cols) in Fortran order — the memory layout of
    grid_map::Matrix (Eigen::MatrixXf): element (i, j) at i + j*rows, i <-> x / easting, j <-> y / northing.

    Host mode (default) keeps these arrays authoritative: every process() uploads what it reads and downloads what
    it writes, like the reference mutating the caller's map in place.  to_device() switches to resident mode: the
    layers live in HBM across process() calls (the incremental pipeline's access pattern) and come back with
    download().
    """

    def __init__(self, layer_names=LAYER_NAMES, pinned=False):
        self.layer_names = tuple(layer_names)
        self.pinned = bool(pinned)  # allocate the host layers in page-locked memory (full-rate PCIe copies)
        self._pinned_ptrs = []
        self.geometry = None
        self.layers = {}
        self.frame_id = ""
        self._ctx = None
        self._resident = False
        self._device = 0
        self._col_range = None

    # --- grid_map::GridMap surface used by the reference ---
    def setFrameId(self, frame_id):
        self.frame_id = frame_id

    def setGeometry(self, length, resolution, position=(0.0, 0.0)):
        g = Geometry()
        check(lib().amb_geometry_init(float(length[0]), float(length[1]), float(resolution), float(position[0]),
                                      float(position[1]), C.byref(g)))
        self.geometry = g
        self._free_pinned()
        self.layers = {}
        for name in self.layer_names:
            a = self._alloc_layer(g.rows, g.cols)
            a[...] = np.float32(np.nan)  # grid_map::GridMap::setGeometry -> clearAll()
            self.layers[name] = a
        self._release()

    def _alloc_layer(self, rows, cols):
        if not self.pinned:
            return np.empty((rows, cols), dtype=np.float32, order="F")
        p = C.c_void_p()
        nbytes = rows * cols * 4
        check(lib().amb_host_alloc(C.byref(p), nbytes))
        self._pinned_ptrs.append(p)
        buf = (C.c_float * (rows * cols)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float32).reshape((rows, cols), order="F")

    def _free_pinned(self):
        self.layers = {}
        for p in self._pinned_ptrs:
            lib().amb_host_free(p)
        self._pinned_ptrs = []

    def getSize(self):
        return (self.geometry.rows, self.geometry.cols)

    def getLength(self):
        return (self.geometry.length_x, self.geometry.length_y)

    def getResolution(self):
        return self.geometry.resolution

    def getPosition(self, index=None):
        if index is None:
            return (self.geometry.pos_x, self.geometry.pos_y)
        x, y = C.c_double(), C.c_double()
        check(lib().amb_geometry_position(C.byref(self.geometry), int(index[0]), int(index[1]), C.byref(x),
                                          C.byref(y)))
        return (x.value, y.value)

    def __getitem__(self, layer):
        return self.layers[layer]

    def __setitem__(self, layer, value):
        self.layers[layer][...] = value

    # --- device residency ---
    def context(self, device=None, col_range=None):
        if self._ctx is None:
            if device is not None:
                self._device = int(device)
            g = self.geometry
            c0, c1 = (0, g.cols) if col_range is None else col_range
            self._col_range = (int(c0), int(c1))
            ctx = C.c_void_p()
            check(lib().amb_create(C.byref(g), self._device, int(c0), int(c1), C.byref(ctx)))
            self._ctx = ctx
        return self._ctx

    def _slab(self, name):
        c0, c1 = self._col_range
        a = self.layers[name]
        assert a.dtype == np.float32 and a.flags.f_contiguous
        return a[:, c0:c1]  # contiguous in F order

    def upload(self, names=HOT_LAYERS):
        ctx = self.context()
        for name in names:
            slab = self._slab(name)
            check(lib().amb_upload_layer(ctx, LAYER_ID[name], slab.ctypes.data_as(C.c_void_p)), ctx)

    def download(self, names=HOT_LAYERS):
        ctx = self.context()
        for name in names:
            slab = self._slab(name)
            check(lib().amb_download_layer(ctx, LAYER_ID[name], slab.ctypes.data_as(C.c_void_p)), ctx)

    def download_async(self, names=HOT_LAYERS):
        """Enqueue device->host copies that overlap later work; sync() completes them (pinned layers recommended)."""
        ctx = self.context()
        for name in names:
            slab = self._slab(name)
            check(lib().amb_download_layer_async(ctx, LAYER_ID[name], slab.ctypes.data_as(C.c_void_p)), ctx)

    def set_mirrors(self, names=HOT_LAYERS, enable=True):
        """Register the host layers as mirrors (amb_set_host_mirror): process() streams every result layer back as
        soon as it is final, overlapping later stages; sync() completes the copies.  Use pinned layers."""
        ctx = self.context()
        # one-byte transport of `ortho` / `observation_index` (amb_set_host_mirror_compact; AMB_COMPACT_MIRRORS=0 turns it off)
        compact = os.environ.get("AMB_COMPACT_MIRRORS", "1") not in ("", "0")   # default on (exact-bits transport)
        for name in names:
            slab = self._slab(name)
            ptr = slab.ctypes.data_as(C.c_void_p) if enable else None
            check(lib().amb_set_host_mirror(ctx, LAYER_ID[name], ptr), ctx)
            if name in ("ortho", "observation_index"):
                check(lib().amb_set_host_mirror_compact(ctx, LAYER_ID[name], 1 if (enable and compact) else 0), ctx)

    def to_device(self, device=0, col_range=None, names=HOT_LAYERS):
        """Make the layers device-resident (uploads the current host values once)."""
        self._release()
        self.context(device, col_range)
        self.upload(names)
        self._resident = True
        return self

    def is_resident(self):
        return self._resident

    def sync(self):
        if self._ctx is not None:
            check(lib().amb_sync(self._ctx), self._ctx)

    def timings(self):
        t = _lib.Timings()
        check(lib().amb_get_timings(self.context(), C.byref(t)), self._ctx)
        return t.as_dict()

    def _release(self):
        if self._ctx is not None:
            lib().amb_destroy(self._ctx)
            self._ctx = None
        self._resident = False

    def __del__(self):
        try:
            self._release()
            self._free_pinned()
        except Exception:
            pass


class AerialGridMap(object):
    """grid_map::AerialGridMap (aerial-mapper-grid-map.cc:23-49) without the ROS publisher."""

    def __init__(self, settings, pinned=False, layer_names=LAYER_NAMES):
        self.settings_ = settings
        self.map_ = GridMap(layer_names, pinned=pinned)  # :25-28
        self.map_.setFrameId("world")
        self.map_.setGeometry((settings.delta_easting, settings.delta_northing), settings.resolution,
                              (settings.center_easting, settings.center_northing))  # :30-33
        self.reset()

    INITIAL_VALUES = {"ortho": 255.0, "elevation": np.nan, "elevation_angle": 0.0,
                      "elevation_angle_first_view": np.nan, "num_observations": 0.0, "observation_index": np.nan,
                      "observation_index_first": np.nan, "delta": np.nan, "colored_ortho": np.nan}

    def reset(self):
        """The setConstant() calls of AerialGridMap::initialize (aerial-mapper-grid-map.cc:40-48)."""
        for name in self.map_.layer_names:
            self.map_[name] = self.INITIAL_VALUES[name]

    def getMutable(self):
        return self.map_

    get_mutable = getMutable


class DsmSettings(object):
    """dsm::Settings (dsm.h:25-32).  interpolation_radius is an int compared against SQUARED distances (m^2)."""

    def __init__(self, interpolation_radius=1, adaptive_interpolation=False, center_easting=0.0,
                 center_northing=0.0, use_multi_threads=True):
        self.interpolation_radius = int(interpolation_radius)  # `int interpolation_radius = 1.0;`
        self.adaptive_interpolation = adaptive_interpolation   # never read by the reference (dsm.cc)
        self.center_easting = center_easting
        self.center_northing = center_northing
        self.use_multi_threads = use_multi_threads             # both reference twins compute the same values


class Dsm(object):
    """dsm::Dsm (dsm.h:34-72)."""

    def __init__(self, settings, map):
        if map is None:
            raise AmbError(_lib.AMB_ERR_INVALID_ARGUMENT, "CHECK(map) (dsm.cc:22)")
        self.settings_ = settings
        self.debug = False
        self.last_debug = None
        # arithmetic of the gather's weights and sums (amb_dsm_set_precision): None = the library default ("f32":
        # float32 weights and sums, neighbour sets exact; the environment variable AMB_DSM_PRECISION=f64|f32, read by
        # the library itself, overrides that default process-wide — how the tests run both), or "f32" / "f64"
        self.precision = None
        # chunked evaluation + early mirroring of finished columns (amb_dsm_set_stream_chunks); 1 = off, default 4
        self.stream_chunks = max(1, int(os.environ.get("AMB_DSM_STREAM_CHUNKS", "4") or 4))

    def process(self, point_cloud, map):
        """point_cloud: float64 [n, 3] (the AoS layout of std::vector<Eigen::Vector3d>).  Mutates map['elevation']."""
        pc = np.ascontiguousarray(point_cloud, dtype=np.float64)
        n = pc.size // 3
        if n == 0:
            log.warning("Passed empty point cloud to DSM module")  # dsm.cc:189-192
            return
        if map is None:
            raise AmbError(_lib.AMB_ERR_INVALID_ARGUMENT, "CHECK(map) (dsm.cc:194)")
        ctx = map.context()
        s = self.settings_
        if not map.is_resident():
            map.upload(("elevation",))
        check(lib().amb_dsm_enable_debug(ctx, 1 if self.debug else 0), ctx)
        check(lib().amb_dsm_set_stream_chunks(ctx, int(self.stream_chunks)), ctx)
        if self.precision is not None:
            check(lib().amb_dsm_set_precision(ctx, _lib.DSM_F32 if self.precision == "f32" else _lib.DSM_F64), ctx)
        check(lib().amb_dsm_process(ctx, pc.ctypes.data_as(C.c_void_p), n, int(s.interpolation_radius),
                                    float(s.center_easting), float(s.center_northing)), ctx)
        self._fetch_debug(map)
        if not map.is_resident():
            map.download(("elevation",))

    def process_device(self, d_xyz, n, map, d_ids=None):
        """Points already in HBM (device pointer as int); asynchronous — map.sync() before reading results.
        d_ids: optional device pointer to one uint64 id per point (a sharded cloud: global point ids)."""
        if n == 0:
            log.warning("Passed empty point cloud to DSM module")
            return
        ctx = map.context()
        s = self.settings_
        check(lib().amb_dsm_enable_debug(ctx, 1 if self.debug else 0), ctx)
        check(lib().amb_dsm_set_stream_chunks(ctx, int(self.stream_chunks)), ctx)
        if self.precision is not None:
            check(lib().amb_dsm_set_precision(ctx, _lib.DSM_F32 if self.precision == "f32" else _lib.DSM_F64), ctx)
        if d_ids is None:
            check(lib().amb_dsm_process_device(ctx, C.c_void_p(int(d_xyz)), int(n), int(s.interpolation_radius),
                                               float(s.center_easting), float(s.center_northing)), ctx)
        else:
            check(lib().amb_dsm_process_device_ids(ctx, C.c_void_p(int(d_xyz)), C.c_void_p(int(d_ids)), int(n),
                                                   int(s.interpolation_radius), float(s.center_easting),
                                                   float(s.center_northing)), ctx)

    def process_sharded_device(self, d_xyz, d_ids, n_local, map, halo_capacity):
        """A cloud sharded by stripe, the exchange step included (amb_dsm_process_sharded_device): this rank's points
        and global ids (device pointers) -> border halo compaction -> one NCCL all-gather inside the library -> the
        stripe's elevation.  Collective over the communicator the map's context joined (sharding.init_comm);
        asynchronous — map.sync() before reading results."""
        ctx = map.context()
        s = self.settings_
        check(lib().amb_dsm_enable_debug(ctx, 1 if self.debug else 0), ctx)
        if self.precision is not None:
            check(lib().amb_dsm_set_precision(ctx, _lib.DSM_F32 if self.precision == "f32" else _lib.DSM_F64), ctx)
        check(lib().amb_dsm_process_sharded_device(ctx, C.c_void_p(int(d_xyz)), C.c_void_p(int(d_ids)), int(n_local),
                                                   int(s.interpolation_radius), float(s.center_easting),
                                                   float(s.center_northing), int(halo_capacity)), ctx)

    def process_sharded(self, point_cloud, ids, map, halo_capacity):
        """Same with this rank's points (float64 [n, 3]) and global ids (uint64 [n]) in host memory."""
        pc = np.ascontiguousarray(point_cloud, dtype=np.float64)
        idv = np.ascontiguousarray(ids, dtype=np.uint64)
        n = pc.size // 3
        ctx = map.context()
        s = self.settings_
        if self.precision is not None:
            check(lib().amb_dsm_set_precision(ctx, _lib.DSM_F32 if self.precision == "f32" else _lib.DSM_F64), ctx)
        check(lib().amb_dsm_process_sharded(ctx, pc.ctypes.data_as(C.c_void_p), idv.ctypes.data_as(C.c_void_p), n,
                                            int(s.interpolation_radius), float(s.center_easting),
                                            float(s.center_northing), int(halo_capacity)), ctx)

    def _fetch_debug(self, map):
        if not self.debug:
            self.last_debug = None
            return
        c0, c1 = map._col_range
        rows = map.geometry.rows
        cnt = np.empty((rows, c1 - c0), dtype=np.int32, order="F")
        lvl = np.empty((rows, c1 - c0), dtype=np.int8, order="F")
        check(lib().amb_dsm_download_debug(map.context(), cnt.ctypes.data_as(C.c_void_p),
                                           lvl.ctypes.data_as(C.c_void_p)), map.context())
        self.last_debug = (cnt, lvl)


def dsm_thresholds(interpolation_radius):
    buf = (C.c_double * 64)()
    n = lib().amb_dsm_thresholds(int(interpolation_radius), C.cast(buf, C.c_void_p), 64)
    if n < 0:
        raise AmbError(n)
    return [buf[k] for k in range(n)]


class OrthoSettings(object):
    """ortho::Settings of ortho-backward-grid.h:32-41.  Only colored_ortho affects results."""

    def __init__(self, show_orthomosaic_opencv=True, save_orthomosaic_jpg=True, orthomosaic_jpg_filename="",
                 orthomosaic_elevation_m=0.0, use_digital_elevation_map=True, colored_ortho=False,
                 use_multi_threads=True):
        self.show_orthomosaic_opencv = show_orthomosaic_opencv
        self.save_orthomosaic_jpg = save_orthomosaic_jpg
        self.orthomosaic_jpg_filename = orthomosaic_jpg_filename
        self.orthomosaic_elevation_m = orthomosaic_elevation_m
        self.use_digital_elevation_map = use_digital_elevation_map
        self.colored_ortho = colored_ortho
        self.use_multi_threads = use_multi_threads


class NCamera(object):
    """The part of aslam::NCamera the path uses: camera 0 (pinhole + distortion) and T_C_B(0)
    (ortho-backward-grid.cc:131,232)."""

    def __init__(self, width, height, fu, fv, cu, cv, dist_type=_lib.DIST_NONE, dist=(0.0, 0.0, 0.0, 0.0),
                 q_C_B=(1.0, 0.0, 0.0, 0.0), t_C_B=(0.0, 0.0, 0.0)):
        cam = Camera()
        cam.width, cam.height = int(width), int(height)
        cam.fu, cam.fv, cam.cu, cam.cv = float(fu), float(fv), float(cu), float(cv)
        cam.dist_type = int(dist_type)
        cam.dist = (C.c_double * 4)(*[float(d) for d in dist])
        cam.q_C_B = (C.c_double * 4)(*[float(d) for d in q_C_B])
        cam.t_C_B = (C.c_double * 3)(*[float(d) for d in t_C_B])
        self.camera = cam

    def imageWidth(self):
        return self.camera.width

    def imageHeight(self):
        return self.camera.height


class OrthoBackwardGrid(object):
    """ortho::OrthoBackwardGrid (ortho-backward-grid.h:43-71)."""

    def __init__(self, ncameras, settings, map=None):
        if ncameras is None:
            raise AmbError(_lib.AMB_ERR_INVALID_ARGUMENT, "CHECK(ncameras_) (ortho-backward-grid.cc:27)")
        self.ncameras_ = ncameras
        self.settings_ = settings
        self.brute_force = False
        # per-tile dominance cull of the frame list (amb_ortho_set_dominance_cull; same output bits, measured 4.70 ->
        # 2.72 ms at joint_10k): on by default, AMB_ORTHO_DOMINANCE=0 selects the plain conservative list
        self.dominance_cull = os.environ.get("AMB_ORTHO_DOMINANCE", "1") not in ("", "0")

    def _layers_written(self):
        out = "colored_ortho" if self.settings_.colored_ortho else "ortho"
        return ("elevation_angle", "observation_index", out)

    def process(self, T_G_Bs, images, map):
        """T_G_Bs: float64 [n, 7] rows x y z qw qx qy qz; images: n uint8 arrays [H, W] (gray) or [H, W, 3] (BGR)."""
        T = np.ascontiguousarray(T_G_Bs, dtype=np.float64).reshape(-1, 7)
        n = T.shape[0]
        if n == 0:
            raise AmbError(_lib.AMB_ERR_EMPTY, "CHECK(!T_G_Bs.empty()) (ortho-backward-grid.cc:225)")
        if n != len(images):
            raise AmbError(_lib.AMB_ERR_SIZE_MISMATCH,
                           "CHECK(T_G_Bs.size() == images.size()) (ortho-backward-grid.cc:226)")
        if map is None:
            raise AmbError(_lib.AMB_ERR_INVALID_ARGUMENT, "CHECK(map) (ortho-backward-grid.cc:227)")
        colored = bool(self.settings_.colored_ortho)
        channels = 3 if colored else 1
        cam = self.ncameras_.camera
        imgs = []
        for im in images:
            im = np.ascontiguousarray(im, dtype=np.uint8)
            ok = im.shape[:2] == (cam.height, cam.width) and (
                (im.ndim == 3 and im.shape[2] == 3) if colored else im.ndim == 2)
            if not ok:
                raise AmbError(_lib.AMB_ERR_SIZE_MISMATCH, "image shape %r" % (im.shape,))
            imgs.append(im)
        row_step = cam.width * channels
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        ctx = map.context()
        if not map.is_resident():
            map.upload(("elevation",) + self._layers_written())
        check(lib().amb_ortho_set_brute_force(ctx, 1 if self.brute_force else 0), ctx)
        check(lib().amb_ortho_set_dominance_cull(ctx, 1 if self.dominance_cull else 0), ctx)
        check(lib().amb_ortho_process(ctx, C.byref(cam), T.ctypes.data_as(C.c_void_p), C.cast(ptrs, C.c_void_p), n,
                                      channels, row_step, 1 if colored else 0), ctx)
        if not map.is_resident():
            map.download(self._layers_written())

    def process_device(self, T_G_Bs, d_image_ptrs, row_step, map):
        """Frames already in HBM: d_image_ptrs is a sequence of device pointers (ints).  Asynchronous."""
        T = np.ascontiguousarray(T_G_Bs, dtype=np.float64).reshape(-1, 7)
        n = T.shape[0]
        if n == 0 or n != len(d_image_ptrs):
            raise AmbError(_lib.AMB_ERR_SIZE_MISMATCH)
        colored = bool(self.settings_.colored_ortho)
        ptrs = (C.c_void_p * n)(*[int(p) for p in d_image_ptrs])
        ctx = map.context()
        check(lib().amb_ortho_set_brute_force(ctx, 1 if self.brute_force else 0), ctx)
        check(lib().amb_ortho_set_dominance_cull(ctx, 1 if self.dominance_cull else 0), ctx)
        check(lib().amb_ortho_process_device(ctx, C.byref(self.ncameras_.camera), T.ctypes.data_as(C.c_void_p),
                                             C.cast(ptrs, C.c_void_p), n, 3 if colored else 1, int(row_step),
                                             1 if colored else 0), ctx)


class OrthoFromPclSettings(object):
    """ortho::Settings of ortho-from-pcl.h:28-35 ("next" row N1)."""

    def __init__(self, show_orthomosaic_opencv=False, interpolation_radius=2, use_adaptive_interpolation=False,
                 save_orthomosaic_jpg=False, orthomosaic_jpg_filename=""):
        self.show_orthomosaic_opencv = show_orthomosaic_opencv
        self.interpolation_radius = int(interpolation_radius)
        self.use_adaptive_interpolation = use_adaptive_interpolation
        self.save_orthomosaic_jpg = save_orthomosaic_jpg
        self.orthomosaic_jpg_filename = orthomosaic_jpg_filename


class OrthoFromPcl(object):
    """ortho::OrthoFromPcl (ortho-from-pcl.h:37-52): IDW of point intensities into map['ortho']."""

    def __init__(self, settings):
        self.settings_ = settings

    def process(self, pointcloud, intensities, map):
        pc = np.ascontiguousarray(pointcloud, dtype=np.float64)
        n = pc.size // 3
        if n == 0:
            raise AmbError(_lib.AMB_ERR_EMPTY, "CHECK(!pointcloud.empty()) (ortho-from-pcl.cc:23)")
        if map is None:
            raise AmbError(_lib.AMB_ERR_INVALID_ARGUMENT, "CHECK(map) (ortho-from-pcl.cc:24)")
        inten = np.ascontiguousarray(intensities, dtype=np.int32)
        if inten.size < n:
            raise AmbError(_lib.AMB_ERR_SIZE_MISMATCH, "CHECK(i < intensities.size()) (ortho-from-pcl.cc:32)")
        ctx = map.context()
        s = self.settings_
        if not map.is_resident():
            map.upload(("ortho",))
        check(lib().amb_ortho_from_pcl_process(ctx, pc.ctypes.data_as(C.c_void_p), inten.ctypes.data_as(C.c_void_p),
                                               n, int(s.interpolation_radius),
                                               1 if s.use_adaptive_interpolation else 0), ctx)
        if not map.is_resident():
            map.download(("ortho",))


K_MAX_INVALID_DISPARITY = 1  # stereo::Densifier::kMaxInvalidDisparity (densifier.h:49)


def compute_point_cloud(disparity_map, image_left, K, baseline, R_G_C, t_G_C1, device=0,
                        max_invalid_disparity=K_MAX_INVALID_DISPARITY):
    """stereo::Densifier::computePointCloud (densifier.cpp:25-108; "next" row N3): disparity map (float32 [H, W]) +
    left rectified image (uint8 [H, W]) -> (point_cloud_eigen float64 [n, 3], point_cloud_intensities int32 [n]) in
    raster order.  K is the 3x3 camera matrix of the rectified pair (or (fx, fy, cx, cy))."""
    disp = np.ascontiguousarray(disparity_map, dtype=np.float32)
    img = np.ascontiguousarray(image_left, dtype=np.uint8)
    h, w = disp.shape
    if img.shape != (h, w):
        raise AmbError(_lib.AMB_ERR_SIZE_MISMATCH, "CHECK_EQ(image_resolution_, disparity_map.size())")
    K = np.asarray(K, dtype=np.float64)
    k4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]]) if K.shape == (3, 3) else K.reshape(4).copy()
    R = np.ascontiguousarray(R_G_C, dtype=np.float64).reshape(9)
    t = np.ascontiguousarray(t_G_C1, dtype=np.float64).reshape(3)
    cap = h * w
    xyz = np.empty((cap, 3), dtype=np.float64)
    inten = np.empty(cap, dtype=np.int32)
    n = C.c_size_t(0)
    check(lib().amb_stereo_reproject(int(device), disp.ctypes.data_as(C.c_void_p), w, img.ctypes.data_as(C.c_void_p),
                                     w, w, h, k4.ctypes.data_as(C.c_void_p), float(baseline),
                                     R.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),
                                     float(max_invalid_disparity), xyz.ctypes.data_as(C.c_void_p),
                                     inten.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    return xyz[:n.value].copy(), inten[:n.value].copy()


def rectify_stereo_setup(K, R_G_C1, R_G_C2, t_G_C1, t_G_C2):
    """Host half of stereo::Rectifier::rectifyStereoPair (rectifier.cpp:43-79; "next" row N3): Fusiello's compact
    rectification.  Returns (baseline, R_G_C_rect float64 [3, 3], T1_inv float32 [3, 3], T2_inv float32 [3, 3])."""
    a = [np.ascontiguousarray(m, dtype=np.float64).reshape(n) for m, n in
         ((K, 9), (R_G_C1, 9), (R_G_C2, 9), (t_G_C1, 3), (t_G_C2, 3))]
    baseline = C.c_double(0.0)
    R = np.zeros(9, np.float64)
    T1 = np.zeros(9, np.float32)
    T2 = np.zeros(9, np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    check(lib().amb_stereo_rectify_setup(p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]), C.byref(baseline), p(R), p(T1),
                                         p(T2)))
    return baseline.value, R.reshape(3, 3), T1.reshape(3, 3), T2.reshape(3, 3)


def rectify_stereo_maps(T1_inv, T2_inv, width, height, device=0):
    """Device half (rectifier.cpp:80-104): the four CV_32FC1 rectification maps cv::remap consumes.
    Returns (map_rectify_1_x, map_rectify_1_y, map_rectify_2_x, map_rectify_2_y), float32 [H, W] each."""
    T1 = np.ascontiguousarray(T1_inv, dtype=np.float32).reshape(9)
    T2 = np.ascontiguousarray(T2_inv, dtype=np.float32).reshape(9)
    maps = [np.empty((int(height), int(width)), np.float32) for _ in range(4)]
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    check(lib().amb_stereo_rectify_maps(int(device), p(T1), p(T2), int(width), int(height), int(width), p(maps[0]),
                                        p(maps[1]), p(maps[2]), p(maps[3])))
    return tuple(maps)
