"""Seeded synthetic inputs for the hot path: point clouds, terrain, lawn-mower poses, frames.

The reference ships no sample data (its flag files point at /tmp/simulation); every workload here is
generated.  numpy.random.Generator(PCG64(seed)), float64 throughout.
"""
import numpy as np


def terrain(x, y):
    """Analytic terrain z(x, y) used by every synthetic workload."""
    return 100.0 + 10.0 * np.sin(0.01 * x) * np.cos(0.01 * y)


def grid_positions(rows, cols, resolution, pos_x=0.0, pos_y=0.0):
    """Cell-centre coordinates with grid_map's getPosition rule (index (0,0) = max-x / max-y corner)."""
    len_x, len_y = rows * resolution, cols * resolution
    qx = (pos_x + (0.5 * len_x - 0.5 * resolution)) + resolution * (-np.arange(rows, dtype=np.float64))
    qy = (pos_y + (0.5 * len_y - 0.5 * resolution)) + resolution * (-np.arange(cols, dtype=np.float64))
    return qx, qy


def analytic_elevation(rows, cols, resolution, pos_x=0.0, pos_y=0.0):
    """float32 column-major elevation layer filled with terrain() at the cell centres (config C3/C5)."""
    qx, qy = grid_positions(rows, cols, resolution, pos_x, pos_y)
    z = terrain(qx[:, None], qy[None, :])
    return np.asfortranarray(z.astype(np.float32))


def point_cloud(n, half_x, half_y, seed, noise=0.05, holes=0, hole_seed=3, hole_sides=(5.0, 60.0),
                center=(0.0, 0.0)):
    """n points, x ~ U(-half_x, half_x), y ~ U(-half_y, half_y) around `center`, z = terrain + N(0, noise).

    holes > 0 deletes the points inside that many random axis-aligned rectangles (variant C2h) to exercise the
    expanding-radius fallback and permanently empty cells.  Returns float64 [m, 3] (AoS, 24-byte stride)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = np.empty((n, 3), dtype=np.float64)
    xyz[:, 0] = rng.uniform(-half_x, half_x, n) + center[0]
    xyz[:, 1] = rng.uniform(-half_y, half_y, n) + center[1]
    xyz[:, 2] = terrain(xyz[:, 0], xyz[:, 1]) + rng.normal(0.0, noise, n)
    if holes:
        hr = np.random.Generator(np.random.PCG64(hole_seed))
        keep = np.ones(n, dtype=bool)
        for _ in range(holes):
            cx = hr.uniform(-half_x, half_x) + center[0]
            cy = hr.uniform(-half_y, half_y) + center[1]
            sx, sy = hr.uniform(hole_sides[0], hole_sides[1], 2)
            keep &= ~((np.abs(xyz[:, 0] - cx) < 0.5 * sx) & (np.abs(xyz[:, 1] - cy) < 0.5 * sy))
        xyz = np.ascontiguousarray(xyz[keep])
    return xyz


def _quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx])


def _axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def lawnmower_poses(lines, per_line, half_x, half_y, agl, seed, mean_terrain=100.0, jitter_rp_deg=2.0,
                    jitter_yaw_deg=5.0, jitter_pos=3.0, center=(0.0, 0.0)):
    """Body poses T_G_B of a lawn-mower survey: `lines` flight lines along x, `per_line` frames each, `agl` metres
    above the mean terrain, nadir-looking (camera z-axis = -Z world) with per-frame attitude / position jitter so
    that no two frames tie exactly in observation angle.  Rows are x y z qw qx qy qz (aerial-mapper-io.cc:110)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q_nadir = np.array([0.0, 1.0, 0.0, 0.0])  # rotation by pi about X: x->x, y->-y, z->-z
    poses = np.empty((lines * per_line, 7), dtype=np.float64)
    k = 0
    for line in range(lines):
        y = -half_y + (line + 0.5) * (2.0 * half_y / lines)
        forward = (line % 2 == 0)
        for f in range(per_line):
            x = -half_x + (f + 0.5) * (2.0 * half_x / per_line)
            if not forward:
                x = -x
            pos = np.array([x + center[0], y + center[1], mean_terrain + agl]) + rng.normal(0.0, jitter_pos, 3)
            yaw = (0.0 if forward else np.pi) + np.deg2rad(rng.normal(0.0, jitter_yaw_deg))
            roll = np.deg2rad(rng.normal(0.0, jitter_rp_deg))
            pitch = np.deg2rad(rng.normal(0.0, jitter_rp_deg))
            q = _quat_mul(_axis_angle((0, 0, 1), yaw), q_nadir)
            q = _quat_mul(q, _axis_angle((1, 0, 0), roll))
            q = _quat_mul(q, _axis_angle((0, 1, 0), pitch))
            q /= np.linalg.norm(q)
            if q[0] < 0:
                q = -q
            poses[k, :3] = pos
            poses[k, 3:] = q
            k += 1
    return poses


# constants of the procedural frames; channel c of frame k at pixel (u, v):
#   (A[c]*u + B[c]*v + C[c]*k + ((u*v) >> 3)) & 0xFF
IMAGE_A = (131, 73, 199)
IMAGE_B = (241, 151, 37)
IMAGE_C = (17, 29, 43)


_BASE_CACHE = {}


def _image_base(width, height, c):
    """(A[c]*u + B[c]*v + ((u*v) >> 3)) & 0xFF as uint8 [H, W]; frame k adds C[c]*k modulo 256."""
    key = (width, height, c)
    if key not in _BASE_CACHE:
        u = np.arange(width, dtype=np.int64)[None, :]
        v = np.arange(height, dtype=np.int64)[:, None]
        _BASE_CACHE.clear() if len(_BASE_CACHE) > 6 else None
        _BASE_CACHE[key] = ((IMAGE_A[c] * u + IMAGE_B[c] * v + ((u * v) >> 3)) & 0xFF).astype(np.uint8)
    return _BASE_CACHE[key]


def procedural_image(k, width, height, channels=1):
    """Frame k as uint8 [H, W] (gray) or [H, W, 3] (B, G, R byte order like cv::Mat CV_8UC3)."""
    if channels == 1:
        return _image_base(width, height, 0) + np.uint8((IMAGE_C[0] * k) & 0xFF)  # uint8 add wraps modulo 256
    img = np.empty((height, width, 3), dtype=np.uint8)
    for c in range(3):
        img[:, :, c] = _image_base(width, height, c) + np.uint8((IMAGE_C[c] * k) & 0xFF)
    return img


def procedural_images_torch(n, width, height, channels, device):
    """All n frames as one uint8 torch tensor [n, H, W(, 3)] generated on `device` (bench plumbing: 250 frames of
    4000x3000 are 3-9 GB, too slow to build with numpy on the host every run)."""
    import torch
    u = torch.arange(width, dtype=torch.int32, device=device)[None, :]
    v = torch.arange(height, dtype=torch.int32, device=device)[:, None]
    uv = (u * v) >> 3
    shape = (n, height, width) if channels == 1 else (n, height, width, 3)
    out = torch.empty(shape, dtype=torch.uint8, device=device)
    bases = [((IMAGE_A[c] * u + IMAGE_B[c] * v + uv) & 0xFF).to(torch.uint8) for c in range(channels)]
    for k in range(n):
        if channels == 1:
            out[k] = bases[0] + ((IMAGE_C[0] * k) & 0xFF)  # uint8 add wraps modulo 256
        else:
            for c in range(3):
                out[k, :, :, c] = bases[c] + ((IMAGE_C[c] * k) & 0xFF)
    return out


# The camera of configs C3-C5: pinhole f=3000, monotone rad-tan.
C3_CAMERA = dict(width=4000, height=3000, fu=3000.0, fv=3000.0, cu=2000.0, cv=1500.0, dist_type=1,
                 dist=(-0.05, 0.01, 1e-4, 1e-4))


def scaled_camera(scale, dist_type=1, dist=(-0.05, 0.01, 1e-4, 1e-4)):
    """The C3 camera with the raster shrunk by `scale` (same field of view): small parity-test frames."""
    return dict(width=int(round(4000 * scale)), height=int(round(3000 * scale)), fu=3000.0 * scale,
                fv=3000.0 * scale, cu=2000.0 * scale, cv=1500.0 * scale, dist_type=dist_type, dist=tuple(dist))
