"""CPU study for the next round: a per-tile DOMINANCE cull of orthomosaic frames.

Today `ortho_kernel` evaluates, for every cell, every frame that survives the tile's view-cone / view-rectangle tests
(~8.5 of 250 at `joint_10k`) although only one wins: the frame with the largest observation angle
alpha = asin(|z_c| / |X - c|) = pi/2 - theta, theta = angle between the ray and the camera's optical axis.
For a tile with bounding sphere (C, rho) and a frame with camera centre c and axis a:
    theta_c = angle(a, C - c),  delta = asin(rho / |C - c|)   =>   theta in [max(0, theta_c - delta), theta_c + delta]
If some frame g is VISIBLE FROM EVERY CELL of the tile (sphere inside the view rectangle shrunk by the distortion
bound E, in front of the camera), every frame f with theta_min(f) > theta_max(g) + margin loses to g in every cell and
can be dropped without changing any output bit: the reference's sequential float32 running-max recurrence ends in the
same (angle, index) whether or not f was visited (g, visited before or after f, exceeds f by more than any float32
rounding of the running best).

This script (numpy + the CPU oracle; no GPU)
  1. applies the cull per 32x32 tile on a scaled copy of the benchmark scene and checks that the oracle restricted to the
     surviving frames reproduces the full oracle bit for bit (observation_index, elevation_angle, ortho), and
  2. reports how many frames per tile survive at the full `joint_10k` geometry (statistics only).

    python tools/ortho_dominance_study.py [--check-size 640] [--stats-tiles 4000]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aerial_mapper_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

TILE = 32
MARGIN = 1e-4  # rad; float32 rounding of an angle <= pi/2 is < 1.2e-7


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def view_rect(camd):
    """Outer and inner bounds of the undistorted normalised coordinates of visible rays (pinhole / rad-tan):
    the model moves a keypoint by at most E = |k1| r^3 + |k2| r^5 + 4 (|p1| + |p2|) r^2 (ortho_kernels.cu
    compute_view_rect).  outer: no visible ray outside; inner: every ray inside is visible."""
    W, H, fu, fv, cu, cv = (camd[k] for k in ("width", "height", "fu", "fv", "cu", "cv"))
    k1, k2, p1, p2 = camd["dist"] if camd["dist_type"] == 1 else (0, 0, 0, 0)
    lo_u, hi_u, lo_v, hi_v = -cu / fu, (W - cu) / fu, -cv / fv, (H - cv) / fv
    r = 1.3 * np.hypot(max(-lo_u, hi_u), max(-lo_v, hi_v))
    E = abs(k1) * r ** 3 + abs(k2) * r ** 5 + 4 * (abs(p1) + abs(p2)) * r ** 2
    outer = (lo_u - E, hi_u + E, lo_v - E, hi_v + E)
    eps = 1e-9
    inner = (lo_u + E + eps, hi_u - E - eps, lo_v + E + eps, hi_v - E - eps)
    return outer, inner


def tile_frames(C, rho, cams, Rs, outer, inner):
    """(candidate mask, survivors-after-dominance mask) for one tile sphere against all frames (vectorised)."""
    d = C[None, :] - cams                                    # [F, 3]
    pc = np.einsum("fij,fi->fj", Rs, d)                      # R_G_C^T d: camera coordinates of the sphere centre
    xp, yp, zp = pc[:, 0], pc[:, 1], pc[:, 2]
    dist = np.linalg.norm(d, axis=1)
    slack = rho * (1 + 1e-9) + 1e-9 * (np.abs(xp) + np.abs(yp) + np.abs(zp))

    def outside(val, bound, sign):       # sphere entirely beyond the plane  sign * (coord - bound * z) > 0
        n = np.sqrt(1.0 + bound * bound)
        return sign * (val - bound * zp) > slack * n

    cand = zp + rho > 0
    cand &= ~outside(xp, outer[1], +1) & ~outside(xp, outer[0], -1)
    cand &= ~outside(yp, outer[3], +1) & ~outside(yp, outer[2], -1)

    def inside(val, bound, sign):        # sphere entirely on the inner side of the plane
        n = np.sqrt(1.0 + bound * bound)
        return sign * (val - bound * zp) < -slack * n

    full = (zp - rho > 1e-6) & inside(xp, inner[1], +1) & inside(xp, inner[0], -1)
    full &= inside(yp, inner[3], +1) & inside(yp, inner[2], -1)

    perp = np.sqrt(np.maximum(dist * dist - zp * zp, 0.0))
    theta_c = np.arctan2(perp, zp)
    delta = np.arcsin(np.minimum(1.0, rho / np.maximum(dist, 1e-300)))
    th_min = np.maximum(0.0, theta_c - delta)
    th_max = theta_c + delta
    keep = cand.copy()
    if (cand & full).any():
        best = th_max[cand & full].min()
        keep &= ~(th_min > best + MARGIN)
    return cand, keep


def scene(rows, cols, res, lines, per_line, agl, cam_scale, seed=4):
    half_x, half_y = rows * res / 2, cols * res / 2
    camd = synth.scaled_camera(cam_scale)
    poses = synth.lawnmower_poses(lines, per_line, half_x, half_y, agl, seed=seed)
    cams = poses[:, :3].copy()
    Rs = np.stack([quat_to_R(p[3:7]) for p in poses])
    return camd, poses, cams, Rs


def tile_sphere(qx, qy, elev, i0, j0):
    i1, j1 = min(i0 + TILE, len(qx)) - 1, min(j0 + TILE, len(qy)) - 1
    e = elev[i0:i1 + 1, j0:j1 + 1]
    e = e[~np.isnan(e)]
    if e.size == 0:
        return None
    zmin, zmax = float(e.min()), float(e.max())
    C = np.array([0.5 * (qx[i0] + qx[i1]), 0.5 * (qy[j0] + qy[j1]), 0.5 * (zmin + zmax)])
    ext = np.array([0.5 * (qx[i0] - qx[i1]), 0.5 * (qy[j0] - qy[j1]), 0.5 * (zmax - zmin)])
    return C, float(np.linalg.norm(ext)) * (1 + 1e-9) + 1e-6


def check(size):
    """Exactness on a scaled scene: same flight pattern density as joint_10k (10 x 25 frames), 1/5 of its extent."""
    rows = cols = size
    res = 0.25
    scale = size * res / 2500.0
    lines, per_line = 10, 25
    camd, poses, cams, Rs = scene(rows, cols, res, lines, per_line, 400.0 * scale, max(0.02, scale))
    outer, inner = view_rect(camd)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(len(poses))]
    elev = synth.analytic_elevation(rows, cols, res)
    elev[5:9, 40:47] = np.nan
    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)

    def fresh():
        return {"elevation": elev, "elevation_angle": np.zeros((rows, cols), np.float32, order="F"),
                "observation_index": np.full((rows, cols), np.nan, np.float32, order="F"),
                "ortho": np.full((rows, cols), 255.0, np.float32, order="F")}

    truth = fresh()
    st, _ = po.ortho_process(g, truth, cam, poses, imgs)
    assert st == 0
    qx, qy = synth.grid_positions(rows, cols, res)
    culled = fresh()
    n_cand, n_keep = [], []
    for j0 in range(0, cols, TILE):
        for i0 in range(0, rows, TILE):
            sph = tile_sphere(qx, qy, elev, i0, j0)
            if sph is None:
                continue
            cand, keep = tile_frames(sph[0], sph[1], cams, Rs, outer, inner)
            n_cand.append(int(cand.sum()))
            n_keep.append(int(keep.sum()))
            sub = np.nonzero(keep)[0]
            if len(sub) == 0:
                continue
            tile_layers = fresh()
            for j in range(j0, min(j0 + TILE, cols)):
                k0 = rows * j + i0
                st, _ = po.ortho_process(g, tile_layers, cam, poses[sub], [imgs[s] for s in sub], num_threads=-1,
                                         cell_range=(k0, k0 + min(TILE, rows - i0)))
                assert st == 0
            sl = (slice(i0, i0 + TILE), slice(j0, j0 + TILE))
            oi = tile_layers["observation_index"][sl]
            culled["observation_index"][sl] = np.where(np.isnan(oi), np.nan, sub[np.nan_to_num(oi).astype(int)])
            culled["elevation_angle"][sl] = tile_layers["elevation_angle"][sl]
            culled["ortho"][sl] = tile_layers["ortho"][sl]
    for k in ("observation_index", "elevation_angle", "ortho"):
        same = np.array_equal(truth[k].view(np.uint32), culled[k].view(np.uint32))
        print("  %-18s %s" % (k, "bit-identical" if same else "DIFFERS in %d cells" %
                              int((truth[k].view(np.uint32) != culled[k].view(np.uint32)).sum())))
        assert same, k
    print("  check scene %dx%d cells, %d frames of %dx%d: candidates/tile %.2f -> survivors/tile %.2f"
          % (rows, cols, len(poses), camd["width"], camd["height"], np.mean(n_cand), np.mean(n_keep)))


def stats(n_tiles, seed=0):
    """Survivor statistics at the full joint_10k geometry (sphere tests only; analytic terrain for the tile range)."""
    rows = cols = 10000
    res = 0.25
    camd, poses, cams, Rs = scene(rows, cols, res, 10, 25, 400.0, 1.0)
    outer, inner = view_rect(camd)
    qx, qy = synth.grid_positions(rows, cols, res)
    rng = np.random.default_rng(seed)
    n_cand, n_keep = [], []
    for _ in range(n_tiles):
        i0 = TILE * int(rng.integers(0, rows // TILE))
        j0 = TILE * int(rng.integers(0, cols // TILE))
        X, Y = np.meshgrid(qx[i0:i0 + TILE], qy[j0:j0 + TILE], indexing="ij")
        z = synth.terrain(X, Y).astype(np.float32)
        C = np.array([0.5 * (qx[i0] + qx[i0 + TILE - 1]), 0.5 * (qy[j0] + qy[j0 + TILE - 1]),
                      0.5 * (float(z.min()) + float(z.max()))])
        ext = np.array([0.5 * (qx[i0] - qx[i0 + TILE - 1]), 0.5 * (qy[j0] - qy[j0 + TILE - 1]),
                        0.5 * (float(z.max()) - float(z.min()))])
        cand, keep = tile_frames(C, float(np.linalg.norm(ext)) * (1 + 1e-9) + 1e-6, cams, Rs, outer, inner)
        n_cand.append(int(cand.sum()))
        n_keep.append(int(keep.sum()))
    n_cand, n_keep = np.array(n_cand), np.array(n_keep)
    print("  joint_10k geometry, %d random tiles: candidates/tile mean %.2f (max %d) -> survivors mean %.2f (max %d); "
          "tiles with no fully-visible frame: %.1f %%"
          % (n_tiles, n_cand.mean(), n_cand.max(), n_keep.mean(), n_keep.max(), 100.0 * np.mean(n_keep == n_cand)))
    print("  histogram of survivors:", np.bincount(n_keep)[:12].tolist())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check-size", type=int, default=640)
    ap.add_argument("--stats-tiles", type=int, default=4000)
    a = ap.parse_args()
    print("exactness check (oracle restricted to the surviving frames vs full oracle):")
    check(a.check_size)
    print("survivor statistics:")
    stats(a.stats_tiles)
