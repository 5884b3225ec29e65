"""Randomised cases of the DSM kernels (count / scan / scatter / bucket order / gather / warp-per-cell) on the CPU EMULATION of the kernel source (tests/emu) against the oracle:
sizes from 1x1, every distortion model / radius / density / offset the generators below draw.  No GPU involved.

    python tools/emu_fuzz_dsm.py <seed> <cases>      # exits non-zero on any mismatch
"""
import ctypes as C, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["AMB_TEST_EMU"] = "1"
import numpy as np
import conftest  # swaps the library
import aerial_mapper_b200 as amb
from aerial_mapper_b200 import synth
from oracle import pyoracle as po
from common import ulp_diff, fresh_layers
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
fails = 0
t0 = time.time()
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    rows, cols = int(rng.integers(1, 90)), int(rng.integers(1, 90))
    res = float(rng.choice([0.1, 0.25, 0.4, 0.5, 1.0, 2.0, 3.0]))
    radius = int(rng.choice([1, 1, 1, 2, 3, 5, 8]))
    pos = (float(rng.uniform(-1000, 1000)), float(rng.uniform(-1000, 1000)))
    ce, cn = (float(rng.uniform(-20, 20)), float(rng.uniform(-20, 20))) if rng.random() < 0.5 else (0.0, 0.0)
    dens = float(rng.choice([0.02, 0.3, 2.0, 20.0, 200.0]))
    n = max(1, min(200000, int(dens * rows * res * cols * res)))
    hx, hy = rows * res / 2, cols * res / 2
    mode = rng.integers(0, 4)
    xyz = np.c_[rng.uniform(-hx - 3, hx + 3, n), rng.uniform(-hy - 3, hy + 3, n), rng.normal(100, 5, n)]
    if mode == 1:   # clustered: a dense blob -> dense tiles
        k = n // 2
        xyz[:k, 0] = rng.normal(0, min(hx, 2.0) / 3, k); xyz[:k, 1] = rng.normal(0, min(hy, 2.0) / 3, k)
    if mode == 2:   # a hole in the middle
        keep = ~((np.abs(xyz[:, 0]) < hx / 2) & (np.abs(xyz[:, 1]) < hy / 2))
        xyz = xyz[keep] if keep.sum() > 0 else xyz[:1]
    xyz[:, 0] += pos[0] + cn; xyz[:, 1] += pos[1] + ce
    W = np.sqrt(radius) / res
    if W > 100: continue
    try:
        gm = amb.AerialGridMap(amb.GridMapSettings(pos[0], pos[1], rows * res, cols * res, res)).getMutable()
        if gm.getSize() != (rows, cols): continue
        d = amb.Dsm(amb.DsmSettings(interpolation_radius=radius, center_easting=ce, center_northing=cn), gm); d.debug = True
        d.process(xyz, gm)
        st_ok = True
    except amb.AmbError as e:
        print("it", it, "AmbError", e, (rows, cols, res, radius, n, mode)); st_ok = False
    if not st_ok: continue
    g = po.make_geometry(rows, cols, res, pos[0], pos[1])
    e = np.full((rows, cols), np.nan, np.float32, order="F")
    st, cnt, lvl, _ = po.dsm_process(g, e, xyz, radius, ce, cn, num_threads=-1, debug=True)
    # heights: f64 mode <= 1 float32 ulp; f32 mode (library default: float32 weights / sums relative to the tile's height
    # offset, this generator's heights scatter with sd 5 m inside a tile) 2e-6 relative — north_star allows 1e-4
    f64 = os.environ.get("AMB_DSM_PRECISION", "f32").lower().startswith("f6")
    fin = ~np.isnan(e)
    close = ulp_diff(gm["elevation"], e).max() <= 1 if f64 else np.allclose(gm["elevation"][fin], e[fin], rtol=2e-6, atol=0)
    ok = st == 0 and np.array_equal(d.last_debug[1], lvl) and np.array_equal(np.isnan(e), np.isnan(gm["elevation"])) and close
    touched = lvl >= 0
    ok = ok and np.array_equal(d.last_debug[0][touched], cnt[touched])
    if not ok:
        fails += 1
        print("MISMATCH it", it, (rows, cols, res, radius, n, mode, pos, ce, cn), "st", st)
print("done", it + 1, "cases, fails", fails, "%.0fs" % (time.time() - t0))
sys.exit(1 if fails else 0)
