"""Randomised adaptive-interpolation cases of OrthoFromPcl (ortho-from-pcl.cc:63-72) on the CPU EMULATION of the kernel
source (tests/emu) against the REFERENCE'S OWN ortho-from-pcl.cc compiled verbatim (oracle/_ref): sparse and
corner-concentrated clouds, so that cells need the 10r / 100r / 1000r ... balls.  No GPU involved.

    python tools/emu_fuzz_pcl.py <seed> <cases>
"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["AMB_TEST_EMU"] = "1"
import numpy as np
import conftest
import aerial_mapper_b200 as amb
from oracle import pyoracle as po
from common import ulp_diff
rng = np.random.default_rng(int(sys.argv[1]))
fails = 0
for it in range(int(sys.argv[2])):
    rows, cols = int(rng.integers(1, 120)), int(rng.integers(1, 120))
    res = float(rng.choice([0.25, 0.5, 1.0, 2.0])); radius = int(rng.choice([1, 2, 2, 3]))
    n = max(1, int(rng.choice([0.002, 0.02, 0.3, 3.0]) * rows * res * cols * res))
    hx, hy = rows * res / 2, cols * res / 2
    xyz = np.c_[rng.uniform(-hx, hx, n) * 0.999, rng.uniform(-hy, hy, n) * 0.999, rng.uniform(0, 5, n)]
    if rng.random() < 0.5 and n > 10:   # concentrate the points in a corner: large empty regions -> high levels
        xyz[:, 0] = -hx * 0.999 + (xyz[:, 0] + hx) * 0.2; xyz[:, 1] = -hy * 0.999 + (xyz[:, 1] + hy) * 0.2
    inten = rng.integers(0, 256, n).astype(np.int32)
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    if gm.getSize() != (rows, cols): continue
    try:
        amb.OrthoFromPcl(amb.OrthoFromPclSettings(interpolation_radius=radius, use_adaptive_interpolation=True)).process(xyz, inten, gm)
    except amb.AmbError as e:
        print("it", it, "AmbError", e, (rows, cols, res, radius, n)); continue
    o = np.full((rows, cols), 255.0, np.float32, order="F")
    st = po.refsrc_ortho_from_pcl_process(po.make_geometry(rows, cols, res), o, xyz, inten, radius, True)
    ok = st == 0 and np.isfinite(gm["ortho"]).all() and ulp_diff(gm["ortho"], o).max() <= 1
    if not ok:
        fails += 1; print("MISMATCH", it, (rows, cols, res, radius, n), st, ulp_diff(gm["ortho"], o).max())
print("done, fails", fails)
sys.exit(1 if fails else 0)
