"""Generate the committed golden fixtures (small seeded inputs -> oracle outputs).

The reference has no golden vectors for this path, so these are outputs of THIS repository's
oracle — the DSM one produced with oracle/_ref (the reference's own nanoflann.hpp compiled verbatim + restated loop,
which also records neighbour counts and retry levels) when /root/reference is present, which is how the committed
files were generated.  tests/test_oracle_refsrc.py then checks that the reference's OWN dsm.cc and
ortho-backward-grid.cc (compiled verbatim into oracle/_ref) reproduce every committed layer bit for bit, and
tests/test_golden.py that the portable restatement does.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from aerial_mapper_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def dsm_case():
    rows, cols, res = 96, 80, 0.5
    xyz = synth.point_cloud(9000, rows * res / 2, cols * res / 2, seed=21, holes=3, hole_seed=22,
                            hole_sides=(3.0, 9.0))
    g = po.make_geometry(rows, cols, res)
    elev = np.full((rows, cols), np.nan, np.float32, order="F")
    use_ref = po.have_ref()
    st, cnt, lvl, _ = po.dsm_process(g, elev, xyz, num_threads=-1, debug=True, use_ref=use_ref)
    assert st == 0
    np.savez_compressed(os.path.join(HERE, "dsm_96x80.npz"), rows=rows, cols=cols, res=res, xyz=xyz,
                        elevation=elev, neighbour_count=cnt, threshold_index=lvl,
                        generator="oracle/_ref (nanoflann verbatim)" if use_ref else "oracle restatement")
    print("dsm golden:", "ref" if use_ref else "port", "nan cells", int(np.isnan(elev).sum()),
          "levels", np.bincount(lvl[lvl >= 0].ravel().astype(np.int64)).tolist())


def ortho_case(colored):
    rows, cols, res = 96, 80, 0.5
    camd = synth.scaled_camera(0.04, dist_type=1)
    poses = synth.lawnmower_poses(2, 4, rows * res / 2, cols * res / 2, agl=40.0, seed=23, jitter_pos=1.0)
    ch = 3 if colored else 1
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], ch) for k in range(len(poses))]
    g = po.make_geometry(rows, cols, res)
    elev = synth.analytic_elevation(rows, cols, res)
    elev[10:14, 20:30] = np.nan  # cells without elevation stay untouched
    L = {"elevation": elev, "elevation_angle": np.zeros((rows, cols), np.float32, order="F"),
         "observation_index": np.full((rows, cols), np.nan, np.float32, order="F"),
         "ortho": np.full((rows, cols), 255.0, np.float32, order="F"),
         "colored_ortho": np.full((rows, cols), np.nan, np.float32, order="F")}
    st, _ = po.ortho_process(g, L, po.make_camera(**camd), poses, imgs, colored=colored, num_threads=-1)
    assert st == 0
    name = "ortho_color_96x80.npz" if colored else "ortho_gray_96x80.npz"
    np.savez_compressed(os.path.join(HERE, name), rows=rows, cols=cols, res=res, poses=poses,
                        cam_scale=0.04, elevation=elev, elevation_angle=L["elevation_angle"],
                        observation_index=L["observation_index"],
                        out=(L["colored_ortho"] if colored else L["ortho"]).view(np.uint32))
    print("ortho golden colored=%d covered=%d" % (colored, int((~np.isnan(L["observation_index"])).sum())))


if __name__ == "__main__":
    dsm_case()
    ortho_case(False)
    ortho_case(True)
