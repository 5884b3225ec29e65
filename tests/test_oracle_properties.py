"""Property tests of the oracle (hypothesis; CPU): the invariants SURVEY.md §4 lists for the path."""
import numpy as np
from hypothesis import given, settings, strategies as st

from common import fresh_layers, ulp_diff
from aerial_mapper_b200 import synth
from oracle import pyoracle as po


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), rows=st.integers(5, 24), cols=st.integers(5, 24),
       res=st.sampled_from([0.25, 0.5, 1.0]), radius=st.integers(1, 3))
def test_dsm_point_order_does_not_change_the_decisions(seed, rows, cols, res, radius):
    xyz = synth.point_cloud(300, rows * res / 2 + 1, cols * res / 2 + 1, seed)
    g = po.make_geometry(rows, cols, res)
    a = np.full((rows, cols), np.nan, np.float32, order="F")
    b = a.copy(order="F")
    st1, c1, l1, _ = po.dsm_process(g, a, xyz, radius=radius, num_threads=-1, debug=True)
    perm = np.random.default_rng(seed).permutation(len(xyz))
    st2, c2, l2, _ = po.dsm_process(g, b, xyz[perm], radius=radius, num_threads=-1, debug=True)
    assert st1 == 0 and st2 == 0
    assert np.array_equal(c1, c2) and np.array_equal(l1, l2)       # neighbour sets / retry levels: exact
    assert ulp_diff(a, b).max() <= 1                               # heights: summation order only
    ok = ~np.isnan(a)
    if ok.any():                                                   # IDW is a convex combination of heights
        assert a[ok].min() >= xyz[:, 2].min() - 1e-3 and a[ok].max() <= xyz[:, 2].max() + 1e-3


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10_000), split=st.integers(1, 5), colored=st.booleans())
def test_ortho_batch_split_invariance(seed, split, colored):
    rows, cols, res = 40, 32, 0.5
    camd = synth.scaled_camera(0.04)
    poses = synth.lawnmower_poses(2, 3, rows * res / 2, cols * res / 2, 40.0, seed=seed, jitter_pos=0.5)
    ch = 3 if colored else 1
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], ch) for k in range(6)]
    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)
    elev = synth.analytic_elevation(rows, cols, res)
    A = fresh_layers(rows, cols, elev)
    B = fresh_layers(rows, cols, elev)
    assert po.ortho_process(g, A, cam, poses, imgs, colored=colored, num_threads=-1)[0] == 0
    assert po.ortho_process(g, B, cam, poses[:split], imgs[:split], colored=colored, num_threads=-1)[0] == 0
    assert po.ortho_process(g, B, cam, poses[split:], imgs[split:], colored=colored, num_threads=-1)[0] == 0
    key = "colored_ortho" if colored else "ortho"
    assert np.array_equal(A[key].view(np.uint32), B[key].view(np.uint32))
    assert np.array_equal(A["elevation_angle"], B["elevation_angle"])
    late = A["observation_index"] >= split                         # batch-relative frame indices
    assert np.array_equal(B["observation_index"][late], A["observation_index"][late] - split)
