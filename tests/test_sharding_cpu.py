"""Host-side sharding logic that needs no GPU: stripe y-intervals tile the axis, every point has one owner."""
import numpy as np

import aerial_mapper_b200 as amb
from aerial_mapper_b200 import sharding, synth


def test_y_intervals_tile_the_map_and_owners_partition_the_cloud():
    gm = amb.GridMap()
    gm.setGeometry((37.0, 53.0), 0.5, (10.0, -20.0))
    rows, cols = gm.getSize()
    world = 4
    qx, qy = synth.grid_positions(rows, cols, 0.5, 10.0, -20.0)
    prev_lo = None
    y = np.random.default_rng(0).uniform(-60.0, 20.0, 5000)
    owners = np.zeros(y.size, int)
    for r in range(world):
        c0, c1 = sharding.stripe_range(cols, r, world)
        lo, hi = sharding.stripe_y_interval(gm.geometry, c0, c1)
        assert np.isclose(hi, qy[c0] + 0.25) and np.isclose(lo, qy[c1 - 1] - 0.25)
        if prev_lo is not None:
            assert hi == prev_lo            # column 0 is the max-y side: intervals descend and abut exactly
        prev_lo = lo
        owners += sharding.owner_mask(y, lo, hi, r, world)
    assert (owners == 1).all()
    assert amb.lib().amb_dsm_halo_reach(gm.geometry, 1) > 2.59
