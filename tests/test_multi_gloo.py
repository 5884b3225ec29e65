"""The N>1 path on CPU: world_size-2 (and 3) gloo groups run the stripe sharding + single all-gather exactly as
bench.py does on GPUs, with the oracle standing in for the per-rank compute; the assembled map must be bit-identical
to the single-process result."""
import os
import sys

import numpy as np
import pytest

from common import ROOT, fresh_layers


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from aerial_mapper_b200 import sharding, synth
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, cols, res = 48, 70, 0.5   # 70 columns do not divide by 3: ragged last stripe
    xyz = synth.point_cloud(6000, rows * res / 2, cols * res / 2, seed=61, holes=2, hole_sides=(2.0, 6.0))
    camd = synth.scaled_camera(0.04)
    poses = synth.lawnmower_poses(2, 2, rows * res / 2, cols * res / 2, 40.0, seed=62, jitter_pos=0.5)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(len(poses))]
    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)
    c0, c1 = sharding.stripe_range(cols, rank, world)
    L = fresh_layers(rows, cols)
    if c1 > c0:
        kr = (rows * c0, rows * c1)
        assert po.dsm_process(g, L["elevation"], xyz, num_threads=-1, cell_range=kr)[0] == 0
        assert po.ortho_process(g, L, cam, poses, imgs, num_threads=-1, cell_range=kr)[0] == 0
    names = ("ortho", "elevation", "elevation_angle", "observation_index")
    width = sharding.stripe_width(cols, world)
    slabs = [torch.from_numpy(np.ascontiguousarray(L[n][:, c0:c1].T).ravel()) for n in names]
    packed = torch.empty((len(names), rows * width), dtype=torch.float32)
    sharding.pack_slabs(torch, slabs, rows, width, packed)
    gathered = sharding.all_gather_stripes(torch, dist, packed, world)
    full = sharding.unpack_full(gathered, rows, cols, world, len(names))
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), **{n: f for n, f in zip(names, full)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stripes_plus_one_allgather_equal_the_single_process_map(tmp_path, world):
    import torch.multiprocessing as mp
    from aerial_mapper_b200 import sharding, synth
    from oracle import pyoracle as po
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows, cols, res = 48, 70, 0.5
    xyz = synth.point_cloud(6000, rows * res / 2, cols * res / 2, seed=61, holes=2, hole_sides=(2.0, 6.0))
    camd = synth.scaled_camera(0.04)
    poses = synth.lawnmower_poses(2, 2, rows * res / 2, cols * res / 2, 40.0, seed=62, jitter_pos=0.5)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(len(poses))]
    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)
    L = fresh_layers(rows, cols)
    assert po.dsm_process(g, L["elevation"], xyz, num_threads=-1)[0] == 0
    assert po.ortho_process(g, L, cam, poses, imgs, num_threads=-1)[0] == 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for n in ("ortho", "elevation", "elevation_angle", "observation_index"):
            assert np.array_equal(z[n].view(np.uint32), L[n].view(np.uint32)), (r, n)
    # stripes tile the columns exactly once
    covered = np.zeros(cols, int)
    for r in range(world):
        c0, c1 = sharding.stripe_range(cols, r, world)
        covered[c0:c1] += 1
    assert (covered == 1).all()


def test_stripe_ranges_edge_cases():
    from aerial_mapper_b200 import sharding
    assert sharding.stripe_range(10, 0, 1) == (0, 10)
    assert [sharding.stripe_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [sharding.stripe_range(3, r, 8) for r in range(8)][:4] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert sharding.stripe_range(10000, 7, 8) == (8750, 10000)
