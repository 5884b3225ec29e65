"""The N>1 path on CPU: world_size-2 (and 3) gloo groups run the stripe sharding + single all-gather exactly as
bench.py does on GPUs, with the oracle standing in for the per-rank compute; the assembled map must be bit-identical
to the single-process result.  A second test runs the sharded-CLOUD path of bench.py (global point ids, border-halo
compaction kernel, one all-gather of the halos, DSM + orthomosaic per stripe) through the product's own
sharding.HaloExchange and C ABI on the emulated kernels (tests/emu), two gloo ranks."""
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


# ---- the same N>1 path with the REAL per-rank code on the emulated kernels (tests/emu) ---------------------------------
def _emu_worker(rank, world, port, tmp):
    """What bench.py does per rank at N > 1 — cloud sharded by stripe with global ids, border halo compacted by
    amb_dsm_extract_halo, ONE all-gather (gloo here, NCCL there), DSM on [all halos | own points], orthomosaic on the
    rank's own stripe — through the product's sharding.HaloExchange and C ABI, with the CUDA sources compiled as plain C++
    (AMB_TEST_EMU=1: host memory is device memory, CPU tensors are device tensors)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["AMB_TEST_EMU"] = "1"
    import ctypes as C
    import torch
    import torch.distributed as dist
    import conftest  # noqa: F401  (swaps in tests/emu/_build/libamb_emu.so)
    import aerial_mapper_b200 as amb
    from aerial_mapper_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, cols, res = 96, 140, 0.5
    xyz_np = synth.point_cloud(40000, rows * res / 2 + 3.0, cols * res / 2 + 3.0, seed=81, holes=3, hole_sides=(3.0, 9.0))
    camd = synth.scaled_camera(0.05)
    poses = synth.lawnmower_poses(2, 3, rows * res / 2, cols * res / 2, 50.0, seed=82, jitter_pos=0.5)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(len(poses))]
    c0, c1 = sharding.stripe_range(cols, rank, world)
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    gm.to_device(0, col_range=(c0, c1))
    xyz = torch.from_numpy(xyz_np)
    ids = torch.arange(xyz.shape[0], dtype=torch.int64)
    y_lo, y_hi = sharding.stripe_y_interval(gm.geometry, c0, c1)
    m = sharding.owner_mask(xyz[:, 1], y_lo, y_hi, rank, world)         # this rank's share of the cloud
    hx = sharding.HaloExchange(torch, world, rank, 40000, xyz[m], ids[m], torch.device("cpu"))
    reach = amb.lib().amb_dsm_halo_reach(C.byref(gm.geometry), 1)
    amb.check(amb.lib().amb_dsm_set_density_hint(gm.context(), xyz.shape[0] / float(rows * cols)), gm.context())
    hx.extract(gm.context(), y_lo, y_hi, reach)
    hx.exchange(dist)                                                    # the one collective of the DSM stage
    assert (hx.counts() <= hx.cap).all()
    hx.assemble()
    d = amb.Dsm(amb.DsmSettings(), gm)
    d.process_device(hx.big_xyz.data_ptr(), hx.n_total, gm, d_ids=hx.big_ids.data_ptr())
    o = amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gm)
    o.process_device(poses, [im.ctypes.data for im in imgs], camd["width"], gm)   # frames resident on every rank
    gm.sync()
    gm.download()
    np.savez(os.path.join(tmp, "emu_rank%d.npz" % rank), c0=c0, c1=c1, n_local=hx.n_local,
             **{n: gm[n][:, c0:c1] for n in ("ortho", "elevation", "elevation_angle", "observation_index")})
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_cloud_halo_allgather_on_the_emulated_kernels(tmp_path):
    import torch.multiprocessing as mp
    from aerial_mapper_b200 import synth
    from oracle import pyoracle as po
    from common import ulp_diff
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_emu_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows, cols, res = 96, 140, 0.5
    xyz = synth.point_cloud(40000, rows * res / 2 + 3.0, cols * res / 2 + 3.0, seed=81, holes=3, hole_sides=(3.0, 9.0))
    camd = synth.scaled_camera(0.05)
    poses = synth.lawnmower_poses(2, 3, rows * res / 2, cols * res / 2, 50.0, seed=82, jitter_pos=0.5)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(len(poses))]
    g, cam = po.make_geometry(rows, cols, res), po.make_camera(**camd)
    L = fresh_layers(rows, cols)
    st, _, lvl, _ = po.dsm_process(g, L["elevation"], xyz, num_threads=-1, debug=True)
    assert st == 0 and (lvl > 0).any() and (lvl < 0).any()
    n_local = 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "emu_rank%d.npz" % r))
        c0, c1 = int(z["c0"]), int(z["c1"])
        n_local += int(z["n_local"])
        assert np.array_equal(np.isnan(z["elevation"]), np.isnan(L["elevation"][:, c0:c1]))
        assert ulp_diff(z["elevation"], L["elevation"][:, c0:c1]).max() <= 1
        Lr = fresh_layers(rows, cols)
        Lr["elevation"][:, c0:c1] = z["elevation"]                      # the rank's orthomosaic ran on its own heights
        assert po.ortho_process(g, Lr, cam, poses, imgs, num_threads=-1, cell_range=(rows * c0, rows * c1))[0] == 0
        for n in ("ortho", "observation_index"):
            assert np.array_equal(z[n].view(np.uint32), Lr[n][:, c0:c1].view(np.uint32)), (r, n)
        assert ulp_diff(z["elevation_angle"], Lr["elevation_angle"][:, c0:c1]).max() <= 1
    assert n_local == len(xyz)                                          # the stripes partition the cloud
