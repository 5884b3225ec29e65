"""Developer check run on the GPU box: parity statistics against the oracle + stage timings.  Not a test."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import aerial_mapper_b200 as amb
from aerial_mapper_b200 import synth
from oracle import pyoracle as po


def dsm_case(rows, cols, res, n, seed, holes=0, radius=1, big=False):
    half_x, half_y = rows * res / 2, cols * res / 2
    xyz = synth.point_cloud(n, half_x, half_y, seed, holes=holes)
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    d = amb.Dsm(amb.DsmSettings(interpolation_radius=radius), gm)
    d.debug = not big
    t = time.time(); d.process(xyz, gm); t_gpu = time.time() - t
    tim = gm.timings()
    out = dict(case="dsm %dx%d@%g n=%d holes=%d r=%d" % (rows, cols, res, len(xyz), holes, radius), wall_s=t_gpu,
               **{k: v for k, v in tim.items() if k.startswith("dsm")})
    if not big:
        g = po.make_geometry(rows, cols, res)
        e = np.full((rows, cols), np.nan, np.float32, order="F")
        st, cnt, lvl, sec = po.dsm_process(g, e, xyz, radius=radius, debug=True)
        gc, gl = d.last_debug
        ge = gm["elevation"]
        out["oracle_s"] = sec.tolist()
        out["count_mismatch"] = int((gc != cnt).sum())
        out["level_mismatch"] = int((gl != lvl).sum())
        out["nan_mismatch"] = int((np.isnan(ge) != np.isnan(e)).sum())
        ok = ~np.isnan(e) & ~np.isnan(ge)
        ulp = np.abs(ge.view(np.int32)[ok].astype(np.int64) - e.view(np.int32)[ok].astype(np.int64))
        out["ulp_max"] = int(ulp.max()) if ulp.size else 0
        out["ulp_nonzero"] = int((ulp != 0).sum())
        out["levels"] = np.bincount(lvl[lvl >= 0].ravel().astype(np.int64)).tolist()
        out["nan_cells"] = int(np.isnan(e).sum())
    print(json.dumps(out), flush=True)


def ortho_case(rows, cols, res, lines, per_line, agl, scale, colored, dist_type=1, brute=False, seed=4):
    half_x, half_y = rows * res / 2, cols * res / 2
    camd = synth.scaled_camera(scale, dist_type=dist_type, dist=(-0.05, 0.01, 1e-4, 1e-4) if dist_type == 1 else
                               ((0.01, -0.002, 0.0005, -0.0001) if dist_type == 2 else (0, 0, 0, 0)))
    poses = synth.lawnmower_poses(lines, per_line, half_x, half_y, agl, seed)
    ch = 3 if colored else 1
    imgs = [synth.procedural_image(k, camd["width"], camd["height"], ch) for k in range(len(poses))]
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    gm.layers["elevation"][...] = synth.analytic_elevation(rows, cols, res)
    o = amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(colored_ortho=colored), gm)
    o.brute_force = brute
    t = time.time(); o.process(poses, imgs, gm); wall = time.time() - t
    tim = gm.timings()
    g = po.make_geometry(rows, cols, res)
    L = {"elevation": synth.analytic_elevation(rows, cols, res),
         "elevation_angle": np.zeros((rows, cols), np.float32, order="F"),
         "observation_index": np.full((rows, cols), np.nan, np.float32, order="F"),
         "ortho": np.full((rows, cols), 255.0, np.float32, order="F"),
         "colored_ortho": np.full((rows, cols), np.nan, np.float32, order="F")}
    st, sec = po.ortho_process(g, L, po.make_camera(**camd), poses, imgs, colored=colored)
    out = dict(case="ortho %dx%d@%g frames=%d scale=%g colored=%d dist=%d brute=%d" % (
        rows, cols, res, len(poses), scale, colored, dist_type, brute), wall_s=wall, oracle_s=sec,
        **{k: v for k, v in tim.items() if k.startswith("ortho")})
    oi, goi = L["observation_index"], gm["observation_index"]
    out["obs_mismatch"] = int((~((oi == goi) | (np.isnan(oi) & np.isnan(goi)))).sum())
    out["covered"] = int((~np.isnan(oi)).sum())
    lay = "colored_ortho" if colored else "ortho"
    out["pixel_mismatch"] = int((L[lay].view(np.uint32) != gm[lay].view(np.uint32)).sum())
    ea, gea = L["elevation_angle"], gm["elevation_angle"]
    out["angle_ulp_max"] = int(np.abs(ea.view(np.int32).astype(np.int64) - gea.view(np.int32).astype(np.int64)).max())
    print(json.dumps(out), flush=True)


def ortho_big(rows, cols, res, lines, per_line, agl, colored=False, brute=False, reps=3):
    import torch
    half_x, half_y = rows * res / 2, cols * res / 2
    camd = dict(synth.C3_CAMERA)
    poses = synth.lawnmower_poses(lines, per_line, half_x, half_y, agl, 4)
    n = len(poses)
    ch = 3 if colored else 1
    dev = torch.device("cuda:0")
    imgs = synth.procedural_images_torch(n, camd["width"], camd["height"], ch, dev)
    torch.cuda.synchronize()
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    gm.layers["elevation"][...] = synth.analytic_elevation(rows, cols, res)
    gm.to_device(0)
    o = amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(colored_ortho=colored), gm)
    o.brute_force = brute
    ptrs = [imgs[k].data_ptr() for k in range(n)]
    for r in range(reps):
        amb.lib().amb_init_layers(gm.context())
        gm.upload(("elevation",))
        o.process_device(poses, ptrs, camd["width"] * ch, gm)
        gm.sync()
        tim = gm.timings()
        print(json.dumps(dict(case="ortho_big %dx%d@%g frames=%d colored=%d brute=%d rep=%d" % (
            rows, cols, res, n, colored, brute, r), **{k: v for k, v in tim.items() if k.startswith("ortho")})),
            flush=True)
    gm.download(("observation_index",))
    oi = gm["observation_index"]
    print(json.dumps(dict(covered=int((~np.isnan(oi)).sum()), cells=rows * cols)), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which in ("small", "all"):
        dsm_case(256, 256, 1.0, 100000, 1)
        dsm_case(300, 200, 0.25, 40000, 5, holes=6)
        dsm_case(100, 130, 0.5, 20000, 6, radius=2)
        dsm_case(1000, 1000, 0.25, 500000, 7, holes=10)
        ortho_case(200, 160, 0.5, 3, 4, 60.0, 0.1, False)
        ortho_case(200, 160, 0.5, 3, 4, 60.0, 0.1, True)
        ortho_case(200, 160, 0.5, 3, 4, 60.0, 0.1, False, dist_type=0)
        ortho_case(200, 160, 0.5, 3, 4, 60.0, 0.1, False, dist_type=2)
        ortho_case(200, 160, 0.5, 3, 4, 60.0, 0.1, False, brute=True)
        ortho_case(640, 480, 0.5, 4, 6, 60.0, 0.1, False)
    if which in ("big", "all"):
        dsm_case(4000, 4000, 0.25, 8000000, 2, big=True)
        dsm_case(10000, 10000, 0.25, 50000000, 2, big=True)
        dsm_case(10000, 10000, 0.25, 50000000, 2, big=True)
    if which in ("obig", "all"):
        ortho_big(8000, 8000, 0.5, 10, 25, 600.0)
        ortho_big(8000, 8000, 0.5, 10, 25, 600.0, colored=True, reps=2)
        ortho_big(2000, 2000, 0.5, 10, 25, 600.0, brute=True, reps=1)
