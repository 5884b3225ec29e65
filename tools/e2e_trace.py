"""Developer tool: wall-clock trace of the host-buffer (e2e) step of bench.py, call by call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import aerial_mapper_b200 as amb
from aerial_mapper_b200 import synth
import bench

wl = bench.WORKLOADS["joint_10k"]
rows, cols, res = wl["rows"], wl["cols"], wl["res"]
dev = torch.device("cuda:0")
xyz_d = bench.device_point_cloud(torch, wl["n_points"], rows * res / 2, cols * res / 2, dev)
camd = dict(synth.C3_CAMERA)
poses = synth.lawnmower_poses(wl["lines"], wl["per_line"], rows * res / 2, cols * res / 2, wl["agl"], seed=4)
n_frames = len(poses); W, H = camd["width"], camd["height"]
imgs_d = synth.procedural_images_torch(n_frames, W, H, 1, dev)
xyz_h = torch.empty(xyz_d.shape, dtype=torch.float64, pin_memory=True); xyz_h.copy_(xyz_d)
imgs_h = torch.empty((n_frames, H, W), dtype=torch.uint8, pin_memory=True); imgs_h.copy_(imgs_d)
torch.cuda.synchronize(); del xyz_d, imgs_d
xyz_np = xyz_h.numpy(); img_np = [imgs_h[k].numpy() for k in range(n_frames)]
names = ("ortho", "elevation", "elevation_angle", "observation_index")
gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res), pinned=True, layer_names=names).getMutable()
gm.to_device(0, names=names)
ctx = gm.context()
dsm = amb.Dsm(amb.DsmSettings(), gm)
ortho = amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gm)

def step(use_async):
    marks = [("start", time.perf_counter())]
    amb.check(amb.lib().amb_init_layers(ctx), ctx); marks.append(("init", time.perf_counter()))
    dsm.process(xyz_np, gm); marks.append(("dsm.process", time.perf_counter()))
    if use_async:
        gm.download_async(("elevation",)); marks.append(("dl_async(elev)", time.perf_counter()))
    ortho.process(poses, img_np, gm); marks.append(("ortho.process", time.perf_counter()))
    if use_async:
        gm.download(("ortho", "elevation_angle", "observation_index")); marks.append(("dl x3", time.perf_counter()))
        gm.sync(); marks.append(("sync", time.perf_counter()))
    else:
        gm.download(names); marks.append(("dl x4", time.perf_counter()))
    return marks

def step_mirror():
    marks = [("start", time.perf_counter())]
    amb.check(amb.lib().amb_init_layers(ctx), ctx); marks.append(("init", time.perf_counter()))
    dsm.process(xyz_np, gm); marks.append(("dsm.process", time.perf_counter()))
    ortho.process(poses, img_np, gm); marks.append(("ortho.process", time.perf_counter()))
    gm.sync(); marks.append(("sync", time.perf_counter()))
    return marks

gm.set_mirrors(names)
for _ in range(3):
    m = step_mirror()
    print("mirror total %.1f ms: " % ((m[-1][1] - m[0][1]) * 1e3) +
          ", ".join("%s %.1f" % (m[k][0], (m[k][1] - m[k - 1][1]) * 1e3) for k in range(1, len(m))), flush=True)
gm.set_mirrors(names, enable=False)
for use_async in (False, True):
    step(use_async)
    m = step(use_async)
    print("async=%d total %.1f ms: " % (use_async, (m[-1][1] - m[0][1]) * 1e3) +
          ", ".join("%s %.1f" % (m[k][0], (m[k][1] - m[k - 1][1]) * 1e3) for k in range(1, len(m))), flush=True)
    print("   timings", {k: round(v, 2) if isinstance(v, float) else v for k, v in gm.timings().items()}, flush=True)
