"""What the reference's batch demo (main-ortho-backward-grid.cc:119-141) gets through the C++ drop-in headers at the
benchmark size: tests/cpp/shim_demo.cc built against aerial_mapper_b200/shim + the CUDA library, the joint_10k scenario
(50 M points, 250 frames of 4000x3000 — the data of bench.py) read from a file into ordinary pageable memory
(std::vector<Eigen::Vector3d>, cv::Mat, Eigen matrices), the (DSM, orthomosaic) sequence repeated and timed by the demo
itself.  Prints one JSON line per configuration.  Developer tool; the numbers go to DESIGN.md / profiles/."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from aerial_mapper_b200 import synth

workload = sys.argv[1] if len(sys.argv) > 1 else "joint_10k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.WORKLOADS[workload]
rows, cols, res = wl["rows"], wl["cols"], wl["res"]
dev = torch.device("cuda:0")
tmp = tempfile.mkdtemp(prefix="amb_shim_bench_")
libdir = os.path.join(ROOT, "aerial_mapper_b200")
exe = os.path.join(tmp, "shim_demo")
subprocess.check_call(["g++", "-std=c++11", "-O2", "-DAMB_SHIM_MINI", "-I" + os.path.join(libdir, "shim"),
                       os.path.join(ROOT, "tests", "cpp", "shim_demo.cc"), "-o", exe, "-L" + libdir,
                       "-laerial_mapper_b200", "-Wl,-rpath," + libdir])

camd = dict(synth.C3_CAMERA) if wl["cam_scale"] == 1.0 else synth.scaled_camera(wl["cam_scale"])
poses = synth.lawnmower_poses(wl["lines"], wl["per_line"], rows * res / 2, cols * res / 2, wl["agl"], seed=4)
W, H = camd["width"], camd["height"]
scen = os.path.join(tmp, "scenario.bin")
with open(scen, "wb") as f:
    np.array([rows * res, cols * res, res, 0.0, 0.0], np.float64).tofile(f)
    np.array([wl["n_points"], len(poses), W, H, 1, camd["dist_type"]], np.int64).tofile(f)
    np.array([camd["fu"], camd["fv"], camd["cu"], camd["cv"]] + list(camd["dist"]), np.float64).tofile(f)
    xyz = bench.device_point_cloud(torch, wl["n_points"], rows * res / 2, cols * res / 2, dev)
    xyz.cpu().numpy().tofile(f)
    del xyz
    np.asarray(poses, np.float64).tofile(f)
    imgs = synth.procedural_images_torch(len(poses), W, H, 1, dev)
    for k in range(len(poses)):
        imgs[k].cpu().numpy().tofile(f)
    del imgs
torch.cuda.empty_cache()

for label, env in (("host layers authoritative (default)", {}),
                   ("AMB_SHIM_RESIDENT_LAYERS=1", {"AMB_SHIM_RESIDENT_LAYERS": "1"}),
                   ("driver staging (AMB_STAGING_OFF=1), default layers", {"AMB_STAGING_OFF": "1"})):
    r = subprocess.run([exe, scen, os.path.join(tmp, "layers.bin"), str(reps)], capture_output=True, text=True,
                       env=dict(os.environ, AMB_SHIM_TRACE="1", **env))
    print("\n".join(r.stderr.splitlines()[-9:]), flush=True)   # the last repetition's step trace
    lines = [l for l in r.stdout.splitlines() if l.startswith("shim_demo rep")]
    if r.returncode != 0 or not lines:
        print(json.dumps({"config": label, "error": (r.stdout + r.stderr)[-500:]}))
        continue
    vals = [[float(x.split(" ms")[0].split()[-1]) for x in l.split(",")] for l in lines]
    best = min(vals[1:] or vals, key=lambda v: v[2])
    print(json.dumps({"workload": workload, "api": "C++ drop-in classes (dsm::Dsm, ortho::OrthoBackwardGrid), pageable "
                      "caller memory", "config": label, "reps": [dict(zip(("dsm_ms", "ortho_ms", "total_ms"), v)) for v in vals],
                      "best_after_first": dict(zip(("dsm_ms", "ortho_ms", "total_ms"), best)),
                      "cells_per_s": rows * cols / (best[2] * 1e-3)}), flush=True)
os.remove(scen)
