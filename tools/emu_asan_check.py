"""Memory-safety pass over the kernel source: the emulated library (tests/emu) rebuilt with AddressSanitizer — global
("device") buffers are heap blocks, static and dynamic shared memory are globals / heap with red zones — and driven through
random DSM, orthomosaic (plain + dominance), adaptive OrthoFromPcl and rectification workloads.

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
        python tools/emu_asan_check.py
"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import aerial_mapper_b200 as amb
from aerial_mapper_b200 import _lib, synth
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import build_emu
build_emu.build()
import subprocess
ASAN_LIB = "/tmp/libamb_emu_asan.so"
B = build_emu.OUT
subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-pthread",
                       "-fsanitize=address", "-fno-omit-frame-pointer", "-I" + os.path.join(build_emu.HERE, "include"), "-I" + B,
                       "-I" + build_emu.CSRC, "-I" + os.path.join(ROOT, "include")] +
                      sorted(os.path.join(B, f) for f in os.listdir(B) if f.endswith("_emu.cc")) +
                      sorted(os.path.join(build_emu.CSRC, f) for f in os.listdir(build_emu.CSRC) if f.endswith(".cc")) +
                      [os.path.join(build_emu.HERE, "emu_runtime.cc"), "-ldl", "-o", ASAN_LIB])
L = C.CDLL(ASAN_LIB)
for name, (restype, argtypes) in _lib.SYMBOLS.items():
    fn = getattr(L, name); fn.restype = restype; fn.argtypes = argtypes
_lib._lib = L
rng = np.random.default_rng(3)
for case in range(6):
    rows, cols = int(rng.integers(1, 120)), int(rng.integers(1, 120))
    res = float(rng.choice([0.25, 0.5, 1.0])); radius = int(rng.choice([1, 2, 3]))
    n = max(1, int(rng.choice([0.3, 2.0, 30.0]) * rows * res * cols * res))
    xyz = np.c_[rng.uniform(-rows*res/2-3, rows*res/2+3, n), rng.uniform(-cols*res/2-3, cols*res/2+3, n), rng.normal(100, 5, n)]
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, rows * res, cols * res, res)).getMutable()
    d = amb.Dsm(amb.DsmSettings(interpolation_radius=radius), gm); d.debug = True
    d.process(xyz, gm)
    camd = synth.scaled_camera(0.05)
    poses = synth.lawnmower_poses(2, 3, rows * res / 2, cols * res / 2, 50.0, int(rng.integers(0, 99)), jitter_pos=0.5)
    imgs = [synth.procedural_image(k, camd["width"], camd["height"]) for k in range(len(poses))]
    for dom in ("0", "1"):
        os.environ["AMB_ORTHO_DOMINANCE"] = dom
        o = amb.OrthoBackwardGrid(amb.NCamera(**camd), amb.OrthoSettings(), gm)
        o.process(poses, imgs, gm)
    inten = rng.integers(0, 256, n).astype(np.int32)
    inside = (np.abs(xyz[:, 0]) < rows*res/2) & (np.abs(xyz[:, 1]) < cols*res/2)
    if inside.sum() > 0:
        amb.OrthoFromPcl(amb.OrthoFromPclSettings(interpolation_radius=radius, use_adaptive_interpolation=True)).process(xyz[inside], inten[inside], gm)
    print("case", case, rows, cols, res, radius, n, "ok")
T1 = np.eye(3, dtype=np.float32); T1[0, 2] = 0.3
print("rectify", amb.rectify_stereo_maps(T1, np.eye(3, dtype=np.float32), 333, 77)[0].shape)
