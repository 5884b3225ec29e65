"""The C-ABI shared library loads on a box WITHOUT a GPU and exports every symbol include/*.h declares; the
GPU-free entry points work; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from common import ROOT
import aerial_mapper_b200 as amb
from aerial_mapper_b200 import _lib

HEADER = os.path.join(ROOT, "include", "aerial_mapper_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 25
    L = amb.lib()
    for n in names:
        assert hasattr(L, n), "libaerial_mapper_b200.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", amb.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (amb_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_struct_layouts_match_the_c_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(amb_geometry),sizeof(amb_camera),sizeof(amb_timings),offsetof(amb_camera,dist),'
                   'offsetof(amb_timings,dsm_points_binned));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(_lib.Geometry), C.sizeof(_lib.Camera), C.sizeof(_lib.Timings),
                   _lib.Camera.dist.offset, _lib.Timings.dsm_points_binned.offset]


def test_header_compiles_as_c_and_cpp(tmp_path):
    for comp, ext in (("gcc", "c"), ("g++", "cc")):
        src = tmp_path / ("t." + ext)
        src.write_text('#include "%s"\nint main(void){amb_geometry g; g.rows = AMB_NUM_LAYERS; return g.rows == 0;}\n' % HEADER)
        subprocess.check_call([comp, "-Wall", "-Werror", "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_gpu_free_entry_points():
    L = amb.lib()
    assert L.amb_abi_version() == 2
    assert b"coincide" in L.amb_status_string(-3)
    g = amb.Geometry()
    # setGeometry: size = round(length / resolution), length = size * resolution
    assert L.amb_geometry_init(100.3, 50.0, 0.5, 10.0, -4.0, C.byref(g)) == 0
    assert (g.rows, g.cols) == (201, 100) and g.length_x == 100.5 and g.length_y == 50.0
    x, y = C.c_double(), C.c_double()
    assert L.amb_geometry_position(C.byref(g), 0, 0, C.byref(x), C.byref(y)) == 0
    assert x.value == (10.0 + (0.5 * 100.5 - 0.25)) and y.value == (-4.0 + (25.0 - 0.25))  # max-x / max-y corner
    assert L.amb_geometry_position(C.byref(g), 201, 0, C.byref(x), C.byref(y)) == _lib.AMB_ERR_SIZE_MISMATCH
    assert L.amb_geometry_init(1.0, 1.0, 0.0, 0, 0, C.byref(g)) == _lib.AMB_ERR_INVALID_ARGUMENT


def test_geometry_matches_the_oracle_restatement():
    from oracle import pyoracle as po
    from aerial_mapper_b200 import synth
    gm = amb.GridMap()
    gm.setGeometry((37.3, 21.9), 0.3, (1234.5, -987.25))
    rows, cols = gm.getSize()
    qx, qy = synth.grid_positions(rows, cols, 0.3, 1234.5, -987.25)
    for i, j in [(0, 0), (rows - 1, cols - 1), (5, 7), (rows // 2, 1)]:
        assert gm.getPosition((i, j)) == (qx[i], qy[j])


def test_compute_entry_points_fail_loudly_without_a_gpu(gpu_count):
    if gpu_count > 0:
        pytest.skip("a GPU is visible")
    g = amb.Geometry()
    amb.lib().amb_geometry_init(8.0, 8.0, 1.0, 0, 0, C.byref(g))
    ctx = C.c_void_p()
    assert amb.lib().amb_create(C.byref(g), 0, 0, 8, C.byref(ctx)) == _lib.AMB_ERR_NO_DEVICE
    gm = amb.AerialGridMap(amb.GridMapSettings(0, 0, 8, 8, 1.0)).getMutable()
    with pytest.raises(amb.AmbError) as ei:
        amb.Dsm(amb.DsmSettings(), gm).process(np.zeros((4, 3)), gm)
    assert ei.value.status == _lib.AMB_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "aerial_mapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cc", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                code = "\n".join(ln for ln in text.splitlines() if "oracle" in ln and not ln.lstrip().startswith(("//", "#", "*", '"')))
                assert "oracle" not in code.replace("never links, loads or calls anything under oracle/", ""), \
                    "%s references oracle/" % os.path.join(dirpath, f)
