"""Kernel-source logic on a box without a GPU: the product's CUDA sources compiled as plain C++ on a CPU stand-in for the
CUDA runtime (tests/emu/README.md — test infrastructure only, never a fallback of the product) and driven through the
normal Python mirror + C ABI by the GPU tests themselves, in a child pytest process with AMB_TEST_EMU=1.

  * every `gpu_pending` test (kernels written after the round's GPU budget was spent: the orthomosaic's dominance cull,
    OrthoFromPcl's adaptive interpolation, the stereo rectification maps), and
  * a sample of already validated `gpu` tests, which shows the emulation reproduces what the B200 produced.
This does not replace a GPU run (fibers run one after the other: no races, no memory model, no performance)."""
import os
import subprocess
import sys

import pytest

from common import ROOT


def run_child(marker, files, extra=()):
    env = dict(os.environ, AMB_TEST_EMU="1")
    env.pop("AMB_ORTHO_DOMINANCE", None)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", marker, "-p", "no:cacheprovider"] + list(extra) + \
          [os.path.join(ROOT, "tests", f) for f in files]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


def test_pending_gpu_tests_pass_on_the_emulated_kernels():
    tail = run_child("gpu_pending", ["test_gpu_ortho_dominance.py", "test_ortho_from_pcl.py", "test_stereo_rectify.py"])
    assert " passed" in tail and "failed" not in tail


def test_validated_gpu_tests_pass_on_the_emulated_kernels_too():
    tail = run_child("gpu", ["test_gpu_smoke.py", "test_gpu_refsrc.py", "test_stereo_reproject.py"])
    assert " passed" in tail and "failed" not in tail


@pytest.mark.parametrize("tool,seed,cases", [("emu_fuzz_dsm.py", 11, 6), ("emu_fuzz_ortho.py", 12, 8)])
def test_random_cases_on_the_emulated_kernels(tool, seed, cases):
    # tools/emu_fuzz_*.py: random geometry / density / cameras (from 1x1 maps up) against the oracle
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(seed), str(cases)], cwd=ROOT,
                       capture_output=True, text=True)
    assert r.returncode == 0 and "fails 0" in r.stdout, (r.stdout + r.stderr)[-2000:]


def test_the_product_never_reaches_for_the_emulated_library():
    pkg = os.path.join(ROOT, "aerial_mapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".inc", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "libamb_emu" not in text and "tests/emu/" .replace(" ", "") not in text.replace("tests/emu (", "").replace("tests/emu only", ""), f
    assert "AMB_TEST_EMU" not in open(os.path.join(ROOT, "bench.py")).read()
    assert "emu" not in open(os.path.join(ROOT, "__graft_entry__.py")).read()
