import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `pytest -m gpu` on the B200 box)")


EMULATED = os.environ.get("AMB_TEST_EMU", "0") not in ("", "0")


def _swap_in_the_emulated_library():
    """AMB_TEST_EMU=1 (set only by tests/test_emulated_kernels.py for a child pytest process): the tests of this process
    talk to tests/emu/_build/libamb_emu.so — the product's CUDA SOURCES compiled as plain C++ on a CPU stand-in for
    the CUDA runtime (tests/emu/README.md) — so that kernel logic, incl. kernels that have not yet run on a GPU, is
    exercised on a box without one.  The product package itself never does this."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from aerial_mapper_b200 import _lib
    L = ctypes.CDLL(build_emu.build())
    for name, (restype, argtypes) in _lib.SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib._lib = L


if EMULATED:
    _swap_in_the_emulated_library()


def _gpu_count():
    try:
        import aerial_mapper_b200 as amb
        n = amb.lib().amb_device_count()
        return max(n, 0)
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu_count():
    return _gpu_count()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must FAIL (not skip) when the CUDA library is missing on a GPU box; without any device they
    # are skipped so that an accidental plain `pytest` on the CPU container stays green.
    n = None
    for item in items:
        if "gpu" in item.keywords:
            if n is None:
                n = _gpu_count()
            if n == 0 and not os.environ.get("AMB_REQUIRE_GPU"):
                item.add_marker(pytest.mark.skip(reason="no CUDA device visible"))
