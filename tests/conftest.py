import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_addoption(parser):
    parser.addoption("--emulated-library", action="store_true", default=False,
                     help="run the `gpu` tests on the CPU against tests/emul/_build/libdvbt_b200_emul.so (the library's own "
                          "sources compiled for the host on a stand-in CUDA runtime; test infrastructure, see tests/emul/)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if config.getoption("--emulated-library"):
        import ctypes
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
        import build_vit_emul
        import gr_dvbt_b200.capi as capi
        asan = os.environ.get("DVBT_EMUL_ASAN") == "1"   # AddressSanitizer build: needs LD_PRELOAD of libasan, see build_all_asan()
        tsan = os.environ.get("DVBT_EMUL_TSAN") == "1"   # ThreadSanitizer build (racecheck): LD_PRELOAD of libtsan
        path = build_vit_emul.build_all_asan() if asan else build_vit_emul.build_all_tsan() if tsan else build_vit_emul.build_all()
        capi._lib = capi.declare(ctypes.CDLL(path))
        global HAVE_GPU
        HAVE_GPU = True


def _have_gpu():
    try:
        import gr_dvbt_b200.capi as capi
        return capi.lib().dvbt_b200_device_count() > 0
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
