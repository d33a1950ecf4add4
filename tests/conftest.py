import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _device_count() -> int:
    try:
        from powerserve_b200 import capi
        return int(capi.load_library().ps_cuda_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests SKIP (instead of erroring) on a box without a CUDA device; with a device but without the built
    library they still fail loudly — there is no CPU fallback to fall back to."""
    if not any("gpu" in it.keywords for it in items):
        return
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "powerserve_b200", "libps_cuda.so")
    if os.path.exists(lib) and _device_count() == 0:
        skip = pytest.mark.skip(reason="no CUDA device (ps_cuda_device_count() == 0)")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)
