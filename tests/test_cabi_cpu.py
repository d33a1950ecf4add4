"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/ps_cuda.h declares, fails LOUDLY without a GPU (no CPU fallback), and its host-side restatement of glibc's
expf / ggml_v_expf (shared with the device code) matches the platform libm and the oracle bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from powerserve_b200 import build, capi
from tests import _libs as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return capi.load_library()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ps_cuda.h")).read()
    declared = set(re.findall(r"\b(ps_cuda_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(lib, name), f"libps_cuda.so does not export {name}"
    assert set(capi.exported_symbols()) == declared
    assert lib.ps_cuda_abi_version() == 1


def test_spec_library_exports_every_declared_symbol(lib):
    build.build_spec()
    hdr = open(os.path.join(ROOT, "include", "ps_spec.h")).read()
    declared = set(re.findall(r"\b(ps_spec_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 6
    sl = capi.load_spec_library()
    for name in sorted(declared):
        assert hasattr(sl, name), f"libps_spec.so does not export {name}"


def test_no_cpu_fallback_without_device(lib):
    if lib.ps_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    desc = capi.ModelDesc(512, 1536, 2, 8, 2, 64, 1024, 512, 1e-5, 64, 0, 5e5, 1.0, 1.0, 0, 8, 0, 1)
    h = C.c_void_p()
    rc = lib.ps_cuda_create(C.byref(h), 0, C.byref(desc))
    assert rc == 1  # PS_CUDA_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.ps_cuda_last_error(None)
    with pytest.raises(capi.PsCudaError):
        capi.CudaBackend(desc)


def test_expf_restatement_matches_libm(lib):
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    rng = np.random.default_rng(0)
    xs = np.concatenate([
        rng.uniform(-110, 90, 200000), rng.standard_normal(100000) * 5, rng.standard_normal(50000) * 1e-3,
        [0.0, -0.0, 88.0, 88.72, 88.73, -103.9, -103.98, -87.3, -87.4, 1e-30, -1e-30, np.inf, -np.inf],
    ]).astype(np.float32)
    bad = 0
    for x in xs:
        a, b = np.float32(libm.expf(float(x))), np.float32(lib.ps_cuda_host_expf_ref(float(x)))
        bad += a.view(np.uint32) != b.view(np.uint32)
    assert bad == 0


def test_v_expf_matches_oracle(lib):
    o = L.oracle()
    xs = np.concatenate([np.linspace(-200, 100, 60001), [-np.inf, 0.0, -0.0, -87.33, -88.0, -126 * 0.693147, 88.3]]).astype(np.float32)
    for x in xs:
        a, b = np.float32(o.ps_or_v_expf(float(x))), np.float32(lib.ps_cuda_host_v_expf(float(x)))
        assert a.view(np.uint32) == b.view(np.uint32), x


def test_product_never_imports_the_oracle():
    """The product tree must not reference oracle/ (a product path routed through the oracle voids every parity claim)."""
    for base, _, files in os.walk(os.path.join(ROOT, "powerserve_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(base, f), errors="ignore").read()
                assert "ps_oracle" not in src and "libps_oracle" not in src and "oracle/" not in src.replace("SURVEY", ""), f
    assert "oracle" not in open(os.path.join(ROOT, "include", "ps_cuda.h")).read()
