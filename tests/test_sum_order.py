"""DESIGN.md section 2, last bullet: the CUDA kernels add the double-precision sums of RMSNorm in a tree order (lane
partials -> warp butterfly -> warps) while the reference adds sequentially (ggml.c:12691-12697).  The doubles differ in
their last bits; the claim is that the ONE fp32 value derived from them (the norm scale) does not.  This test pins the
claim on CPU with a numpy model of the kernel's order (ps_rw.cuh prologue) against the sequential order."""
import numpy as np
import pytest


def _sequential(sq):
    s = 0.0
    for v in sq:
        s += v
    return s


def _kernel_order(sq):
    nb = sq.size // 256
    blk = sq.reshape(nb, 256)
    lane = np.zeros((16, 32))                                  # [warp][lane] running double partials
    for i in range(nb):                                        # block i belongs to warp i % 16, visited in increasing i
        e = np.concatenate([blk[i, :128].reshape(32, 4), blk[i, 128:].reshape(32, 4)], axis=1)   # lane l: 4l..4l+3, 128+4l..+3
        for t in range(8):
            lane[i % 16] += e[:, t]
    idx = np.arange(32)
    for o in (16, 8, 4, 2, 1):                                 # warp butterfly
        lane = lane + lane[:, idx ^ o]
    v = np.zeros(32)
    v[:16] = lane[:, 0]
    for o in (16, 8, 4, 2, 1):                                 # the 16 warp totals
        v = v + v[idx ^ o]
    return v[0]


def _scale(s, n, eps=np.float32(1e-5)):
    mean = np.float32(s / n)
    return np.float32(1.0) / np.sqrt(np.float32(mean + eps))


@pytest.mark.parametrize("kind", ["normal", "wide", "outliers"])
def test_rms_scale_is_order_independent(kind):
    rng = np.random.default_rng({"normal": 1, "wide": 2, "outliers": 3}[kind])
    differ = 0
    for _ in range(40):
        n = int(rng.choice([2048, 4096, 14336]))
        x = rng.standard_normal(n).astype(np.float32)
        if kind == "wide":
            x = (x * np.exp(rng.uniform(-6, 6, n))).astype(np.float32)
        elif kind == "outliers":
            x = (x * np.where(rng.random(n) < 0.01, 100.0, 1.0)).astype(np.float32)
        sq = (x * x).astype(np.float32).astype(np.float64)     # fp32 product, then widened: (ggml_float)(x[i] * x[i])
        a, b = _sequential(sq), _kernel_order(sq)
        differ += a != b
        assert _scale(a, n) == _scale(b, n)
    assert differ > 0                                          # the model really is a different order
