"""Pin the oracle (oracle/ps_oracle.c) against the COMPILED REFERENCE (oracle/_ref, built from /root/reference by
`make -C oracle ref`): every operator must agree BIT FOR BIT on seeded random inputs, including ragged sizes.
Skipped when oracle/_ref is absent (a checkout without the build container's artefacts); the committed golden
vectors (tests/test_oracle_golden.py) cover that case."""
import ctypes as C

import numpy as np
import pytest

from powerserve_b200 import synth
from tests import _libs as L

pytestmark = pytest.mark.skipif(not L.have_ref(), reason="oracle/_ref not built (needs /root/reference)")

QTYPES = [(L.Q4_K, 256), (L.Q6_K, 256), (L.Q4_0, 32), (L.Q8_0, 32)]
# Q5_K (real Q4_K_S / Q5_K_M files): pinned at the block level - the reference's own DataType layer cannot carry it
BLOCK_TYPES = QTYPES + [(L.Q5_K, 256)]


def rand_act(rng, n, kind):
    if kind == "normal":
        return rng.standard_normal(n).astype(np.float32)
    if kind == "wide":
        return (rng.standard_normal(n) * np.exp(rng.uniform(-6, 6, n))).astype(np.float32)
    if kind == "ties":  # values that land exactly on .5 after scaling, plus zeros blocks
        x = rng.integers(-254, 255, n).astype(np.float32) * 0.5
        x[: n // 8] = 0.0
        return x
    raise ValueError(kind)


def test_fp16_roundtrip_all_halfs():
    o, r = L.oracle(), L.ref_ggml()
    for h in range(0, 65536, 1):
        a, b = o.ps_or_fp16_to_fp32(h), r.ggml_fp16_to_fp32(h)
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32) or (a != a and b != b), h


def test_fp32_to_fp16_random():
    o, r = L.oracle(), L.ref_ggml()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rand_act(rng, 20000, "wide"), np.float32([0, -0.0, 65504, 65519.99, 65520, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5])])
    for x in xs:
        assert o.ps_or_fp32_to_fp16(float(x)) == r.ggml_fp32_to_fp16(float(x)), x


@pytest.mark.parametrize("kind", ["normal", "wide", "ties"])
def test_quantize_q8_K(kind):
    o, r = L.oracle(), L.ref_ggml()
    rng = np.random.default_rng(1)
    k = 256 * 9
    x = rand_act(rng, k, kind)
    nb = o.ps_or_row_size(L.Q8_K, k)
    a, b = np.zeros(nb, np.uint8), np.zeros(nb, np.uint8)
    o.ps_or_quantize_row(L.Q8_K, L.fptr(x), L.vptr(a), k)
    r.quantize_row_q8_K(L.fptr(x), L.vptr(b), k)
    assert (a == b).all()


@pytest.mark.parametrize("kind", ["normal", "wide", "ties"])
def test_quantize_q8_0(kind):
    o, r = L.oracle(), L.ref_ggml()
    rng = np.random.default_rng(2)
    k = 32 * 67
    x = rand_act(rng, k, kind)
    nb = o.ps_or_row_size(L.Q8_0, k)
    a, b = np.zeros(nb, np.uint8), np.zeros(nb, np.uint8)
    o.ps_or_quantize_row(L.Q8_0, L.fptr(x), L.vptr(a), k)
    r.quantize_row_q8_0(L.fptr(x), L.vptr(b), k)
    assert (a == b).all()


@pytest.mark.parametrize("t,blk", BLOCK_TYPES)
def test_dequantize(t, blk):
    o, r = L.oracle(), L.ref_ggml()
    rng = np.random.default_rng(3)
    k = blk * 8
    w = synth.random_blocks(rng, t, 5, k, 0.05).reshape(-1)
    # also fully random bytes (every bit pattern a file could hold), with the fp16 scale fields kept finite
    w2 = rng.integers(0, 256, w.size, dtype=np.uint8)
    for ww in (w, w2):
        a, b = np.zeros(5 * k, np.float32), np.zeros(5 * k, np.float32)
        o.ps_or_dequantize_row(t, L.vptr(ww), L.fptr(a), 5 * k)
        getattr(r, {L.Q4_0: "dequantize_row_q4_0", L.Q8_0: "dequantize_row_q8_0", L.Q4_K: "dequantize_row_q4_K", L.Q5_K: "dequantize_row_q5_K", L.Q6_K: "dequantize_row_q6_K"}[t])(L.vptr(ww), L.fptr(b), 5 * k)
        fin = np.isfinite(b)
        L.assert_bit_equal(np.where(fin, a, 0), np.where(fin, b, 0), f"dequant {t}")


@pytest.mark.parametrize("t,blk", BLOCK_TYPES)
@pytest.mark.parametrize("kind", ["normal", "wide"])
def test_vec_dot(t, blk, kind):
    o, r = L.oracle(), L.ref_ggml()
    rng = np.random.default_rng(4)
    fn = {L.Q4_0: "ggml_vec_dot_q4_0_q8_0", L.Q8_0: "ggml_vec_dot_q8_0_q8_0", L.Q4_K: "ggml_vec_dot_q4_K_q8_K", L.Q5_K: "ggml_vec_dot_q5_K_q8_K", L.Q6_K: "ggml_vec_dot_q6_K_q8_K"}[t]
    qt = L.Q8_K if blk == 256 else L.Q8_0
    for k in (blk, blk * 3, blk * 16, blk * 56):
        for _ in range(8):
            w = synth.random_blocks(rng, t, 1, k, k ** -0.5).reshape(-1)
            x = rand_act(rng, k, kind)
            xq = np.zeros(o.ps_or_row_size(qt, k), np.uint8)
            o.ps_or_quantize_row(qt, L.fptr(x), L.vptr(xq), k)
            out = np.zeros(1, np.float32)
            getattr(r, fn)(k, L.fptr(out), 0, L.vptr(w), 0, L.vptr(xq), 0, 1)
            got = np.float32(o.ps_or_vec_dot(t, k, L.vptr(w), L.vptr(xq)))
            L.assert_bit_equal(np.array([got]), out, f"{fn} k={k}")


@pytest.mark.parametrize("t,blk", QTYPES)
@pytest.mark.parametrize("bs", [1, 3, 17])
def test_matmul(t, blk, bs):
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(5)
    K, N = blk * 6, 37
    w = synth.random_blocks(rng, t, N, K, K ** -0.5).reshape(-1)
    x = rand_act(rng, K * bs, "normal")
    a, b = np.zeros(N * bs, np.float32), np.zeros(N * bs, np.float32)
    o.ps_or_matmul(t, L.vptr(w), K, N, L.fptr(x), bs, L.fptr(a))
    ro.L.ref_matmul(ro.h, t, L.vptr(w), K, N, L.fptr(x), bs, L.fptr(b))
    L.assert_bit_equal(a, b, "matmul")


@pytest.mark.parametrize("dim,bs", [(64, 1), (896, 3), (4096, 2), (100, 5)])
def test_rmsnorm(dim, bs):
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(6)
    x = rand_act(rng, dim * bs, "wide" if dim == 100 else "normal")
    w = (1 + 0.1 * rng.standard_normal(dim)).astype(np.float32)
    a, b = np.zeros_like(x), np.zeros_like(x)
    o.ps_or_rmsnorm(L.fptr(a), L.fptr(x), L.fptr(w), dim, bs, 1e-5)
    ro.L.ref_rmsnorm(ro.h, L.fptr(b), L.fptr(x), L.fptr(w), dim, bs, 1e-5)
    L.assert_bit_equal(a, b, "rmsnorm")


@pytest.mark.parametrize("mode,base", [(0, 5e5), (2, 1e6), (0, 1e4)])
@pytest.mark.parametrize("hs", [64, 128])
def test_rope(mode, base, hs):
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(7)
    nh, bs = 6, 9
    x = rand_act(rng, hs * nh * bs, "normal")
    pos = np.array([0, 1, 2, 3, 100, 1000, 2047, 4095, 7], np.int32)
    a, b = np.zeros_like(x), np.zeros_like(x)
    o.ps_or_rope(L.fptr(a), L.fptr(x), hs, nh, bs, L.iptr(pos), hs, mode, base, 1.0, 1.0)
    ro.L.ref_rope(ro.h, L.fptr(b), L.fptr(x), hs, nh, bs, L.iptr(pos), hs, mode, base, 1.0, 1.0)
    L.assert_bit_equal(a, b, "rope")


@pytest.mark.parametrize("n_kv,bs,nh", [(1, 1, 4), (7, 1, 4), (8, 1, 2), (33, 3, 4), (300, 5, 2), (2049, 1, 3)])
def test_softmax_ext(n_kv, bs, nh):
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(8)
    x = (rng.standard_normal(n_kv * bs * nh) * 8).astype(np.float32)
    pos = np.arange(n_kv - bs, n_kv, dtype=np.int32)
    mask = np.zeros(n_kv * bs, np.float32)
    o.ps_or_get_mask(L.fptr(mask), n_kv, bs, L.iptr(pos))
    a, b = np.zeros_like(x), np.zeros_like(x)
    o.ps_or_softmax_ext(L.fptr(a), L.fptr(x), L.fptr(mask), n_kv, bs, nh, 0.125)
    ro.L.ref_softmax_ext(ro.h, L.fptr(b), L.fptr(x), L.fptr(mask), n_kv, bs, nh, 0.125)
    L.assert_bit_equal(a, b, "softmax_ext")


def test_v_expf_range():
    """ggml_v_expf lanes incl. the overflow / underflow branches, through the reference softmax (max = 0 row)."""
    o, ro = L.oracle(), L.ref_ops()
    xs = np.concatenate([np.linspace(-110, 0, 4096), [-np.inf, -87.3, -88.5, -103.9, -104.1, -126.0 * 0.6931, -200.0]]).astype(np.float32)
    xs = np.concatenate([xs, np.zeros((-len(xs)) % 8, np.float32)])
    xs[0] = 0.0
    n = len(xs)
    mask = np.zeros(n, np.float32)
    a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
    o.ps_or_softmax_ext(L.fptr(a), L.fptr(xs), L.fptr(mask), n, 1, 1, 1.0)
    ro.L.ref_softmax_ext(ro.h, L.fptr(b), L.fptr(xs), L.fptr(mask), n, 1, 1, 1.0)
    L.assert_bit_equal(a, b, "v_expf")


def test_add_and_bias_broadcast():
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(9)
    a_, b_ = rand_act(rng, 896 * 3, "normal"), rand_act(rng, 896, "normal")
    a, b = np.zeros_like(a_), np.zeros_like(a_)
    o.ps_or_add(L.fptr(a), L.fptr(a_), L.fptr(b_), 896 * 3, 896)
    ro.L.ref_add(ro.h, L.fptr(b), L.fptr(a_), L.fptr(b_), 896, 3, 1)
    L.assert_bit_equal(a, b, "add bias")


def test_silu_hadamard():
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(10)
    g = np.concatenate([rand_act(rng, 5000, "normal") * 4, np.float32([0, -0.0, 88, -88, 100, -104, 1e-20, -1e-20, 20, -20])])
    u = rand_act(rng, g.size, "normal")
    a, b = np.zeros_like(g), np.zeros_like(g)
    o.ps_or_silu_hadamard(L.fptr(a), L.fptr(g), L.fptr(u), g.size)
    ro.L.ref_silu_hadamard(ro.h, L.fptr(b), L.fptr(g), L.fptr(u), g.size)
    L.assert_bit_equal(a, b, "silu")


@pytest.mark.parametrize("t,blk", QTYPES)
def test_get_embedding(t, blk):
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(11)
    dim, vocab = blk * 4, 50
    w = synth.random_blocks(rng, t, vocab, dim, 1.0).reshape(-1)
    toks = np.array([0, 49, 7, 7, 23], np.int32)
    a, b = np.zeros(dim * 5, np.float32), np.zeros(dim * 5, np.float32)
    o.ps_or_get_embedding(L.fptr(a), L.vptr(w), t, dim, L.iptr(toks), 5)
    ro.L.ref_get_embedding(ro.h, L.fptr(b), L.vptr(w), t, dim, vocab, L.iptr(toks), 5)
    L.assert_bit_equal(a, b, "embedding")


@pytest.mark.parametrize("hs,nh,nkv,n_kv,bs", [(64, 4, 2, 1, 1), (64, 8, 2, 37, 1), (128, 8, 2, 100, 3), (64, 14, 2, 65, 2)])
def test_attention_matmuls(hs, nh, nkv, n_kv, bs):
    o, ro = L.oracle(), L.ref_ops()
    rng = np.random.default_rng(12)
    n_ctx, kv_dim = 128, hs * nkv
    kc = rand_act(rng, n_ctx * kv_dim, "normal")
    vt = rand_act(rng, kv_dim * n_ctx, "normal")
    q = rand_act(rng, hs * nh * bs, "normal")
    a, b = np.zeros(n_kv * bs * nh, np.float32), np.zeros(n_kv * bs * nh, np.float32)
    o.ps_or_attn_scores(L.fptr(a), L.fptr(kc), L.fptr(q), hs, nh, nkv, n_kv, bs)
    ro.L.ref_attn_scores(ro.h, L.fptr(b), L.fptr(kc), L.fptr(q), hs, nh, nkv, n_kv, bs)
    L.assert_bit_equal(a, b, "attn scores")
    p = np.abs(rand_act(rng, n_kv * bs * nh, "normal"))
    a2, b2 = np.zeros(hs * nh * bs, np.float32), np.zeros(hs * nh * bs, np.float32)
    o.ps_or_attn_pv(L.fptr(a2), L.fptr(vt), L.fptr(p), hs, nh, nkv, n_kv, n_ctx, bs)
    ro.L.ref_attn_pv(ro.h, L.fptr(b2), L.fptr(vt), L.fptr(p), hs, nh, nkv, n_kv, n_ctx, bs)
    L.assert_bit_equal(a2, b2, "attn pv")
