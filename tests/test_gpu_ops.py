"""GPU parity, operator level: every CUDA table op (through the C ABI) must equal the ORACLE bit for bit on seeded
inputs — integer block dots, fp32 chains, libm-restated expf, everything.  Edge cases: ragged block counts
(nb % 4 != 0), zero blocks, ties, batch columns, GQA, masks, softmax tails (n_kv % 8), PV leftovers (n_kv % 32)."""
import numpy as np
import pytest

from powerserve_b200 import capi, synth
from tests import _libs as L

pytestmark = pytest.mark.gpu

QTYPES = [(L.Q4_K, 256), (L.Q6_K, 256), (L.Q4_0, 32), (L.Q8_0, 32), (L.Q5_K, 256)]   # Q5_K: real-file coverage


@pytest.fixture(scope="module")
def be():
    desc = capi.ModelDesc(4096, 14336, 1, 32, 8, 128, 1024, 4096, 1e-5, 128, 0, 5e5, 1.0, 1.0, 0, 32, 0, 1)
    b = capi.CudaBackend(desc)
    yield b
    b.close()


@pytest.fixture(scope="module")
def be64():
    desc = capi.ModelDesc(512, 1536, 1, 8, 2, 64, 1024, 4096, 1e-5, 64, 2, 1e6, 1.0, 1.0, 0, 32, 0, 1)
    b = capi.CudaBackend(desc)
    yield b
    b.close()


def act(rng, n, kind="normal"):
    if kind == "normal":
        return rng.standard_normal(n).astype(np.float32)
    if kind == "wide":
        return (rng.standard_normal(n) * np.exp(rng.uniform(-6, 6, n))).astype(np.float32)
    x = rng.integers(-254, 255, n).astype(np.float32) * 0.5   # ties + zero blocks
    x[: n // 4] = 0.0
    return x


@pytest.mark.parametrize("t,blk", QTYPES)
@pytest.mark.parametrize("bs", [1, 2, 3, 8, 13])
@pytest.mark.parametrize("kind", ["normal", "ties"])
def test_matmul_bit_exact(be, t, blk, bs, kind):
    o = L.oracle()
    rng = np.random.default_rng(100 + bs)
    for K, N in [(blk * 1, 5), (blk * 6, 37), (blk * 16, 64), (blk * 56 if blk == 256 else blk * 152, 19)]:
        w = synth.random_blocks(rng, t, N, K, K ** -0.5).reshape(-1)
        x = act(rng, K * bs, kind)
        ref = np.zeros(N * bs, np.float32)
        o.ps_or_matmul(t, L.vptr(w), K, N, L.fptr(x), bs, L.fptr(ref))
        wd = be.register_weight(w, t, K, N)
        xd, dd = be.upload(x), be.empty(N * bs)
        be.matmul(dd, wd, t, K, N, xd, bs)
        L.assert_bit_equal(dd.numpy(), ref, f"matmul type={t} K={K} N={N} bs={bs}")
        xd.free(); dd.free(); be.unregister_weight(w)


def test_matmul_real_shapes_q4k(be):
    """Llama-3.1-8B / 3.2-1B shapes: sampled rows only on the oracle side to stay fast."""
    o = L.oracle()
    rng = np.random.default_rng(7)
    for K, N in [(4096, 4096), (4096, 1024), (14336, 4096), (2048, 8192)]:
        w = synth.random_blocks(rng, L.Q4_K, N, K, K ** -0.5)
        x = act(rng, K, "normal")
        wd = be.register_weight(w.reshape(-1), L.Q4_K, K, N)
        xd, dd = be.upload(x), be.empty(N)
        be.matmul(dd, wd, L.Q4_K, K, N, xd, 1)
        got = dd.numpy()
        rows = rng.choice(N, 64, replace=False)
        sub = np.ascontiguousarray(w[rows]).reshape(-1)
        ref = np.zeros(64, np.float32)
        o.ps_or_matmul(L.Q4_K, L.vptr(sub), K, 64, L.fptr(x), 1, L.fptr(ref))
        L.assert_bit_equal(got[rows], ref, f"matmul {K}x{N}")
        xd.free(); dd.free(); be.unregister_weight(w.reshape(-1))


@pytest.mark.parametrize("dim,bs", [(64, 1), (896, 3), (4096, 2), (100, 5), (14336, 1)])
def test_rmsnorm(be, dim, bs):
    o = L.oracle()
    rng = np.random.default_rng(6)
    x = act(rng, dim * bs, "wide" if dim == 100 else "normal")
    w = (1 + 0.1 * rng.standard_normal(dim)).astype(np.float32)
    ref = np.zeros_like(x)
    o.ps_or_rmsnorm(L.fptr(ref), L.fptr(x), L.fptr(w), dim, bs, 1e-5)
    xd, wd, dd = be.upload(x), be.upload(w), be.empty(x.size)
    be.rmsnorm(dd, xd, wd, dim, bs, 1e-5)
    L.assert_bit_equal(dd.numpy(), ref, "rmsnorm")


@pytest.mark.parametrize("which", ["norm128", "neox64"])
def test_rope(be, be64, which):
    o = L.oracle()
    b, hs, mode, base = (be, 128, 0, 5e5) if which == "norm128" else (be64, 64, 2, 1e6)
    rng = np.random.default_rng(7)
    nh, pos = 6, np.array([0, 1, 2, 3, 100, 1000, 2047, 4095, 7], np.int32)
    x = act(rng, hs * nh * len(pos))
    ref = np.zeros_like(x)
    o.ps_or_rope(L.fptr(ref), L.fptr(x), hs, nh, len(pos), L.iptr(pos), hs, mode, base, 1.0, 1.0)
    xd, dd = b.upload(x), b.empty(x.size)
    b.rope(dd, xd, hs, nh, len(pos), pos)
    L.assert_bit_equal(dd.numpy(), ref, "rope")


@pytest.mark.parametrize("n_kv,bs,nh", [(1, 1, 4), (7, 1, 4), (8, 1, 2), (33, 3, 4), (300, 5, 2), (2049, 1, 3), (4096, 2, 2)])
def test_softmax_ext(be, n_kv, bs, nh):
    o = L.oracle()
    rng = np.random.default_rng(8)
    x = (rng.standard_normal(n_kv * bs * nh) * 8).astype(np.float32)
    pos = np.arange(n_kv - bs, n_kv, dtype=np.int32)
    mask = np.zeros(n_kv * bs, np.float32)
    o.ps_or_get_mask(L.fptr(mask), n_kv, bs, L.iptr(pos))
    ref = np.zeros_like(x)
    o.ps_or_softmax_ext(L.fptr(ref), L.fptr(x), L.fptr(mask), n_kv, bs, nh, 0.125)
    xd, md, dd = be.upload(x), be.empty(mask.size), be.empty(x.size)
    be.get_mask(md, n_kv, bs, pos)
    L.assert_bit_equal(md.numpy(), mask, "get_mask")
    be.softmax_ext(dd, xd, md, n_kv, bs, nh, 0.125)
    L.assert_bit_equal(dd.numpy(), ref, "softmax_ext")


def test_softmax_exp_ranges(be):
    o = L.oracle()
    xs = np.concatenate([np.linspace(-110, 0, 4099), [-np.inf, -87.3, -88.5, -103.9, -104.1, -87.33655, -200.0]]).astype(np.float32)
    xs[0] = 0.0
    n = len(xs)
    mask = np.zeros(n, np.float32)
    ref = np.zeros(n, np.float32)
    o.ps_or_softmax_ext(L.fptr(ref), L.fptr(xs), L.fptr(mask), n, 1, 1, 1.0)
    xd, md, dd = be.upload(xs), be.upload(mask), be.empty(n)
    be.softmax_ext(dd, xd, md, n, 1, 1, 1.0)
    L.assert_bit_equal(dd.numpy(), ref, "softmax exp ranges")


def test_silu_hadamard_and_add(be):
    o = L.oracle()
    rng = np.random.default_rng(10)
    g = np.concatenate([act(rng, 50000) * 4, np.float32([0, -0.0, 88, -88, 100, -104, 1e-20, -1e-20, 20, -20, 88.8, -103.99])])
    u = act(rng, g.size)
    ref = np.zeros_like(g)
    o.ps_or_silu_hadamard(L.fptr(ref), L.fptr(g), L.fptr(u), g.size)
    gd, ud, dd = be.upload(g), be.upload(u), be.empty(g.size)
    be.silu_hadamard(dd, gd, ud, g.size)
    L.assert_bit_equal(dd.numpy(), ref, "silu_hadamard")
    a_, b_ = act(rng, 896 * 3), act(rng, 896)
    ref2 = np.zeros_like(a_)
    o.ps_or_add(L.fptr(ref2), L.fptr(a_), L.fptr(b_), a_.size, 896)
    ad, bd, d2 = be.upload(a_), be.upload(b_), be.empty(a_.size)
    be.add(d2, ad, bd, a_.size, 896)
    L.assert_bit_equal(d2.numpy(), ref2, "add broadcast")


@pytest.mark.parametrize("t,blk", QTYPES)
def test_get_embedding(be, t, blk):
    o = L.oracle()
    rng = np.random.default_rng(11)
    dim, vocab = blk * 4, 50
    w = synth.random_blocks(rng, t, vocab, dim, 1.0).reshape(-1)
    toks = np.array([0, 49, 7, 7, 23], np.int32)
    ref = np.zeros(dim * 5, np.float32)
    o.ps_or_get_embedding(L.fptr(ref), L.vptr(w), t, dim, L.iptr(toks), 5)
    wd = be.register_weight(w, t, dim, vocab)
    dd = be.empty(dim * 5)
    be.get_embedding(dd, wd, t, dim, toks)
    L.assert_bit_equal(dd.numpy(), ref, "get_embedding")
    be.unregister_weight(w)


@pytest.mark.parametrize("hs,nh,nkv,n_kv,bs", [(64, 4, 2, 1, 1), (64, 8, 2, 37, 1), (128, 8, 2, 100, 3), (64, 14, 2, 65, 2), (128, 32, 8, 2048, 1)])
def test_attention_matmuls(be, hs, nh, nkv, n_kv, bs):
    o = L.oracle()
    rng = np.random.default_rng(12)
    n_ctx, kv_dim = max(128, n_kv), hs * nkv
    kc, vt, q = act(rng, n_ctx * kv_dim), act(rng, kv_dim * n_ctx), act(rng, hs * nh * bs)
    ref = np.zeros(n_kv * bs * nh, np.float32)
    o.ps_or_attn_scores(L.fptr(ref), L.fptr(kc), L.fptr(q), hs, nh, nkv, n_kv, bs)
    kd, vd, qd, sd = be.upload(kc), be.upload(vt), be.upload(q), be.empty(ref.size)
    be.attn_scores(sd, kd, qd, hs, nh, nkv, n_kv, bs)
    L.assert_bit_equal(sd.numpy(), ref, "attn scores")
    p = np.abs(act(rng, n_kv * bs * nh))
    ref2 = np.zeros(hs * nh * bs, np.float32)
    o.ps_or_attn_pv(L.fptr(ref2), L.fptr(vt), L.fptr(p), hs, nh, nkv, n_kv, n_ctx, bs)
    pd, od = be.upload(p), be.empty(ref2.size)
    be.attn_pv(od, vd, pd, hs, nh, nkv, n_kv, n_ctx, bs)
    L.assert_bit_equal(od.numpy(), ref2, "attn pv")
    for b in (kd, vd, qd, sd, pd, od):
        b.free()


def test_error_behaviour(be):
    """Errors surface as exceptions carrying the reference's assertion text (C-ABI status != 0), never silently."""
    x = be.empty(64)
    with pytest.raises(capi.PsCudaError):
        be.rmsnorm(x, x, x, 64, 1, 0.0)          # GGML_ASSERT(eps > 0)
    with pytest.raises(capi.PsCudaError):
        be.rope(x, x, 64, 1, 1, [0])             # head_size mismatch with the model's rope dims
    with pytest.raises(capi.PsCudaError):
        be.kv_rollback(1)                        # POWERSERVE_ASSERT_KVCACHE(position >= n_tokens)


def test_get_embedding_f16_table(be):
    """an F16 token_embd table (real GGUFs of the F16 / some K-quant recipes): GGML_FP16_TO_FP32 per element"""
    rng = np.random.default_rng(31)
    dim, vocab = 96, 40
    w16 = rng.standard_normal(vocab * dim).astype(np.float16)
    toks = np.array([0, 39, 5, 5], np.int32)
    w = w16.view(np.uint8)
    wd = be.register_weight(w, 1, dim, vocab)
    dd = be.empty(dim * 4)
    be.get_embedding(dd, wd, 1, dim, toks)
    L.assert_bit_equal(dd.numpy(), w16.reshape(vocab, dim)[toks].astype(np.float32).reshape(-1), "f16 embedding")
    be.unregister_weight(w)


def test_rope_freq_factors(be):
    """rope_freqs.weight (ggml_rope_cache_init with freq_factors, ggml.c:15342-15356: rope_yarn(theta / ff, ...)); off by default
    because the reference never passes it.  Checked against a float32 restatement evaluated with the platform
    libm (the table is built on the host with cosf / sinf / powf), then switched off again (bit-exact vs the oracle)."""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    for f in ("cosf", "sinf"):
        getattr(libm, f).restype = C.c_float
        getattr(libm, f).argtypes = [C.c_float]
    libm.powf.restype = C.c_float
    libm.powf.argtypes = [C.c_float, C.c_float]
    hs, nh, base = 128, 3, 5e5
    rng = np.random.default_rng(17)
    ff = np.concatenate([np.ones(20), np.linspace(1.0, 8.0, 24), np.full(20, 8.0)]).astype(np.float32)     # llama3-style: low frequencies stretched
    pos = np.array([0, 1, 5, 100, 2047, 4095], np.int32)
    x = act(rng, hs * nh * len(pos))
    be.set_rope_freq_factors(ff)
    xd, dd = be.upload(x), be.empty(x.size)
    be.rope(dd, xd, hs, nh, len(pos), pos)
    got = dd.numpy().reshape(len(pos), nh, hs)
    theta_scale = np.float32(libm.powf(np.float32(base), np.float32(-2.0 / hs)))
    want = np.zeros_like(got)
    X = x.reshape(len(pos), nh, hs)
    for a, p in enumerate(pos):
        theta = np.float32(p)
        for i0 in range(0, hs, 2):
            th = np.float32(np.float32(1.0) * np.float32(theta / ff[i0 // 2]))
            c, s = np.float32(libm.cosf(th)), np.float32(libm.sinf(th))
            x0, x1 = X[a, :, i0], X[a, :, i0 + 1]
            want[a, :, i0] = (x0 * c).astype(np.float32) - (x1 * s).astype(np.float32)
            want[a, :, i0 + 1] = (x0 * s).astype(np.float32) + (x1 * c).astype(np.float32)
            theta = np.float32(theta * theta_scale)
    L.assert_bit_equal(got.reshape(-1), want.reshape(-1), "rope with freq factors")
    be.set_rope_freq_factors(None)
    o = L.oracle()
    ref = np.zeros_like(x)
    o.ps_or_rope(L.fptr(ref), L.fptr(x), hs, nh, len(pos), L.iptr(pos), hs, 0, base, 1.0, 1.0)
    be.rope(dd, xd, hs, nh, len(pos), pos)
    L.assert_bit_equal(dd.numpy(), ref, "rope, factors off again")
