"""The golden cases, written once against an abstract backend so that the SAME code produces the golden vectors from
the compiled reference (make_golden.py), checks the oracle against them (CPU) and checks the CUDA library (GPU)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from powerserve_b200 import synth
from tests import _libs as L

QTYPES = {"q4_K": (L.Q4_K, 256), "q6_K": (L.Q6_K, 256), "q4_0": (L.Q4_0, 32), "q8_0": (L.Q8_0, 32)}
MODEL_CASES = [("tiny-llama", 20, 8, 8), ("tiny-qwen2", 20, 8, 8), ("tiny-q8", 41, 128, 6), ("tiny-mixed", 1, 128, 6), ("tiny-llama-hs128", 41, 128, 6)]
# long-context cases (tests/golden/long_ctx.npz, make_golden_long.py): (preset, prompt tokens, prefill batch, decode steps)
LONG_CASES = [("slice-8b-long", 2100, 128, 10), ("slice-1b-long", 2100, 128, 10), ("slice-1b-long", 4400, 128, 6)]


def act(seed, n, kind="normal"):
    rng = np.random.default_rng(seed)
    if kind == "wide":
        return (rng.standard_normal(n) * np.exp(rng.uniform(-6, 6, n))).astype(np.float32)
    if kind == "ties":
        x = rng.integers(-254, 255, n).astype(np.float32) * 0.5
        x[: n // 8] = 0.0
        return x
    return rng.standard_normal(n).astype(np.float32)


def weights(seed, t, n_rows, k):
    rng = np.random.Generator(np.random.SFC64([seed, t]))
    return np.ascontiguousarray(synth.random_blocks(rng, t, n_rows, k, k ** -0.5).reshape(-1))


class RefBackend:
    """the compiled reference (oracle/_ref)"""

    def __init__(self):
        self.g, self.o = L.ref_ggml(), L.ref_ops(3)

    def quantize(self, t, x):
        out = np.zeros(L.oracle().ps_or_row_size(t, x.size), np.uint8)
        (self.g.quantize_row_q8_K if t == L.Q8_K else self.g.quantize_row_q8_0)(L.fptr(x), L.vptr(out), x.size)
        return out

    def matmul(self, t, w, k, n, x, bs):
        out = np.zeros((bs, n), np.float32)
        self.o.L.ref_matmul(self.o.h, t, L.vptr(w), k, n, L.fptr(x), bs, L.fptr(out))
        return out

    def rmsnorm(self, x, w, dim, bs, eps):
        out = np.zeros_like(x)
        self.o.L.ref_rmsnorm(self.o.h, L.fptr(out), L.fptr(x), L.fptr(w), dim, bs, eps)
        return out

    def rope(self, x, hs, nh, bs, pos, mode, base):
        out = np.zeros_like(x)
        self.o.L.ref_rope(self.o.h, L.fptr(out), L.fptr(x), hs, nh, bs, L.iptr(pos), hs, mode, base, 1.0, 1.0)
        return out

    def softmax_ext(self, x, mask, ne0, ne1, ne2, scale):
        out = np.zeros_like(x)
        self.o.L.ref_softmax_ext(self.o.h, L.fptr(out), L.fptr(x), L.fptr(mask), ne0, ne1, ne2, scale)
        return out

    def silu_hadamard(self, g, u):
        out = np.zeros_like(g)
        self.o.L.ref_silu_hadamard(self.o.h, L.fptr(out), L.fptr(g), L.fptr(u), g.size)
        return out

    def get_embedding(self, w, t, dim, vocab, tokens):
        out = np.zeros((len(tokens), dim), np.float32)
        self.o.L.ref_get_embedding(self.o.h, L.fptr(out), L.vptr(w), t, dim, vocab, L.iptr(tokens), len(tokens))
        return out


class OracleBackend:
    """oracle/ps_oracle.c"""

    def __init__(self):
        self.o = L.oracle()

    def quantize(self, t, x):
        out = np.zeros(self.o.ps_or_row_size(t, x.size), np.uint8)
        self.o.ps_or_quantize_row(t, L.fptr(x), L.vptr(out), x.size)
        return out

    def matmul(self, t, w, k, n, x, bs):
        out = np.zeros((bs, n), np.float32)
        self.o.ps_or_matmul(t, L.vptr(w), k, n, L.fptr(x), bs, L.fptr(out))
        return out

    def rmsnorm(self, x, w, dim, bs, eps):
        out = np.zeros_like(x)
        self.o.ps_or_rmsnorm(L.fptr(out), L.fptr(x), L.fptr(w), dim, bs, eps)
        return out

    def rope(self, x, hs, nh, bs, pos, mode, base):
        out = np.zeros_like(x)
        self.o.ps_or_rope(L.fptr(out), L.fptr(x), hs, nh, bs, L.iptr(pos), hs, mode, base, 1.0, 1.0)
        return out

    def softmax_ext(self, x, mask, ne0, ne1, ne2, scale):
        out = np.zeros_like(x)
        self.o.ps_or_softmax_ext(L.fptr(out), L.fptr(x), L.fptr(mask), ne0, ne1, ne2, scale)
        return out

    def silu_hadamard(self, g, u):
        out = np.zeros_like(g)
        self.o.ps_or_silu_hadamard(L.fptr(out), L.fptr(g), L.fptr(u), g.size)
        return out

    def get_embedding(self, w, t, dim, vocab, tokens):
        out = np.zeros((len(tokens), dim), np.float32)
        self.o.ps_or_get_embedding(L.fptr(out), L.vptr(w), t, dim, L.iptr(tokens), len(tokens))
        return out


def case_quantize(be):
    res = {}
    for kind in ("normal", "wide", "ties"):
        x = act(11, 256 * 5, kind)
        res[f"q8_K/{kind}"] = be.quantize(L.Q8_K, x)
        res[f"q8_0/{kind}"] = be.quantize(L.Q8_0, x[:32 * 7])
    return res


def case_matmul(be):
    res = {}
    for name, (t, blk) in QTYPES.items():
        k, n = blk * (3 if blk == 256 else 20), 37
        w = weights(21, t, n, k)
        for bs in (1, 3):
            x = act(22 + bs, k * bs, "wide" if bs == 3 else "normal")
            res[f"{name}/bs{bs}"] = L.bits(be.matmul(t, w, k, n, x, bs))
    return res


def case_small_ops(be):
    res = {}
    dim, bs = 640, 3
    x = act(31, dim * bs).reshape(bs, dim)
    w = (1.0 + 0.1 * act(32, dim)).astype(np.float32)
    res["rmsnorm"] = L.bits(be.rmsnorm(x, w, dim, bs, 1e-5))
    hs, nh = 64, 6
    xr = act(33, hs * nh * bs)
    pos = np.asarray([0, 17, 333], np.int32)
    res["rope/norm"] = L.bits(be.rope(xr, hs, nh, bs, pos, 0, 5e5))
    res["rope/neox"] = L.bits(be.rope(xr, hs, nh, bs, pos, 2, 1e6))
    ne0, ne1, ne2 = 45, 3, 4
    s = act(34, ne0 * ne1 * ne2)
    mask = np.where(np.arange(ne0)[None, :] <= np.asarray([42, 43, 44])[:, None], 0.0, -np.inf).astype(np.float32)
    res["softmax_ext"] = L.bits(be.softmax_ext(s, np.ascontiguousarray(mask), ne0, ne1, ne2, 0.125))
    g, u = act(35, 1000, "wide"), act(36, 1000)
    res["silu_hadamard"] = L.bits(be.silu_hadamard(np.clip(g, -100, 100).astype(np.float32), u))
    toks = np.asarray([0, 5, 63, 17], np.int32)
    for name, (t, blk) in QTYPES.items():
        dimw = blk * 2
        wt = weights(37, t, 64, dimw)
        res[f"get_embedding/{name}"] = L.bits(be.get_embedding(wt, t, dimw, 64, toks))
    return res


OP_CASES = {"quantize": case_quantize, "matmul": case_matmul, "small": case_small_ops}

# (gguf-py type name, ggml type id, block elements, rows, blocks per row, seed, scale) for make_golden_gguf_py.py: the
# reference's Python dequantisers (pinned against libggml by its own gguf-py/tests/test_quants.py) on seeded random blocks
GGUF_PY_CASES = [("Q4_K", L.Q4_K, 256, 24, 4, 101, 1.0), ("Q4_K", L.Q4_K, 256, 3, 56, 102, 0.02), ("Q6_K", L.Q6_K, 256, 24, 4, 103, 1.0),
                 ("Q4_0", L.Q4_0, 32, 24, 16, 104, 1.0), ("Q8_0", L.Q8_0, 32, 24, 16, 105, 1.0), ("Q8_0", L.Q8_0, 32, 5, 152, 106, 30.0),
                 ("Q5_K", L.Q5_K, 256, 24, 4, 107, 1.0), ("Q5_K", L.Q5_K, 256, 3, 56, 108, 0.02)]
# Q5_K dot products (real-file coverage): (K, rows, seed) - golden values from the compiled reference's ggml_vec_dot_q5_K_q8_K
# (tests/golden/make_golden_q5k.py -> q5k_dot.npz); the reference's DataType layer cannot carry Q5_K, so its operator
# table is not involved
Q5K_DOT_CASES = [(256, 9, 201), (256 * 16, 7, 202), (256 * 56, 5, 203)]
