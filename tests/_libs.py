"""ctypes loaders for the CHECKERS used by the tests: oracle/libps_oracle.so (our restatement) and, when present,
oracle/_ref/*.so (the reference compiled from /root/reference).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

F32, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K, Q8_K = 0, 2, 8, 12, 13, 14, 15

c_f = C.POINTER(C.c_float)
c_i32 = C.POINTER(C.c_int32)


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_f)


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_i32)


def vptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class OrConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "ffn_dim", "n_layers", "n_heads", "n_kv_heads", "head_size", "vocab_size", "n_ctx")] + \
               [("norm_eps", C.c_float), ("rope_n_dims", C.c_int32), ("rope_type", C.c_int32),
                ("rope_freq_base", C.c_float), ("rope_freq_scale", C.c_float), ("rope_attn_factor", C.c_float),
                ("qkv_bias", C.c_int32)]


class OrTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("type", C.c_int32), ("_pad", C.c_int32)]


class OrLayer(C.Structure):
    _fields_ = [(n, OrTensor) for n in ("attn_norm", "ffn_norm", "attn_q", "attn_k", "attn_v", "attn_output",
                                        "ffn_gate", "ffn_up", "ffn_down", "q_bias", "k_bias", "v_bias")]


class OrWeights(C.Structure):
    _fields_ = [("token_embd", OrTensor), ("output_norm", OrTensor), ("output", OrTensor), ("layers", C.POINTER(OrLayer))]


_oracle: Optional[C.CDLL] = None


def build_oracle() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is not None:
        return _oracle
    path = os.path.join(ORACLE_DIR, "libps_oracle.so")
    src = os.path.join(ORACLE_DIR, "ps_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build_oracle()
    L = C.CDLL(path)
    L.ps_or_fp16_to_fp32.restype = C.c_float
    L.ps_or_fp16_to_fp32.argtypes = [C.c_uint16]
    L.ps_or_fp32_to_fp16.restype = C.c_uint16
    L.ps_or_fp32_to_fp16.argtypes = [C.c_float]
    L.ps_or_v_expf.restype = C.c_float
    L.ps_or_v_expf.argtypes = [C.c_float]
    L.ps_or_row_size.restype = C.c_size_t
    L.ps_or_row_size.argtypes = [C.c_int, C.c_int64]
    L.ps_or_quantize_row.argtypes = [C.c_int, c_f, C.c_void_p, C.c_int64]
    L.ps_or_dequantize_row.argtypes = [C.c_int, C.c_void_p, c_f, C.c_int64]
    L.ps_or_vec_dot.restype = C.c_float
    L.ps_or_vec_dot.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.ps_or_vec_dot_f32.restype = C.c_float
    L.ps_or_vec_dot_f32.argtypes = [C.c_int64, c_f, c_f]
    L.ps_or_matmul.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, c_f, C.c_int64, c_f]
    L.ps_or_rmsnorm.argtypes = [c_f, c_f, c_f, C.c_int64, C.c_int64, C.c_float]
    L.ps_or_rope.argtypes = [c_f, c_f, C.c_int64, C.c_int64, C.c_int64, c_i32, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    L.ps_or_get_mask.argtypes = [c_f, C.c_int64, C.c_int64, c_i32]
    L.ps_or_softmax_ext.argtypes = [c_f, c_f, c_f, C.c_int64, C.c_int64, C.c_int64, C.c_float]
    L.ps_or_add.argtypes = [c_f, c_f, c_f, C.c_int64, C.c_int64]
    L.ps_or_silu_hadamard.argtypes = [c_f, c_f, c_f, C.c_int64]
    L.ps_or_get_embedding.argtypes = [c_f, C.c_void_p, C.c_int, C.c_int64, c_i32, C.c_int64]
    L.ps_or_attn_scores.argtypes = [c_f, c_f, c_f] + [C.c_int64] * 5
    L.ps_or_attn_pv.argtypes = [c_f, c_f, c_f] + [C.c_int64] * 6
    L.ps_or_model_create.restype = C.c_void_p
    L.ps_or_model_create.argtypes = [C.POINTER(OrConfig), C.POINTER(OrWeights)]
    L.ps_or_model_free.argtypes = [C.c_void_p]
    L.ps_or_model_reset.argtypes = [C.c_void_p]
    L.ps_or_model_position.argtypes = [C.c_void_p]
    L.ps_or_model_set_position.argtypes = [C.c_void_p, C.c_int]
    L.ps_or_model_forward.argtypes = [C.c_void_p, c_i32, c_i32, C.c_int, C.c_int, c_f]
    L.ps_or_model_tap.restype = C.c_int64
    L.ps_or_model_tap.argtypes = [C.c_void_p, C.c_int, C.c_int, c_f]
    L.ps_or_model_k_cache.restype = c_f
    L.ps_or_model_k_cache.argtypes = [C.c_void_p, C.c_int]
    L.ps_or_model_v_cache.restype = c_f
    L.ps_or_model_v_cache.argtypes = [C.c_void_p, C.c_int]
    _oracle = L
    return L


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("libggml_ref.so", "libps_ref_ops.so", "ps_ref_run"))


_ref_ggml: Optional[C.CDLL] = None
_ref_ops = None


def ref_ggml() -> C.CDLL:
    """The reference's vendored ggml: raw quantisers / dequantisers / vec_dot kernels."""
    global _ref_ggml
    if _ref_ggml is None:
        L = C.CDLL(os.path.join(REF_DIR, "libggml_ref.so"))

        class InitParams(C.Structure):
            _fields_ = [("mem_size", C.c_size_t), ("mem_buffer", C.c_void_p), ("no_alloc", C.c_bool)]

        # ggml_init fills the fp16->fp32 lookup table every block kernel reads (ggml.c:3700-3712)
        L.ggml_init.restype = C.c_void_p
        L.ggml_init.argtypes = [InitParams]
        L._ctx = L.ggml_init(InitParams(1 << 20, None, True))
        for n in ("quantize_row_q8_K", "quantize_row_q8_0"):
            getattr(L, n).argtypes = [c_f, C.c_void_p, C.c_int64]
        for n in ("dequantize_row_q4_0", "dequantize_row_q8_0", "dequantize_row_q4_K", "dequantize_row_q5_K", "dequantize_row_q6_K"):
            getattr(L, n).argtypes = [C.c_void_p, c_f, C.c_int64]
        for n in ("ggml_vec_dot_q4_K_q8_K", "ggml_vec_dot_q5_K_q8_K", "ggml_vec_dot_q6_K_q8_K", "ggml_vec_dot_q4_0_q8_0", "ggml_vec_dot_q8_0_q8_0"):
            getattr(L, n).argtypes = [C.c_int, c_f, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ggml_fp16_to_fp32.restype = C.c_float
        L.ggml_fp16_to_fp32.argtypes = [C.c_uint16]
        L.ggml_fp32_to_fp16.restype = C.c_uint16
        L.ggml_fp32_to_fp16.argtypes = [C.c_float]
        _ref_ggml = L
    return _ref_ggml


class RefOps:
    """The reference's GGMLBackend operator table behind oracle/ref_ops_shim.cpp."""

    def __init__(self, n_threads: int = 4):
        L = C.CDLL(os.path.join(REF_DIR, "libps_ref_ops.so"))
        L.ref_backend_create.restype = C.c_void_p
        L.ref_backend_create.argtypes = [C.c_int]
        L.ref_backend_destroy.argtypes = [C.c_void_p]
        L.ref_matmul.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, c_f, C.c_int64, c_f]
        L.ref_rmsnorm.argtypes = [C.c_void_p, c_f, c_f, c_f, C.c_int64, C.c_int64, C.c_float]
        L.ref_rope.argtypes = [C.c_void_p, c_f, c_f, C.c_int64, C.c_int64, C.c_int64, c_i32, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_softmax_ext.argtypes = [C.c_void_p, c_f, c_f, c_f, C.c_int64, C.c_int64, C.c_int64, C.c_float]
        L.ref_add.argtypes = [C.c_void_p, c_f, c_f, c_f, C.c_int64, C.c_int64, C.c_int64]
        L.ref_silu_hadamard.argtypes = [C.c_void_p, c_f, c_f, c_f, C.c_int64]
        L.ref_get_embedding.argtypes = [C.c_void_p, c_f, C.c_void_p, C.c_int, C.c_int64, C.c_int64, c_i32, C.c_int64]
        L.ref_attn_scores.argtypes = [C.c_void_p, c_f, c_f, c_f] + [C.c_int64] * 5
        L.ref_attn_pv.argtypes = [C.c_void_p, c_f, c_f, c_f] + [C.c_int64] * 6
        self.L = L
        self.h = L.ref_backend_create(n_threads)

    def close(self):
        if self.h:
            self.L.ref_backend_destroy(self.h)
            self.h = None


def ref_ops(n_threads: int = 4) -> RefOps:
    global _ref_ops
    if _ref_ops is None:
        _ref_ops = RefOps(n_threads)
    return _ref_ops


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(a: np.ndarray, b: np.ndarray, what: str = "") -> None:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ne = bits(a) != bits(b)
    # +0 / -0 and NaN payloads are compared bitwise on purpose
    if ne.any():
        idx = np.flatnonzero(ne.reshape(-1))
        i = idx[0]
        raise AssertionError(
            f"{what}: {idx.size}/{a.size} elements differ bitwise; first @ {i}: {a.reshape(-1)[i]!r} vs {b.reshape(-1)[i]!r} "
            f"(max abs diff {np.nanmax(np.abs(a.reshape(-1)[idx] - b.reshape(-1)[idx])):.3e})")
