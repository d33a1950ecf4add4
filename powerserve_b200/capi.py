"""ctypes binding of libps_cuda.so (include/ps_cuda.h) — the same entry points a PowerServe `CUDABackend` C++ class
binds.  Python is only the test / bench harness here: every call below crosses the C ABI, nothing
is computed in Python, and there is NO fallback — a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import gguf

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libps_cuda.so")

F32, Q4_0, Q8_0, Q4_K, Q6_K = 0, 2, 8, 12, 14


class PsCudaError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "ffn_dim", "n_layers", "n_heads", "n_kv_heads", "head_size", "vocab_size", "n_ctx")] + [
        ("norm_eps", C.c_float), ("rope_n_dims", C.c_int32), ("rope_type", C.c_int32), ("rope_freq_base", C.c_float),
        ("rope_freq_scale", C.c_float), ("rope_attn_factor", C.c_float), ("qkv_bias", C.c_int32), ("max_batch", C.c_int32),
        ("tp_rank", C.c_int32), ("tp_size", C.c_int32)]


class Tensor(C.Structure):
    _fields_ = [("host", C.c_void_p), ("type", C.c_int32), ("_pad", C.c_int32)]


class LayerWeights(C.Structure):
    _fields_ = [(n, Tensor) for n in ("attn_norm", "ffn_norm", "attn_q", "attn_k", "attn_v", "attn_output", "ffn_gate", "ffn_up",
                                      "ffn_down", "attn_q_bias", "attn_k_bias", "attn_v_bias")]


class ModelWeights(C.Structure):
    _fields_ = [("token_embd", Tensor), ("output_norm", Tensor), ("output", Tensor), ("layers", C.POINTER(LayerWeights))]


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load libps_cuda.so; raises if it has not been built (python -m powerserve_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PsCudaError(f"{LIB_PATH} is missing — build it with `python -m powerserve_b200.build`; there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32p, fp, i64, ci, sz = C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_int64, C.c_int, C.c_size_t
    sig = {
        "ps_cuda_abi_version": (ci, []),
        "ps_cuda_device_count": (ci, []),
        "ps_cuda_create": (ci, [C.POINTER(vp), ci, C.POINTER(ModelDesc)]),
        "ps_cuda_destroy": (None, [vp]),
        "ps_cuda_last_error": (C.c_char_p, [vp]),
        "ps_cuda_sync": (ci, [vp]),
        "ps_cuda_stream": (vp, [vp]),
        "ps_cuda_malloc": (ci, [vp, sz, C.POINTER(vp)]),
        "ps_cuda_free": (ci, [vp, vp]),
        "ps_cuda_memcpy_h2d": (ci, [vp, vp, vp, sz]),
        "ps_cuda_memcpy_d2h": (ci, [vp, vp, vp, sz]),
        "ps_cuda_register_weight": (ci, [vp, vp, ci, i64, i64, C.POINTER(vp)]),
        "ps_cuda_lookup_weight": (vp, [vp, vp]),
        "ps_cuda_unregister_weight": (ci, [vp, vp]),
        "ps_cuda_get_embedding": (ci, [vp, fp, vp, ci, i64, i32p, i64]),
        "ps_cuda_rmsnorm": (ci, [vp, fp, fp, fp, i64, i64, C.c_float]),
        "ps_cuda_matmul": (ci, [vp, fp, vp, ci, i64, i64, fp, i64]),
        "ps_cuda_rope": (ci, [vp, fp, fp, i64, i64, i64, i32p]),
        "ps_cuda_add": (ci, [vp, fp, fp, fp, i64, i64]),
        "ps_cuda_silu_hadamard": (ci, [vp, fp, fp, fp, i64]),
        "ps_cuda_get_mask": (ci, [vp, fp, i64, i64, i32p]),
        "ps_cuda_softmax_ext": (ci, [vp, fp, fp, fp, i64, i64, i64, C.c_float]),
        "ps_cuda_attn_scores": (ci, [vp, fp, fp, fp, i64, i64, i64, i64, i64]),
        "ps_cuda_attn_pv": (ci, [vp, fp, fp, fp, i64, i64, i64, i64, i64, i64]),
        "ps_cuda_copy_2d": (ci, [vp, vp, i64, i64, vp, i64, i64, i64, i64]),
        "ps_cuda_copy_4d": (ci, [vp, vp, C.POINTER(i64), C.POINTER(i64), vp, C.POINTER(i64), C.POINTER(i64)]),
        "ps_cuda_matmul_f32": (ci, [vp, fp, vp, i64, i64, i64, i64, i64, vp, i64, i64, i64, i64]),
        "ps_cuda_softmax": (ci, [vp, fp, fp, i64, i64]),
        "ps_cuda_kv_position": (ci, [vp]),
        "ps_cuda_kv_reset": (ci, [vp]),
        "ps_cuda_kv_truncate": (ci, [vp, ci]),
        "ps_cuda_kv_rollback": (ci, [vp, ci]),
        "ps_cuda_kv_advance": (ci, [vp, ci]),
        "ps_cuda_kv_copy_slot": (ci, [vp, ci, ci]),
        "ps_cuda_kv_move_slot": (ci, [vp, ci, ci]),
        "ps_cuda_kv_mask_slot": (ci, [vp, ci]),
        "ps_cuda_kv_unmask_slot": (ci, [vp, ci]),
        "ps_cuda_kv_k": (vp, [vp, ci]),
        "ps_cuda_kv_v": (vp, [vp, ci]),
        "ps_cuda_bind_model": (ci, [vp, C.POINTER(ModelWeights)]),
        "ps_cuda_forward": (ci, [vp, i32p, i32p, ci, ci, C.c_void_p]),
        "ps_cuda_forward_tree": (ci, [vp, i32p, i32p, ci, C.c_void_p, ci, C.c_void_p]),
        "ps_cuda_decode_greedy": (ci, [vp, C.c_int32, ci, i32p]),
        "ps_cuda_sample_topk": (ci, [vp, ci, ci, fp, i32p]),
        "ps_cuda_set_rope_freq_factors": (ci, [vp, fp, ci]),
        "ps_cuda_session_create": (ci, [vp, C.POINTER(ci)]),
        "ps_cuda_session_destroy": (ci, [vp, ci]),
        "ps_cuda_session_select": (ci, [vp, ci]),
        "ps_cuda_session_current": (ci, [vp]),
        "ps_cuda_session_position": (ci, [vp, ci]),
        "ps_cuda_forward_sessions": (ci, [vp, i32p, i32p, ci, ci, fp, i32p]),
        "ps_cuda_logits_dev": (vp, [vp]),
        "ps_cuda_tp_unique_id": (ci, [vp]),
        "ps_cuda_tp_init": (ci, [vp, vp]),
        "ps_cuda_tp_export": (ci, [vp, vp]),
        "ps_cuda_tp_import": (ci, [vp, vp, ci]),
        "ps_cuda_set_option": (ci, [vp, C.c_char_p, ci]),
        "ps_cuda_get_counter": (i64, [vp, C.c_char_p]),
        "ps_cuda_read_trace": (ci, [vp, C.c_void_p, ci]),
        "ps_cuda_host_expf_ref": (C.c_float, [C.c_float]),
        "ps_cuda_host_v_expf": (C.c_float, [C.c_float]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError here == the library does not export what include/ps_cuda.h declares
        fn.restype = res
        fn.argtypes = args
    L._ps_signatures = sig
    _lib = L
    return L


def exported_symbols() -> List[str]:
    return list(load_library()._ps_signatures.keys())


def _i32(a: Sequence[int]) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class DeviceBuffer:
    """A CUDABuffer: device memory owned by a context (ps_cuda_malloc / ps_cuda_free)."""

    def __init__(self, be: "CudaBackend", nbytes: int):
        self.be, self.nbytes = be, int(nbytes)
        p = C.c_void_p()
        be._ck(be.L.ps_cuda_malloc(be.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def from_numpy(cls, be: "CudaBackend", a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = cls(be, a.nbytes)
        be._ck(be.L.ps_cuda_memcpy_h2d(be.h, b.ptr, a.ctypes.data, a.nbytes))
        return b

    def numpy(self, dtype=np.float32, shape=None) -> np.ndarray:
        out = np.empty(self.nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self.be._ck(self.be.L.ps_cuda_memcpy_d2h(self.be.h, out.ctypes.data, self.ptr, out.nbytes))
        return out.reshape(shape) if shape is not None else out

    def free(self):
        if self.ptr:
            self.be.L.ps_cuda_free(self.be.h, self.ptr)
            self.ptr = None


class CudaBackend:
    """Mirror of `powerserve::ggml::GGMLBackend`'s operator table (src/backend/ggml/ggml.hpp:216-244) over the C ABI.
    Method names, argument meaning and error behaviour follow the reference; tensors are DeviceBuffers."""

    def __init__(self, desc: ModelDesc, device: int = 0):
        self.L = load_library()
        self.desc = desc
        h = C.c_void_p()
        rc = self.L.ps_cuda_create(C.byref(h), device, C.byref(desc))
        if rc != 0:
            raise PsCudaError(f"ps_cuda_create failed ({rc}): {self.L.ps_cuda_last_error(None).decode()}")
        self.h = h

    def _ck(self, rc: int):
        if rc != 0:
            raise PsCudaError(f"libps_cuda error {rc}: {self.L.ps_cuda_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.L.ps_cuda_destroy(self.h)
            self.h = None

    # ---- memory
    def empty(self, n_floats: int) -> DeviceBuffer:
        return DeviceBuffer(self, 4 * int(n_floats))

    def upload(self, a: np.ndarray) -> DeviceBuffer:
        return DeviceBuffer.from_numpy(self, a)

    def register_weight(self, host: np.ndarray, ggml_type: int, ne0: int, ne1: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.ps_cuda_register_weight(self.h, host.ctypes.data, ggml_type, ne0, ne1, C.byref(p)))
        return p.value

    def unregister_weight(self, host: np.ndarray):
        self._ck(self.L.ps_cuda_unregister_weight(self.h, host.ctypes.data))

    def sync(self):
        self._ck(self.L.ps_cuda_sync(self.h))

    def counter(self, name: str) -> int:
        return int(self.L.ps_cuda_get_counter(self.h, name.encode()))

    def set_option(self, name: str, value: int):
        self._ck(self.L.ps_cuda_set_option(self.h, name.encode(), int(value)))

    # ---- operator table
    def get_embedding(self, dst, weight_dev, wtype, dim, tokens):
        t = _i32(tokens)
        self._ck(self.L.ps_cuda_get_embedding(self.h, dst.ptr, weight_dev, wtype, dim, t.ctypes.data_as(C.POINTER(C.c_int32)), len(t)))

    def rmsnorm(self, out, x, weight, dim, bs, eps):
        self._ck(self.L.ps_cuda_rmsnorm(self.h, out.ptr, x.ptr, weight.ptr, dim, bs, eps))

    def matmul(self, dst, weight_dev, wtype, K, N, x, bs):
        self._ck(self.L.ps_cuda_matmul(self.h, dst.ptr, weight_dev, wtype, K, N, x.ptr, bs))

    def rope(self, out, src, head_size, n_heads, bs, pos):
        p = _i32(pos)
        self._ck(self.L.ps_cuda_rope(self.h, out.ptr, src.ptr, head_size, n_heads, bs, p.ctypes.data_as(C.POINTER(C.c_int32))))

    def add(self, dst, a, b, n, nb):
        self._ck(self.L.ps_cuda_add(self.h, dst.ptr, a.ptr, b.ptr, n, nb))

    def silu_hadamard(self, out, hb, hb2, n):
        self._ck(self.L.ps_cuda_silu_hadamard(self.h, out.ptr, hb.ptr, hb2.ptr, n))

    def get_mask(self, mask, n_kv, bs, pos):
        p = _i32(pos)
        self._ck(self.L.ps_cuda_get_mask(self.h, mask.ptr, n_kv, bs, p.ctypes.data_as(C.POINTER(C.c_int32))))

    def softmax_ext(self, out, x, mask, ne0, ne1, ne2, scale):
        self._ck(self.L.ps_cuda_softmax_ext(self.h, out.ptr, x.ptr, mask.ptr, ne0, ne1, ne2, scale))

    def attn_scores(self, kq, k_cache, q, hs, n_heads, n_kv_heads, n_kv, bs):
        self._ck(self.L.ps_cuda_attn_scores(self.h, kq.ptr, k_cache.ptr, q.ptr, hs, n_heads, n_kv_heads, n_kv, bs))

    def attn_pv(self, out, v_cache_t, p, hs, n_heads, n_kv_heads, n_kv, n_ctx, bs):
        self._ck(self.L.ps_cuda_attn_pv(self.h, out.ptr, v_cache_t.ptr, p.ptr, hs, n_heads, n_kv_heads, n_kv, n_ctx, bs))

    def copy_2d(self, dst, ds0, ds1, src, ss0, ss1, ne0, ne1):
        """GGMLBackend::copy / cont on 2-D fp32 views (byte strides; powerserve_compute_forward_dup)."""
        self._ck(self.L.ps_cuda_copy_2d(self.h, dst.ptr, ds0, ds1, src.ptr, ss0, ss1, ne0, ne1))

    def set_rope_freq_factors(self, factors):
        """rope_freqs.weight (off by default: the reference ignores it); None restores the default"""
        if factors is None:
            self._ck(self.L.ps_cuda_set_rope_freq_factors(self.h, None, 0))
        else:
            f = np.ascontiguousarray(factors, dtype=np.float32)
            self._ck(self.L.ps_cuda_set_rope_freq_factors(self.h, f.ctypes.data_as(C.POINTER(C.c_float)), len(f)))

    def copy_4d(self, dst, dst_ne, dst_nb, src, src_ne, src_nb, dst_off=0, src_off=0):
        """GGMLBackend::copy / cont on views of up to four dims whose shapes may differ (shapes in elements, strides in bytes)."""
        a = lambda v: (C.c_int64 * 4)(*[int(x) for x in v])
        self._ck(self.L.ps_cuda_copy_4d(self.h, dst.ptr + dst_off, a(dst_ne), a(dst_nb), src.ptr + src_off, a(src_ne), a(src_nb)))

    def matmul_f32(self, dst, src0, ne00, ne01, ne02, nb01, nb02, src1, ne11, ne12, nb11, nb12):
        """GGMLBackend::matmul with an FP32 src0 over strided views (the attention products of the unfused graph)."""
        self._ck(self.L.ps_cuda_matmul_f32(self.h, dst.ptr, src0.ptr, ne00, ne01, ne02, nb01, nb02, src1.ptr, ne11, ne12, nb11, nb12))

    def softmax(self, out, x, ne0, n_rows):
        self._ck(self.L.ps_cuda_softmax(self.h, out.ptr, x.ptr, ne0, n_rows))

    def read_device(self, dev_ptr: int, n_floats: int, dtype=np.float32) -> np.ndarray:
        out = np.empty(n_floats, dtype=dtype)
        self._ck(self.L.ps_cuda_memcpy_d2h(self.h, out.ctypes.data, dev_ptr, out.nbytes))
        return out

    # ---- kv
    def kv_advance(self, n: int):
        self._ck(self.L.ps_cuda_kv_advance(self.h, n))

    # KVCacheInterface slot operations (speculative decode, kv_cache.hpp:120-143)
    def kv_copy_slot(self, dst_cache_index: int, src_token_index: int):
        self._ck(self.L.ps_cuda_kv_copy_slot(self.h, dst_cache_index, src_token_index))

    def kv_move_slot(self, dst_cache_index: int, src_cache_index: int):
        self._ck(self.L.ps_cuda_kv_move_slot(self.h, dst_cache_index, src_cache_index))

    def kv_mask_slot(self, cache_index: int):
        self._ck(self.L.ps_cuda_kv_mask_slot(self.h, cache_index))

    def kv_unmask_slot(self, cache_index: int):
        self._ck(self.L.ps_cuda_kv_unmask_slot(self.h, cache_index))

    def kv_k(self, layer: int) -> int:
        return self.L.ps_cuda_kv_k(self.h, layer)

    def kv_v(self, layer: int) -> int:
        return self.L.ps_cuda_kv_v(self.h, layer)

    def logits_dev(self) -> int:
        return self.L.ps_cuda_logits_dev(self.h)

    @property
    def kv_position(self) -> int:
        return self.L.ps_cuda_kv_position(self.h)

    def reset_kv(self):
        self._ck(self.L.ps_cuda_kv_reset(self.h))

    def kv_rollback(self, n: int):
        self._ck(self.L.ps_cuda_kv_rollback(self.h, n))

    def kv_truncate(self, n: int):
        self._ck(self.L.ps_cuda_kv_truncate(self.h, n))


def desc_from_model_json(cfg: dict, max_batch: int = 128, n_ctx: Optional[int] = None, qkv_bias: bool = False,
                         tp_rank: int = 0, tp_size: int = 1) -> ModelDesc:
    llm, rope = cfg["llm_config"], cfg["llm_config"]["rope_config"]
    return ModelDesc(llm["embed_dim"], llm["ffn_dim"], llm["n_layers"], llm["n_attn_heads"], llm["n_attn_kv_heads"], llm["head_size"],
                     llm["vocab_size"], n_ctx or llm["n_ctx"], llm["norm_eps"], rope["rope_dim"], rope["rope_type"],
                     rope["rope_freq_base"], rope["rope_freq_scale"], rope["rope_attn_factor"], int(qkv_bias), max_batch, tp_rank, tp_size)


def tp_unique_id() -> bytes:
    """Rank 0 of a tensor-parallel group: the 128-byte NCCL id every rank passes to CudaModel(nccl_id=...)."""
    buf = C.create_string_buffer(128)
    rc = load_library().ps_cuda_tp_unique_id(buf)
    if rc != 0:
        raise PsCudaError(f"ps_cuda_tp_unique_id failed ({rc}): NCCL not available")
    return buf.raw


class CudaModel:
    """A PowerServe model directory (model.json + ggml/weights.gguf) bound to the CUDA backend.
    forward / decode follow LlamaModel::forward / decode (src/model/llama/llama_model.cpp:52-132)."""

    def __init__(self, path: Optional[str] = None, *, desc: Optional[ModelDesc] = None,
                 tensors: Optional[Dict[str, gguf.GGUFTensor]] = None, max_batch: int = 128, device: int = 0,
                 tp_rank: int = 0, tp_size: int = 1, nccl_id: Optional[bytes] = None):
        if path is not None:
            cfg = json.load(open(os.path.join(path, "model.json")))
            self._gguf = gguf.GGUFFile(os.path.join(path, "ggml", "weights.gguf"))
            tensors = self._gguf.tensors
            desc = desc_from_model_json(cfg, max_batch=max_batch, qkv_bias="blk.0.attn_q.bias" in tensors, tp_rank=tp_rank, tp_size=tp_size)
        assert desc is not None and tensors is not None
        self.desc, self.tensors = desc, tensors
        self.vocab = desc.vocab_size
        self.be = CudaBackend(desc, device)
        L = self.be.L
        if desc.tp_size > 1:
            assert nccl_id is not None and len(nccl_id) == 128, "tensor parallel contexts need the group's NCCL id (capi.tp_unique_id on rank 0)"
            self.be._ck(L.ps_cuda_tp_init(self.be.h, nccl_id))

        def T(name: Optional[str]) -> Tensor:
            if name is None:
                return Tensor(None, 0, 0)
            t = tensors[name]
            return Tensor(t.host_ptr, t.ggml_type, 0)

        bias = bool(desc.qkv_bias)
        self._layers = (LayerWeights * desc.n_layers)()
        for i in range(desc.n_layers):
            p = f"blk.{i}."
            self._layers[i] = LayerWeights(T(p + "attn_norm.weight"), T(p + "ffn_norm.weight"), T(p + "attn_q.weight"),
                                           T(p + "attn_k.weight"), T(p + "attn_v.weight"), T(p + "attn_output.weight"),
                                           T(p + "ffn_gate.weight"), T(p + "ffn_up.weight"), T(p + "ffn_down.weight"),
                                           T(p + "attn_q.bias" if bias else None), T(p + "attn_k.bias" if bias else None),
                                           T(p + "attn_v.bias" if bias else None))
        out = "output.weight" if "output.weight" in tensors else "token_embd.weight"  # weights.hpp:67
        self._w = ModelWeights(T("token_embd.weight"), T("output_norm.weight"), T(out), self._layers)
        self.be._ck(L.ps_cuda_bind_model(self.be.h, C.byref(self._w)))

    def tp_export(self) -> bytes:
        """CUDA-IPC handle (64 bytes) of this rank's exchange heap; gather them in rank order and call tp_import."""
        buf = C.create_string_buffer(64)
        self.be._ck(self.be.L.ps_cuda_tp_export(self.be.h, buf))
        return buf.raw

    def tp_import(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.desc.tp_size
        self.be._ck(self.be.L.ps_cuda_tp_import(self.be.h, blob, self.desc.tp_size))

    @property
    def position(self) -> int:
        return self.be.kv_position

    def reset(self):
        self.be.reset_kv()

    def forward(self, tokens, pos=None, lm_head: bool = True) -> Optional[np.ndarray]:
        t = _i32(tokens)
        bs = len(t)
        p = _i32(pos) if pos is not None else np.arange(self.position, self.position + bs, dtype=np.int32)
        logits = np.empty((bs, self.vocab), np.float32) if lm_head else None
        self.be._ck(self.be.L.ps_cuda_forward(self.be.h, t.ctypes.data_as(C.POINTER(C.c_int32)), p.ctypes.data_as(C.POINTER(C.c_int32)),
                                              bs, int(lm_head), logits.ctypes.data if lm_head else None))
        return logits

    def forward_tree(self, tokens, pos, tree_mask=None, lm_head: bool = True) -> Optional[np.ndarray]:
        """The speculative path's forward (ps_cuda_forward_tree): arbitrary positions, in-batch tree mask [bs][bs] (None = causal)."""
        t, p = _i32(tokens), _i32(pos)
        bs = len(t)
        m = None if tree_mask is None else np.ascontiguousarray(np.asarray(tree_mask, dtype=np.uint8).reshape(bs, bs))
        logits = np.empty((bs, self.vocab), np.float32) if lm_head else None
        self.be._ck(self.be.L.ps_cuda_forward_tree(self.be.h, t.ctypes.data_as(C.POINTER(C.c_int32)), p.ctypes.data_as(C.POINTER(C.c_int32)), bs,
                                                   m.ctypes.data if m is not None else None, int(lm_head), logits.ctypes.data if lm_head else None))
        return logits

    def decode_greedy(self, first_token: int, n_steps: int) -> np.ndarray:
        ids = np.zeros(n_steps, np.int32)
        self.be._ck(self.be.L.ps_cuda_decode_greedy(self.be.h, int(first_token), n_steps, ids.ctypes.data_as(C.POINTER(C.c_int32))))
        return ids

    # ---- sessions (server-side batching): independent KV sets over this model's weights
    def session_create(self) -> int:
        sid = C.c_int(0)
        self.be._ck(self.be.L.ps_cuda_session_create(self.be.h, C.byref(sid)))
        return sid.value

    def session_destroy(self, sid: int):
        self.be._ck(self.be.L.ps_cuda_session_destroy(self.be.h, sid))

    def session_select(self, sid: int):
        self.be._ck(self.be.L.ps_cuda_session_select(self.be.h, sid))

    def session_position(self, sid: int) -> int:
        return self.be.L.ps_cuda_session_position(self.be.h, sid)

    def forward_sessions(self, session_ids, tokens, lm_head: bool = True, want_logits: bool = True):
        """one token of each session in ONE forward pass -> (logits [n][vocab] or None, device arg-max ids [n])"""
        sids, toks = _i32(session_ids), _i32(tokens)
        n = len(sids)
        logits = np.empty((n, self.desc.vocab_size), np.float32) if (lm_head and want_logits) else None
        ids = np.zeros(n, np.int32)
        self.be._ck(self.be.L.ps_cuda_forward_sessions(self.be.h, sids.ctypes.data_as(C.POINTER(C.c_int32)), toks.ctypes.data_as(C.POINTER(C.c_int32)), n,
                                                       1 if lm_head else 0, logits.ctypes.data_as(C.POINTER(C.c_float)) if logits is not None else None,
                                                       ids.ctypes.data_as(C.POINTER(C.c_int32)) if lm_head else None))
        return logits, ids

    def forward_lazy(self, tokens, pos=None):
        """forward with lm_head whose logits stay on the device (read them with sample_topk / be.logits_dev)"""
        t = _i32(tokens)
        bs = len(t)
        p = _i32(pos) if pos is not None else np.arange(self.position, self.position + bs, dtype=np.int32)
        self.be._ck(self.be.L.ps_cuda_forward(self.be.h, t.ctypes.data_as(C.POINTER(C.c_int32)), p.ctypes.data_as(C.POINTER(C.c_int32)), bs, 1, None))

    def sample_topk(self, k: int, row: int = 0):
        """TopKSampler on the device: (logits [k] descending, token ids [k]) of row `row` of the last forward pass"""
        vals, ids = np.empty(k, np.float32), np.empty(k, np.int32)
        self.be._ck(self.be.L.ps_cuda_sample_topk(self.be.h, row, k, vals.ctypes.data_as(C.POINTER(C.c_float)), ids.ctypes.data_as(C.POINTER(C.c_int32))))
        return vals, ids

    def prefill(self, prompt, batch_size: int = 128):
        """ModelTokenIterator's prefill loop (src/model/model.hpp:147-160): prompt[:-1] in chunks, lm_head = false."""
        prompt = list(map(int, prompt))
        i = 0
        while i < len(prompt) - 1:
            bs = min(batch_size, len(prompt) - 1 - i)
            self.forward(prompt[i:i + bs], lm_head=False)
            i += bs

    def generate(self, prompt, n_decode: int, batch_size: int = 128, forced=None):
        self.reset()
        self.prefill(prompt, batch_size)
        ids, logits, tok = [], [], int(prompt[-1])
        for step in range(n_decode):
            lg = self.forward([tok])[0]
            best = int(np.argmax(lg))
            ids.append(best)
            logits.append(lg)
            tok = int(forced[step]) if forced is not None and step < len(forced) else best
        return ids, np.stack(logits)

    def close(self):
        self.be.close()


# ---------------------------------------------------------------------------------------------------------------- speculative
SPEC_LIB_PATH = os.path.join(HERE, "libps_spec.so")


class SpecConfig(C.Structure):
    """ps_spec_config == SpeculativeConfig (src/speculative/speculative_config.hpp:21-36)"""
    _fields_ = [("draft_batch_size", C.c_int32), ("top_k", C.c_int32), ("temperature", C.c_float), ("p_base", C.c_float), ("max_fan_out", C.c_int32),
                ("min_prob", C.c_float), ("early_stop", C.c_int32), ("n_stop", C.c_int32), ("stop_tokens", C.c_int32 * 8)]


class SpecStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_draft_times", "n_draft_tokens", "n_accepted_tokens", "n_iterations", "n_generated_tokens")] + \
               [(n, C.c_double) for n in ("prefill_s", "draft_s", "verify_s", "total_s")]


_spec_lib: Optional[C.CDLL] = None


def load_spec_library() -> C.CDLL:
    global _spec_lib
    if _spec_lib is None:
        load_library()   # libps_spec.so links libps_cuda.so (rpath $ORIGIN)
        if not os.path.exists(SPEC_LIB_PATH):
            raise PsCudaError(f"{SPEC_LIB_PATH} is missing - build it with `python -m powerserve_b200.build`")
        L = C.CDLL(SPEC_LIB_PATH)
        vp, ci = C.c_void_p, C.c_int
        L.ps_spec_default_config.argtypes = [C.POINTER(SpecConfig)]
        L.ps_spec_default_config.restype = None
        L.ps_spec_create.argtypes = [C.POINTER(vp), vp, vp, C.POINTER(SpecConfig)]
        L.ps_spec_destroy.argtypes = [vp]
        L.ps_spec_destroy.restype = None
        L.ps_spec_set_vocab.argtypes = [vp, ci]
        L.ps_spec_generate.argtypes = [vp, C.POINTER(C.c_int32), ci, ci, ci, C.POINTER(C.c_int32), C.POINTER(SpecStats)]
        L.ps_spec_last_error.argtypes = [vp]
        L.ps_spec_last_error.restype = C.c_char_p
        _spec_lib = L
    return _spec_lib


class SpecDecoder:
    """Token-tree speculative decoding over a target and a draft CudaModel (include/ps_spec.h)."""

    def __init__(self, target: CudaModel, draft: CudaModel, **overrides):
        assert target.vocab == draft.vocab, "target and draft must share the vocabulary"
        self.L = load_spec_library()
        self.cfg = SpecConfig()
        self.L.ps_spec_default_config(C.byref(self.cfg))
        for k, v in overrides.items():
            setattr(self.cfg, k, v)
        self.h = C.c_void_p()
        rc = self.L.ps_spec_create(C.byref(self.h), target.be.h, draft.be.h, C.byref(self.cfg))
        if rc:
            raise PsCudaError(f"ps_spec_create failed ({rc})")
        self.L.ps_spec_set_vocab(self.h, target.vocab)
        self.target, self.draft = target, draft

    def generate(self, prompt, n_tokens: int, prefill_batch: int = 128):
        p = _i32(prompt)
        out = np.zeros(n_tokens, np.int32)
        st = SpecStats()
        rc = self.L.ps_spec_generate(self.h, p.ctypes.data_as(C.POINTER(C.c_int32)), len(p), n_tokens, prefill_batch,
                                     out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st))
        if rc:
            raise PsCudaError(f"ps_spec_generate failed ({rc}): {self.L.ps_spec_last_error(self.h).decode()}")
        return out, {n: getattr(st, n) for n, _ in SpecStats._fields_}

    def close(self):
        if self.h:
            self.L.ps_spec_destroy(self.h)
            self.h = None
