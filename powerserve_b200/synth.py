"""Synthetic PowerServe model directories (there are no real GGUF files and no network on the build / GPU boxes).

A model directory is what `powerserve-run --work-folder` expects (/root/reference/README.md:118-151,
src/core/config.cpp:68-120): `<dir>/model.json` + `<dir>/ggml/weights.gguf`.  Tensor names follow
src/model/llama/llama_weight.hpp:24-34, qwen2_weight.hpp:24-37 and common/weights.hpp:64-70 (a missing
`output.weight` means tied embeddings).

Weights are drawn DIRECTLY in the quantised block format (random 4-/6-/8-bit quants, random 6-bit sub-block
scales, fp16 super-block scales sized so that the de-quantised weights have std ~ gain/sqrt(K)); this is orders of
magnitude faster than quantising fp32 Gaussians (an 8B model is 4.7 GB of blocks) and every byte pattern that can
occur in a real file can occur here.  Block layouts: libs/ggml/src/ggml-common.h:158-162 (Q4_0), :200-204 (Q8_0),
:299-310 (Q4_K), :335-340 (Q6_K).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import gguf
from .gguf import GGML_F32, GGML_Q4_0, GGML_Q8_0, GGML_Q4_K, GGML_Q5_K, GGML_Q6_K


@dataclass
class ModelShape:
    name: str
    arch: str              # "llama" | "qwen2"
    dim: int
    ffn_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    head_size: int
    vocab_size: int
    tied: bool
    rope_type: int         # 0 = NORM (adjacent pairs), 2 = NEOX   (ggml.c:15440; tools/gguf_config_to_json/main.cpp:61-69)
    rope_freq_base: float
    norm_eps: float
    wtype: int             # ggml type of every matmul weight
    n_ctx: int = 4096      # capped: the CPU reference reserves n_ctx fp32 KV per layer (ggml_kv_cache.cpp:35-58)
    qkv_bias: bool = False
    embd_type: Optional[int] = None    # token_embd type (defaults to wtype)
    output_type: Optional[int] = None  # output.weight type (defaults to wtype)

    @property
    def kv_dim(self) -> int:
        return self.n_kv_heads * self.head_size


PRESETS: Dict[str, ModelShape] = {
    # BASELINE.json configs
    "qwen2-0.5b": ModelShape("qwen2-0.5b", "qwen2", 896, 4864, 24, 14, 2, 64, 151936, True, 2, 1e6, 1e-6, GGML_Q4_0, qkv_bias=True),
    "llama-3.2-1b": ModelShape("llama-3.2-1b", "llama", 2048, 8192, 16, 32, 8, 64, 128256, True, 0, 5e5, 1e-5, GGML_Q4_K),
    "llama-3.1-8b": ModelShape("llama-3.1-8b", "llama", 4096, 14336, 32, 32, 8, 128, 128256, False, 0, 5e5, 1e-5, GGML_Q4_K),
    # small shapes for parity tests (same structure, seconds on the CPU oracle)
    "tiny-llama": ModelShape("tiny-llama", "llama", 512, 1536, 2, 8, 2, 64, 1024, False, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=512),
    "tiny-llama-hs128": ModelShape("tiny-llama-hs128", "llama", 512, 1024, 2, 4, 2, 128, 768, True, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=384),
    # head sizes off the templated fast paths of the attention kernels (32 / 96 / 256: generic step loops, round-1 batch kernels)
    "tiny-hs32": ModelShape("tiny-hs32", "llama", 512, 1024, 2, 16, 4, 32, 512, True, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=256),
    "tiny-hs96": ModelShape("tiny-hs96", "llama", 512, 1024, 2, 8, 2, 96, 512, True, 2, 5e5, 1e-5, GGML_Q4_K, n_ctx=256),
    "tiny-hs256": ModelShape("tiny-hs256", "llama", 512, 1024, 2, 2, 1, 256, 512, True, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=256),
    "tiny-qwen2": ModelShape("tiny-qwen2", "qwen2", 256, 608, 2, 4, 2, 64, 512, True, 2, 1e6, 1e-6, GGML_Q4_0, n_ctx=256, qkv_bias=True),
    # query heads per kv head that are not a power of two (Qwen2-0.5B has 14 / 2 = 7): the decode attention kernels run
    # with the next power-of-two template and r2 active heads
    "tiny-qwen2-r7": ModelShape("tiny-qwen2-r7", "qwen2", 448, 608, 2, 7, 1, 64, 512, True, 2, 1e6, 1e-6, GGML_Q4_0, n_ctx=256, qkv_bias=True),
    "tiny-q8-r3": ModelShape("tiny-q8-r3", "llama", 384, 512, 2, 6, 2, 64, 512, False, 0, 1e4, 1e-5, GGML_Q8_0, n_ctx=256),
    "tiny-q5k": ModelShape("tiny-q5k", "llama", 512, 1024, 2, 8, 2, 64, 768, True, 0, 5e5, 1e-5, GGML_Q5_K, n_ctx=256),   # real-file coverage
    "tiny-q8": ModelShape("tiny-q8", "llama", 256, 512, 2, 4, 4, 64, 512, False, 0, 1e4, 1e-5, GGML_Q8_0, n_ctx=256),
    # real per-layer shapes of the BASELINE models with few layers / small vocab (exercise nb = 8/32 and 16/56 paths)
    "slice-1b": ModelShape("slice-1b", "llama", 2048, 8192, 2, 32, 8, 64, 2048, True, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=1024),
    "slice-8b": ModelShape("slice-8b", "llama", 4096, 14336, 1, 32, 8, 128, 2048, False, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=1024),
    # the same slices with room for the benchmarked context (2048+) and beyond the shared-memory-resident attention range
    "slice-1b-long": ModelShape("slice-1b-long", "llama", 2048, 8192, 2, 32, 8, 64, 2048, True, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=8192),
    "slice-8b-long": ModelShape("slice-8b-long", "llama", 4096, 14336, 1, 32, 8, 128, 2048, False, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=8192),
    # many row octets per CTA (several rounds per compute warp in the persistent step kernel) / many layers
    "tiny-bigvocab": ModelShape("tiny-bigvocab", "llama", 512, 1536, 2, 8, 2, 64, 40000, False, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=256),
    "tiny-deep": ModelShape("tiny-deep", "llama", 512, 4096 + 2048, 9, 8, 2, 64, 1024, True, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=256),
    "tiny-mixed": ModelShape("tiny-mixed", "llama", 512, 1024, 2, 8, 4, 64, 768, False, 0, 5e5, 1e-5, GGML_Q4_K, n_ctx=256,
                             output_type=GGML_Q6_K),
}


# ----------------------------------------------------------------------------------------------------------------
# random blocks
# ----------------------------------------------------------------------------------------------------------------
def _f16_bytes(x: np.ndarray) -> np.ndarray:
    return x.astype(np.float16).view(np.uint8).reshape(x.shape + (2,))


def random_blocks(rng: np.random.Generator, ggml_type: int, n_rows: int, k: int, std: float) -> np.ndarray:
    """Return uint8 array [n_rows, row_bytes] of random blocks whose de-quantised values have std ~= `std`."""
    blk, nbytes = gguf.TYPE_INFO[ggml_type]
    assert k % blk == 0, (k, blk)
    nb = n_rows * (k // blk)
    # One pass of raw random bytes fills the whole tensor (fast path of the bit generator); the few structured
    # fields (fp16 scales, 6-bit sub-block scales) are then overwritten using bit tricks on those same bytes.
    out = rng.integers(0, 256, size=(nb, nbytes), dtype=np.uint8)
    jitter = 0.6 + 0.8 * (rng.integers(0, 256, size=nb, dtype=np.uint8).astype(np.float32) / 255.0)
    if ggml_type == GGML_Q4_0:
        # w = (q - 8) * d ; q uniform 0..15 -> std(q-8) ~ 4.63
        out[:, 0:2] = _f16_bytes(std / 4.63 * jitter)
    elif ggml_type == GGML_Q8_0:
        # w = q * d ; q uniform int8 (-128 remapped to -127, the quantiser never emits -128) -> std ~ 73.6
        out[:, 0:2] = _f16_bytes(std / 73.6 * jitter)
        q = out[:, 2:]
        q[q == 0x80] = 0x81
    elif ggml_type == GGML_Q4_K:
        # w = d*sc_j*q - dmin*m_j.  sc_j in [16,62], m_j ~ 7.5*sc_j*(d/dmin) with dmin = 8 d so that every sub-block is
        # (nearly) zero-mean; std(w) ~ d * sqrt(E[sc^2]) * 4.61 ~ 190 d.
        d = std / 190.0 * jitter
        r = out[:, 4:12].copy()                                   # 8 random bytes per block drive the 8 (sc, m) pairs
        sc = (16 + (r & 31) + ((r >> 5) & 7) * 2).astype(np.uint8)             # 16 .. 61
        m = np.minimum(63, ((sc.astype(np.uint16) * 15 + 8) >> 4) + ((r >> 3) & 3)).astype(np.uint8) - 1
        out[:, 0:2] = _f16_bytes(d)
        out[:, 2:4] = _f16_bytes(8.0 * d)
        # inverse of get_scale_min_k4 (ggml-quants.c:1912-1919)
        out[:, 4:8] = (sc[:, 0:4] & 63) | ((sc[:, 4:8] >> 4) << 6)
        out[:, 8:12] = (m[:, 0:4] & 63) | ((m[:, 4:8] >> 4) << 6)
        out[:, 12:16] = (sc[:, 4:8] & 0xF) | ((m[:, 4:8] & 0xF) << 4)
    elif ggml_type == GGML_Q5_K:
        # block_q5_K = d, dmin, scales[12], qh[32], qs[128]: w = d*sc_j*q - dmin*m_j with q 5-bit uniform (mean 15.5, std 9.23);
        # m_j ~ 15.5*sc_j*(d/dmin) with dmin = 16 d keeps every sub-block (nearly) zero-mean; std(w) ~ d * 41 * 9.23 ~ 380 d
        d = std / 380.0 * jitter
        r = out[:, 4:12].copy()
        sc = (16 + (r & 31) + ((r >> 5) & 7) * 2).astype(np.uint8)             # 16 .. 61
        m = np.minimum(63, ((sc.astype(np.uint16) * 31 + 16) >> 5) + ((r >> 3) & 3)).astype(np.uint8) - 1
        out[:, 0:2] = _f16_bytes(d)
        out[:, 2:4] = _f16_bytes(16.0 * d)
        out[:, 4:8] = (sc[:, 0:4] & 63) | ((sc[:, 4:8] >> 4) << 6)
        out[:, 8:12] = (m[:, 0:4] & 63) | ((m[:, 4:8] >> 4) << 6)
        out[:, 12:16] = (sc[:, 4:8] & 0xF) | ((m[:, 4:8] & 0xF) << 4)
    elif ggml_type == GGML_Q6_K:
        # w = d * sc_j * (q - 32), q 6-bit uniform -> std(q-32) ~ 18.5 ; |sc| in [24, 87], random sign
        r = out[:, 192:208]
        mag = (24 + (r & 63)).astype(np.int8)
        out[:, 192:208] = np.where(r & 0x80, -mag, mag).astype(np.int8).view(np.uint8)
        out[:, 208:210] = _f16_bytes(std / (18.5 * 58.0) * jitter)
    else:
        raise ValueError(f"no random generator for ggml type {ggml_type}")
    return out.reshape(n_rows, (k // blk) * nbytes)


# ----------------------------------------------------------------------------------------------------------------
# model directories
# ----------------------------------------------------------------------------------------------------------------
def model_json(shape: ModelShape, model_id: Optional[str] = None) -> dict:
    return {
        "model_arch": shape.arch,
        "model_id": model_id or shape.name,
        "version": 1,
        "llm_config": {
            "embed_dim": shape.dim, "ffn_dim": shape.ffn_dim, "n_layers": shape.n_layers,
            "n_attn_heads": shape.n_heads, "n_attn_kv_heads": shape.n_kv_heads, "n_ctx": shape.n_ctx,
            "vocab_size": shape.vocab_size, "kv_dim": shape.kv_dim, "head_size": shape.head_size,
            "norm_eps": shape.norm_eps,
            "rope_config": {"rope_dim": shape.head_size, "n_rope_ctx_orig": shape.n_ctx,
                            "rope_freq_base": shape.rope_freq_base, "rope_freq_scale": 1.0,
                            "rope_attn_factor": 1.0, "rope_type": shape.rope_type},
        },
    }


def tensor_plan(shape: ModelShape) -> List[Tuple[str, int, Tuple[int, ...], float]]:
    """(name, ggml type, ggml-order shape, target std) for every tensor of the model, in file order."""
    s, wt = shape, shape.wtype
    plan: List[Tuple[str, int, Tuple[int, ...], float]] = []
    plan.append(("token_embd.weight", s.embd_type or wt, (s.dim, s.vocab_size), 1.0 if not s.tied else s.dim ** -0.5))
    plan.append(("output_norm.weight", GGML_F32, (s.dim,), 0.0))
    if not s.tied:
        plan.append(("output.weight", s.output_type or wt, (s.dim, s.vocab_size), s.dim ** -0.5))
    for L in range(s.n_layers):
        p = f"blk.{L}."
        plan += [
            (p + "attn_norm.weight", GGML_F32, (s.dim,), 0.0),
            (p + "ffn_norm.weight", GGML_F32, (s.dim,), 0.0),
            (p + "attn_q.weight", wt, (s.dim, s.n_heads * s.head_size), s.dim ** -0.5),
            (p + "attn_k.weight", wt, (s.dim, s.kv_dim), s.dim ** -0.5),
            (p + "attn_v.weight", wt, (s.dim, s.kv_dim), s.dim ** -0.5),
            (p + "attn_output.weight", wt, (s.n_heads * s.head_size, s.dim), s.dim ** -0.5),
            (p + "ffn_gate.weight", wt, (s.dim, s.ffn_dim), s.dim ** -0.5),
            (p + "ffn_up.weight", wt, (s.dim, s.ffn_dim), s.dim ** -0.5),
            (p + "ffn_down.weight", wt, (s.ffn_dim, s.dim), s.ffn_dim ** -0.5),
        ]
        if s.qkv_bias:
            plan += [
                (p + "attn_q.bias", GGML_F32, (s.n_heads * s.head_size,), 0.02),
                (p + "attn_k.bias", GGML_F32, (s.kv_dim,), 0.02),
                (p + "attn_v.bias", GGML_F32, (s.kv_dim,), 0.02),
            ]
    return plan


def _make_tensor(shape_seed):
    seed, idx, name, t, shp, std = shape_seed
    rng = np.random.Generator(np.random.SFC64([seed, idx]))   # one independent stream per tensor
    if t == GGML_F32:
        if name.endswith("norm.weight"):
            data = (1.0 + 0.1 * rng.standard_normal(shp[0])).astype(np.float32)
        else:
            data = (std * rng.standard_normal(shp[0])).astype(np.float32)
        return (name, t, shp, data.view(np.uint8))
    return (name, t, shp, random_blocks(rng, t, shp[1], shp[0], std).reshape(-1))


def generate_tensors(shape: ModelShape, seed: int = 0, workers: Optional[int] = None) -> List[Tuple[str, int, Tuple[int, ...], np.ndarray]]:
    """Materialise every tensor of `shape`.  Deterministic in `seed` (each tensor has its own random stream, so the
    result does not depend on `workers`); numpy releases the GIL in the generators, so threads scale."""
    jobs = [(seed, idx, name, t, shp, std) for idx, (name, t, shp, std) in enumerate(tensor_plan(shape))]
    workers = workers or min(32, os.cpu_count() or 1)
    if workers <= 1 or len(jobs) < 4:
        return [_make_tensor(j) for j in jobs]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(workers) as ex:
        return list(ex.map(_make_tensor, jobs))


def write_model_dir(path: str, shape: ModelShape, seed: int = 0, model_id: Optional[str] = None) -> str:
    os.makedirs(os.path.join(path, "ggml"), exist_ok=True)
    with open(os.path.join(path, "model.json"), "w") as f:
        json.dump(model_json(shape, model_id), f, indent=1)
    gguf.write_gguf(os.path.join(path, "ggml", "weights.gguf"), generate_tensors(shape, seed), arch=shape.arch)
    with open(os.path.join(path, "shape.json"), "w") as f:   # our own side-car (not read by the reference)
        json.dump(asdict(shape), f, indent=1)
    return path


def random_prompt(vocab_size: int, n: int, seed: int = 1234) -> np.ndarray:
    """Token ids uniform in [0, vocab) — SURVEY.md section 8(d) 'synthetic prompts'."""
    return np.random.Generator(np.random.SFC64(seed)).integers(0, vocab_size, size=n, dtype=np.int32)
