"""Test support: tensor-parallel sharding plan (host-side mirror of what `ps_cuda_bind_model` does for a context with tp_size > 1;
used by the gloo tests of the sharding logic, not by the product).

Every matrix is ROW-sharded — rank r owns rows [r * rows / N, (r + 1) * rows / N) of W{K, rows} — so each output element
is still one full-K dot product and the sharded model is bit-identical to the unsharded one; the sharded outputs are
exchanged with all-gathers (4 per layer: attention output, x after Wo, FFN hidden, x after Wdown; + arg-max partials or
logits per token) instead of the K-split all-reduces of Megatron-style sharding, which would change the fp32 summation
order.  q/k/v rows are whole heads, so RoPE, the KV cache and attention are local to a rank.
"""
from __future__ import annotations

from typing import Dict, Tuple

from powerserve_b200 import gguf

ROW_SHARDED = ("attn_q.weight", "attn_k.weight", "attn_v.weight", "attn_output.weight", "ffn_gate.weight", "ffn_up.weight",
               "ffn_down.weight", "attn_q.bias", "attn_k.bias", "attn_v.bias")


def validate(n_heads: int, n_kv_heads: int, ffn: int, vocab: int, dim: int, size: int) -> None:
    if n_heads % size or n_kv_heads % size or ffn % (8 * size) or vocab % (8 * size) or dim % (8 * size):
        raise ValueError(f"tensor parallel size {size} does not divide heads {n_heads}/{n_kv_heads}, ffn {ffn}, vocab {vocab} or dim {dim} into octets")


def row_range(rows: int, rank: int, size: int) -> Tuple[int, int]:
    assert rows % size == 0
    return rank * rows // size, (rank + 1) * rows // size


def shard_tensor(name: str, t: gguf.GGUFTensor, rank: int, size: int, lm_head: bool = False):
    """(byte offset, n_rows) of rank's shard inside tensor `t` (ggml order: shape[0] = K contiguous, shape[1] = rows)."""
    sharded = lm_head or any(name.endswith(s) for s in ROW_SHARDED)
    if name.endswith(".bias"):
        rows = t.shape[0]
        r0, r1 = row_range(rows, rank, size) if sharded else (0, rows)
        return r0 * 4, r1 - r0
    rows = t.shape[1] if len(t.shape) > 1 else 1
    if not sharded or size == 1:
        return 0, rows
    r0, r1 = row_range(rows, rank, size)
    return r0 * gguf.tensor_bytes(t.ggml_type, (t.shape[0], 1)), r1 - r0


def gathers_per_token(n_layers: int, pick: bool = True) -> int:
    """exchanges of one decoded token: four per layer + one per token (the arg-max partials, or the logits)"""
    return 4 * n_layers + 1
