"""CPU simulation of the tensor-parallel decode step: each rank runs the ORACLE's operators on its row shards and the
shards are exchanged with an `all_gather` callback (torch.distributed gloo in tests/test_tp_gloo.py).  Test
infrastructure: it mirrors decode_step_fused of powerserve_b200/csrc/ps_cuda.cu phase by phase."""
from __future__ import annotations

import json
import os

import numpy as np

from powerserve_b200 import gguf
from tests import _tp_plan as tp
from tests import _libs as L


class InbandExchange:
    """CPU model of the in-band-flag exchange of the CUDA path (ps_tp_ll_store / ps_tp_ll_done / ps_rw_load8_ll in
    powerserve_b200/csrc): every rank stores its shard into EVERY rank's mirror of the gathered vector as naturally
    aligned 64-bit words {value bits, epoch}; nobody fences or signals; the consumer polls the words it is about to read
    until all of them carry the epoch of this phase.  One epoch counter per phase (`slot`), local, advanced once per
    exchange, in lock step on every rank; x-after-Wo and x-after-Wdown use separate mirrors because their epochs are
    equal numbers.  `mirrors[slot][p]` is rank p's mirror (np.uint64 view of memory shared by all ranks)."""
    SLOTS = ("att", "x1", "h", "x2", "logits")

    def __init__(self, rank: int, size: int, mirrors):
        self.rank, self.size, self.mirrors = rank, size, mirrors
        self.epoch = {s: 0 for s in self.SLOTS}
        self.polls = 0

    def __call__(self, a, slot):
        import time
        a32 = np.ascontiguousarray(a, np.float32)
        e = np.uint64(self.epoch[slot] + 1)
        words = (e << np.uint64(32)) | a32.view(np.uint32).astype(np.uint64)
        off = self.rank * a32.size
        for p in range(self.size):                                   # peer stores, own rank included
            self.mirrors[slot][p][off:off + a32.size] = words
        self.epoch[slot] = int(e)
        mine, n, t0 = self.mirrors[slot][self.rank], a32.size * self.size, time.time()
        while True:                                                  # the consumer's poll
            w = mine[:n].copy()
            self.polls += 1
            if np.all((w >> np.uint64(32)) == e):
                return (w & np.uint64(0xffffffff)).astype(np.uint32).view(np.float32)
            if time.time() - t0 > 120:
                raise RuntimeError(f"in-band flag wait gave up (slot {slot}, epoch {int(e)})")


class ShardedOracle:
    def __init__(self, path: str, rank: int, size: int, all_gather):
        self.o = L.oracle()
        self.rank, self.size, self.all_gather = rank, size, all_gather
        cfg = json.load(open(os.path.join(path, "model.json")))["llm_config"]
        self.cfg, self.rope = cfg, cfg["rope_config"]
        self.g = gguf.GGUFFile(os.path.join(path, "ggml", "weights.gguf"))
        self.dim, self.ffn, self.nl = cfg["embed_dim"], cfg["ffn_dim"], cfg["n_layers"]
        self.nh, self.nkv, self.hs, self.vocab, self.n_ctx = cfg["n_attn_heads"], cfg["n_attn_kv_heads"], cfg["head_size"], cfg["vocab_size"], cfg["n_ctx"]
        tp.validate(self.nh, self.nkv, self.ffn, self.vocab, self.dim, size)
        self.nh_l, self.nkv_l = self.nh // size, self.nkv // size
        kvd_l = self.nkv_l * self.hs
        self.kc = [np.zeros((self.n_ctx, kvd_l), np.float32) for _ in range(self.nl)]
        self.vct = [np.zeros((kvd_l, self.n_ctx), np.float32) for _ in range(self.nl)]
        self.pos = 0
        self.out_name = "output.weight" if "output.weight" in self.g else "token_embd.weight"

    def _w(self, name, lm_head=False):
        """(uint8 view of rank's row shard, ggml type, K, rows)"""
        t = self.g[name]
        off, rows = tp.shard_tensor(name, t, self.rank, self.size, lm_head)
        k = t.shape[0]
        nbytes = gguf.tensor_bytes(t.ggml_type, (k, rows))
        return np.ascontiguousarray(t.data[off:off + nbytes]), t.ggml_type, k, rows

    def _f32(self, name):
        return np.ascontiguousarray(self.g[name].data.view(np.float32))

    def _matmul(self, name, x, lm_head=False):
        w, t, k, rows = self._w(name, lm_head)
        out = np.zeros(rows, np.float32)
        self.o.ps_or_matmul(t, L.vptr(w), k, rows, L.fptr(np.ascontiguousarray(x)), 1, L.fptr(out))
        return out

    def _bias(self, y, name):
        t = self.g[name]
        off, rows = tp.shard_tensor(name, t, self.rank, self.size)
        b = np.ascontiguousarray(t.data[off:off + 4 * rows].view(np.float32))
        out = np.zeros_like(y)
        self.o.ps_or_add(L.fptr(out), L.fptr(y), L.fptr(b), rows, rows)
        return out

    def _rmsnorm(self, x, wname):
        out = np.zeros_like(x)
        self.o.ps_or_rmsnorm(L.fptr(out), L.fptr(x), L.fptr(self._f32(wname)), self.dim, 1, self.cfg["norm_eps"])
        return out

    def _rope(self, v, n_heads):
        out = np.zeros_like(v)
        pos = np.asarray([self.pos], np.int32)
        self.o.ps_or_rope(L.fptr(out), L.fptr(np.ascontiguousarray(v)), self.hs, n_heads, 1, L.iptr(pos), self.rope["rope_dim"], self.rope["rope_type"],
                          self.rope["rope_freq_base"], self.rope["rope_freq_scale"], self.rope["rope_attn_factor"])
        return out

    def step(self, token: int, lm_head: bool = True):
        o, hs, r = self.o, self.hs, self.rank
        x = np.zeros(self.dim, np.float32)
        emb = self.g["token_embd.weight"]
        o.ps_or_get_embedding(L.fptr(x), L.vptr(np.ascontiguousarray(emb.data)), emb.ggml_type, self.dim, L.iptr(np.asarray([token], np.int32)), 1)
        n_kv = self.pos + 1
        for l in range(self.nl):
            p = f"blk.{l}."
            xn = self._rmsnorm(x, p + "attn_norm.weight")
            q, k, v = self._matmul(p + "attn_q.weight", xn), self._matmul(p + "attn_k.weight", xn), self._matmul(p + "attn_v.weight", xn)
            if p + "attn_q.bias" in self.g:  # Qwen2: row-sharded F32 bias, added before RoPE (qwen2_model / norm_attention)
                q, k, v = self._bias(q, p + "attn_q.bias"), self._bias(k, p + "attn_k.bias"), self._bias(v, p + "attn_v.bias")
            q, k = self._rope(q, self.nh_l), self._rope(k, self.nkv_l)
            self.kc[l][self.pos] = k
            self.vct[l][:, self.pos] = v
            kq = np.zeros((self.nh_l, 1, n_kv), np.float32)
            o.ps_or_attn_scores(L.fptr(kq), L.fptr(self.kc[l]), L.fptr(q), hs, self.nh_l, self.nkv_l, n_kv, 1)
            mask = np.zeros((1, n_kv), np.float32)
            o.ps_or_get_mask(L.fptr(mask), n_kv, 1, L.iptr(np.asarray([self.pos], np.int32)))
            pr = np.zeros_like(kq)
            o.ps_or_softmax_ext(L.fptr(pr), L.fptr(kq), L.fptr(mask), n_kv, 1, self.nh_l, 1.0 / np.sqrt(np.float32(hs)))
            att = np.zeros(self.nh_l * hs, np.float32)
            o.ps_or_attn_pv(L.fptr(att), L.fptr(self.vct[l]), L.fptr(pr), hs, self.nh_l, self.nkv_l, n_kv, self.n_ctx, 1)
            att_full = self.all_gather(att, "att")                                  # gather 1: attention output
            d0, d1 = tp.row_range(self.dim, r, self.size)
            xs = np.zeros(d1 - d0, np.float32)
            wo = self._matmul(p + "attn_output.weight", att_full)
            o.ps_or_add(L.fptr(xs), L.fptr(np.ascontiguousarray(x[d0:d1])), L.fptr(wo), d1 - d0, d1 - d0)
            x = self.all_gather(xs, "x1")                                           # gather 2: x after Wo
            xn = self._rmsnorm(x, p + "ffn_norm.weight")
            g, u = self._matmul(p + "ffn_gate.weight", xn), self._matmul(p + "ffn_up.weight", xn)
            h = np.zeros_like(g)
            o.ps_or_silu_hadamard(L.fptr(h), L.fptr(g), L.fptr(u), g.size)
            h_full = self.all_gather(h, "h")                                        # gather 3: FFN hidden
            dn = self._matmul(p + "ffn_down.weight", h_full)
            o.ps_or_add(L.fptr(xs), L.fptr(np.ascontiguousarray(x[d0:d1])), L.fptr(dn), d1 - d0, d1 - d0)
            x = self.all_gather(xs, "x2")                                           # gather 4: x after Wdown
        self.pos += 1
        if not lm_head:
            return None
        xn = self._rmsnorm(x, "output_norm.weight")
        return self.all_gather(self._matmul(self.out_name, xn, lm_head=True), "logits")  # logits (or arg-max partials on the GPU)
