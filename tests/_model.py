"""Test helpers: load a synthetic model dir into the ORACLE (ctypes structs over GGUF bytes) and run the compiled
reference driver.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile
from typing import List, Optional, Tuple

import numpy as np

from powerserve_b200 import gguf, synth
from tests import _libs as L

_CACHE = os.environ.get("PS_TEST_MODEL_CACHE", os.path.join(tempfile.gettempdir(), "ps_b200_test_models"))


def model_dir(preset: str, seed: int = 0) -> str:
    d = os.path.join(_CACHE, f"{preset}-s{seed}")
    if not os.path.exists(os.path.join(d, "ggml", "weights.gguf")):
        synth.write_model_dir(d, synth.PRESETS[preset], seed=seed)
    return d


class OracleModel:
    def __init__(self, path: str):
        self.path = path
        self.cfg = json.load(open(os.path.join(path, "model.json")))
        llm, rope = self.cfg["llm_config"], self.cfg["llm_config"]["rope_config"]
        self.g = gguf.GGUFFile(os.path.join(path, "ggml", "weights.gguf"))
        self.lib = L.oracle()
        self.vocab = llm["vocab_size"]
        self.dim = llm["embed_dim"]
        self.n_layers = llm["n_layers"]
        bias = "blk.0.attn_q.bias" in self.g
        c = L.OrConfig(llm["embed_dim"], llm["ffn_dim"], llm["n_layers"], llm["n_attn_heads"], llm["n_attn_kv_heads"],
                       llm["head_size"], llm["vocab_size"], llm["n_ctx"], llm["norm_eps"], rope["rope_dim"],
                       rope["rope_type"], rope["rope_freq_base"], rope["rope_freq_scale"], rope["rope_attn_factor"], int(bias))

        def T(name: Optional[str]) -> L.OrTensor:
            if name is None:
                return L.OrTensor(None, 0, 0)
            t = self.g[name]
            return L.OrTensor(t.host_ptr, t.ggml_type, 0)

        self._layers = (L.OrLayer * llm["n_layers"])()
        for i in range(llm["n_layers"]):
            p = f"blk.{i}."
            self._layers[i] = L.OrLayer(T(p + "attn_norm.weight"), T(p + "ffn_norm.weight"), T(p + "attn_q.weight"),
                                        T(p + "attn_k.weight"), T(p + "attn_v.weight"), T(p + "attn_output.weight"),
                                        T(p + "ffn_gate.weight"), T(p + "ffn_up.weight"), T(p + "ffn_down.weight"),
                                        T(p + "attn_q.bias" if bias else None), T(p + "attn_k.bias" if bias else None),
                                        T(p + "attn_v.bias" if bias else None))
        out = "output.weight" if "output.weight" in self.g else "token_embd.weight"   # weights.hpp:67
        self._w = L.OrWeights(T("token_embd.weight"), T("output_norm.weight"), T(out), self._layers)
        self.h = self.lib.ps_or_model_create(C.byref(c), C.byref(self._w))

    def reset(self):
        self.lib.ps_or_model_reset(self.h)

    @property
    def position(self) -> int:
        return self.lib.ps_or_model_position(self.h)

    def forward(self, tokens, pos=None, lm_head=True) -> Optional[np.ndarray]:
        tokens = np.ascontiguousarray(tokens, np.int32)
        bs = len(tokens)
        if pos is None:
            pos = np.arange(self.position, self.position + bs, dtype=np.int32)
        pos = np.ascontiguousarray(pos, np.int32)
        logits = np.zeros((bs, self.vocab), np.float32) if lm_head else np.zeros((1, 1), np.float32)
        rc = self.lib.ps_or_model_forward(self.h, L.iptr(tokens), L.iptr(pos), bs, int(lm_head), L.fptr(logits))
        assert rc == 0
        return logits if lm_head else None

    def tap(self, layer: int, which: int, bs: int) -> np.ndarray:
        out = np.zeros((bs, self.dim), np.float32)
        n = self.lib.ps_or_model_tap(self.h, layer, which, L.fptr(out))
        assert n == out.size
        return out

    def generate(self, prompt, n_decode: int, batch_size: int = 128, forced=None) -> Tuple[List[int], np.ndarray]:
        """Same loop as the reference (model.hpp:141-183): prefill prompt[:-1] in chunks, then greedy decode."""
        self.reset()
        prompt = list(map(int, prompt))
        i = 0
        while i < len(prompt) - 1:
            bs = min(batch_size, len(prompt) - 1 - i)
            self.forward(prompt[i:i + bs], lm_head=False)
            i += bs
        ids, logits, tok = [], [], prompt[-1]
        for step in range(n_decode):
            lg = self.forward([tok])[0]
            best = int(np.argmax(lg))   # first max wins
            ids.append(best)
            logits.append(lg)
            tok = int(forced[step]) if forced is not None and step < len(forced) else best
        return ids, np.stack(logits)

    def close(self):
        if self.h:
            self.lib.ps_or_model_free(self.h)
            self.h = None


def run_reference(path: str, prompt, n_decode: int, batch_size: int = 128, n_threads: int = 4, dump_logits: int = 0,
                  forced=None, timeout: int = 3600):
    """Run oracle/_ref/ps_ref_run (the reference's own model stack). Returns (ids, logits[n,vocab] or None, timings)."""
    exe = os.path.join(L.REF_DIR, "ps_ref_run")
    with tempfile.TemporaryDirectory() as td:
        pf = os.path.join(td, "prompt.txt")
        open(pf, "w").write(" ".join(str(int(t)) for t in prompt))
        cmd = [exe, path, str(n_threads), str(batch_size), pf, str(n_decode), os.path.join(td, "out")]
        if dump_logits:
            cmd += ["--dump-logits", str(dump_logits)]
        if forced is not None:
            ff = os.path.join(td, "forced.txt")
            open(ff, "w").write(" ".join(str(int(t)) for t in forced))
            cmd += ["--force", ff]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"ps_ref_run failed: {r.stderr[-2000:]}")
        timings = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        ids = [int(x) for x in open(os.path.join(td, "out.ids")).read().split()]
        logits = None
        if dump_logits:
            vocab = json.load(open(os.path.join(path, "model.json")))["llm_config"]["vocab_size"]
            logits = np.fromfile(os.path.join(td, "out.logits"), dtype=np.float32).reshape(-1, vocab)
        return ids, logits, timings
