"""Small driver for ncu captures: an N-layer slice of a BASELINE model shape (real per-layer shapes, small vocab so the
profiler's memory save/restore stays cheap), `ctx` tokens prefetched into the KV cache, then a few fused decode steps.

    python tools/prof_decode.py [model] [n_layers] [ctx] [steps]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b"
shape = synth.PRESETS[model]
shape.n_layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 64
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
shape.vocab_size = 4096
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=128, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
m.prefill(synth.random_prompt(shape.vocab_size, ctx + 1), 128)
ids = m.decode_greedy(1, steps)
print("ids", list(ids), "device ms/step", m.be.counter("last_device_ns") / 1e6 / steps)
m.close()
