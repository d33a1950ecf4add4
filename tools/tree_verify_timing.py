"""Measurement: device time of one ps_cuda_forward_tree verify batch (12 nodes, causal chain) on an N-layer slice of a BASELINE shape.

    python tools/tree_verify_timing.py [model] [n_layers] [ctx]
"""
import sys
import numpy as np
sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

model, n_layers, ctx_len = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
shape = synth.PRESETS[model]
shape.n_layers = n_layers
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=128, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
m.prefill(synth.random_prompt(shape.vocab_size, ctx_len + 1), 128)
base = m.position
import os
for kv in filter(None, os.environ.get("PS_OPTS", "").split(",")):  # e.g. PS_OPTS=tc_min=16,pv_batch_min=2
    m.be.set_option(kv.split("=")[0], int(kv.split("=")[1]))
for bs in ([int(a) for a in sys.argv[4:]] or [1, 4, 8, 12, 16]):
    best = 1e9
    for _ in range(3):
        m.be.kv_truncate(base)
        m.forward_tree(list(range(1, bs + 1)), list(range(base, base + bs)), None, lm_head=True)
        best = min(best, m.be.counter("last_device_ns") / 1e3)
    print(f"bs={bs:3d}: {best:9.1f} us per verify batch = {best / n_layers:7.1f} us per layer (incl. lm_head share)", flush=True)
m.close()
