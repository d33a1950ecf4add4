"""ncu target: a few persistent decode steps (ps_k_step) of a model slice.  python tools/prof_step.py [model] [n_layers] [ctx] [steps]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
model = args[0] if len(args) > 0 else "llama-3.1-8b"
shape = synth.PRESETS[model]
if len(args) > 1:
    shape.n_layers = int(args[1])
ctx_len = int(args[2]) if len(args) > 2 else 2048
steps = int(args[3]) if len(args) > 3 else 6
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=128, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
for a in sys.argv[1:]:
    if a.startswith("--opt="):
        k, v = a[6:].split("=")
        m.be.set_option(k, int(v))
m.prefill(synth.random_prompt(shape.vocab_size, ctx_len + 1), 128)
ids = m.decode_greedy(1, steps)
print("ids", list(ids), "ms/step", m.be.counter("last_device_ns") / 1e6 / steps)
m.close()
