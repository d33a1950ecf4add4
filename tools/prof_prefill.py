"""Driver for prefill launch lists: an N-layer slice of a BASELINE model shape, prefill `ctx` tokens in chunks of `batch`."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b"
shape = synth.PRESETS[model]
shape.n_layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 128
shape.vocab_size = 4096
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=batch, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
p = synth.random_prompt(shape.vocab_size, ctx + 1)
m.prefill(p[:batch + 1], batch)
m.reset()
t0 = time.perf_counter()
m.prefill(p, batch)
dt = time.perf_counter() - t0
print(f"prefill {ctx} tokens, {shape.n_layers} layers, batch {batch}: {dt * 1e3:.1f} ms -> {ctx / dt:.0f} tok/s ({dt * 1e3 / shape.n_layers / (ctx / batch):.2f} ms per layer-chunk)")
m.close()
