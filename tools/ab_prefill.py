"""A/B of the batched attention kernels on a prefill: same slice of a BASELINE model, option attn_tile 0 / 1, wall clock
through ps_cuda_forward (chunks of `batch`), logits of the last token compared bit for bit."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b"
shape = synth.PRESETS[model]
shape.n_layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 128
shape.vocab_size = 4096
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=batch, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
p = synth.random_prompt(shape.vocab_size, ctx + 1)
ref = None
for tile in (0, 1, 0, 1):
    m.be.set_option("attn_tile", tile)
    m.reset()
    m.prefill(p[:batch + 1], batch)
    m.reset()
    t0 = time.perf_counter()
    m.prefill(p, batch)
    dt = time.perf_counter() - t0
    lg = np.asarray(m.forward([p[-1]])).reshape(-1)[-shape.vocab_size:].copy()
    if ref is None:
        ref = lg
    same = bool((lg.view(np.uint32) == ref.view(np.uint32)).all())
    print(f"attn_tile={tile}: prefill {ctx} tokens, {shape.n_layers} layers, batch {batch}: {dt * 1e3:.1f} ms ({dt * 1e3 / shape.n_layers / (ctx / batch):.3f} ms per layer-chunk), logits bit-equal to first run: {same}")
m.close()
