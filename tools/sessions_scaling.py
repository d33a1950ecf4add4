"""Measurement: time ps_cuda_forward_sessions for several batch widths on an N-layer slice of a BASELINE shape."""
import sys
import numpy as np
sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth

model, n_layers = sys.argv[1], int(sys.argv[2])
shape = synth.PRESETS[model]
shape.n_layers = n_layers
shape.n_ctx = 4096
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=128, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
n_max = max([int(a) for a in sys.argv[3:]] or [64])
sids = [0] + [m.session_create() for _ in range(n_max - 1)]
for sid in sids:
    m.session_select(sid); m.reset()
    m.prefill(synth.random_prompt(shape.vocab_size, 33, seed=sid), 32)
m.session_select(0)
for n in ([int(a) for a in sys.argv[3:]] or [1, 2, 4, 8, 15, 16, 32, 64]):
    toks = [1] * n
    best = 1e9
    for _ in range(4):
        _, ids = m.forward_sessions(sids[:n], toks, want_logits=False)
        best = min(best, m.be.counter("last_device_ns") / 1e3)
    print(f"n={n:3d}: {best:9.1f} us per pass = {best / n_layers:7.1f} us per layer; aggregate {n / best * 1e6 * n_layers / 32:8.0f} tok/s at 32 layers", flush=True)
m.close()
