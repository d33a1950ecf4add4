"""Debug: group-synchronised decode attention vs the two-kernel path, per step, at several context lengths."""
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
for n_prompt in [int(a) for a in sys.argv[3:]]:
    prompt = synth.random_prompt(shape.vocab_size, n_prompt, seed=3)
    out = {}
    for ag in (1, 0):
        m.be.set_option("attn_group", ag)
        ids, lg = m.generate(prompt, 12, batch_size=128)
        out[ag] = (ids, lg)
    d = np.abs(out[0][1] - out[1][1]).max(axis=1)
    print(f"prompt {n_prompt}: ids equal {out[0][0] == out[1][0]}, per-step max |dlogit| {[float(x) for x in d]}, step_error {m.be.counter('step_error')}", flush=True)
m.close()
