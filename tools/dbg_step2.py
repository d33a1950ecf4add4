"""Debug: persistent step kernel on 8B-like shapes with overrides: python tools/dbg_step2.py n_layers vocab dim ffn [persist]"""
import sys, dataclasses
import numpy as np
sys.path.insert(0, ".")
from powerserve_b200 import capi, gguf, synth
n_layers, vocab, dim, ffn = (int(x) for x in sys.argv[1:5])
persist = int(sys.argv[5]) if len(sys.argv) > 5 else 1
shape = dataclasses.replace(synth.PRESETS["llama-3.1-8b"], n_layers=n_layers, vocab_size=vocab, dim=dim, ffn_dim=ffn, n_ctx=4096)
tensors = synth.generate_tensors(shape, 0)
tmap = {n: gguf.GGUFTensor(n, t, tuple(s), np.ascontiguousarray(d).view(np.uint8).reshape(-1)) for n, t, s, d in tensors}
desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=16, n_ctx=4096)
m = capi.CudaModel(desc=desc, tensors=tmap)
m.be.set_option("tc", 0)
m.prefill(synth.random_prompt(shape.vocab_size, 9), 16)
m.be.set_option("persist", 0)
a = list(m.decode_greedy(1, 4)); m.be.kv_rollback(4)
m.be.set_option("persist", persist)
try:
    b = list(m.decode_greedy(1, 4))
except Exception as e:
    print(str(e)[-200:], "dbg", [m.be.counter(f"step_dbg{k}") for k in range(14)])
    raise
print(sys.argv[1:], "ok" if a == b else "MISMATCH", a, b, "step_error", m.be.counter("step_error"))
m.close()
